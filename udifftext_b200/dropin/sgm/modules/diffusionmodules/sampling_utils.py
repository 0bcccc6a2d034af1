from udifftext_b200.host.schedule import NoDynamicThresholding, to_d  # noqa: F401
