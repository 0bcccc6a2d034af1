from udifftext_b200.host.schedule import DiscreteSampling  # noqa: F401
