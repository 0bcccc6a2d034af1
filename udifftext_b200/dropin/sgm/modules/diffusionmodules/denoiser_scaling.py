from udifftext_b200.host.schedule import EpsScaling  # noqa: F401
