from udifftext_b200.host.schedule import EpsWeighting  # noqa: F401
