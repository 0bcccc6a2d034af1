from udifftext_b200.host.engine import DiffusionEngine  # noqa: F401
