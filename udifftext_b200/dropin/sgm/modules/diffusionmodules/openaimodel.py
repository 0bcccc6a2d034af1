from udifftext_b200.host.network import UnifiedUNetModel  # noqa: F401
