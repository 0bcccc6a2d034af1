from udifftext_b200.host.loss import FullLoss  # noqa: F401
