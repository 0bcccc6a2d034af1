from udifftext_b200.host.predictor import ParseqPredictor  # noqa: F401
