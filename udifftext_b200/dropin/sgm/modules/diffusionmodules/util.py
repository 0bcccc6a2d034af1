from udifftext_b200.host.schedule import make_beta_schedule  # noqa: F401
