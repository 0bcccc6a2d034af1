from udifftext_b200.host.schedule import Discretization, LegacyDDPMDiscretization  # noqa: F401
