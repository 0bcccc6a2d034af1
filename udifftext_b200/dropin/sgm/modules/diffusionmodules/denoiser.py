from udifftext_b200.host.schedule import Denoiser, DiscreteDenoiser  # noqa: F401
