from .autoencoder import AutoencoderKL, AutoencoderKLInferenceWrapper  # noqa: F401
from .diffusion import DiffusionEngine  # noqa: F401
