from udifftext_b200.host.autoencoder import AutoencoderKL, AutoencoderKLInferenceWrapper, DiagonalGaussianDistribution  # noqa: F401
