from udifftext_b200.host.autoencoder import DiagonalGaussianDistribution  # noqa: F401
