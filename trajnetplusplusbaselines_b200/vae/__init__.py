from .vae import VAE, VAEDecoder, VAEEncoder, VAEPredictor, sample_multivariate_distribution  # noqa: F401
