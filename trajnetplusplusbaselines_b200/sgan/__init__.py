from .sgan import SGAN, LSTMDiscriminator, LSTMGenerator, SGANPredictor, drop_distant, get_noise, make_mlp  # noqa: F401
