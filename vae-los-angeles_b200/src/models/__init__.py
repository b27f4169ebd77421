"""Models package: the names the reference exports (src/models/__init__.py:4-8) for the VAE hot path."""
from .vae import MultiModalVAE, reparameterize
from .encoders import EncoderA, EncoderB, EncoderC
from .decoders import DecoderA, DecoderB, DecoderC
from .directional_vae import RNA2DNAVAE, DNA2RNAVAE

__all__ = [
    'MultiModalVAE', 'reparameterize',
    'EncoderA', 'EncoderB', 'EncoderC',
    'DecoderA', 'DecoderB', 'DecoderC',
    'RNA2DNAVAE', 'DNA2RNAVAE',
]
