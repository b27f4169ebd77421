"""Models package: the names the reference exports (src/models/__init__.py:4-8) for the VAE hot path."""
from .vae import MultiModalVAE, reparameterize
from .encoders import EncoderA, EncoderB, EncoderC
from .decoders import DecoderA, DecoderB, DecoderC
from .directional_vae import RNA2DNAVAE, DNA2RNAVAE
from .directional_ae import RNA2DNAAE, DNA2RNAAE   # (the reference imports these from the submodule: vae_cross_modality_cv.py:34)

__all__ = [
    'MultiModalVAE', 'reparameterize',
    'EncoderA', 'EncoderB', 'EncoderC',
    'DecoderA', 'DecoderB', 'DecoderC',
    'RNA2DNAVAE', 'DNA2RNAVAE',
    'RNA2DNAAE', 'DNA2RNAAE',
]
