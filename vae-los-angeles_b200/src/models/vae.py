"""Tri-modal VAE behind the reference's interface (reference src/models/vae.py:11-79)."""
import torch

from vla_b200.core import VaeModule

from .decoders import DecoderA, DecoderB, DecoderC
from .encoders import EncoderA, EncoderB, EncoderC


def reparameterize(mu, logvar):
    """z = mu + eps * exp(logvar / 2), eps ~ N(0, 1) (reference src/models/vae.py:11-15).

    Standalone helper kept for API completeness; the VAE modules do this inside the fused latent kernel."""
    return torch.addcmul(mu, torch.randn_like(logvar), torch.exp(0.5 * logvar))


class _StackFactory:
    ENC = {"A": EncoderA, "B": EncoderB, "C": EncoderC}
    DEC = {"A": DecoderA, "B": DecoderB, "C": DecoderC}

    def _make_stack(self, role, t, feature_dim):
        if role == "enc":
            if t == "C":
                return EncoderC(feature_dim, self.latent_dim, embed_dim=self.embed_dim)
            return self.ENC[t](feature_dim, self.latent_dim)
        return self.DEC[t](self.latent_dim, feature_dim)


class MultiModalVAE(_StackFactory, VaeModule):
    """RNA (a) + DNA methylation (b) + primary site encoders, mean-fused latent, three decoders.

    forward(a=None, b=None, site=None) -> (out_a, out_b, out_c, mu, logvar); any subset of modalities may be
    given, all three decoders always run; no modality -> five Nones."""

    kind = "multimodal"

    def __init__(self, input_dim_a, input_dim_b, n_sites, latent_dim, embed_dim=32):
        super().__init__(input_dim_a, input_dim_b, n_sites, latent_dim, embed_dim)

    def forward(self, a=None, b=None, site=None):
        if a is None and b is None and site is None:
            return None, None, None, None, None
        return self._run(a, b, site)
