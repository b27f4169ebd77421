"""Directional autoencoders behind the reference's interface (reference src/models/directional_ae.py:10-134)."""
from vla_b200.core import AeModule

from .vae import _StackFactory


class RNA2DNAAE(_StackFactory, AeModule):
    """RNA + site -> DNA methylation, deterministic latent.  forward(rna=None, site=None) -> (recon_dna, latent)."""

    kind = "rna2dna_ae"

    def __init__(self, rna_dim, dna_dim, n_sites, latent_dim, embed_dim=32):
        super().__init__(rna_dim, dna_dim, n_sites, latent_dim, embed_dim)

    def forward(self, rna=None, site=None):
        if rna is None and site is None:
            return None, None
        return self._run_ae(rna, None, site)


class DNA2RNAAE(_StackFactory, AeModule):
    """DNA methylation + site -> RNA, deterministic latent.  forward(dna=None, site=None) -> (recon_rna, latent)."""

    kind = "dna2rna_ae"

    def __init__(self, rna_dim, dna_dim, n_sites, latent_dim, embed_dim=32):
        super().__init__(rna_dim, dna_dim, n_sites, latent_dim, embed_dim)

    def forward(self, dna=None, site=None):
        if dna is None and site is None:
            return None, None
        return self._run_ae(None, dna, site)
