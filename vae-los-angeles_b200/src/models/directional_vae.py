"""Directional VAEs behind the reference's interface (reference src/models/directional_vae.py:12-111)."""
from vla_b200.core import VaeModule

from .vae import _StackFactory


class RNA2DNAVAE(_StackFactory, VaeModule):
    """RNA + site -> DNA methylation.  forward(rna=None, site=None) -> (recon_dna, mu, logvar)."""

    kind = "rna2dna"

    def __init__(self, rna_dim, dna_dim, n_sites, latent_dim, embed_dim=32):
        super().__init__(rna_dim, dna_dim, n_sites, latent_dim, embed_dim)

    def forward(self, rna=None, site=None):
        if rna is None and site is None:
            return None, None, None
        return self._run(rna, None, site)


class DNA2RNAVAE(_StackFactory, VaeModule):
    """DNA methylation + site -> RNA.  forward(dna=None, site=None) -> (recon_rna, mu, logvar)."""

    kind = "dna2rna"

    def __init__(self, rna_dim, dna_dim, n_sites, latent_dim, embed_dim=32):
        super().__init__(rna_dim, dna_dim, n_sites, latent_dim, embed_dim)

    def forward(self, dna=None, site=None):
        if dna is None and site is None:
            return None, None, None
        return self._run(None, dna, site)
