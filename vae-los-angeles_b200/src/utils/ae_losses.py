"""Losses of the directional autoencoders (reference src/utils/ae_losses.py:8-39) on the fused loss kernel:
reconstruction only, no KL term."""
from vla_b200.losses import fused_vae_loss


def rna2dna_ae_loss(recon_dna, dna):
    """BCE_sum(recon_dna, dna).  Returns (total tensor, recon float)."""
    total, stats = fused_vae_loss(recon_b=recon_dna, b=dna, kl=False)
    return total, stats[1].item()


def dna2rna_ae_loss(recon_rna, rna):
    """MSE_sum(recon_rna, rna).  Returns (total tensor, recon float)."""
    total, stats = fused_vae_loss(recon_a=recon_rna, a=rna, kl=False)
    return total, stats[1].item()
