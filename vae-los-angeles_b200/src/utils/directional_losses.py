"""Losses of the directional VAEs (reference src/utils/directional_losses.py:8-55) on the fused loss kernel."""
from vla_b200.losses import fused_vae_loss


def rna2dna_loss(recon_dna, dna, mu, logvar, beta=1e-3):
    """BCE_sum(recon_dna, dna) + beta * KL.  Returns (total tensor, recon float, kld float)."""
    total, stats = fused_vae_loss(recon_b=recon_dna, b=dna, mu=mu, logvar=logvar, beta=beta)
    _, recon, _, kld = stats.tolist()
    return total, recon, kld


def dna2rna_loss(recon_rna, rna, mu, logvar, beta=1e-3):
    """MSE_sum(recon_rna, rna) + beta * KL.  Returns (total tensor, recon float, kld float)."""
    total, stats = fused_vae_loss(recon_a=recon_rna, a=rna, mu=mu, logvar=logvar, beta=beta)
    _, recon, _, kld = stats.tolist()
    return total, recon, kld
