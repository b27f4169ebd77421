"""`vae_loss` with the reference's signature and return convention (reference src/utils/losses.py:8-46),
computed by the fused loss kernel: one launch for MSE + BCE + weighted CE + KL and their gradients, one
16-byte device->host read for the three Python floats."""
from vla_b200.losses import fused_vae_loss


def vae_loss(recon_a, a, recon_b, b, recon_c, site, mu, logvar, beta=1e-3, gamma=1.0, class_weights=None):
    """total = [MSE_sum(recon_a, a) + BCE_sum(recon_b, b)] + gamma * CE_sum(recon_c, site; class_weights)
              + beta * KL(mu, logvar).

    Returns (total 0-d tensor with grad, recon float, class float, kld float).  A term whose tensors are None
    is skipped (the reference raises in that case when it calls `.item()` on the int 0; this is strictly
    more permissive)."""
    total, stats = fused_vae_loss(recon_a, a, recon_b, b, recon_c, site, mu, logvar, beta=beta, gamma=gamma,
                                  class_weights=class_weights)
    _, recon, cls, kld = stats.tolist()
    return total, recon, cls, kld
