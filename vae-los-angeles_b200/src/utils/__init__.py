"""Utils package (reference src/utils/__init__.py:4)."""
from .losses import vae_loss

__all__ = ['vae_loss']
