"""Decoder stacks (reference src/models/decoders.py:8-50): parameter containers; see encoders.py."""
from vla_b200.core import Stack


class DecoderA(Stack):
    """latent -> 128 -> ReLU -> output_dim (linear RNA reconstruction)."""

    def __init__(self, latent_dim, output_dim):
        super().__init__("dec", "A", output_dim, latent_dim)


class DecoderB(Stack):
    """latent -> 256 -> ReLU -> 512 -> ReLU -> output_dim -> sigmoid (beta values in (0, 1))."""

    def __init__(self, latent_dim, output_dim):
        super().__init__("dec", "B", output_dim, latent_dim)


class DecoderC(Stack):
    """latent -> 64 -> ReLU -> n_sites logits."""

    def __init__(self, latent_dim, n_sites):
        super().__init__("dec", "C", n_sites, latent_dim)
