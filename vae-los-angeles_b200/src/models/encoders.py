"""Encoder stacks of the VAE: parameter containers with the reference's constructor signatures and
state_dict names (reference src/models/encoders.py:8-61).  Their arithmetic is executed by the owning
VAE module as fused tcgen05 GEMM + BatchNorm/ReLU/dropout kernels (vla_b200.core)."""
from vla_b200.core import Stack


class EncoderA(Stack):
    """RNA expression: input_dim -> 128 (BatchNorm, ReLU, Dropout 0.1) -> {mu, logvar}."""

    def __init__(self, input_dim, latent_dim):
        super().__init__("enc", "A", input_dim, latent_dim)


class EncoderB(Stack):
    """DNA methylation: input_dim -> 512 -> 256 (BatchNorm, ReLU, Dropout 0.1 each) -> {mu, logvar}."""

    def __init__(self, input_dim, latent_dim):
        super().__init__("enc", "B", input_dim, latent_dim)


class EncoderC(Stack):
    """Primary site: Embedding(n_sites, embed_dim) -> {mu, logvar}."""

    def __init__(self, n_sites, latent_dim, embed_dim=32):
        super().__init__("enc", "C", n_sites, latent_dim, embed_dim)
