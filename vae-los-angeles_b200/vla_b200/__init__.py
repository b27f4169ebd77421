"""vla_b200: B200 (sm_100a) implementation of the vae-los-angeles hot path behind the reference's Python surface.

Everything numeric runs in libvla_b200.so (hand-written CUDA, C ABI in include/vla_b200.h); PyTorch provides
device memory, streams, autograd plumbing and torch.distributed.  There is no CPU fallback.
"""
from .core import VaeModule, Stack, KINDS
from .dp import Layout, allreduce_gradients
from .engine import DeviceDataset, FusedAdamW, Trainer
from .losses import fused_vae_loss
from .metrics import recon_metrics
from .population import Population, shard

__all__ = ["VaeModule", "Stack", "KINDS", "DeviceDataset", "FusedAdamW", "Trainer", "fused_vae_loss", "Layout",
           "allreduce_gradients", "Population", "shard", "recon_metrics"]
