"""Data-parallel plumbing: flat-arena layout queries (no GPU needed) and the one collective of the step.

The reference has no distributed code at all; the partitioning is new: batch-sharded replicas, one
process per GPU, and ONE all-reduce(SUM) per step over [flat gradient arena | 4 loss scalars].  SUM, not mean,
because every reference loss is reduction='sum' (src/utils/losses.py:31-42): the summed shard gradients are the
gradients of the concatenated batch, up to BatchNorm, which normalises with per-shard statistics."""
import ctypes as C

import torch

from . import _lib


class Layout:
    """Parameter-arena layout of a model kind (names, offsets, shapes), obtained from the library without a device."""

    def __init__(self, kind, dim_a, dim_b, n_sites, latent, embed=32):
        L = _lib.lib()
        cfg = _lib.Config(_lib.KIND[kind], dim_a, dim_b, n_sites, latent, embed)
        handle = C.c_void_p()
        _lib.check(L.vla_model_create_layout_only(C.byref(cfg), C.byref(handle)), "vla_model_create_layout_only")
        try:
            self.n_params = L.vla_param_count(handle)
            self.n_buffers = L.vla_buffer_count(handle)
            self.entries = []
            info = _lib.TensorInfo()
            for i in range(L.vla_num_tensors(handle)):
                _lib.check(L.vla_tensor_info(handle, i, C.byref(info)), "vla_tensor_info")
                self.entries.append((info.name.decode(), info.kind, info.offset, tuple(info.shape[: info.ndim])))
        finally:
            L.vla_model_destroy(handle)
        self.params = {n: (off, shape) for n, kind, off, shape in self.entries if kind == _lib.TENSOR_PARAM}

    def pack(self, named, extra=0, dtype=torch.float32, device="cpu"):
        """{name: tensor} -> flat [n_params + extra]; missing entries stay zero."""
        flat = torch.zeros(self.n_params + extra, dtype=dtype, device=device)
        for name, (off, shape) in self.params.items():
            if name in named and named[name] is not None:
                t = torch.as_tensor(named[name], dtype=dtype, device=device)
                flat[off:off + t.numel()] = t.reshape(-1)
        return flat

    def unpack(self, flat):
        out = {}
        for name, (off, shape) in self.params.items():
            n = 1
            for s in shape:
                n *= s
            out[name] = flat[off:off + n].view(shape)
        return out


def allreduce_gradients(flat, group=None):
    """The step's single collective: in-place SUM over ranks of [gradients | loss scalars]."""
    import torch.distributed as dist
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
