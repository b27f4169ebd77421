"""Fused loss (values + gradients in one launch) behind the reference's loss signatures.

`fused_vae_loss` backs `vae_loss` (reference src/utils/losses.py:8-46) and the directional losses
(src/utils/directional_losses.py:8-30, 33-55): sum-reduced MSE + BCE, weighted cross-entropy, KL, and
total = recon + gamma * class + beta * kld, with one device->host read for the Python floats instead
of the reference's three `.item()` syncs.
"""
import ctypes as C

import torch

from . import _lib
from .core import _ptr, _stream

_workspaces = {}


def _workspace(device, nbytes):
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _prep(t, dtype, device, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"vla_b200: loss input `{name}` is on {t.device}; there is no CPU fallback")
    if t.device != device:
        raise RuntimeError(f"vla_b200: loss inputs live on different devices ({t.device} vs {device})")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class _LossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, recon_a, a, recon_b, b, recon_c, site, mu, logvar, beta, gamma, class_weights):
        L = _lib.lib()
        ref = next(t for t in (recon_a, recon_b, recon_c, mu) if t is not None)
        dev = ref.device
        recon_a = _prep(recon_a, torch.float32, dev, "recon_a")
        recon_b = _prep(recon_b, torch.float32, dev, "recon_b")
        recon_c = _prep(recon_c, torch.float32, dev, "recon_c")
        mu = _prep(mu, torch.float32, dev, "mu")
        logvar = _prep(logvar, torch.float32, dev, "logvar")
        a = _prep(a, torch.float32, dev, "a")
        b = _prep(b, torch.float32, dev, "b")
        site = _prep(site, torch.long, dev, "site")
        cw = _prep(class_weights, torch.float32, dev, "class_weights")
        batch = ref.shape[0]
        dim_a = recon_a[0].numel() if recon_a is not None else 0
        dim_b = recon_b[0].numel() if recon_b is not None else 0
        n_sites = recon_c.shape[1] if recon_c is not None else 0
        latent = mu.shape[1] if mu is not None else 0
        if recon_a is not None and a.numel() != recon_a.numel():
            raise RuntimeError("vla_b200: recon_a / a shape mismatch")
        if recon_b is not None and b.numel() != recon_b.numel():
            raise RuntimeError("vla_b200: recon_b / b shape mismatch")
        needs = [t is not None and t.requires_grad for t in (recon_a, recon_b, recon_c, mu, logvar)]
        # ctx.needs_input_grad indices: recon_a 0, recon_b 2, recon_c 4, mu 6, logvar 7
        want = [ctx.needs_input_grad[i] for i in (0, 2, 4, 6, 7)]
        grads = [torch.empty_like(t) if (w and t is not None) else None
                 for t, w in zip((recon_a, recon_b, recon_c, mu, logvar), want)]
        out = torch.empty(4, dtype=torch.float32, device=dev)
        nbytes = L.vla_loss_workspace_bytes(batch, dim_a, dim_b, n_sites, latent)
        ws = _workspace(dev, nbytes)
        args = _lib.LossArgs(
            recon_a=_ptr(recon_a), a=_ptr(a), dim_a=dim_a, recon_b=_ptr(recon_b), b=_ptr(b), dim_b=dim_b,
            recon_c=_ptr(recon_c), site=_ptr(site), class_weights=_ptr(cw), n_sites=n_sites,
            mu=_ptr(mu), logvar=_ptr(logvar), latent=latent, batch=batch, beta=float(beta), gamma=float(gamma),
            g_recon_a=_ptr(grads[0]), g_recon_b=_ptr(grads[1]), g_recon_c=_ptr(grads[2]), g_mu=_ptr(grads[3]),
            g_logvar=_ptr(grads[4]), out=_ptr(out), workspace=_ptr(ws))
        with torch.cuda.device(dev):
            _lib.check(L.vla_loss(C.byref(args), _stream()), "vla_loss")
        ctx.grads = grads
        del needs
        return out

    @staticmethod
    def backward(ctx, gout):
        # only d/d(total) flows (the recon / class / kld entries are reported values): every stored gradient is scaled by it
        # in place, all tensors in ONE launch of the library (vla_scale_inplace) instead of one ATen multiply per tensor
        live = [t for t in ctx.grads if t is not None]
        if live:
            g = gout.contiguous()
            L = _lib.lib()
            ptrs = (C.c_void_p * len(live))(*[t.data_ptr() for t in live])
            counts = (C.c_longlong * len(live))(*[t.numel() for t in live])
            with torch.cuda.device(g.device):
                _lib.check(L.vla_scale_inplace(ptrs, counts, len(live), _ptr(g), _stream()), "vla_scale_inplace")
        gs = ctx.grads
        return (gs[0], None, gs[1], None, gs[2], None, gs[3], gs[4], None, None, None)


def fused_vae_loss(recon_a=None, a=None, recon_b=None, b=None, recon_c=None, site=None, mu=None, logvar=None,
                   beta=1e-3, gamma=1.0, class_weights=None, kl=True):
    """Returns (total 0-d tensor with grad, stats tensor [total, recon, class, kld] on the device).
    kl=False: reconstruction terms only (the autoencoder losses, reference src/utils/ae_losses.py)."""
    if not kl:
        mu = logvar = None
    elif mu is None or logvar is None:
        raise RuntimeError("vla_b200: mu and logvar are required")
    if recon_a is None or a is None:
        recon_a = a = None
    if recon_b is None or b is None:
        recon_b = b = None
    if recon_c is None or site is None:
        recon_c = site = None
    out = _LossFunction.apply(recon_a, a, recon_b, b, recon_c, site, mu, logvar, beta, gamma, class_weights)
    return out[0], out.detach()
