"""Graph-replay time of the fused step and of single kernels repeated in a graph (per-launch steady state)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-los-angeles_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from src.models import RNA2DNAVAE  # noqa: E402
from vla_b200 import DeviceDataset, Trainer, _lib  # noqa: E402
from vla_b200.core import _ptr, _stream  # noqa: E402

L = _lib.lib()
torch.manual_seed(0)
m = RNA2DNAVAE(782, 572, 24, 20).cuda().train()
ds = DeviceDataset.synthetic(4096 * 16, 782, 572, 24, "cuda", seed=1)
tr = Trainer(m, ds, 4096)
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    tr.graphs[0].replay()
e1.record()
torch.cuda.synchronize()
print(f"PDL {'off' if os.environ.get('VLA_NO_PDL') == '1' else 'on'}: fused step graph replay {e0.elapsed_time(e1) * 1e3 / 200:.1f} us/step")


def graph_time(fn, n_in_graph=20, replays=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n_in_graph):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (n_in_graph * replays)


core = tr.core
args = _lib.AdamWArgs(params=_ptr(core.arena), grads=_ptr(tr.grads), exp_avg=_ptr(tr.exp_avg), exp_avg_sq=_ptr(tr.exp_avg_sq),
                      lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=5)
print(f"adamw kernel alone: {graph_time(lambda: _lib.check(L.vla_adamw(core.handle, C.byref(args), _stream()), 'adamw')):.2f} us/launch")
from src.utils.directional_losses import rna2dna_loss  # noqa: E402
recon = torch.rand(4096, 572, device="cuda").clamp_(0.01, 0.99)
tgt = torch.rand(4096, 572, device="cuda")
mu = torch.randn(4096, 20, device="cuda"); lv = torch.randn(4096, 20, device="cuda")
from vla_b200.losses import _workspace  # noqa: E402
out = torch.empty(4, device="cuda"); ws = _workspace(recon.device, 1 << 16)
la = _lib.LossArgs(recon_a=None, a=None, dim_a=0, recon_b=_ptr(recon), b=_ptr(tgt), dim_b=572, recon_c=None, site=None, class_weights=None,
                   n_sites=0, mu=_ptr(mu), logvar=_ptr(lv), latent=20, batch=4096, beta=1e-3, gamma=1.0, g_recon_a=None, g_recon_b=None,
                   g_recon_c=None, g_mu=None, g_logvar=None, out=_ptr(out), workspace=_ptr(ws))
print(f"loss kernel alone (BCE 4096x572 + KL, no grads): {graph_time(lambda: _lib.check(L.vla_loss(C.byref(la), _stream()), 'loss')):.2f} us/launch")
x = torch.rand(4096, 572, device="cuda"); y = torch.empty_like(x)
print(f"torch copy 4096x572 fp32: {graph_time(lambda: y.copy_(x)):.2f} us/launch")
