"""Per-launch device times of the eval forward model(a=x) at a large batch (BASELINE configs[3]): vla_profile_* event pairs
around every launch of one eager vla_forward, repeated.  Usage: python profiles/infer_profile.py [batch]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-los-angeles_b200"))
import torch
from src.models import MultiModalVAE
from vla_b200 import _lib
from vla_b200.core import _ptr, _stream

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MultiModalVAE(782, 572, 24, 20).to(dev).eval()
core = model._ensure_core()
x = torch.rand(batch, 782, device=dev)
outs = [torch.empty(batch, n, device=dev) for n in (782, 572, 24, 20, 20)]
L = _lib.lib()


def call(refresh):
    args = _lib.ForwardArgs(params=_ptr(core.arena), buffers=_ptr(core.buffers), counters=_ptr(core.counters), x_a=_ptr(x),
                            x_b=None, site=None, batch=batch, train=0, refresh_shadows=refresh, eps=None, keep_masks=None,
                            seed=1, offset=0, recon_a=_ptr(outs[0]), recon_b=_ptr(outs[1]), recon_c=_ptr(outs[2]),
                            mu=_ptr(outs[3]), logvar=_ptr(outs[4]))
    _lib.check(L.vla_forward(core.handle, C.byref(args), _stream()), "vla_forward")


call(1)
call(0)
torch.cuda.synchronize()
buf = (_lib.ProfEntry * 512)()
agg = {}
for rep in range(5):
    _lib.check(L.vla_profile_begin(core.handle), "begin")
    call(0)
    torch.cuda.synchronize()
    n = L.vla_profile_collect(core.handle, buf, 512)
    for i in range(n):
        a = agg.setdefault((i, buf[i].name.decode()), [0, 0.0, float(buf[i].flops), float(buf[i].bytes)])
        a[0] += 1; a[1] += float(buf[i].ms)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    call(0)
e1.record()
torch.cuda.synchronize()
rows = [dict(i=k[0], name=k[1], ms=v[1] / v[0], gflop=v[2] / 1e9, mb=v[3] / 1e6, gbps=v[3] / (v[1] / v[0] * 1e-3) / 1e9 if v[1] else 0,
             tflops=v[2] / (v[1] / v[0] * 1e-3) / 1e12 if v[1] else 0) for k, v in sorted(agg.items())]
print(json.dumps(dict(batch=batch, total_ms=e0.elapsed_time(e1) / 5, launches=rows), indent=1))
