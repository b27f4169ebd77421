"""Per-op timeline inside the row-chain launch of one train step (VLA_RC_TIMELINE=1: %globaltimer stamps of every CTA).
Usage (GPU box): python profiles/rowchain_timeline.py [workload] [batch] > gpurun_out/rowchain_timeline.log"""
import ctypes as C
import os
import sys

os.environ["VLA_RC_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-los-angeles_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from src.models import RNA2DNAVAE  # noqa: E402
from vla_b200 import DeviceDataset, Trainer, _lib  # noqa: E402

B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
torch.manual_seed(0)
m = RNA2DNAVAE(782, 572, 24, 20).cuda().train()
ds = DeviceDataset.synthetic(B * 16, 782, 572, 24, "cuda", seed=1)
tr = Trainer(m, ds, B, use_graph=False)
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
buf = (C.c_ulonglong * (148 * 32 * 4))()
kinds, subs = (C.c_int * 32)(), (C.c_int * 32)()
n = _lib.lib().vla_rowchain_timeline(m._ensure_core().handle, buf, kinds, subs)
t = np.frombuffer(buf, dtype=np.uint64).reshape(148, 32, 4).astype(np.int64)
ctas = min(148, (B + 127) // 128)
t = t[:ctas, :n]
t0 = t[t > 0].min()
KN = {1: "load_a", 2: "bn_act", 3: "gemm", 4: "epi"}
SN = {1: "latent", 2: "relu", 3: "loss", 4: "mask", 5: "latent_bwd", 6: "dgrad_enc"}
print(f"{ctas} CTAs; times in us relative to the first stamp, mean over CTAs")
print(f"{'op':3s} {'kind':18s} {'start':>8s} {'issued':>8s} {'dur':>7s}")
for i in range(n):
    k = kinds[i]
    if k == 3:
        a, b = t[:, i, 0], t[:, i, 1]
        print(f"{i:3d} {'gemm':18s} {(a.mean() - t0) / 1e3:8.2f} {(b.mean() - t0) / 1e3:8.2f} {(b - a).mean() / 1e3:7.2f}")
    elif k in (2, 4):
        a, b = t[:, i, 2], t[:, i, 3]
        name = KN[k] if k == 2 else "epi " + SN.get(subs[i], "?")
        print(f"{i:3d} {name:18s} {(a.mean() - t0) / 1e3:8.2f} {(b.mean() - t0) / 1e3:8.2f} {(b - a).mean() / 1e3:7.2f}")
print(f"span {(t.max() - t0) / 1e3:.2f} us")
