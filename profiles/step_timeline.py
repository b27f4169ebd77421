"""Per-phase breakdown of the whole-step kernel from its in-kernel %globaltimer stamps (one replayed step).
usage: python profiles/step_timeline.py [workload] [batch]
Stamps per unit: 0 start, 1 deps resolved, 3 first operands landed, 4 MMAs issued, 5 accumulator ready, 6 published."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["VLA_FUSED_STEP"] = "1"

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vae-los-angeles_b200")]
from src.models import DNA2RNAVAE, MultiModalVAE, RNA2DNAVAE  # noqa: E402
from vla_b200 import DeviceDataset, Trainer, _lib  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "rna2dna"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cls = {"rna2dna": RNA2DNAVAE, "dna2rna": DNA2RNAVAE, "multimodal": MultiModalVAE}[wl]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = cls(782, 572, 24, 20).to(dev).train()
ds = DeviceDataset.synthetic(B * 8, 782, 572, 24, dev, seed=1)
tr = Trainer(model, ds, B)
for _ in range(5):
    tr.step()
torch.cuda.synchronize()
L = _lib.lib()
_lib.check(L.vla_step_timeline(tr.core.handle, 1), "timeline on")
for _ in range(3):
    tr.step()
torch.cuda.synchronize()
n_units = L.vla_step_timeline_units(tr.core.handle)
n_ph = L.vla_step_timeline_phases(tr.core.handle)
buf = (C.c_ulonglong * (8 * n_units))()
L.vla_step_timeline_read(tr.core.handle, buf, n_units)
t = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 8).astype(np.int64)
t0 = t[:, 0].min()
print(f"{wl} batch {B}: step {(t[:, 6].max() - t0) / 1e3:.1f} us, {n_units} units")
print(f"{'phase':14s} {'units':>5s} {'first':>7s} {'last':>7s} | {'wait':>6s} {'load1':>6s} {'mma':>6s} {'drain':>6s} {'epi_w2':>6s} {'tail':>6s} | {'unit':>6s}  (us, means)")
for p in range(n_ph):
    name = C.create_string_buffer(48)
    nu, ub = C.c_int(), C.c_int()
    L.vla_step_phase_info(tr.core.handle, p, name, C.byref(nu), C.byref(ub), None, None)
    r = t[ub.value:ub.value + nu.value]
    gemm = (r[:, 3] > 0).all() and (r[:, 3] >= r[:, 0]).all() and name.value.decode().startswith(("gemm", "dgrad", "wgrad"))
    f = lambda a, b: float((r[:, a] - r[:, b]).mean()) / 1e3
    if gemm:
        cols = f"{f(1, 0):6.2f} {f(3, 1):6.2f} {f(4, 3):6.2f} {f(5, 4):6.2f} {f(2, 5):6.2f} {f(6, 2):6.2f}"
    elif name.value.decode() in ("ingest", "adamw"):
        cols = f"{f(1, 0):6.2f} {f(2, 1):6.2f} {f(3, 2):6.2f} {f(4, 3):6.2f} {f(5 if name.value.decode() == 'ingest' else 4, 4):6.2f} {f(6, 5 if name.value.decode() == 'ingest' else 4):6.2f}"
    else:
        cols = f"{f(1, 0):6.2f} {'':6s} {'':6s} {'':6s} {f(6, 1):6.2f} {'':6s}"
    print(f"{name.value.decode():14s} {nu.value:5d} {(r[:, 0].min() - t0) / 1e3:7.2f} {(r[:, 6].max() - t0) / 1e3:7.2f} | {cols} | {f(6, 0):6.2f}")

if os.environ.get("VLA_TL_UNITS"):
    # per-unit rows of one phase: which = phase name; shows first vs later units on the same CTA
    which = os.environ["VLA_TL_UNITS"]
    for p in range(n_ph):
        name = C.create_string_buffer(48)
        nu, ub = C.c_int(), C.c_int()
        L.vla_step_phase_info(tr.core.handle, p, name, C.byref(nu), C.byref(ub), None, None)
        if name.value.decode() != which:
            continue
        r = t[ub.value:ub.value + nu.value]
        G = 148
        first = r[:min(G, nu.value)]
        later = r[G:]
        for lab, rr in (("first unit on its CTA", first), ("later units", later)):
            if len(rr):
                d = lambda a, b: float((rr[:, a] - rr[:, b]).mean()) / 1e3
                print(f"{which} {lab:22s} n={len(rr):4d}: 1-0 {d(1,0):6.2f}  2-1 {d(2,1):6.2f}  3-2 {d(3,2):6.2f}  4-3 {d(4,3):6.2f}  5-4 {d(5,4):6.2f}  6-5 {d(6,5):6.2f}  total {d(6,0):6.2f}")
