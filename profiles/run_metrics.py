"""A few launches of the reconstruction-metrics kernel at BASELINE configs[3]'s batch: the command ncu wraps.
usage: python profiles/run_metrics.py [rows] [dim]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vae-los-angeles_b200")]
from vla_b200 import recon_metrics  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 782
g = torch.Generator(device="cuda"); g.manual_seed(0)
t = torch.rand(rows, dim, device="cuda", generator=g)
p = t + 0.1 * torch.randn(rows, dim, device="cuda", generator=g)
for _ in range(3):
    r = recon_metrics(t, p, per_sample=True)
torch.cuda.synchronize()
print({k: v for k, v in r.items() if not k.startswith("_")})
