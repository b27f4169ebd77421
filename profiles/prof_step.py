"""Small eager (no CUDA graph) run of the fused train step for ncu: 3 steps of the BASELINE configs[1] workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-los-angeles_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from src.models import DNA2RNAVAE, MultiModalVAE, RNA2DNAVAE  # noqa: E402
from vla_b200 import DeviceDataset, Trainer  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rna2dna"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cls = {"rna2dna": RNA2DNAVAE, "dna2rna": DNA2RNAVAE, "multimodal": MultiModalVAE}[workload]
torch.manual_seed(0)
m = cls(782, 572, 24, 20).cuda().train()
ds = DeviceDataset.synthetic(4096 * 4, 782, 572, 24, "cuda", seed=1)
tr = Trainer(m, ds, 4096, use_graph=False)
for _ in range(steps):
    tr.step()
torch.cuda.synchronize()
print("losses", tr.losses())
