"""A few eager train steps (no CUDA graph) of one workload: the command ncu wraps.
usage: python profiles/run_steps.py [workload] [batch] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vae-los-angeles_b200")]
from src.models import DNA2RNAVAE, MultiModalVAE, RNA2DNAVAE  # noqa: E402
from vla_b200 import DeviceDataset, Trainer  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "rna2dna"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cls = {"rna2dna": RNA2DNAVAE, "dna2rna": DNA2RNAVAE, "multimodal": MultiModalVAE}[wl]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = cls(782, 572, 24, 20).to(dev).train()
ds = DeviceDataset.synthetic(B * 8, 782, 572, 24, dev, seed=1)
pg = None
if os.environ.get("VLA_FORCE_DP") == "1":      # world-1 data parallel: the peer-memory exchange + AdamW launch, for ncu
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29544")
    dist.init_process_group("gloo", rank=0, world_size=1)
    pg = dist.group.WORLD
tr = Trainer(model, ds, B, use_graph=False, process_group=pg)
for _ in range(steps):
    tr.step()
torch.cuda.synchronize()
print("losses", tr.losses())
