"""Per-launch times of one lock-step population step (VLA_GROUP_PROF=1 makes vla_train_step_group print an event-pair time per
merged launch).  usage: VLA_GROUP_PROF=1 python profiles/pop_profile.py [batch] [members]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vae-los-angeles_b200"))
os.environ.setdefault("VLA_GROUP_PROF", "1")
import torch
import bench
from vla_b200 import DeviceDataset, Population

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
ds = DeviceDataset.synthetic(B * 8, 782, 572, 24, dev, seed=9)
pop = Population(bench.population_specs(n), ds, B, device=dev, grouped=True, use_graph=False)
for i in range(3):
    sys.stderr.write(f"--- step {i}\n")
    pop.step(1)
pop.synchronize()
pop.close()
