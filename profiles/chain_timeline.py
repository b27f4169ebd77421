"""Per-phase timeline inside the chain launches of one train step (Trainer.timeline(): %globaltimer stamps of every CTA).
Usage (GPU box): python profiles/chain_timeline.py [workload] [batch] > gpurun_out/chain_timeline.log"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-los-angeles_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from src.models import DNA2RNAVAE, MultiModalVAE, RNA2DNAVAE  # noqa: E402
from vla_b200 import DeviceDataset, Trainer  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rna2dna"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cls = {"rna2dna": RNA2DNAVAE, "dna2rna": DNA2RNAVAE, "multimodal": MultiModalVAE}[workload]
torch.manual_seed(0)
m = cls(782, 572, 24, 20).cuda().train()
ds = DeviceDataset.synthetic(B * 16, 782, 572, 24, "cuda", seed=1)
tr = Trainer(m, ds, B)
for _ in range(10):
    tr.step()
torch.cuda.synchronize()
for rep in range(2):
    chains = tr.timeline()
for c in chains:
    print(f"== {c['name']}: {c['ctas']} CTAs, {c['span_us']:.1f} us, {c['flops'] / 1e9:.2f} GFLOP, {c['bytes'] / 1e6:.1f} MB")
    print(f"   {'phase':16s} {'start':>7s} {'busy':>5s} {'operands':>8s} {'mma_iss':>8s} {'acc_rdy':>8s} {'epi_1st':>8s} {'epi_last':>8s} {'bar_in':>7s} {'bar_out':>7s} {'span':>7s}   (us, relative to the CTA's phase start)")
    for p in c["phases"]:
        print(f"   {p['name']:16s} {p['start_us']:7.2f} {p['busy_ctas']:5d} {p['operands_us']:8.2f} {p['mma_issued_us']:8.2f} {p['acc_ready_us']:8.2f} "
              f"{p['epi_first_us']:8.2f} {p['epi_last_us']:8.2f} {p['barrier_in_us']:7.2f} {p['barrier_out_us']:7.2f} {p['span_us']:7.2f}")
