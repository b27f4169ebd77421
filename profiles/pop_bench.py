"""Population throughput lines (bench.population_throughput) on their own.  usage: python profiles/pop_bench.py [cmp] [B:n:steps ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vae-los-angeles_b200"))
import torch
import bench

dev = torch.device("cuda:0")
cases = [a for a in sys.argv[1:] if ":" in a] or ["32:40:60", "4096:40:12"]
for c in cases:
    B, n, steps = (int(x) for x in c.split(":"))
    print(json.dumps(bench.population_throughput(dev, B, n, steps, 3, compare="cmp" in sys.argv)), flush=True)
