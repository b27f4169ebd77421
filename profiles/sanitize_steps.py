"""Small train steps, forwards, the functional loss, the metrics and the gather kernel of every model kind: the command
compute-sanitizer wraps (one tool per gpurun call).  usage: python profiles/sanitize_steps.py [batch]
VLA_FORCE_DP=1 adds the world-1 peer-memory exchange + AdamW launch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vae-los-angeles_b200")]
from src.models import DNA2RNAVAE, MultiModalVAE, RNA2DNAVAE  # noqa: E402
from src.models.directional_ae import RNA2DNAAE  # noqa: E402
from src.utils.losses import vae_loss  # noqa: E402
from vla_b200 import DeviceDataset, Trainer, recon_metrics  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 300          # ragged: 2 full row blocks + 44 rows
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pg = None
if os.environ.get("VLA_FORCE_DP") == "1":
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29545")
    dist.init_process_group("gloo", rank=0, world_size=1)
    pg = dist.group.WORLD
ds = DeviceDataset.synthetic(B * 2 + 7, 782, 572, 24, dev, seed=1)
for cls in (RNA2DNAVAE, DNA2RNAVAE, MultiModalVAE, RNA2DNAAE):
    model = cls(782, 572, 24, 20).to(dev).train()
    tr = Trainer(model, ds, B, use_graph=False, process_group=pg)
    for _ in range(2):
        tr.step()
    if pg is None:
        tr.run_epoch()                                       # includes the 7-row tail
    torch.cuda.synchronize()
    print(cls.__name__, "losses", tr.losses())
    tr.close()
# the per-call autograd path of the scripts (forward, functional loss, backward) and eval inference
m = MultiModalVAE(782, 572, 24, 20).to(dev).train()
a, b, s = ds.tpm[:33], ds.beta[:33], ds.site[:33]
ra, rb, rc, mu, lv = m(a=a, b=b, site=s)
total, *_ = vae_loss(ra, a, rb, b, rc, s, mu, lv, beta=1e-3, gamma=1.0)
total.backward()
m.eval()
with torch.no_grad():
    out = m(a=ds.tpm[:1500])
print("autograd path", float(total), "metrics", sorted(recon_metrics(ds.tpm[:1500], out[0]).items())[:3])
slot = DeviceDataset.synthetic(64, 782, 572, 24, dev, seed=2)
ds.gather_into(torch.randperm(len(ds), device=dev)[:64], slot)
torch.cuda.synchronize()
print("sanitize run complete")
