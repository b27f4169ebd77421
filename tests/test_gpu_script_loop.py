"""The call pattern of the reference's train_rna2dna.py / train_dna2rna.py (which cannot be shipped to the GPU box) against
the drop-in `src` surface: DataLoader at batch 32 with shuffle + drop_last, model.train()/eval(), torch.optim.AdamW,
ReduceLROnPlateau, beta warm-up, validation under no_grad with a ragged last batch, torch.save / load_state_dict."""
import io

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from oracle import vae_oracle as vo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("direction", ["rna2dna", "dna2rna"])
def test_training_loop_like_the_reference_scripts(direction):
    from src.config import Config
    from src.data import MultiModalDataset
    from src.models import DNA2RNAVAE, RNA2DNAVAE
    from src.utils.directional_losses import dna2rna_loss, rna2dna_loss
    dims = dict(A=782, B=572, S=24, L=20, E=32)
    tpm, beta_v, site = vo.synthetic_batch(330, dims, seed=5)
    train = MultiModalDataset.from_numpy(tpm[:256], beta_v[:256], site[:256])
    val = MultiModalDataset.from_numpy(tpm[256:], beta_v[256:], site[256:])            # 74 rows: ragged last batch (10)
    train_dl = DataLoader(train, batch_size=Config.BATCH_SIZE, shuffle=True, drop_last=True)
    val_dl = DataLoader(val, batch_size=Config.BATCH_SIZE, shuffle=False)
    torch.manual_seed(0)
    dev = torch.device("cuda")
    cls = RNA2DNAVAE if direction == "rna2dna" else DNA2RNAVAE
    model = cls(dims["A"], dims["B"], dims["S"], Config.LATENT_DIM).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=Config.LEARNING_RATE, weight_decay=Config.WEIGHT_DECAY)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=Config.LR_SCHEDULER_FACTOR,
                                                       patience=Config.LR_SCHEDULER_PATIENCE)

    def run(dl, epoch, training):
        model.train(training)
        beta = min(1.0, epoch / Config.BETA_WARMUP_EPOCHS) * Config.BETA_START
        total = 0.0
        ctx = torch.enable_grad() if training else torch.no_grad()
        with ctx:
            for t, b, s in dl:
                t, b, s = t.to(dev), b.to(dev), s.to(dev)
                if direction == "rna2dna":
                    recon, mu, lv = model(rna=t, site=s)
                    loss, rl, kl = rna2dna_loss(recon, b, mu, lv, beta=beta)
                else:
                    recon, mu, lv = model(dna=b, site=s)
                    loss, rl, kl = dna2rna_loss(recon, t, mu, lv, beta=beta)
                assert isinstance(rl, float) and isinstance(kl, float)
                if training:
                    opt.zero_grad()
                    loss.backward()
                    opt.step()
                total += loss.item()
        return total / len(dl)

    hist = []
    for epoch in range(6):
        tr = run(train_dl, epoch, True)
        va = run(val_dl, epoch, False)
        sched.step(va)
        hist.append((tr, va))
    assert all(np.isfinite(x) for h in hist for x in h)
    assert hist[-1][0] < hist[0][0], hist          # the training loss falls (the synthetic targets are noise: no claim on validation)
    sd = model.state_dict()
    assert int(sd[[k for k in sd if k.endswith("num_batches_tracked")][0]]) == 6 * len(train_dl)
    buf = io.BytesIO()
    torch.save(sd, buf)
    buf.seek(0)
    model2 = cls(dims["A"], dims["B"], dims["S"], Config.LATENT_DIM).to(dev)
    model2.load_state_dict(torch.load(buf))
    model.eval(); model2.eval()
    x = torch.from_numpy(tpm[:16] if direction == "rna2dna" else beta_v[:16]).to(dev)
    eps = torch.randn(16, Config.LATENT_DIM, device=dev)
    kw = dict(rna=x) if direction == "rna2dna" else dict(dna=x)
    with torch.no_grad(), model.inject(eps=eps), model2.inject(eps=eps):
        o1 = model(site=None, **kw)
        o2 = model2(site=None, **kw)
    for a, b in zip(o1, o2):
        assert torch.equal(a, b)
