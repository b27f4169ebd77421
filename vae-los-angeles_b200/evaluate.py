#!/usr/bin/env python
"""evaluate.py -- the evaluation entry point the reference documents (README.md:89-101, run_pipeline.sh:23-24) and that its
downstream_task.py imports (`from evaluate import get_run_id, load_model_and_data`, downstream_task.py:18, 399-400) but
does not ship.  Loads checkpoints/best_multivae.pt, runs the cross-modal reconstructions of the validation split
(`model(a=tpm)` -> DNA, `model(b=beta)` -> RNA, as downstream_task.py:32, 48 do) at a large batch, and computes the
metrics of compare_directional_imputation.py:167-210 on the device (vla_b200.recon_metrics).  Writes
plots/evaluation_results.json.
"""
import json
import os
import pickle
import sys

import pandas as pd
import torch
from sklearn.model_selection import train_test_split
from torch.utils.data import DataLoader

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from src.config import Config  # noqa: E402
from src.data import MultiModalDataset  # noqa: E402
from src.models import MultiModalVAE  # noqa: E402
from vla_b200 import recon_metrics  # noqa: E402

EVAL_BATCH = int(os.getenv("EVAL_BATCH", 262144))        # BASELINE configs[3]: inference at batch 262 144


def get_run_id():
    """The tri-modal pipeline keeps a single best checkpoint (Config.BEST_MODEL_NAME); a RUN_ID may be given from outside."""
    return os.environ.get("RUN_ID", "")


def load_model_and_data():
    """-> (model in eval mode on Config.DEVICE, validation DataLoader, run id)."""
    for name in ("INPUT_DIM_A", "INPUT_DIM_B", "LATENT_DIM", "BATCH_SIZE"):
        setattr(Config, name, int(os.getenv(name, getattr(Config, name))))
    Config.DEVICE = torch.device(os.environ.get("DEVICE", "cuda"))
    merged_df = pd.read_pickle("data/processed_data.pkl")
    with open("data/label_encoder.pkl", "rb") as f:
        n_sites = len(pickle.load(f).classes_)
    _, val_df = train_test_split(merged_df, test_size=Config.TRAIN_TEST_SPLIT, random_state=Config.RANDOM_SEED)
    val_dataloader = DataLoader(MultiModalDataset(val_df), batch_size=Config.BATCH_SIZE, shuffle=False)
    model = MultiModalVAE(Config.INPUT_DIM_A, Config.INPUT_DIM_B, n_sites, Config.LATENT_DIM)
    state = torch.load(os.path.join(Config.CHECKPOINT_DIR, Config.BEST_MODEL_NAME), map_location="cpu")
    model.load_state_dict(state)
    return model.to(Config.DEVICE).eval(), val_dataloader, get_run_id()


@torch.no_grad()
def cross_modal_reconstructions(model, ds, device, batch=EVAL_BATCH):
    """RNA -> (RNA, DNA, site) and DNA -> (RNA, DNA, site) for every row of `ds`, in chunks of `batch` rows."""
    tpm = torch.as_tensor(ds.tpm_data).to(device)
    beta = torch.as_tensor(ds.beta_data).to(device)
    site = torch.as_tensor(ds.primary_site).to(device)
    outs = {k: [] for k in ("a2a", "a2b", "a2c", "b2a", "b2b", "b2c")}
    for lo in range(0, len(site), batch):
        ra, rb, rc, _, _ = model(a=tpm[lo:lo + batch])
        outs["a2a"].append(ra); outs["a2b"].append(rb); outs["a2c"].append(rc)
        ra, rb, rc, _, _ = model(b=beta[lo:lo + batch])
        outs["b2a"].append(ra); outs["b2b"].append(rb); outs["b2c"].append(rc)
    return tpm, beta, site, {k: torch.cat(v) for k, v in outs.items()}


def main():
    os.makedirs("plots", exist_ok=True)
    model, val_dataloader, run_id = load_model_and_data()
    tpm, beta, site, rec = cross_modal_reconstructions(model, val_dataloader.dataset, Config.DEVICE)
    results = []
    for key, true, modality, name in (("a2a", tpm, "RNA", "RNA -> RNA"), ("b2a", tpm, "RNA", "DNA -> RNA"),
                                      ("b2b", beta, "DNA", "DNA -> DNA"), ("a2b", beta, "DNA", "RNA -> DNA")):
        r = recon_metrics(true, rec[key], modality, name, per_sample=False)
        results.append(r)
        print(f"{name}: MSE {r['MSE']:.5f}  MAE {r['MAE']:.5f}  cosine {r['CosineSimilarity']:.4f}  Pearson r {r['PearsonMean']:.4f}")
    acc = {"RNA -> site": float((rec["a2c"].argmax(1) == site).float().mean().item()),
           "DNA -> site": float((rec["b2c"].argmax(1) == site).float().mean().item())}
    print("primary-site accuracy:", acc)
    with open(os.path.join("plots", "evaluation_results.json"), "w") as f:
        json.dump(dict(run_id=run_id, metrics=results, site_accuracy=acc, n_val=int(len(site))), f, indent=1)
    print("Saved plots/evaluation_results.json")


if __name__ == "__main__":
    main()
