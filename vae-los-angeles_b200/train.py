#!/usr/bin/env python
"""train.py -- the tri-modal training entry point the reference documents (README.md:73-87, run_pipeline.sh:19-20) but
does not ship.  Same skeleton as its train_rna2dna.py (:20-245): data/processed_data.pkl + data/label_encoder.pkl,
train_test_split(TRAIN_TEST_SPLIT, RANDOM_SEED), shuffled full batches (drop_last), beta warm-up, balanced class weights
(optimize_hyperparameters.py:33-44), validation every epoch, ReduceLROnPlateau, early stopping with PATIENCE, best model to
checkpoints/best_multivae.pt.  The loop body (fwd + loss + bwd + AdamW) is one CUDA-graph replay per batch on a
device-resident dataset (vla_b200.Trainer); the host reads one loss per epoch.

Env overrides as in the reference scripts: INPUT_DIM_A, INPUT_DIM_B, LATENT_DIM (+ NUM_EPOCHS, BATCH_SIZE for short runs).
"""
import json
import os
import pickle
import sys
import time

import numpy as np
import pandas as pd
import torch
from sklearn.model_selection import train_test_split

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

from src.config import Config  # noqa: E402
from src.data import MultiModalDataset  # noqa: E402
from src.models import MultiModalVAE  # noqa: E402
from vla_b200 import DeviceDataset, Trainer, fused_vae_loss  # noqa: E402


def setup_directories():
    os.makedirs(Config.CHECKPOINT_DIR, exist_ok=True)
    os.makedirs("plots", exist_ok=True)


def load_data():
    merged_df = pd.read_pickle("data/processed_data.pkl")
    with open("data/label_encoder.pkl", "rb") as f:
        label_encoder = pickle.load(f)
    print(f"Data shape: {merged_df.shape}\nNumber of primary sites: {len(label_encoder.classes_)}")
    return merged_df, label_encoder


def apply_env_overrides():
    for name in ("INPUT_DIM_A", "INPUT_DIM_B", "LATENT_DIM", "NUM_EPOCHS", "BATCH_SIZE"):
        setattr(Config, name, int(os.getenv(name, getattr(Config, name))))
    Config.DEVICE = torch.device(os.environ.get("DEVICE", "cuda"))
    if Config.DEVICE.type != "cuda":
        raise RuntimeError("this implementation runs on a B200 (sm_100a) only; there is no CPU fallback")


def balanced_class_weights(sites, n_sites):
    counts = np.bincount(sites, minlength=n_sites).astype(np.float64)
    return torch.tensor(len(sites) / (n_sites * np.maximum(counts, 1.0)), dtype=torch.float32)


@torch.no_grad()
def validate(model, val, beta, gamma, class_weights, batch):
    """Mean per-batch validation loss (the reference divides the summed batch losses by len(val_dataloader))."""
    model.eval()
    total = torch.zeros((), device=val.site.device)
    n = 0
    for lo in range(0, len(val), batch):
        a, b, s = val.tpm[lo:lo + batch], val.beta[lo:lo + batch], val.site[lo:lo + batch]
        ra, rb, rc, mu, lv = model(a=a, b=b, site=s)
        total += fused_vae_loss(ra, a, rb, b, rc, s, mu, lv, beta=beta, gamma=gamma, class_weights=class_weights)[0]
        n += 1
    model.train()
    return float(total.item()) / max(n, 1)


def main():
    setup_directories()
    apply_env_overrides()
    merged_df, label_encoder = load_data()
    n_sites = len(label_encoder.classes_)
    train_df, val_df = train_test_split(merged_df, test_size=Config.TRAIN_TEST_SPLIT, random_state=Config.RANDOM_SEED)
    print(f"Train set size: {len(train_df)}\nValidation set size: {len(val_df)}")
    train_ds, val_ds = MultiModalDataset(train_df), MultiModalDataset(val_df)
    dev = Config.DEVICE
    train = DeviceDataset(train_ds.tpm_data, train_ds.beta_data, train_ds.primary_site, dev)
    val = DeviceDataset(val_ds.tpm_data, val_ds.beta_data, val_ds.primary_site, dev)
    cw = balanced_class_weights(train_ds.primary_site, n_sites).to(dev)
    batch = min(Config.BATCH_SIZE, len(train))
    steps_per_epoch = len(train) // batch                                   # shuffle=True, drop_last=True
    torch.manual_seed(Config.RANDOM_SEED)
    model = MultiModalVAE(Config.INPUT_DIM_A, Config.INPUT_DIM_B, n_sites, Config.LATENT_DIM).to(dev).train()
    trainer = Trainer(model, train, batch, lr=Config.LEARNING_RATE, weight_decay=Config.WEIGHT_DECAY, beta_kl=0.0,
                      gamma=Config.GAMMA, class_weights=cw, seed=Config.RANDOM_SEED)
    dummy = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=Config.LEARNING_RATE)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(dummy, mode="min", factor=Config.LR_SCHEDULER_FACTOR,
                                                           patience=Config.LR_SCHEDULER_PATIENCE)
    best, trigger, hist = float("inf"), 0, dict(train=[], val=[])
    path = os.path.join(Config.CHECKPOINT_DIR, Config.BEST_MODEL_NAME)
    t0 = time.time()
    for epoch in range(Config.NUM_EPOCHS):
        beta = min(1.0, epoch / Config.BETA_WARMUP_EPOCHS) * Config.BETA_START
        trainer.set_hyper(lr=dummy.param_groups[0]["lr"], beta_kl=beta)
        train.shuffle_()                                                    # new row order in place: the captured graph stays valid
        trainer.reset_counters(trainer.steps, 0)                            # start at the first resident batch again
        for _ in range(steps_per_epoch):
            trainer.step()
        train_loss = trainer.losses()[0]
        val_loss = validate(model, val, beta, Config.GAMMA, cw, batch)
        scheduler.step(val_loss)
        hist["train"].append(train_loss); hist["val"].append(val_loss)
        print(f"Epoch [{epoch + 1}/{Config.NUM_EPOCHS}] | Train Loss (last batch): {train_loss:.2f} | Val Loss: {val_loss:.2f} | beta={beta:.5f}")
        if val_loss < best:
            best, trigger = val_loss, 0
            torch.save(model.state_dict(), path)
            print(f"Best model saved (val_loss: {val_loss:.2f})")
        else:
            trigger += 1
            if trigger >= Config.PATIENCE:
                print(f"Early stopping triggered at epoch {epoch + 1}!")
                break
    trainer.close()
    with open(os.path.join("plots", "training_losses.json"), "w") as f:
        json.dump(dict(hist, seconds=time.time() - t0, best_val_loss=best), f)
    try:                                                                    # the plot is optional: matplotlib may be absent
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        plt.figure(figsize=(8, 5)); plt.plot(hist["train"], label="train (last batch)"); plt.plot(hist["val"], label="validation")
        plt.xlabel("epoch"); plt.ylabel("loss"); plt.legend(); plt.savefig(os.path.join("plots", "training_losses.png")); plt.close()
    except Exception:
        pass
    print(f"Training complete. Best validation loss: {best:.2f}. Checkpoint: {path}")


if __name__ == "__main__":
    main()
