"""train.py / evaluate.py (the entry points the reference documents, README.md:73-101, and downstream_task.py:18 imports)
end to end on a synthetic data/processed_data.pkl in a scratch directory: a checkpoint with the reference's state_dict
keys is written, the loss goes down, evaluate.py loads it and reports finite metrics through the on-device kernel."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest
import torch
from sklearn.preprocessing import LabelEncoder

from oracle import vae_oracle as vo

pytestmark = pytest.mark.gpu
PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vae-los-angeles_b200")


def test_train_then_evaluate(tmp_path):
    dims = dict(A=64, B=48, S=6, L=8, E=32)
    n = 640
    tpm, beta, site = vo.synthetic_batch(n, dims, seed=1)
    tpm = tpm + 0.5 * site[:, None]                      # make the modalities informative about the site
    os.makedirs(tmp_path / "data")
    names = np.array([f"site_{i}" for i in range(dims["S"])])
    df = pd.DataFrame({"tpm_unstranded": list(tpm.astype(np.float32)), "beta_value": list(beta), "primary_site": names[site],
                       "primary_site_encoded": site})
    df.to_pickle(tmp_path / "data" / "processed_data.pkl")
    with open(tmp_path / "data" / "label_encoder.pkl", "wb") as f:
        pickle.dump(LabelEncoder().fit(names), f)
    env = dict(os.environ, INPUT_DIM_A=str(dims["A"]), INPUT_DIM_B=str(dims["B"]), LATENT_DIM=str(dims["L"]), NUM_EPOCHS="6",
               BATCH_SIZE="64", EVAL_BATCH="100", PYTHONPATH=PKG + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for script in ("train.py", "evaluate.py"):
        res = subprocess.run([sys.executable, os.path.join(PKG, script)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    hist = json.load(open(tmp_path / "plots" / "training_losses.json"))
    assert len(hist["val"]) == 6 and hist["val"][-1] < hist["val"][0]
    sd = torch.load(tmp_path / "checkpoints" / "best_multivae.pt", map_location="cpu")
    assert list(sd) == list(vo.param_shapes("multimodal", dims))                  # the reference's state_dict keys, in order
    ev = json.load(open(tmp_path / "plots" / "evaluation_results.json"))
    assert ev["n_val"] == 128 and len(ev["metrics"]) == 4
    for m in ev["metrics"]:
        assert all(np.isfinite(m[k]) for k in ("MAE", "MSE", "RMSE", "R2", "CosineSimilarity", "PearsonMean", "PearsonStd"))
    assert 0.0 <= ev["site_accuracy"]["RNA -> site"] <= 1.0
