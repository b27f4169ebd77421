"""The reference's train_rna2dna.py and train_dna2rna.py, UNMODIFIED (byte-for-byte copies staged by
tests/golden/stage_reference_scripts.py into the git-ignored tests/_ref_scripts/), executed by path on the GPU against this
repository's drop-in `src` package: BASELINE.json north_star "train_rna2dna.py and train_dna2rna.py run unmodified".
Needs what the scripts read: data/processed_data.pkl (a DataFrame with the reference's columns), data/label_encoder.pkl and
an importable matplotlib (a no-op stub package here: the plots are not under test).  The run's log is kept under
gpurun_out/ when that directory exists."""
import hashlib
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = os.path.join(ROOT, "tests", "_ref_scripts")


def _workdir(tmp_path):
    import pandas as pd
    from sklearn.preprocessing import LabelEncoder
    from oracle import vae_oracle as vo
    dims = dict(A=782, B=572, S=24, L=20, E=32)
    n = 400
    tpm, beta, site = vo.synthetic_batch(n, dims, seed=77)
    names = np.array([f"site_{i:02d}" for i in range(dims["S"])])
    le = LabelEncoder().fit(names)
    df = pd.DataFrame({"case_barcode": [f"TCGA-{i:04d}" for i in range(n)], "tpm_unstranded": list(tpm),
                       "primary_site": names[site], "beta_value": list(beta), "primary_site_encoded": site})
    (tmp_path / "data").mkdir()
    df.to_pickle(tmp_path / "data" / "processed_data.pkl")
    with open(tmp_path / "data" / "label_encoder.pkl", "wb") as f:
        pickle.dump(le, f)
    stub = tmp_path / "stubs" / "matplotlib"
    stub.mkdir(parents=True)
    noop = ("class _N:\n    def __call__(self, *a, **k):\n        return self\n    def __getattr__(self, n):\n        return self\n"
            "    def __iter__(self):\n        return iter((self, self))\n\n\ndef __getattr__(name):\n    return _N()\n")
    (stub / "__init__.py").write_text(noop)
    (stub / "pyplot.py").write_text(noop)
    return tmp_path


@pytest.mark.parametrize("script,prefix", [("train_rna2dna.py", "best_rna2dna_"), ("train_dna2rna.py", "best_dna2rna_")])
def test_reference_training_script_runs_unmodified(script, prefix, tmp_path):
    path = os.path.join(SCRIPTS, script)
    if not os.path.exists(path):
        pytest.skip("tests/_ref_scripts/ not staged (run tests/golden/stage_reference_scripts.py where /root/reference exists)")
    with open(os.path.join(SCRIPTS, "MANIFEST.json")) as f:
        manifest = json.load(f)
    with open(path, "rb") as f:
        assert hashlib.sha256(f.read()).hexdigest() == manifest["sha256"][script], "the staged script is not the reference's file"
    wd = _workdir(tmp_path)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "vae-los-angeles_b200"), str(wd / "stubs")])
    env.update(INPUT_DIM_A="782", INPUT_DIM_B="572", DEVICE="cuda", PYTHONDONTWRITEBYTECODE="1")
    res = subprocess.run([sys.executable, path], cwd=wd, env=env, capture_output=True, text=True, timeout=900)
    log = res.stdout[-6000:] + "\n--- stderr ---\n" + res.stderr[-3000:]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"reference_script_{script}.log"), "w") as f:
            f.write(log)
    assert res.returncode == 0, log
    assert "Training complete!" in res.stdout
    saved = [p for p in os.listdir(wd / "checkpoints") if p.startswith(prefix)]
    assert saved, "no checkpoint written"
    # it trained (the synthetic targets are i.i.d. noise: the training loss can only fall a little before early stopping)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("Epoch [")]
    first = float(lines[0].split("Train Loss:")[1].split("|")[0])
    last = float(lines[-1].split("Train Loss:")[1].split("|")[0])
    assert len(lines) >= 5 and np.isfinite(last) and last < first, (first, last)
    import torch
    sd = torch.load(os.path.join(wd, "checkpoints", saved[0]), map_location="cpu")
    assert all(torch.isfinite(v.float()).all() for v in sd.values())
