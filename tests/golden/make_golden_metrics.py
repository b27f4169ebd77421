"""Golden values of the reconstruction metrics from the LIBRARY CALLS the reference makes
(compare_directional_imputation.py:167-210: sklearn mean_absolute_error / mean_squared_error / r2_score /
cosine_similarity, scipy pearsonr).  The reference module itself cannot be imported here (it needs matplotlib, seaborn and
plotly at import time), so this script repeats its function body verbatim in spirit: same calls, same order, same
NaN handling.  Run in the build container:  python tests/golden/make_golden_metrics.py
"""
import os
import sys

import numpy as np
from scipy.stats import pearsonr
from sklearn.metrics import mean_absolute_error, mean_squared_error, r2_score
from sklearn.metrics.pairwise import cosine_similarity

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import vae_oracle as vo  # noqa: E402

CASES = [dict(name="metrics_rna", n=96, dim=782, seed=1, kind="a"), dict(name="metrics_dna", n=64, dim=572, seed=2, kind="b"),
         dict(name="metrics_edge", n=9, dim=37, seed=3, kind="edge")]


def case_arrays(case):
    n, dim = case["n"], case["dim"]
    dims = dict(A=782, B=572, S=24, L=20, E=32)
    if case["kind"] == "a":
        t = vo.synthetic_batch(n, dims, seed=case["seed"])[0]
    elif case["kind"] == "b":
        t = vo.synthetic_batch(n, dims, seed=case["seed"])[1]
    else:
        t = (vo.hash_uniform(n * dim, case["seed"], 5).reshape(n, dim) * 4 - 1).astype(np.float32)
    noise = vo.hash_normal(n * dim, case["seed"], 6).reshape(n, dim).astype(np.float32)
    p = (0.8 * t + 0.3 * noise).astype(np.float32)
    if case["kind"] == "edge":
        t[0] = 0.0                      # zero true row: cosine 0, Pearson undefined
        p[1] = 0.25                     # constant prediction: Pearson undefined
        p[2] = t[2]                     # perfect row
        p[3] = -t[3]                    # anti-correlated row
    return t, p


def reference_metrics(y_true, y_pred):
    y_true_flat, y_pred_flat = y_true.flatten(), y_pred.flatten()
    mae = mean_absolute_error(y_true_flat, y_pred_flat)
    mse = mean_squared_error(y_true_flat, y_pred_flat)
    r2 = r2_score(y_true_flat, y_pred_flat)
    cos = np.diag(cosine_similarity(y_true, y_pred))
    pearson_all = []
    for i in range(len(y_true)):
        try:
            r, _ = pearsonr(y_true[i], y_pred[i])
            if not np.isnan(r):
                pearson_all.append(r)
        except Exception:
            pass
    return dict(MAE=mae, MSE=mse, RMSE=np.sqrt(mse), R2=r2, CosineSimilarity=float(cos.mean()),
                PearsonMean=np.mean(pearson_all) if pearson_all else 0.0, PearsonStd=np.std(pearson_all) if pearson_all else 0.0,
                PearsonCount=len(pearson_all)), cos, np.array(pearson_all)


def main():
    import warnings
    warnings.simplefilter("ignore")
    for case in CASES:
        t, p = case_arrays(case)
        scal, cos, pear = reference_metrics(t, p)
        np.savez_compressed(os.path.join(HERE, case["name"] + ".npz"), cos=cos.astype(np.float64), pearson=pear.astype(np.float64),
                            **{k: np.float64(v) for k, v in scal.items()})
        print(case["name"], {k: round(float(v), 6) for k, v in scal.items()})


if __name__ == "__main__":
    main()
