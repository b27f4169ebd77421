"""CPU oracle of the reconstruction metrics -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU arm).

Restates `compute_metrics` of the reference (compare_directional_imputation.py:167-210), which calls scikit-learn and scipy:
  MAE / MSE / RMSE / R2 over the flattened arrays (sklearn.metrics.mean_absolute_error, mean_squared_error, r2_score),
  cosine similarity per sample (the diagonal of sklearn's cosine_similarity: rows L2-normalised, a zero row stays zero),
  Pearson r per sample (scipy.stats.pearsonr; NaN for a constant row, skipped), its mean and population std.
Pinned against those library functions themselves by tests/golden/metrics_*.npz (tests/golden/make_golden_metrics.py).
"""
import numpy as np


def recon_metrics(y_true, y_pred):
    """Returns (dict of scalars, per-sample cosine [N], per-sample Pearson r [N] with NaN where undefined)."""
    t = np.asarray(y_true, dtype=np.float64)
    p = np.asarray(y_pred, dtype=np.float64)
    assert t.shape == p.shape and t.ndim == 2
    d = p - t
    mae = np.abs(d).mean()                                           # compare_directional_imputation.py:174
    mse = (d ** 2).mean()                                            # :175
    ss_res = (d ** 2).sum()
    ss_tot = ((t - t.mean()) ** 2).sum()
    if ss_tot > 0:
        r2 = 1.0 - ss_res / ss_tot                                   # :176 (r2_score on the flattened arrays)
    else:
        r2 = 1.0 if ss_res == 0 else 0.0                             # sklearn's force_finite convention
    nt = np.sqrt((t ** 2).sum(1))
    npred = np.sqrt((p ** 2).sum(1))
    cos = (t * p).sum(1) / (np.where(nt == 0, 1.0, nt) * np.where(npred == 0, 1.0, npred))   # :179-180
    tc = t - t.mean(1, keepdims=True)
    pc = p - p.mean(1, keepdims=True)
    vt, vp = (tc ** 2).sum(1), (pc ** 2).sum(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where((vt > 0) & (vp > 0), (tc * pc).sum(1) / np.sqrt(vt * vp), np.nan)       # :184-190
    r = np.clip(r, -1.0, 1.0)
    valid = r[~np.isnan(r)]
    scal = dict(MAE=float(mae), MSE=float(mse), RMSE=float(np.sqrt(mse)), R2=float(r2), CosineSimilarity=float(cos.mean()),
                PearsonMean=float(valid.mean()) if valid.size else 0.0, PearsonStd=float(valid.std()) if valid.size else 0.0,
                PearsonCount=int(valid.size))
    return scal, cos, r
