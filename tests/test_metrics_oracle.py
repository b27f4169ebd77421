"""Pin the metrics oracle (oracle/metrics_oracle.py) against the values scikit-learn / scipy produce for the reference's
compute_metrics (compare_directional_imputation.py:167-210), committed as tests/golden/metrics_*.npz.  CPU only."""
import os

import numpy as np
import pytest

from golden.make_golden_metrics import CASES, case_arrays
from oracle import metrics_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_metrics_oracle_matches_library_values(case):
    fix = np.load(os.path.join(GOLDEN, case["name"] + ".npz"))
    t, p = case_arrays(case)
    scal, cos, r = mo.recon_metrics(t, p)
    for k in ("MAE", "MSE", "RMSE", "R2", "CosineSimilarity", "PearsonMean", "PearsonStd"):
        np.testing.assert_allclose(scal[k], float(fix[k]), rtol=2e-6, atol=1e-9, err_msg=k)     # (sklearn works in float32 here)
    assert scal["PearsonCount"] == int(fix["PearsonCount"])
    np.testing.assert_allclose(cos, fix["cos"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r[~np.isnan(r)], fix["pearson"], rtol=1e-6, atol=1e-7)
