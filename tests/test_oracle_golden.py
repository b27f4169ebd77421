"""Pin the numpy oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os
import re

import numpy as np
import pytest

from golden.make_golden import CASES, SAMPLE_STRIDE, case_inputs
from oracle import vae_oracle as vo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(x, ref):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    return np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30)


def is_pre_bn_bias(name):
    """Linear bias directly followed by a train-mode BatchNorm (encoders.py:13-14, 31-32, 35-36)."""
    m = re.fullmatch(r"encoder_\w+\.fc\.(\d+)\.bias", name)
    if m:
        return int(m.group(1)) % 4 == 0
    m = re.fullmatch(r"encoder_(rna|dna)\.(\d+)\.bias", name)          # autoencoders: bare nn.Sequential, the last Linear
    if m:                                                                # (index 4 * depth) is the head, not followed by BN
        depth = 1 if m.group(1) == "rna" else 2
        return int(m.group(2)) % 4 == 0 and int(m.group(2)) < 4 * depth
    return False


def check_packed(fix, prefix, arr, tol, atol=0.0):
    arr = np.asarray(arr, dtype=np.float64)
    if prefix + "|full" in fix:
        ref = fix[prefix + "|full"]
        assert ref.shape == arr.shape, prefix
        err = np.linalg.norm(arr - ref)
        assert err <= tol * np.linalg.norm(ref) + atol, (prefix, err, np.linalg.norm(ref))
    else:
        ref = fix[prefix + "|sample"]
        got = arr.reshape(-1)[::SAMPLE_STRIDE]
        err = np.linalg.norm(got - ref)
        assert err <= tol * np.linalg.norm(ref) + atol, (prefix, err, np.linalg.norm(ref))
        nrm = np.sqrt(float(fix[prefix + "|sumsq"]))
        assert abs(np.sqrt((arr ** 2).sum()) - nrm) <= tol * nrm + atol, prefix


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_oracle_matches_reference_fixture(case, dtype):
    fix = np.load(os.path.join(GOLDEN, case["name"] + ".npz"))
    state, batch, eps, masks, cw = case_inputs(case)
    state = {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in state.items()}
    batch = {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in batch.items()}
    eps = eps.astype(dtype)
    cw = cw.astype(dtype) if cw is not None else None
    tol = 2e-5 if dtype == np.float64 else 2e-4
    inputs = {k: (batch[k] if k in case["present"] else None) for k in ("a", "b", "site")}
    opt, step = vo.adamw_init(state)
    nsteps = max(case["steps"], 1)
    for it in range(nsteps):
        out, cache = vo.forward(case["kind"], case["dims"], state, inputs, eps, masks, train=case["train"])
        scalars, og = vo.loss_and_output_grads(case["kind"], out, batch, case["beta"], case["gamma"], cw)
        if it == 0:
            for prefix, r in out["recon"].items():
                check_packed(fix, f"out.recon.{prefix}", r, tol)
            check_packed(fix, "out.mu", out["mu"], tol)
            if out["logvar"] is not None:
                check_packed(fix, "out.logvar", out["logvar"], tol)
            ref = fix["loss"]
            got = [scalars["total"], scalars["recon"], scalars["cls"], scalars["kld"]]
            np.testing.assert_allclose(got, ref, rtol=5e-6 if dtype == np.float64 else 5e-5, atol=1e-6)
        if case["steps"] > 0:
            grads = vo.backward(case["kind"], case["dims"], state, cache, og, train=case["train"])
            if it == 0:
                n_checked = 0
                for key in fix.files:
                    if not key.startswith("grad."):
                        continue
                    name, kind = key[5:].split("|")
                    if kind == "none":
                        assert name not in grads, name
                    elif kind in ("full", "sample"):
                        # Linear biases feeding a train-mode BatchNorm have an exactly-zero true
                        # gradient; both sides hold only rounding noise there -> absolute tolerance.
                        gscale = max(np.abs(grads[name]).max(), 1.0)
                        check_packed(fix, "grad." + name, grads[name], tol * 5, atol=2e-4 * gscale ** 0)
                        n_checked += 1
                assert n_checked == len(grads)
            step = vo.adamw_step(state, grads, opt, step)
    if case["steps"] > 0:
        ref = fix["loss_last"]
        got = [scalars["total"], scalars["recon"], scalars["cls"], scalars["kld"]]
        np.testing.assert_allclose(got, ref, rtol=2e-5 if dtype == np.float64 else 1e-4, atol=1e-6)
        for key in fix.files:
            if key.startswith("final.") and key.endswith(("|full", "|sample")):
                name = key[6:].split("|")[0]
                if name.endswith("num_batches_tracked"):
                    assert int(state[name]) == int(fix[key]), name
                elif is_pre_bn_bias(name):
                    # true gradient is exactly zero; Adam turns the rounding noise into +-lr steps
                    check_packed(fix, "final." + name, state[name], 1e-4, atol=2.1 * 5e-4 * nsteps * np.sqrt(state[name].size))
                elif name.endswith("running_mean"):
                    # the batch mean contains the pre-BN bias, which carries the +-lr noise above
                    check_packed(fix, "final." + name, state[name], 1e-4,
                                 atol=2.1 * 5e-4 * nsteps * vo.BN_MOMENTUM * nsteps * np.sqrt(state[name].size))
                else:
                    check_packed(fix, "final." + name, state[name], 1e-4)


def test_param_counts_match_survey():
    dims = dict(A=782, B=572, S=24, L=20, E=32)
    counts = {}
    for kind in vo.MODEL_KINDS:
        shapes = vo.param_shapes(kind, dims)
        counts[kind] = sum(int(np.prod(s)) for k, s in shapes.items() if not vo.is_buffer(k))
    # the autoencoders have one head (L x last + L) per encoder less than their VAE counterparts
    assert counts == {"multimodal": 1081114, "rna2dna": 538124, "dna2rna": 542174,
                      "rna2dna_ae": 538124 - (20 * 128 + 20) - (20 * 32 + 20), "dna2rna_ae": 542174 - (20 * 256 + 20) - (20 * 32 + 20)}
