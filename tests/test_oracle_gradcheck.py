"""The oracle's explicit backward (oracle/vae_oracle.py) against central finite differences of its own forward + loss, in
fp64: an independent check of the test infrastructure besides the reference fixtures.  CPU only."""
import numpy as np
import pytest

from oracle import vae_oracle as vo

DIMS = dict(A=23, B=17, S=5, L=6, E=8)


def _total(kind, state, batch, present, eps, masks, beta, gamma, cw):
    st = {k: v.copy() for k, v in state.items()}
    inputs = {k: (batch[k] if k in present else None) for k in ("a", "b", "site")}
    out, _ = vo.forward(kind, DIMS, st, inputs, eps, masks, train=True, update_running=False)
    scal, _ = vo.loss_and_output_grads(kind, out, batch, beta, gamma, cw)
    return scal["total"]


@pytest.mark.parametrize("kind,present", [("multimodal", ("a", "b", "site")), ("rna2dna", ("a", "site")), ("dna2rna", ("b",)),
                                          ("rna2dna_ae", ("a", "site")), ("dna2rna_ae", ("b", "site"))])
def test_backward_matches_finite_differences(kind, present):
    n = 11
    state = {k: (v.astype(np.float64) if v.dtype.kind == "f" else v) for k, v in vo.init_state(kind, DIMS, seed=3).items()}
    tpm, beta_v, site = vo.synthetic_batch(n, DIMS, seed=3)
    batch = dict(a=tpm.astype(np.float64), b=beta_v.astype(np.float64), site=site)
    eps, masks = vo.synthetic_noise(n, DIMS, kind, seed=3)
    eps = eps.astype(np.float64)
    cw = vo.balanced_class_weights(site, DIMS["S"]).astype(np.float64) if kind == "multimodal" else None
    beta, gamma = 0.05, 1.7
    st = {k: v.copy() for k, v in state.items()}
    inputs = {k: (batch[k] if k in present else None) for k in ("a", "b", "site")}
    out, cache = vo.forward(kind, DIMS, st, inputs, eps, masks, train=True, update_running=False)
    _, og = vo.loss_and_output_grads(kind, out, batch, beta, gamma, cw)
    grads = vo.backward(kind, DIMS, st, cache, og, train=True)
    rng = np.random.default_rng(0)
    checked = 0
    for name, g in grads.items():
        flat_idx = rng.choice(g.size, size=min(3, g.size), replace=False)
        for fi in flat_idx:
            idx = np.unravel_index(fi, g.shape)
            h = 1e-5 * max(1.0, abs(state[name][idx]))
            plus = {k: v.copy() for k, v in state.items()}
            minus = {k: v.copy() for k, v in state.items()}
            plus[name][idx] += h
            minus[name][idx] -= h
            num = (_total(kind, plus, batch, present, eps, masks, beta, gamma, cw)
                   - _total(kind, minus, batch, present, eps, masks, beta, gamma, cw)) / (2 * h)
            scale = max(abs(num), abs(g[idx]), 1e-3 * np.abs(g).max(), 1e-8)
            # (round-off of the difference quotient: ~1e-16 * |total| / h, the totals here are of order 1e3)
            assert abs(num - g[idx]) <= 2e-5 * scale + 1e-6, (name, idx, num, g[idx])
            checked += 1
    assert checked >= 3 * 8
