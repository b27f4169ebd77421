"""The row-chain kernel (csrc/rowchain.cu: the row-local middle of a directional model's train step with every activation
on chip, one CTA per 128-row block) against the same step issued as separate launches (VLA_ROWCHAIN=0), at batch sizes
with many, few and ragged row blocks; the oracle parity of the separate launches (the default path) is in test_gpu_train.py / test_gpu_parity_large.py; the
row-chain step is an opt-in (VLA_ROWCHAIN=1, profiles/r2_rowchain_experiments.md says why)."""
import os

import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo
from parity_util import is_pre_bn_bias, make_module, rel_l2, to_t

pytestmark = pytest.mark.gpu

FULL = dict(A=782, B=572, S=24, L=20, E=32)


def _run(kind, batch, rowchain, n_steps, only_fb, dims=FULL, inject=True):
    from vla_b200 import DeviceDataset, Trainer
    os.environ["VLA_ROWCHAIN"] = "1" if rowchain else "0"      # (opt-in: the default path is the separate launches)
    os.environ["VLA_HEADBLOCK"] = "0"            # like with like: the row chain runs these layers on the tensor cores too
    try:
        state = vo.init_state(kind, dims, seed=3)
        tpm, beta_v, site = vo.synthetic_batch(batch * 2, dims, seed=3)
        eps, masks = vo.synthetic_noise(batch, dims, kind, seed=3)
        m = make_module(kind, dims, state).train()
        ds = DeviceDataset(tpm, beta_v, site, "cuda")
        tr = Trainer(m, ds, batch, lr=5e-4, weight_decay=1e-5, beta_kl=2e-3, gamma=1.5, use_graph=False, seed=5)
        if inject:
            tr.injected = dict(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()])
        if only_fb:
            tr.forward_backward()
            torch.cuda.synchronize()
            return tr.grads.cpu().numpy().copy(), np.array(tr.losses()), None
        losses = []
        for _ in range(n_steps):
            tr.step()
            losses.append(tr.losses())
        torch.cuda.synchronize()
        sd = {k: v.detach().float().cpu().numpy().copy() for k, v in m.state_dict().items()}
        return None, np.array(losses), sd
    finally:
        os.environ.pop("VLA_ROWCHAIN", None)
        os.environ.pop("VLA_HEADBLOCK", None)


@pytest.mark.parametrize("kind,batch,inject", [("rna2dna", 4096, True), ("rna2dna", 1000, True), ("rna2dna", 130, True),
                                               ("rna2dna", 40, True), ("rna2dna_ae", 4096, True), ("rna2dna", 2048, False),
                                               ("rna2dna", 20000, True)])
def test_rowchain_gradients_equal_separate_launches(kind, batch, inject):
    """Same arithmetic (split-bf16 operands, fp32 accumulation, identical Philox streams when nothing is injected); only the
    order of the split-K and loss partial sums differs."""
    g_sep, l_sep, _ = _run(kind, batch, False, 1, True, inject=inject)
    g_rc, l_rc, _ = _run(kind, batch, True, 1, True, inject=inject)
    assert np.isfinite(g_rc).all()
    np.testing.assert_allclose(l_rc, l_sep, rtol=2e-6)
    assert rel_l2(g_rc, g_sep) < 1e-5, rel_l2(g_rc, g_sep)


def test_rowchain_other_dimensions():
    dims = dict(A=782, B=572, S=24, L=48, E=64)
    g_sep, l_sep, _ = _run("rna2dna", 700, False, 1, True, dims=dims)
    g_rc, l_rc, _ = _run("rna2dna", 700, True, 1, True, dims=dims)
    np.testing.assert_allclose(l_rc, l_sep, rtol=2e-6)
    assert rel_l2(g_rc, g_sep) < 1e-5, rel_l2(g_rc, g_sep)


def test_rowchain_steps_equal_separate_launches():
    _, l_sep, sd_sep = _run("rna2dna", 4096, False, 5, False)
    _, l_rc, sd_rc = _run("rna2dna", 4096, True, 5, False)
    np.testing.assert_allclose(l_rc, l_sep, rtol=2e-4)
    for k, v in sd_sep.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd_rc[k]) == int(v) == 5
        elif k.endswith("running_mean"):
            np.testing.assert_allclose(sd_rc[k], v, rtol=1e-3, atol=2 * 5e-4 * 5)
        elif k.endswith("running_var"):
            np.testing.assert_allclose(sd_rc[k], v, rtol=5e-3, atol=1e-4)
        else:
            assert np.abs(sd_rc[k] - v).max() <= 2 * 5e-4 * 5 + 1e-6, k
            if not is_pre_bn_bias(k):
                assert rel_l2(sd_rc[k], v) < 1e-2, (k, rel_l2(sd_rc[k], v))


def test_rowchain_graph_replay_trains(monkeypatch):
    monkeypatch.setenv("VLA_ROWCHAIN", "1")
    from vla_b200 import DeviceDataset, Trainer
    m = make_module("rna2dna", FULL, vo.init_state("rna2dna", FULL, seed=4)).train()
    ds = DeviceDataset.synthetic(512 * 4, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=1)
    tr = Trainer(m, ds, 512, use_graph=True)
    first = None
    for i in range(12):
        tr.step()
        if i == 0:
            first = tr.losses()[0]
    last = tr.losses()
    assert np.isfinite(last).all() and last[0] < first
    launches = {}
    for name, ms, fl, by in tr.profile(1):
        launches[name] = launches.get(name, 0) + 1
    launches.pop("_empty_pair", None)
    assert "rowchain_fwd_bwd" in launches and sum(launches.values()) <= 7, launches
