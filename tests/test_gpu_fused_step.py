"""The whole-step kernel (one persistent launch, units linked by row-block counters) against the same step issued as
separate launches (VLA_FUSED_STEP=0: identical device code per phase, kernel boundaries instead of counters), at batch
sizes with many and with ragged 128-row blocks; and against the oracle through the existing train tests."""
import os

import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo
from parity_util import is_pre_bn_bias, make_module, rel_l2, to_t

pytestmark = pytest.mark.gpu

FULL = dict(A=782, B=572, S=24, L=20, E=32)


def _run(kind, batch, fused, n_steps, phases_only_fb):
    from vla_b200 import DeviceDataset, Trainer
    os.environ["VLA_FUSED_STEP"] = "1" if fused else "0"
    try:
        state = vo.init_state(kind, FULL, seed=3)
        tpm, beta_v, site = vo.synthetic_batch(batch * 2, FULL, seed=3)
        eps, masks = vo.synthetic_noise(batch, FULL, kind, seed=3)
        cw = vo.balanced_class_weights(site, FULL["S"]) if kind == "multimodal" else None
        m = make_module(kind, FULL, state).train()
        ds = DeviceDataset(tpm, beta_v, site, "cuda")
        tr = Trainer(m, ds, batch, lr=5e-4, weight_decay=1e-5, beta_kl=2e-3, gamma=1.5, class_weights=to_t(cw), use_graph=False)
        tr.injected = dict(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()])
        if phases_only_fb:
            tr._call(ds, 1)                       # forward + loss + backward only: gradients stay in tr.grads
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
        os.environ.pop("VLA_FUSED_STEP", None)


@pytest.mark.parametrize("kind,batch", [("rna2dna", 4096), ("rna2dna", 1000), ("dna2rna", 333), ("multimodal", 1500)])
def test_fused_gradients_equal_separate_launches(kind, batch):
    g_sep, l_sep, _ = _run(kind, batch, False, 1, True)
    g_fus, l_fus, _ = _run(kind, batch, True, 1, True)
    np.testing.assert_allclose(l_fus, l_sep, rtol=1e-6)
    # same arithmetic; only the order of the split-K red.add partial sums differs
    assert rel_l2(g_fus, g_sep) < 1e-5, rel_l2(g_fus, g_sep)
    assert np.isfinite(g_fus).all()


@pytest.mark.parametrize("kind,batch", [("rna2dna", 4096), ("multimodal", 700)])
def test_fused_steps_equal_separate_launches(kind, batch):
    _, l_sep, sd_sep = _run(kind, batch, False, 5, False)
    _, l_fus, sd_fus = _run(kind, batch, True, 5, False)
    np.testing.assert_allclose(l_fus, l_sep, rtol=2e-4)
    for k, v in sd_sep.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd_fus[k]) == int(v) == 5
        elif k.endswith(("running_mean", "running_var")):
            np.testing.assert_allclose(sd_fus[k], v, rtol=1e-3, atol=1e-4)
        else:
            # Adam turns rounding-level gradient differences of near-zero gradients into +-lr steps
            assert np.abs(sd_fus[k] - v).max() <= 2 * 5e-4 * 5 + 1e-6, k
            if not is_pre_bn_bias(k):             # exactly-zero true gradient: the sign of rounding noise decides every step
                assert rel_l2(sd_fus[k], v) < 1e-2, (k, rel_l2(sd_fus[k], v))


def test_fused_step_graph_replay_and_timeline():
    """CUDA-graph replays of the whole-step kernel walk the resident batches, and the per-unit timeline is complete."""
    from vla_b200 import DeviceDataset, Trainer
    kind, batch = "rna2dna", 512
    state = vo.init_state(kind, FULL, seed=4)
    m = make_module(kind, FULL, state).train()
    ds = DeviceDataset.synthetic(batch * 4, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=1)
    os.environ["VLA_FUSED_STEP"] = "1"
    try:
        tr = Trainer(m, ds, batch, use_graph=True)
        first = None
        for i in range(12):
            tr.step()
            if i == 0:
                first = tr.losses()[0]
        last = tr.losses()
        assert np.isfinite(last).all() and last[0] < first          # it trains
        step_us, phases = tr.timeline()
    finally:
        os.environ.pop("VLA_FUSED_STEP", None)
    assert step_us > 0 and len(phases) >= 15
    names = [p["name"] for p in phases]
    assert names[0] == "ingest" and names[-1] == "adamw" and "wgrad_all" in names
    for p in phases:
        assert p["end_us"] >= p["start_us"] >= 0 and p["units"] > 0
