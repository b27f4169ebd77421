"""The chain kernel (row-local stretches of the step as one launch per stretch: a 4-CTA cluster per 128-row block, cluster
barriers between the phases) against the same step issued as separate launches (VLA_CHAIN=0: identical device code per
phase, kernel boundaries instead of cluster barriers), at batch sizes with many, few and ragged 128-row blocks; the
oracle parity of the chained step itself is in test_gpu_train.py / test_gpu_parity_large.py (the default path)."""
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
    os.environ["VLA_CHAIN"] = "1" if fused else "0"
    os.environ["VLA_HEADBLOCK"] = "0"            # like with like: the chain runs these layers as tensor-core tiles too
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
            tr.forward_backward()                 # forward + loss + backward only: gradients stay in tr.grads
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
        os.environ.pop("VLA_CHAIN", None)
        os.environ.pop("VLA_HEADBLOCK", None)


@pytest.mark.parametrize("kind,batch", [("rna2dna", 4096), ("rna2dna", 1000), ("dna2rna", 333), ("multimodal", 1500),
                                        ("dna2rna_ae", 4096), ("multimodal", 6000), ("rna2dna", 40)])
def test_chained_gradients_equal_separate_launches(kind, batch):
    g_sep, l_sep, _ = _run(kind, batch, False, 1, True)
    g_fus, l_fus, _ = _run(kind, batch, True, 1, True)
    np.testing.assert_allclose(l_fus, l_sep, rtol=1e-6)
    # same arithmetic; only the order of the split-K red.add partial sums and of the loss partials differs
    assert rel_l2(g_fus, g_sep) < 1e-5, rel_l2(g_fus, g_sep)
    assert np.isfinite(g_fus).all()


@pytest.mark.parametrize("kind,batch", [("rna2dna", 4096), ("multimodal", 700)])
def test_chained_steps_equal_separate_launches(kind, batch):
    _, l_sep, sd_sep = _run(kind, batch, False, 5, False)
    _, l_fus, sd_fus = _run(kind, batch, True, 5, False)
    np.testing.assert_allclose(l_fus, l_sep, rtol=2e-4)
    for k, v in sd_sep.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd_fus[k]) == int(v) == 5
        elif k.endswith("running_mean"):
            # the running mean contains the pre-BatchNorm bias, whose true gradient is exactly zero: Adam moves it by +-lr per
            # step on rounding noise, differently for the two orders of summation
            np.testing.assert_allclose(sd_fus[k], v, rtol=1e-3, atol=2 * 5e-4 * 5)
        elif k.endswith("running_var"):
            np.testing.assert_allclose(sd_fus[k], v, rtol=5e-3, atol=1e-4)
        else:
            # Adam turns rounding-level gradient differences of near-zero gradients into +-lr steps
            assert np.abs(sd_fus[k] - v).max() <= 2 * 5e-4 * 5 + 1e-6, k
            if not is_pre_bn_bias(k):             # exactly-zero true gradient: the sign of rounding noise decides every step
                assert rel_l2(sd_fus[k], v) < 1e-2, (k, rel_l2(sd_fus[k], v))


@pytest.fixture
def chain_on(monkeypatch):
    monkeypatch.setenv("VLA_CHAIN", "1")


def test_chain_graph_replay_and_timeline(chain_on):
    """CUDA-graph replays of the chained step walk the resident batches; the step is a handful of launches; the per-phase
    timeline of the chain launches is complete."""
    from vla_b200 import DeviceDataset, Trainer
    kind, batch = "rna2dna", 512
    state = vo.init_state(kind, FULL, seed=4)
    m = make_module(kind, FULL, state).train()
    ds = DeviceDataset.synthetic(batch * 4, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=1)
    tr = Trainer(m, ds, batch, use_graph=True)
    first = None
    for i in range(12):
        tr.step()
        if i == 0:
            first = tr.losses()[0]
    last = tr.losses()
    assert np.isfinite(last).all() and last[0] < first          # it trains
    chains = tr.timeline()
    assert [c["name"] for c in chains] == ["chain_ingest_enc", "chain_fwd_bwd"]
    mid = chains[1]
    names = [p["name"] for p in mid["phases"]]
    assert names[0] == "bn_act" and "latent_fwd" in names and "latent_bwd" in names and names[-1].startswith("dgrad_enc")
    for c in chains:
        assert c["span_us"] > 0 and c["ctas"] % 4 == 0
        for p in c["phases"]:
            assert p["span_us"] >= 0 and p["start_us"] >= 0
    launches = {}
    for name, ms, fl, by in tr.profile(1):
        launches[name] = launches.get(name, 0) + 1
    launches.pop("_empty_pair", None)
    assert sum(launches.values()) <= 6, launches                # ingest+enc | middle | bn_bwd | wgrad | adamw


def test_repeated_eager_steps_reuse_their_plans(chain_on):
    """The plan cache is keyed by the argument image: identical calls must hit it (an uninitialised byte in a group once made
    every call build a new plan and the captured call fail)."""
    from vla_b200 import DeviceDataset, Trainer, _lib
    m = make_module("multimodal", FULL, vo.init_state("multimodal", FULL, seed=4)).train()
    ds = DeviceDataset.synthetic(1024, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=1)
    tr = Trainer(m, ds, 512, use_graph=False)
    tr.step()
    n1 = _lib.lib().vla_chain_cached_plans(m._ensure_core().handle)
    for _ in range(5):
        tr.step()
    torch.cuda.synchronize()
    assert n1 >= 2 and _lib.lib().vla_chain_cached_plans(m._ensure_core().handle) == n1


def test_pinned_workspace_refuses_to_grow():
    """A live Trainer's captured graphs reference the workspace: a larger-batch call on the same handle must fail loudly
    instead of reallocating it."""
    from vla_b200 import DeviceDataset, Trainer
    m = make_module("rna2dna", FULL, vo.init_state("rna2dna", FULL, seed=4)).train()
    ds = DeviceDataset.synthetic(256, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=1)
    tr = Trainer(m, ds, 128, use_graph=True)
    tr.step()
    with pytest.raises(RuntimeError, match="pinned"):
        m(rna=torch.rand(512, FULL["A"], device="cuda"), site=torch.zeros(512, dtype=torch.long, device="cuda"))
    m(rna=torch.rand(64, FULL["A"], device="cuda"), site=torch.zeros(64, dtype=torch.long, device="cuda"))     # fits: fine
    tr.step()
    tr.close()
    m(rna=torch.rand(512, FULL["A"], device="cuda"), site=torch.zeros(512, dtype=torch.long, device="cuda"))    # released
