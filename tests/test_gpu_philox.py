"""The random streams the benchmark actually runs: epsilon of reparameterize (reference src/models/vae.py:13, torch.randn_like)
and the Dropout(0.1) keep-masks (src/models/encoders.py:16, 34, 38) come from a counter-based Philox4x32-10 generator keyed by
(seed, step, layer, element).  Every parity test injects both; here the generator itself is checked: moments of epsilon,
keep rate, and that steps, layers and seeds (= ranks under data parallelism) draw different streams."""
import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo
from parity_util import make_module

pytestmark = pytest.mark.gpu
FULL = dict(A=782, B=572, S=24, L=20, E=32)


def _trainer(kind, batch, seed, state=None, n_batches=2):
    from vla_b200 import DeviceDataset, Trainer
    state = state or vo.init_state(kind, FULL, seed=3)
    m = make_module(kind, FULL, state).train()
    ds = DeviceDataset.synthetic(batch * n_batches, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=1)
    return m, Trainer(m, ds, batch, lr=1e-4, seed=seed, use_graph=True)


def _corr(a, b):
    a = a.double().flatten() - a.double().mean()
    b = b.double().flatten() - b.double().mean()
    return float((a @ b) / (a.norm() * b.norm()))


def test_epsilon_moments_and_step_streams():
    from vla_b200.engine import workspace_view
    B, L, steps = 4096, FULL["L"], 100
    m, tr = _trainer("rna2dna", B, seed=0)
    s1 = s2 = s4 = 0.0
    inside = 0
    draws = []
    for t in range(steps):
        tr.step()
        torch.cuda.synchronize()
        e = workspace_view(tr.core, "eps", B).double()
        assert e.shape == (B, L)
        s1 += float(e.sum()); s2 += float((e * e).sum()); s4 += float((e ** 4).sum())
        inside += int((e.abs() < 1).sum())
        if t < 4:
            draws.append(e.clone())
    n = B * L * steps
    mean, var = s1 / n, s2 / n - (s1 / n) ** 2
    kurt = (s4 / n) / var ** 2
    assert abs(mean) < 5 / np.sqrt(n), mean
    assert abs(var - 1) < 5 * np.sqrt(2 / n), var
    assert abs(kurt - 3) < 5 * np.sqrt(96 / n) + 2e-3, kurt     # (24-bit uniforms cut the tails at 5.8 sigma)
    assert abs(inside / n - 0.6826895) < 5 * np.sqrt(0.6827 * 0.3173 / n), inside / n
    for i in range(len(draws)):
        for j in range(i + 1, len(draws)):
            assert not torch.equal(draws[i], draws[j])
            assert abs(_corr(draws[i], draws[j])) < 6 / np.sqrt(B * L), (i, j)
    # rows and latent columns are independent draws
    e = draws[0]
    assert abs(_corr(e[:-1], e[1:])) < 6 / np.sqrt(B * L)
    assert abs(_corr(e[:, :-1], e[:, 1:])) < 6 / np.sqrt(B * L)


def test_dropout_keep_rate_layer_and_step_streams():
    """EncoderB has two Dropout(0.1) layers.  BatchNorm bias = +8 keeps ReLU open, so an exact zero is a dropped unit."""
    from vla_b200.engine import workspace_view
    B = 4096
    kind = "dna2rna"
    state = vo.init_state(kind, FULL, seed=3)
    for k in state:
        if k.endswith((".fc.1.bias", ".fc.5.bias")):
            state[k] = np.full_like(state[k], 8.0)
    m, tr = _trainer(kind, B, seed=0, state=state)
    masks = []
    for t in range(3):
        tr.step()
        torch.cuda.synchronize()
        a0 = workspace_view(tr.core, "act", B, 0, 0)[:, :512].float()
        a1 = workspace_view(tr.core, "act", B, 0, 1)[:, :256].float()
        k0, k1 = (a0 != 0), (a1 != 0)
        for k, width in ((k0, 512), (k1, 256)):
            rate = float(k.double().mean())
            assert abs(rate - 0.9) < 5 * np.sqrt(0.09 / (B * width)), rate
        # kept units are scaled by 1 / (1 - p): BN output ~ N(8, 1) -> 8.9 on average
        assert abs(float(a0[k0].mean()) - 8.0 / 0.9) < 0.05
        masks.append((k0.clone(), k1.clone()))
    n = B * 256
    for t in range(3):                                 # the two layers of one step draw different streams
        assert abs(_corr(masks[t][0][:, :256], masks[t][1])) < 6 / np.sqrt(n)
    for i in range(3):                                 # consecutive steps draw different streams
        for j in range(i + 1, 3):
            assert abs(_corr(masks[i][0], masks[j][0])) < 6 / np.sqrt(B * 512)
            assert abs(_corr(masks[i][1], masks[j][1])) < 6 / np.sqrt(n)


def test_seed_selects_the_stream():
    """Data parallel ranks pass seed = rank: different seeds must give unrelated epsilon, the same seed the same epsilon."""
    from vla_b200.engine import workspace_view
    B = 2048
    eps = {}
    for tag, seed in (("a", 0), ("b", 1), ("a2", 0)):
        m, tr = _trainer("rna2dna", B, seed=seed)
        tr.step()
        torch.cuda.synchronize()
        eps[tag] = workspace_view(tr.core, "eps", B).clone()
    assert torch.equal(eps["a"], eps["a2"])
    assert abs(_corr(eps["a"], eps["b"])) < 6 / np.sqrt(B * FULL["L"])


def test_autograd_path_draws_fresh_epsilon_per_call():
    """Module forward (the script path): every call advances the Philox offset (reference: torch.randn_like per call)."""
    from vla_b200.engine import workspace_view
    m = make_module("rna2dna", FULL, vo.init_state("rna2dna", FULL, seed=3)).train()
    x = torch.rand(512, FULL["A"], device="cuda")
    s = torch.randint(0, FULL["S"], (512,), device="cuda")
    seen = []
    for _ in range(3):
        m(rna=x, site=s)
        torch.cuda.synchronize()
        seen.append(workspace_view(m._core, "eps", 512).clone())
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])
    assert abs(_corr(seen[0], seen[1])) < 6 / np.sqrt(512 * FULL["L"])
    assert abs(float(seen[0].mean())) < 5 / np.sqrt(512 * FULL["L"])
