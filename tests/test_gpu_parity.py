"""CUDA path vs the oracle (and vs the committed reference fixtures) through the drop-in `src` surface.

Tolerances (BASELINE.json north_star): bf16 tensor-core paths 2e-2 relative on outputs, losses and gradients;
fp32 kernels (loss, latent, AdamW) 1e-5 relative; argmax of the site logits exact on every
row."""
import os

import numpy as np
import pytest
import torch

from golden.make_golden import CASES, SAMPLE_STRIDE, case_inputs
from oracle import vae_oracle as vo
from parity_util import (MATCHED_Q, MIN_COSINE_VS_EXACT, TOL_BF16, TOL_FP32, TOL_GRAD_VS_EXACT, assert_close, call_module, cosine,
                         is_pre_bn_bias, loss_for, make_module, oracle_step, rel_l2, to_t)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run_cuda(case, state, batch, eps, masks, cw, backward):
    kind, dims = case["kind"], case["dims"]
    m = make_module(kind, dims, state)
    m.train(case["train"])
    bt = {k: to_t(v) for k, v in batch.items()}
    a = bt["a"] if "a" in case["present"] else None
    b = bt["b"] if "b" in case["present"] else None
    s = bt["site"] if "site" in case["present"] else None
    with m.inject(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()]):
        out = call_module(m, kind, a, b, s)
    total, scal = loss_for(kind, out, bt, case["beta"], case["gamma"], to_t(cw))
    grads = None
    if backward:
        total.backward()
        grads = {k: (None if p.grad is None else p.grad.detach().cpu().numpy()) for k, p in m.named_parameters()}
    return m, out, float(total.item()), scal, grads


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_forward_loss_backward_vs_oracle(case):
    state, batch, eps, masks, cw = case_inputs(case)
    o_out, o_scal, o_grads, _ = oracle_step(case["kind"], case["dims"], state, batch, case["present"], eps, masks,
                                            case["beta"], case["gamma"], cw, train=case["train"])
    q_out, q_scal, q_grads, _ = oracle_step(case["kind"], case["dims"], state, batch, case["present"], eps, masks,
                                            case["beta"], case["gamma"], cw, train=case["train"], q=MATCHED_Q)
    m, out, total, (recon, cls, kld), grads = _run_cuda(case, state, batch, eps, masks, cw, backward=case["steps"] > 0)
    for prefix, ref in o_out["recon"].items():
        assert_close("recon." + prefix, out["recon"][prefix].detach().cpu().numpy(), ref, TOL_BF16)
    assert_close("mu", out["mu"].detach().cpu().numpy(), o_out["mu"], TOL_BF16)
    if o_out["logvar"] is not None:                  # (the autoencoders return the latent only)
        assert_close("logvar", out["logvar"].detach().cpu().numpy(), o_out["logvar"], TOL_BF16, atol=1e-3)
    np.testing.assert_allclose([total, recon, kld], [o_scal["total"], o_scal["recon"], o_scal["kld"]], rtol=TOL_BF16)
    if case["kind"] == "multimodal":
        np.testing.assert_allclose(cls, o_scal["cls"], rtol=TOL_BF16)
        # argmax of the site logits: exact on EVERY row (no filtering by margin)
        got = out["recon"]["decoder_c"].detach().cpu().numpy()
        ref = o_out["recon"]["decoder_c"]
        assert int((got.argmax(1) != ref.argmax(1)).sum()) == 0
    # same algorithm at the declared operand precision: everything within the bf16 tolerance
    for prefix, ref in q_out["recon"].items():
        assert_close("matched recon." + prefix, out["recon"][prefix].detach().cpu().numpy(), ref, TOL_BF16)
    np.testing.assert_allclose(total, q_scal["total"], rtol=TOL_BF16)
    if grads is not None:
        for name, ref in q_grads.items():
            g = grads[name]
            assert g is not None, name
            scale = np.linalg.norm(ref)
            if is_pre_bn_bias(name):
                # exactly-zero true gradient (BatchNorm removes the mean): compare against the layer's weight-gradient scale
                wname = name[:-4] + "weight"
                assert np.linalg.norm(g) <= 2e-2 * np.linalg.norm(q_grads[wname]) + 1e-4, name
                continue
            assert_close("matched grad." + name, g, ref, TOL_BF16, atol=1e-5 * max(scale, 1.0))
            # exact arithmetic: documented envelope (parity_util.TOL_GRAD_VS_EXACT)
            exact = o_grads[name]
            if np.linalg.norm(exact) > 1e-6:
                assert rel_l2(g, exact) <= TOL_GRAD_VS_EXACT, ("exact grad." + name, rel_l2(g, exact))
                assert cosine(g, exact) >= MIN_COSINE_VS_EXACT, ("cosine grad." + name, cosine(g, exact))
        for name, g in grads.items():
            if name not in o_grads:
                assert g is None, f"{name} should have no gradient"


@pytest.mark.parametrize("case", [c for c in CASES if c["steps"] > 0], ids=[c["name"] for c in CASES if c["steps"] > 0])
def test_against_reference_fixture(case):
    """Same run compared directly with what the unmodified reference produced (tests/golden/*.npz)."""
    fix = np.load(os.path.join(GOLDEN, case["name"] + ".npz"))
    state, batch, eps, masks, cw = case_inputs(case)
    m, out, total, (recon, cls, kld), grads = _run_cuda(case, state, batch, eps, masks, cw, backward=True)
    np.testing.assert_allclose([total, recon, cls, kld], fix["loss"], rtol=TOL_BF16, atol=1e-6)

    def check(prefix, arr):
        arr = np.asarray(arr, dtype=np.float64)
        if prefix + "|full" in fix:
            ref = fix[prefix + "|full"]
            got = arr.reshape(ref.shape)
        else:
            ref = fix[prefix + "|sample"]
            got = arr.reshape(-1)[::SAMPLE_STRIDE]
        tol = TOL_GRAD_VS_EXACT if prefix.startswith("grad.") else TOL_BF16
        assert rel_l2(got, ref) <= tol or np.linalg.norm(got - ref) <= 1e-5 * max(np.linalg.norm(ref), 1.0), \
            (prefix, rel_l2(got, ref))
        if prefix.startswith("grad.") and np.linalg.norm(ref) > 1e-6:
            assert cosine(got, ref) >= MIN_COSINE_VS_EXACT, (prefix, cosine(got, ref))

    for prefix, t in out["recon"].items():
        check("out.recon." + prefix, t.detach().cpu().numpy())
    check("out.mu", out["mu"].detach().cpu().numpy())
    for name, g in grads.items():
        if g is None:
            assert f"grad.{name}|none" in fix.files, name
        elif not is_pre_bn_bias(name):
            check("grad." + name, g)


def test_fp32_loss_kernel_matches_oracle():
    """The loss kernel alone is pure fp32: 1e-5 relative on values and gradients."""
    rng = np.random.default_rng(0)
    n, A, B, S, L = 257, 782, 572, 24, 20
    ra = rng.standard_normal((n, A)).astype(np.float32)
    a = rng.standard_normal((n, A)).astype(np.float32)
    rb = rng.uniform(0.001, 0.999, (n, B)).astype(np.float32)
    rb[0, :4] = [0.0, 1.0, 1e-30, 1 - 1e-7]          # exercises the -100 clamp and the 1e-12 floor
    b = rng.uniform(0, 1, (n, B)).astype(np.float32)
    rc = (3 * rng.standard_normal((n, S))).astype(np.float32)
    site = rng.integers(0, S, n)
    mu = rng.standard_normal((n, L)).astype(np.float32)
    lv = (0.5 * rng.standard_normal((n, L))).astype(np.float32)
    cw = rng.uniform(0.5, 2.0, S).astype(np.float32)
    beta, gamma = 3e-3, 1.7
    from src.utils.losses import vae_loss
    ts = [to_t(x).requires_grad_(True) for x in (ra, rb, rc, mu, lv)]
    total, recon, cls, kld = vae_loss(ts[0], to_t(a), ts[1], to_t(b), ts[2], to_t(site), ts[3], ts[4], beta=beta, gamma=gamma,
                                      class_weights=to_t(cw))
    total.backward()
    f64 = lambda x: x.astype(np.float64)
    out = dict(recon={"decoder_a": f64(ra), "decoder_b": f64(rb), "decoder_c": f64(rc)}, mu=f64(mu), logvar=f64(lv))
    scal, og = vo.loss_and_output_grads("multimodal", out, dict(a=f64(a), b=f64(b), site=site), beta, gamma, f64(cw))
    np.testing.assert_allclose([total.item(), recon, cls, kld], [scal["total"], scal["recon"], scal["cls"], scal["kld"]],
                               rtol=TOL_FP32)
    assert_close("g_recon_a", ts[0].grad.cpu().numpy(), og["recon"]["decoder_a"], TOL_FP32)
    gb, gb_ref = ts[1].grad.cpu().numpy(), og["recon"]["decoder_b"]
    finite = np.abs(gb_ref) < 1e6       # the clamped entries are +-1e12-scale; compare them separately
    assert_close("g_recon_b", gb[finite], gb_ref[finite], TOL_FP32)
    np.testing.assert_allclose(gb[~finite], gb_ref[~finite], rtol=1e-4)
    assert_close("g_recon_c", ts[2].grad.cpu().numpy(), og["recon"]["decoder_c"], TOL_FP32 * 5)
    assert_close("g_mu", ts[3].grad.cpu().numpy(), og["mu"], TOL_FP32)
    assert_close("g_logvar", ts[4].grad.cpu().numpy(), og["logvar"], TOL_FP32)


def test_no_cpu_fallback():
    from src.models import RNA2DNAVAE
    m = RNA2DNAVAE(50, 36, 5, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback|CPU"):
        m(rna=torch.zeros(4, 50), site=torch.zeros(4, dtype=torch.long))
    m = m.cuda()
    with pytest.raises(RuntimeError, match="CPU"):
        m(rna=torch.zeros(4, 50), site=torch.zeros(4, dtype=torch.long))


def test_eval_is_stochastic_and_none_inputs():
    """reparameterize samples in eval mode too (reference vae.py:11-15); all-None -> Nones."""
    from src.models import MultiModalVAE
    m = MultiModalVAE(50, 36, 5, 8).cuda().eval()
    assert m() == (None, None, None, None, None)
    x = torch.rand(16, 50, device="cuda")
    with torch.no_grad():
        o1 = m(a=x)
        o2 = m(a=x)
    assert torch.equal(o1[3], o2[3]) and torch.equal(o1[4], o2[4])       # mu / logvar deterministic
    assert not torch.equal(o1[0], o2[0])                                  # decoders see a fresh epsilon
    assert o1[0].shape == (16, 50) and o1[1].shape == (16, 36) and o1[2].shape == (16, 5)
    assert float(o1[1].min()) >= 0.0 and float(o1[1].max()) <= 1.0


def test_saturated_logits_fused_vs_functional():
    """for logits beyond +16.6 fp32 sigmoid rounds to exactly 1 and the reference's
    F.binary_cross_entropy(sigmoid(x), t) clamps log(1 - y) at -100 (loss 100 (1 - t), gradient 0 through the sigmoid); the
    loss fused into the last decoder layer works on the logit (loss (1 - t) x, gradient y - t).  The functional path follows
    ATen; this test pins the size of the deliberate difference on columns whose bias is pushed to +30 / -30 (below -27.6
    ATen's backward floor 1e-12 on y (1 - y) shrinks the reference's gradient), and that nothing else differs."""
    from src.utils.directional_losses import rna2dna_loss
    from vla_b200 import DeviceDataset, Trainer
    kind, dims, n = "rna2dna", dict(A=782, B=572, S=24, L=20, E=32), 64
    state = vo.init_state(kind, dims, seed=31)
    bias = state["decoder_dna.fc.4.bias"].copy()
    bias[:5] = 30.0            # saturated positive
    bias[5:10] = -30.0         # saturated negative: the loss values agree, the reference's gradient is floored
    state["decoder_dna.fc.4.bias"] = bias
    tpm, beta_v, site = vo.synthetic_batch(n, dims, seed=31)
    eps, masks = vo.synthetic_noise(n, dims, kind, seed=31)
    inj = dict(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()])
    # fused engine path
    m1 = make_module(kind, dims, state).train()
    tr = Trainer(m1, DeviceDataset(tpm, beta_v, site, "cuda"), n, beta_kl=1e-3, use_graph=False)
    tr.injected = inj
    tr.forward_backward()
    torch.cuda.synchronize()
    fused_total = tr.losses()[0]
    named = {name: (off, int(np.prod(shape))) for (name, k_, off, shape) in m1._ensure_core().infos if name == "decoder_dna.fc.4.bias"}
    off, cnt = named["decoder_dna.fc.4.bias"]
    g_fused = tr.grads[off:off + cnt].cpu().numpy()
    tr.close()
    # functional (autograd) path = the reference's arithmetic
    m2 = make_module(kind, dims, state).train()
    with m2.inject(**inj):
        recon, mu, lv = m2(rna=to_t(tpm), site=to_t(site))
    total, _, _ = rna2dna_loss(recon, to_t(beta_v), mu, lv, beta=1e-3)
    total.backward()
    g_func = dict(m2.named_parameters())["decoder_dna.fc.4.bias"].grad.cpu().numpy()
    t = beta_v[:, :5].astype(np.float64)
    x = 30.0                                                       # the logit is the bias up to O(1)
    # functional: 100 (1 - t) and zero gradient; fused: (1 - t) x and gradient sum(1 - t)
    assert np.all(np.abs(g_func[:5]) < 1e-3)
    np.testing.assert_allclose(g_fused[:5], (1 - t).sum(0), rtol=2e-2)
    expected_gap = ((1 - t) * (100.0 - x)).sum()
    assert abs((float(total) - fused_total) - expected_gap) <= 0.1 * expected_gap
    # negative saturation (x = -30): y (1 - y) = 9e-14 falls below ATen's backward floor 1e-12, so the reference scales the
    # gradient by y (1 - y) / 1e-12 = 0.09; the fused path keeps y - t = -t
    tn = beta_v[:, 5:10].astype(np.float64)
    np.testing.assert_allclose(g_fused[5:10], -tn.sum(0), rtol=2e-2)
    assert np.all(np.abs(g_func[5:10]) < 0.2 * np.abs(g_fused[5:10]))
    # everything that is not saturated agrees
    np.testing.assert_allclose(g_fused[10:], g_func[10:], rtol=2e-2, atol=2e-2 * np.abs(g_func[10:]).max())
