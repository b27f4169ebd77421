"""Fused train step (vla_train_step through vla_b200.Trainer), the AdamW kernel and FusedAdamW against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo
from parity_util import MATCHED_Q, TOL_BF16, TOL_FP32, assert_close, is_pre_bn_bias, make_module, rel_l2, to_t

pytestmark = pytest.mark.gpu

FULL = dict(A=782, B=572, S=24, L=20, E=32)
SMALL = dict(A=50, B=36, S=5, L=10, E=16)
MANY_SITES = dict(A=50, B=36, S=40, L=10, E=16)     # more than 32 classes: the CE term cannot be fused into a 32-column chunk,
                                                    # the step falls back to the stand-alone loss kernel


def _oracle_train(kind, dims, state, data, n_steps, batch, eps, masks, beta, gamma, cw, q):
    st = {k: (v.astype(np.float64) if v.dtype.kind == "f" else v.copy()) for k, v in state.items()}
    opt, step = vo.adamw_init(st)
    losses = []
    n_batches = len(data["site"]) // batch
    for i in range(n_steps):
        lo = (i % n_batches) * batch
        b = {k: (v[lo:lo + batch].astype(np.float64) if v.dtype.kind == "f" else v[lo:lo + batch]) for k, v in data.items()}
        scal, _, _, step = vo.train_step(kind, dims, st, opt, step, b, eps.astype(np.float64), masks, beta=beta, gamma=gamma,
                                         class_weights=None if cw is None else cw.astype(np.float64), q=q)
        losses.append([scal["total"], scal["recon"], scal["cls"], scal["kld"]])
    return st, np.array(losses)


@pytest.mark.parametrize("kind,dims,batch,use_graph", [("rna2dna", FULL, 64, True), ("multimodal", SMALL, 48, True),
                                                         ("dna2rna", FULL, 40, False), ("multimodal", FULL, 64, True),
                                                         ("rna2dna_ae", FULL, 64, True), ("dna2rna_ae", SMALL, 40, False),
                                                         ("multimodal", MANY_SITES, 48, True)])
def test_fused_train_steps_match_oracle(kind, dims, batch, use_graph):
    from vla_b200 import DeviceDataset, Trainer
    n_steps, n_batches = 4, 3
    state = vo.init_state(kind, dims, seed=21)
    tpm, beta_v, site = vo.synthetic_batch(batch * n_batches, dims, seed=21)
    data = dict(a=tpm, b=beta_v, site=site)
    eps, masks = vo.synthetic_noise(batch, dims, kind, seed=21)
    cw = vo.balanced_class_weights(site, dims["S"]) if kind == "multimodal" else None
    beta, gamma = 2e-3, 1.5
    ref_state, ref_losses = _oracle_train(kind, dims, state, data, n_steps, batch, eps, masks, beta, gamma, cw, MATCHED_Q)

    m = make_module(kind, dims, state).train()
    ds = DeviceDataset(tpm, beta_v, site, "cuda")
    tr = Trainer(m, ds, batch, lr=5e-4, weight_decay=1e-5, beta_kl=beta, gamma=gamma, class_weights=to_t(cw), use_graph=use_graph)
    tr.injected = dict(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()])
    got_losses = []
    for _ in range(n_steps):
        tr.step()
        got_losses.append(tr.losses())
    got_losses = np.array(got_losses)
    np.testing.assert_allclose(got_losses[:, 0], ref_losses[:, 0], rtol=TOL_BF16)
    np.testing.assert_allclose(got_losses[:, 1], ref_losses[:, 1], rtol=TOL_BF16)
    np.testing.assert_allclose(got_losses[:, 3], ref_losses[:, 3], rtol=TOL_BF16)
    if kind == "multimodal":
        np.testing.assert_allclose(got_losses[:, 2], ref_losses[:, 2], rtol=TOL_BF16)
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    lr = 5e-4
    for name, ref in ref_state.items():
        got = sd[name]
        if name.endswith("num_batches_tracked"):
            assert int(got) == n_steps, name
        elif name.endswith(("running_mean", "running_var")):
            assert_close(name, got, ref, TOL_BF16, atol=4 * lr * n_steps * np.sqrt(ref.size))
        else:
            # Adam's first steps move every weight by ~ +-lr whatever the gradient scale: compare the displacement
            delta_ref = ref - state[name].astype(np.float64)
            delta = got.astype(np.float64) - state[name].astype(np.float64)
            if is_pre_bn_bias(name):
                assert np.abs(delta).max() <= 1.05 * lr * n_steps + 1e-7, name      # sign of rounding noise: bounded only
                continue
            assert rel_l2(delta, delta_ref) <= 0.2, (name, rel_l2(delta, delta_ref))
            assert_close(name, got, ref, 1e-2, atol=1e-6)


@pytest.mark.parametrize("kind,dims,batch", [("rna2dna", FULL, 64), ("multimodal", FULL, 200), ("dna2rna", FULL, 40), ("rna2dna_ae", FULL, 64)])
def test_head_block_train_steps_match_oracle(kind, dims, batch, monkeypatch):
    """The opt-in head block (VLA_HEADBLOCK=1: BatchNorm apply + heads + latent + first decoder layer, and their backward, as
    one CUDA-core launch each way, fp32) against the same oracle at the same tolerances."""
    monkeypatch.setenv("VLA_HEADBLOCK", "1")
    test_fused_train_steps_match_oracle(kind, dims, batch, False)
    launches = {}
    from vla_b200 import DeviceDataset, Trainer
    m = make_module(kind, dims, vo.init_state(kind, dims, seed=2)).train()
    ds = DeviceDataset.synthetic(batch * 2, dims["A"], dims["B"], dims["S"], "cuda", seed=1)
    tr = Trainer(m, ds, batch, use_graph=True)
    tr.step()
    for name, ms, fl, by in tr.profile(1):
        launches[name] = launches.get(name, 0) + 1
    assert launches.get("head_block_fwd") == 1 and launches.get("head_block_bwd") == 1, launches
    tr.close()


@pytest.mark.parametrize("kind,dims,batch", [("rna2dna", FULL, 200), ("multimodal", FULL, 136), ("dna2rna_ae", SMALL, 40),
                                             ("multimodal", dict(A=50, B=36, S=5, L=100, E=16), 70)])
def test_latent_backward_in_gemm_epilogue(kind, dims, batch, monkeypatch):
    """The latent backward (reparameterisation + KL backward, /reference/src/models/vae.py:11-15, src/utils/losses.py:44) runs
    in the epilogue of the GEMM that produces dL/dz: no latent_bwd launch in the step, and the same parameters after three
    steps as with the separate launch (VLA_FUSE_LATBWD=0) -- ragged row blocks, latent widths 10 / 20 / 100 (four column
    chunks), one and several modalities, autoencoder mode."""
    from vla_b200 import DeviceDataset, Trainer
    state = vo.init_state(kind, dims, seed=9)
    tpm, beta_v, site = vo.synthetic_batch(batch * 2, dims, seed=9)
    eps, masks = vo.synthetic_noise(batch, dims, kind, seed=9)
    out = {}
    for fused in (True, False):
        monkeypatch.setenv("VLA_FUSE_LATBWD", "1" if fused else "0")
        m = make_module(kind, dims, state).train()
        ds = DeviceDataset(tpm, beta_v, site, "cuda")
        tr = Trainer(m, ds, batch, lr=5e-4, weight_decay=1e-5, beta_kl=3e-2, gamma=1.5, use_graph=True)
        tr.injected = dict(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()])
        for _ in range(3):
            tr.step()
        names = [name for name, ms, fl, by in tr.profile(1)]
        assert ("latent_bwd" in names) == (not fused), names
        assert "dgrad_dec_l0" in names
        out[fused] = ({k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}, np.array(tr.losses()))
        tr.close()
    np.testing.assert_allclose(out[True][1], out[False][1], rtol=1e-5)
    for name, ref in out[False][0].items():
        if ref.dtype.kind == "f" and not is_pre_bn_bias(name) and not name.endswith(("running_mean", "running_var")):
            # same arithmetic on the same fp32 dL/dz; the weight gradients' split-K red.add order is the only free variable
            # (it can flip Adam's first +-lr steps of an element whose gradient is rounding noise: compare displacements)
            d_ref = ref.astype(np.float64) - state[name].astype(np.float64)
            d_got = out[True][0][name].astype(np.float64) - state[name].astype(np.float64)
            if np.linalg.norm(d_ref) > 0:
                assert rel_l2(d_got, d_ref) <= 0.02, (name, rel_l2(d_got, d_ref))
            assert_close(name, out[True][0][name], ref, 2e-3, atol=1e-6)


def test_adamw_kernel_fp32():
    """vla_adamw alone (fp32 arithmetic) against the oracle's AdamW: 1e-5 relative on p, and on m, v."""
    from vla_b200 import _lib
    kind, dims = "multimodal", SMALL
    state = vo.init_state(kind, dims, seed=5)
    m = make_module(kind, dims, state)
    core = m._ensure_core()
    n = core.n_params
    rng = np.random.default_rng(0)
    ea = torch.zeros(n, device="cuda")
    eas = torch.zeros(n, device="cuda")
    st = {k: v.astype(np.float64) for k, v in state.items() if not vo.is_buffer(k)}
    opt, step = vo.adamw_init(st)
    named = dict(m.named_parameters())
    for it in range(3):
        grads_np = {k: rng.standard_normal(v.shape) * 10.0 ** rng.integers(-4, 2) for k, v in st.items()}
        flat = torch.zeros(n, device="cuda")
        for (name, kind_, off, shape) in core.infos:
            if kind_ == _lib.TENSOR_PARAM:
                cnt = int(np.prod(shape))
                flat[off:off + cnt] = to_t(grads_np[name].astype(np.float32)).reshape(-1)
        args = _lib.AdamWArgs(params=core.arena.data_ptr(), grads=flat.data_ptr(), exp_avg=ea.data_ptr(), exp_avg_sq=eas.data_ptr(),
                              lr=3e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2, step=it + 1)
        _lib.check(_lib.lib().vla_adamw(core.handle, C.byref(args), None), "vla_adamw")
        step = vo.adamw_step(st, {k: g.astype(np.float32).astype(np.float64) for k, g in grads_np.items()}, opt, step, lr=3e-3,
                             weight_decay=1e-2)
    torch.cuda.synchronize()
    for name, ref in st.items():
        assert_close(name, named[name].detach().cpu().numpy(), ref, TOL_FP32, atol=1e-7)


def test_fused_adamw_matches_torch_adamw():
    """FusedAdamW on the autograd path == torch.optim.AdamW on the same gradients."""
    from src.utils.directional_losses import rna2dna_loss
    from vla_b200 import FusedAdamW
    kind, dims, n = "rna2dna", SMALL, 32
    state = vo.init_state(kind, dims, seed=9)
    tpm, beta_v, site = vo.synthetic_batch(n, dims, seed=9)
    eps, masks = vo.synthetic_noise(n, dims, kind, seed=9)
    models = [make_module(kind, dims, state).train() for _ in range(2)]
    opts = [torch.optim.AdamW(models[0].parameters(), lr=5e-4, weight_decay=1e-5), FusedAdamW(models[1], lr=5e-4, weight_decay=1e-5)]
    for _ in range(3):
        for m, opt in zip(models, opts):
            with m.inject(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()]):
                recon, mu, lv = m(rna=to_t(tpm), site=to_t(site))
            loss, _, _ = rna2dna_loss(recon, to_t(beta_v), mu, lv, beta=1e-3)
            opt.zero_grad()
            loss.backward()
            opt.step()
    a, b = models[0].state_dict(), models[1].state_dict()
    for k in a:
        if is_pre_bn_bias(k) or k.endswith("running_mean"):
            continue
        assert_close(k, b[k].float().cpu().numpy(), a[k].float().cpu().numpy(), 1e-4, atol=2e-6)
