"""Epoch-level behaviour of the Trainer that the reference's loops rely on: the ragged last batch (the DataLoaders keep it:
train_rna2dna.py:57-67, vae_cross_modality_cv.py:121, no drop_last) and batch assembly by index on the device
(src/data/dataset.py:35-39 + DataLoader collate as one kernel)."""
import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo
from parity_util import MATCHED_Q, TOL_BF16, assert_close, is_pre_bn_bias, make_module, rel_l2, to_t

pytestmark = pytest.mark.gpu

FULL = dict(A=782, B=572, S=24, L=20, E=32)


@pytest.mark.parametrize("kind,use_graph", [("rna2dna", True), ("multimodal", False)])
def test_epoch_with_ragged_last_batch_matches_oracle(kind, use_graph):
    from vla_b200 import DeviceDataset, Trainer
    batch, tail, dims = 64, 23, FULL
    n = 2 * batch + tail
    state = vo.init_state(kind, dims, seed=13)
    tpm, beta_v, site = vo.synthetic_batch(n, dims, seed=13)
    eps, masks = vo.synthetic_noise(batch, dims, kind, seed=13)
    beta, gamma, lr = 2e-3, 1.5, 5e-4
    # oracle: two full batches and the 23-row tail, as the DataLoader yields them
    st = {k: (v.astype(np.float64) if v.dtype.kind == "f" else v.copy()) for k, v in state.items()}
    opt, step = vo.adamw_init(st)
    ref_losses = []
    for lo, rows in ((0, batch), (batch, batch), (2 * batch, tail)):
        b = dict(a=tpm[lo:lo + rows].astype(np.float64), b=beta_v[lo:lo + rows].astype(np.float64), site=site[lo:lo + rows])
        mk = {k: v[:rows] for k, v in masks.items()}
        scal, _, _, step = vo.train_step(kind, dims, st, opt, step, b, eps[:rows].astype(np.float64), mk, beta=beta, gamma=gamma, q=MATCHED_Q)
        ref_losses.append(scal["total"])

    m = make_module(kind, dims, state).train()
    ds = DeviceDataset(tpm, beta_v, site, "cuda")
    tr = Trainer(m, ds, batch, lr=lr, weight_decay=1e-5, beta_kl=beta, gamma=gamma, use_graph=use_graph)
    tr.injected = dict(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()])
    assert tr.run_epoch() == 3
    torch.cuda.synchronize()
    np.testing.assert_allclose(tr.losses()[0], ref_losses[-1], rtol=TOL_BF16)       # the tail's loss
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    for name, ref in st.items():
        got = sd[name]
        if name.endswith("num_batches_tracked"):
            assert int(got) == 3, name
        elif name.endswith(("running_mean", "running_var")):
            assert_close(name, got, ref, TOL_BF16, atol=4 * lr * 3 * np.sqrt(ref.size))
        elif is_pre_bn_bias(name):
            assert np.abs(got - state[name]).max() <= 1.05 * lr * 3 + 1e-7, name
        else:
            d_ref, d = ref - state[name].astype(np.float64), got.astype(np.float64) - state[name].astype(np.float64)
            assert rel_l2(d, d_ref) <= 0.2, (name, rel_l2(d, d_ref))
    # a second epoch starts at row 0 again and takes three more steps
    assert tr.run_epoch() == 3 and tr.steps == 6
    tr.close()


def test_one_row_tail_raises_like_batchnorm():
    from vla_b200 import DeviceDataset, Trainer
    m = make_module("rna2dna", FULL, vo.init_state("rna2dna", FULL, seed=1)).train()
    ds = DeviceDataset.synthetic(65, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=1)
    tr = Trainer(m, ds, 64, use_graph=False)
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        tr.run_epoch()
    tr.close()


def test_gather_rows_on_device():
    from vla_b200 import DeviceDataset
    ds = DeviceDataset.synthetic(1000, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=2)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    idx = torch.randperm(1000, device="cuda", generator=g)[:257]
    slot = DeviceDataset.synthetic(257, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=3)
    ds.gather_into(idx, slot)
    torch.cuda.synchronize()
    assert torch.equal(slot.tpm, ds.tpm[idx]) and torch.equal(slot.beta, ds.beta[idx]) and torch.equal(slot.site, ds.site[idx])
    # odd widths take the narrower paths
    odd = DeviceDataset(torch.rand(50, 7), torch.rand(50, 6), torch.arange(50) % 3, "cuda")
    oslot = DeviceDataset(torch.zeros(9, 7), torch.zeros(9, 6), torch.zeros(9, dtype=torch.long), "cuda")
    oi = torch.tensor([3, 3, 49, 0, 17, 5, 8, 21, 2], device="cuda")
    odd.gather_into(oi, oslot)
    assert torch.equal(oslot.tpm, odd.tpm[oi]) and torch.equal(oslot.beta, odd.beta[oi]) and torch.equal(oslot.site, odd.site[oi])


def test_shuffled_epoch_through_batch_slot_trains():
    """DataLoader(shuffle=True) semantics with the dataset left in place: a permutation per epoch, each batch gathered into
    the slot the captured step graph reads."""
    from vla_b200 import DeviceDataset, Trainer
    m = make_module("rna2dna", FULL, vo.init_state("rna2dna", FULL, seed=6)).train()
    ds = DeviceDataset.synthetic(4 * 128, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=4)
    slot = DeviceDataset.synthetic(128, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=5)
    tr = Trainer(m, slot, 128, use_graph=True)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    first = last = None
    for epoch in range(6):
        perm = torch.randperm(len(ds), device="cuda", generator=g)
        for i in range(4):
            ds.gather_into(perm[i * 128:(i + 1) * 128], slot)
            tr.step()
        last = tr.losses()[0]
        first = last if first is None else first
    assert np.isfinite(last) and last < first
    tr.close()


def test_scaled_loss_backward_scales_gradients():
    """loss.backward() on 0.5 * loss: the stored gradients are scaled in one launch of the library (no ATen multiplies)."""
    from parity_util import call_module, loss_for
    kind, n = "rna2dna", 32
    state = vo.init_state(kind, FULL, seed=8)
    tpm, beta_v, site = vo.synthetic_batch(n, FULL, seed=8)
    eps, masks = vo.synthetic_noise(n, FULL, kind, seed=8)
    grads = []
    for scale in (1.0, 0.5):
        m = make_module(kind, FULL, state).train()
        bt = dict(a=to_t(tpm), b=to_t(beta_v), site=to_t(site))
        with m.inject(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()]):
            out = call_module(m, kind, bt["a"], None, bt["site"])
        total, _ = loss_for(kind, out, bt, 1e-3, 1.0, None)
        (total * scale).backward()
        torch.cuda.synchronize()
        grads.append({k: p.grad.clone() for k, p in m.named_parameters()})
    for k in grads[0]:
        assert torch.allclose(grads[1][k], 0.5 * grads[0][k], rtol=1e-5, atol=1e-7), k
