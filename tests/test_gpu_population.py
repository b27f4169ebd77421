"""Population of independent models on one GPU (vla_b200.Population; BASELINE configs[4]): concurrent members equal the same
models trained alone, and the per-epoch control flow (beta warm-up, ReduceLROnPlateau, early stopping, best-state restore)
follows the reference loops (optimize_hyperparameters.py:68-133, vae_cross_modality_cv.py:113-196)."""
import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo
from parity_util import is_pre_bn_bias, make_module, rel_l2

pytestmark = pytest.mark.gpu

SPECS = [("multimodal", dict(A=50, B=36, S=5, L=10, E=16), dict(lr=1e-3, weight_decay=1e-4, beta_start=2e-3, gamma=2.0)),
         ("multimodal", dict(A=50, B=36, S=5, L=37, E=64), dict(lr=3e-4, weight_decay=1e-6, beta_start=5e-4, gamma=0.7)),
         ("multimodal", dict(A=50, B=36, S=5, L=24, E=32), dict(lr=5e-3, weight_decay=1e-5, beta_start=1e-3, gamma=1.0))]


def _members():
    out = []
    for i, (kind, dims, hyper) in enumerate(SPECS):
        state = vo.init_state(kind, dims, seed=40 + i)
        out.append((kind, dims, hyper, state))
    return out


def test_concurrent_members_equal_solo_training():
    from vla_b200 import DeviceDataset, Population, Trainer
    batch, n_steps = 48, 5
    tpm, beta_v, site = vo.synthetic_batch(batch * 3, SPECS[0][1], seed=3)
    ds = DeviceDataset(tpm, beta_v, site, "cuda")
    members = _members()
    specs = [dict(model=make_module(k, d, st, device="cpu"), seed=7 + i, **h) for i, (k, d, h, st) in enumerate(members)]
    pop = Population(specs, ds, batch)
    pop.begin_epoch(50)
    pop.step(n_steps)
    pop_losses = pop.losses()
    for i, (k, d, h, st) in enumerate(members):
        m = make_module(k, d, st).train()
        tr = Trainer(m, ds, batch, lr=h["lr"], weight_decay=h["weight_decay"], beta_kl=h["beta_start"], gamma=h["gamma"], seed=7 + i)
        for _ in range(n_steps):
            tr.step()
        solo = tr.losses()
        np.testing.assert_allclose(pop_losses[i], solo, rtol=1e-4)
        # split-K red.add order is the only non-determinism.  It matters only where the true gradient is exactly zero (Linear
        # biases in front of a train-mode BatchNorm): there the sign of the rounding residue decides Adam's +-lr step.
        got = dict(pop.members[i].model.named_parameters())
        for name, p in m.named_parameters():
            if is_pre_bn_bias(name):
                assert float((got[name] - p).abs().max()) <= 2.1 * h["lr"] * n_steps, name
            else:
                err = rel_l2(got[name].detach().cpu().numpy(), p.detach().cpu().numpy())
                assert err < 1e-4, (i, name, err)
        tr.close()
    pop.close()


def test_epoch_control_flow():
    from vla_b200 import DeviceDataset, Population, fused_vae_loss
    batch = 32
    dims = SPECS[0][1]
    tpm, beta_v, site = vo.synthetic_batch(batch * 2, dims, seed=5)
    ds = DeviceDataset(tpm, beta_v, site, "cuda")
    members = _members()[:2]
    specs = [dict(model=make_module(k, d, st, device="cpu"), **h) for k, d, h, st in members]
    pop = Population(specs, ds, batch, beta_warmup_epochs=4, lr_factor=0.5, lr_patience=1, patience=3)
    val = [(ds.tpm[:batch], ds.beta[:batch], ds.site[:batch])]

    def loss_fn(mem, model, b):
        ra, rb, rc, mu, lv = model(a=b[0], b=b[1], site=b[2])
        return fused_vae_loss(ra, b[0], rb, b[1], rc, b[2], mu, lv, beta=1e-3, gamma=mem.hyper["gamma"])[0]

    pop.begin_epoch(1)
    assert pop.members[0].trainer.hyper[2] == pytest.approx(0.25 * SPECS[0][2]["beta_start"])        # beta warm-up
    pop.step(2)
    v0 = pop.validate(val, loss_fn)
    assert all(np.isfinite(v0)) and len(v0) == 2
    assert pop.end_epoch(v0) == 2
    snap = pop.members[0].best_state[0].clone()
    # member 0 stops improving: lr halves after `lr_patience` bad epochs, training stops after `patience`
    lrs = []
    for epoch in range(2, 6):
        pop.begin_epoch(epoch)
        pop.step(2)
        alive = pop.end_epoch([v0[0] + 1.0, v0[1] - epoch])                                           # member 1 keeps improving
        lrs.append(pop.members[0].trainer.hyper[0])
    assert lrs[0] == pytest.approx(SPECS[0][2]["lr"]) and lrs[1] == pytest.approx(0.5 * SPECS[0][2]["lr"])
    assert pop.members[0].stopped and not pop.members[1].stopped and alive == 1
    assert pop.members[1].trainer.hyper[0] == pytest.approx(SPECS[1][2]["lr"])
    with pytest.raises(ValueError):
        pop.validate([(torch.zeros(batch + 1, 50, device="cuda"),)], loss_fn)
    pop.restore_best()
    assert torch.equal(pop.members[0].trainer.core.arena, snap)
    pop.step(1)                                                                                       # stopped members are skipped
    pop.synchronize()
    assert torch.equal(pop.members[0].trainer.core.arena, snap)
    pop.close()


def test_population_epoch_with_ragged_tail_equals_solo_epochs():
    """Population.run_epoch walks every member's dataset including the ragged last batch, like Trainer.run_epoch alone."""
    from vla_b200 import DeviceDataset, Population, Trainer
    dims = dict(A=782, B=572, S=24, L=20, E=32)
    batch, n = 64, 64 * 2 + 17
    datasets = [DeviceDataset.synthetic(n, dims["A"], dims["B"], dims["S"], "cuda", seed=40 + i) for i in range(2)]
    states = [vo.init_state("multimodal", dims, seed=50 + i) for i in range(2)]
    pop = Population([dict(model=make_module("multimodal", dims, states[i]), seed=i) for i in range(2)], datasets, batch)
    pop.run_epoch()
    pop.synchronize()
    for i in range(2):
        solo_m = make_module("multimodal", dims, states[i]).train()
        tr = Trainer(solo_m, datasets[i], batch, seed=i)
        assert tr.run_epoch() == 3
        torch.cuda.synchronize()
        assert pop.members[i].trainer.steps == 3
        got = {k: v.detach() for k, v in pop.members[i].model.state_dict().items()}
        for k, v in solo_m.state_dict().items():
            if v.dtype.is_floating_point:
                assert float((got[k] - v).abs().max()) <= 2.1 * 5e-4 * 3 + 1e-6, k
        tr.close()
    pop.close()


HETERO = [("multimodal", dict(A=782, B=572, S=24, L=10, E=16), 48), ("multimodal", dict(A=782, B=572, S=24, L=100, E=64), 48),
          ("multimodal", dict(A=782, B=572, S=24, L=37, E=32), 48), ("multimodal", dict(A=782, B=572, S=24, L=64, E=16), 48)]


@pytest.mark.parametrize("use_graph", [True, False])
def test_grouped_members_match_oracle(use_graph):
    """Lock-step population step (vla_train_step_group: one merged launch per step of the sequence) with members of different
    latent / embedding widths, injected eps and dropout masks: every member against the ORACLE trained alone
    (optimize_hyperparameters.py:78-113 ranges: latent 10..100, embed 16 / 32 / 64)."""
    from parity_util import TOL_BF16, assert_close, to_t
    from vla_b200 import DeviceDataset, Population
    n_steps, n_batches = 3, 2
    specs, refs, noise = [], [], []
    datasets = []
    for i, (kind, dims, batch) in enumerate(HETERO):
        state = vo.init_state(kind, dims, seed=60 + i)
        tpm, beta_v, site = vo.synthetic_batch(batch * n_batches, dims, seed=60 + i)
        eps, masks = vo.synthetic_noise(batch, dims, kind, seed=60 + i)
        cw = vo.balanced_class_weights(site, dims["S"])
        hyper = dict(lr=[1e-3, 3e-4, 5e-4, 2e-3][i], weight_decay=[1e-4, 1e-6, 1e-5, 0.0][i], beta_start=2e-3, gamma=[1.5, 0.7, 1.0, 2.0][i])
        refs.append((state, _oracle_train_h(kind, dims, state, dict(a=tpm, b=beta_v, site=site), n_steps, batch, eps, masks,
                                            hyper, cw)))
        datasets.append(DeviceDataset(tpm, beta_v, site, "cuda"))
        specs.append(dict(model=make_module(kind, dims, state, device="cpu"), class_weights=to_t(cw), **hyper))
        noise.append((eps, masks))
    pop = Population(specs, datasets, HETERO[0][2], use_graph=use_graph)
    assert pop.grouped
    for mem, (eps, masks) in zip(pop.members, noise):
        mem.trainer.injected = dict(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()])
    pop.begin_epoch(50)           # beta = beta_start
    got_losses = []
    for _ in range(n_steps):
        pop.step()
        got_losses.append(pop.losses())
    got_losses = np.array(got_losses)                       # [step, member, 4]
    for i, (state, (ref_state, ref_losses)) in enumerate(refs):
        np.testing.assert_allclose(got_losses[:, i, :], ref_losses, rtol=TOL_BF16)
        sd = {k: v.detach().cpu().numpy() for k, v in pop.members[i].model.state_dict().items()}
        lr = specs[i]["lr"]
        for name, ref in ref_state.items():
            got = sd[name]
            if name.endswith("num_batches_tracked"):
                assert int(got) == n_steps, name
            elif name.endswith(("running_mean", "running_var")):
                assert_close(name, got, ref, TOL_BF16, atol=4 * lr * n_steps * np.sqrt(ref.size))
            elif is_pre_bn_bias(name):
                assert np.abs(got - state[name]).max() <= 1.05 * lr * n_steps + 1e-7, name
            else:
                delta_ref = ref - state[name].astype(np.float64)
                delta = got.astype(np.float64) - state[name].astype(np.float64)
                assert rel_l2(delta, delta_ref) <= 0.2, (i, name, rel_l2(delta, delta_ref))
    pop.close()


def _oracle_train_h(kind, dims, state, data, n_steps, batch, eps, masks, hyper, cw):
    from parity_util import MATCHED_Q
    st = {k: (v.astype(np.float64) if v.dtype.kind == "f" else v.copy()) for k, v in state.items()}
    opt, step = vo.adamw_init(st)
    losses = []
    n_batches = len(data["site"]) // batch
    for i in range(n_steps):
        lo = (i % n_batches) * batch
        b = {k: (v[lo:lo + batch].astype(np.float64) if v.dtype.kind == "f" else v[lo:lo + batch]) for k, v in data.items()}
        scal, _, _, step = vo.train_step(kind, dims, st, opt, step, b, eps.astype(np.float64), masks, beta=hyper["beta_start"],
                                         gamma=hyper["gamma"], class_weights=cw.astype(np.float64), q=MATCHED_Q,
                                         lr=hyper["lr"], weight_decay=hyper["weight_decay"])
        losses.append([scal["total"], scal["recon"], scal["cls"], scal["kld"]])
    return st, np.array(losses)


def test_grouped_step_bit_identical_to_separate_steps():
    """vla_train_step_group issues the same tiles / blocks with the same arguments as n separate vla_train_step calls: with
    injected noise and a single k-split per weight-gradient tile the results agree bit for bit; with split-K red.adds
    (the normal case) up to summation order."""
    from vla_b200 import DeviceDataset, Population
    dims_list = [dict(A=782, B=572, S=24, L=20, E=32), dict(A=782, B=572, S=24, L=50, E=64)]
    batch = 256
    for kind in ("rna2dna", "dna2rna_ae"):
        states = [vo.init_state(kind, d, seed=70 + i) for i, d in enumerate(dims_list)]
        ds = DeviceDataset.synthetic(batch * 2, 782, 572, 24, "cuda", seed=9)
        out = {}
        for grouped in (True, False):
            pop = Population([dict(model=make_module(kind, d, st, device="cpu"), seed=i) for i, (d, st) in enumerate(zip(dims_list, states))],
                             ds, batch, grouped=grouped)
            pop.step(3)
            out[grouped] = (pop.losses(), [m.trainer.core.arena.clone() for m in pop.members])
            pop.close()
        np.testing.assert_allclose(out[True][0], out[False][0], rtol=1e-5)
        for a, b in zip(out[True][1], out[False][1]):
            assert float((a - b).abs().max()) <= 2.1 * 5e-4 * 3 + 1e-6
