"""Parity at the sizes BASELINE.json quotes: fused train steps at batch 4096 (configs[1], [2]) for every model kind and the
cross-modal inference forward at batch 262 144 (configs[3]), against the fp64 oracle -- exact arithmetic and the same
algorithm at the CUDA path's declared operand precision.  The per-tensor error tables are written to
gpurun_out/parity_*.json (summaries committed under profiles/)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import vae_oracle as vo
from parity_util import (MATCHED_Q, MIN_COSINE_VS_EXACT, TOL_BF16, TOL_GRAD_VS_EXACT, cosine, grads_by_name, is_pre_bn_bias,
                         make_module, oracle_step, out_dir, rel_l2, to_t)

pytestmark = pytest.mark.gpu
FULL = dict(A=782, B=572, S=24, L=20, E=32)
ENC_INPUTS = {"multimodal": ("a", "b", "site"), "rna2dna": ("a", "site"), "dna2rna": ("b", "site"),
              "rna2dna_ae": ("a", "site"), "dna2rna_ae": ("b", "site")}


def _dump(name, table):
    path = os.path.join(out_dir(), name)
    old = {}
    if os.path.exists(path):
        try:
            old = json.load(open(path))
        except Exception:
            old = {}
    old.update(table)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)


@pytest.mark.parametrize("kind", ["rna2dna", "dna2rna", "multimodal", "rna2dna_ae", "dna2rna_ae"])
def test_train_step_b4096_vs_oracle(kind):
    """One fused train step at batch 4096 with injected epsilon / keep-masks: the four loss scalars, EVERY parameter gradient
    and the AdamW-updated parameters against the oracle (reference loop body train_rna2dna.py:82-99)."""
    from vla_b200 import DeviceDataset, Trainer
    B, dims = 4096, FULL
    state = vo.init_state(kind, dims, seed=41)
    tpm, beta_v, site = vo.synthetic_batch(B, dims, seed=41)
    batch = dict(a=tpm, b=beta_v, site=site)
    eps, masks = vo.synthetic_noise(B, dims, kind, seed=41)
    cw = vo.balanced_class_weights(site, dims["S"]) if kind == "multimodal" else None
    beta, gamma = 1e-3, 1.0
    o_out, o_scal, o_grads, _ = oracle_step(kind, dims, state, batch, ENC_INPUTS[kind], eps, masks, beta, gamma, cw)
    q_out, q_scal, q_grads, q_state = oracle_step(kind, dims, state, batch, ENC_INPUTS[kind], eps, masks, beta, gamma, cw, q=MATCHED_Q)

    m = make_module(kind, dims, state).train()
    ds = DeviceDataset(tpm, beta_v, site, "cuda")
    tr = Trainer(m, ds, B, lr=5e-4, weight_decay=1e-5, beta_kl=beta, gamma=gamma, class_weights=to_t(cw), use_graph=False)
    tr.injected = dict(eps=to_t(eps), keep_masks=[to_t(v) for v in masks.values()])
    tr.forward_backward()                           # phases = 1: gradients stay in the arena
    torch.cuda.synchronize()
    losses = np.array(tr.losses(), dtype=np.float64)
    grads = grads_by_name(tr.core, tr.grads)
    table = dict(losses_rel_exact={}, losses_rel_matched={}, grads={})
    for i, key in enumerate(("total", "recon", "cls", "kld")):
        if abs(o_scal[key]) > 0:
            table["losses_rel_exact"][key] = abs(losses[i] - o_scal[key]) / abs(o_scal[key])
            table["losses_rel_matched"][key] = abs(losses[i] - q_scal[key]) / abs(q_scal[key])
    failures = []
    for name, exact in o_grads.items():
        g = grads[name]
        if is_pre_bn_bias(name):
            # exactly-zero true gradient (BatchNorm removes the mean): bounded by the layer's weight-gradient scale
            wname = name[:-4] + "weight"
            assert np.linalg.norm(g) <= 2e-2 * np.linalg.norm(o_grads[wname]) + 1e-4, name
            continue
        row = dict(rel_exact=rel_l2(g, exact), rel_matched=rel_l2(g, q_grads[name]), cosine_exact=cosine(g, exact),
                   norm=float(np.linalg.norm(exact)))
        table["grads"][name] = row
        if np.linalg.norm(exact) > 1e-6:
            if row["rel_exact"] > TOL_GRAD_VS_EXACT or row["cosine_exact"] < MIN_COSINE_VS_EXACT:
                failures.append(("exact", name, row))
            if row["rel_matched"] > TOL_BF16:
                failures.append(("matched", name, row))
    # ---- finish the step: AdamW on these gradients; parameters against the oracle's AdamW ----
    tr.apply_optimizer()
    torch.cuda.synchronize()
    opt, step = vo.adamw_init(q_state)
    vo.adamw_step(q_state, q_grads, opt, step, lr=5e-4, weight_decay=1e-5)
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    table["params_rel_matched"] = {}
    lr = 5e-4
    for name, ref in q_state.items():
        if vo.is_buffer(name) or is_pre_bn_bias(name):
            continue
        got = sd[name].astype(np.float64)
        table["params_rel_matched"][name] = rel_l2(got, ref)
        assert np.abs(got - ref).max() <= 2.05 * lr, name          # first Adam step: every weight moves by <= lr
        d_ref, d_got = ref - state[name].astype(np.float64), got - state[name].astype(np.float64)
        assert rel_l2(d_got, d_ref) <= 0.2, (name, rel_l2(d_got, d_ref))
    worst = max(table["grads"].items(), key=lambda kv: kv[1]["rel_exact"])
    table["worst_grad_rel_exact"] = [worst[0], worst[1]["rel_exact"]]
    _dump("parity_train_b4096.json", {kind: table})
    for key, v in table["losses_rel_exact"].items():
        assert v <= TOL_BF16, (key, v)
    assert not failures, failures[:5]


def _eval_inputs(n, seed):
    from vla_b200 import DeviceDataset
    ds = DeviceDataset.synthetic(n, FULL["A"], FULL["B"], FULL["S"], "cuda", seed=seed)
    g = torch.Generator(device="cuda")
    g.manual_seed(seed + 1)
    eps = torch.randn(n, FULL["L"], device="cuda", generator=g)
    return ds, eps


@pytest.mark.parametrize("modality", ["a", "b"])
def test_eval_forward_b262144_vs_oracle(modality):
    """Cross-modal inference model(a=x) / model(b=x) (reference downstream_task.py:29-49) at batch 262 144, eval mode, injected
    epsilon.  Eval-mode rows are independent (BatchNorm uses running statistics), so the oracle is run on a row subset for
    all outputs -- and, for model(a=x), on ALL rows for the site logits: argmax mismatches are counted over every row."""
    N, dims, kind = 262144, FULL, "multimodal"
    state = vo.init_state(kind, dims, seed=7)
    rng = np.random.default_rng(7)
    for k in state:                                   # non-trivial running statistics
        if k.endswith("running_mean"):
            state[k] = rng.normal(0, 0.2, state[k].shape).astype(np.float32)
        if k.endswith("running_var"):
            state[k] = rng.uniform(0.5, 1.5, state[k].shape).astype(np.float32)
    ds, eps = _eval_inputs(N, 7)
    m = make_module(kind, dims, state).eval()
    with torch.no_grad(), m.inject(eps=eps):
        out = m(a=ds.tpm) if modality == "a" else m(b=ds.beta)
    torch.cuda.synchronize()
    names = ("decoder_a", "decoder_b", "decoder_c", "mu", "logvar")
    # ---- all outputs on every 8th row ----
    idx = np.arange(0, N, 8)
    sub = dict(a=ds.tpm[::8].cpu().numpy(), b=ds.beta[::8].cpu().numpy(), site=ds.site[::8].cpu().numpy())
    eps_sub = eps[::8].cpu().numpy()
    st64 = {k: (v.astype(np.float64) if v.dtype.kind == "f" else v.copy()) for k, v in state.items()}
    inputs = {modality: sub[modality].astype(np.float64)}
    table = {}
    for label, q in (("exact", None), ("matched", MATCHED_Q)):
        o, _ = vo.forward(kind, dims, st64, inputs, eps_sub.astype(np.float64), None, train=False, q=q)
        ref = dict(o["recon"], mu=o["mu"], logvar=o["logvar"])
        for i, nm in enumerate(names):
            got = out[i][::8].cpu().numpy()
            table[f"{nm}.rel_{label}"] = rel_l2(got, ref[nm])
            assert table[f"{nm}.rel_{label}"] <= TOL_BF16, (nm, label, table[f"{nm}.rel_{label}"])
        if label == "exact":
            sub_mismatch = int((out[2][::8].cpu().numpy().argmax(1) != ref["decoder_c"].argmax(1)).sum())
            table["argmax_mismatch_subset"] = [sub_mismatch, len(idx)]
    # ---- site logits on ALL rows (encoder + decoder_c only, registered as an oracle-only model kind) ----
    if modality == "a":
        vo.MODEL_KINDS["_enc_a_dec_c"] = dict(encoders=[("encoder_a", "A")], decoders=[("decoder_c", "C")])
        try:
            got_c = out[2].cpu().numpy()
            mism, n_close, max_err = 0, 0, 0.0
            chunk = 32768
            for lo in range(0, N, chunk):
                xa = ds.tpm[lo:lo + chunk].cpu().numpy().astype(np.float64)
                o, _ = vo.forward("_enc_a_dec_c", dims, st64, dict(a=xa), eps[lo:lo + chunk].cpu().numpy().astype(np.float64), None,
                                  train=False)
                ref_c = o["recon"]["decoder_c"]
                g = got_c[lo:lo + chunk]
                err = np.abs(g - ref_c).max(1)
                max_err = max(max_err, float(err.max()))
                bad = g.argmax(1) != ref_c.argmax(1)
                top2 = np.sort(ref_c, axis=1)[:, -2:]
                margin = top2[:, 1] - top2[:, 0]
                # a mismatch is only possible where the oracle's own top-2 margin is below the logit error of that row
                assert (margin[bad] <= 2 * err[bad]).all()
                mism += int(bad.sum())
                n_close += int((margin <= 2 * err).sum())
        finally:
            del vo.MODEL_KINDS["_enc_a_dec_c"]
        table["argmax_mismatch_all_rows"] = [mism, N]
        table["rows_with_margin_below_logit_error"] = n_close
        table["max_abs_logit_error"] = max_err
        # site predictions must agree; the only rows allowed to differ are exact near-ties of the reference itself
        assert mism <= max(2, N // 50000), table
    _dump("parity_eval_b262144.json", {f"model({modality}=x)": table})
