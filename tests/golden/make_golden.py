"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's `src` package under the alias `ref_src` (never copied), loads a
deterministic state_dict produced by `oracle.vae_oracle.init_state`, injects a recorded epsilon
(by patching `torch.randn_like`, which `reparameterize` calls: src/models/vae.py:14) and recorded
dropout keep-masks (by swapping the instance's nn.Dropout children for a replay module), then runs
forward, the reference loss function, backward and two torch.optim.AdamW steps.  What it stores per
case: all outputs, the loss scalars, every gradient (large tensors: strided sample + sum + sum of
squares), the post-step parameters (same sampling) and the BatchNorm running statistics.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import vae_oracle as vo  # noqa: E402

REF = "/root/reference/src"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_src", os.path.join(REF, "__init__.py"),
                                                  submodule_search_locations=[REF])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_src"] = mod
    spec.loader.exec_module(mod)
    import ref_src.models as rm
    import ref_src.utils.losses as rl
    import ref_src.utils.directional_losses as rdl
    import ref_src.models.directional_ae as rae          # not re-exported by the reference's models/__init__.py
    import ref_src.utils.ae_losses as rael
    rm.RNA2DNAAE, rm.DNA2RNAAE = rae.RNA2DNAAE, rae.DNA2RNAAE
    rdl.rna2dna_ae_loss, rdl.dna2rna_ae_loss = rael.rna2dna_ae_loss, rael.dna2rna_ae_loss
    return rm, rl, rdl


class ReplayDropout(torch.nn.Module):
    def __init__(self, keep, p):
        super().__init__()
        self.keep, self.p = keep, p

    def forward(self, x):
        if not self.training:
            return x
        return x * self.keep / (1 - self.p)


SAMPLE_STRIDE = 17
FULL_LIMIT = 4096


def pack(prefix, arr, out):
    arr = np.asarray(arr)
    flat = arr.reshape(-1)
    if flat.size <= FULL_LIMIT:
        out[prefix + "|full"] = arr.astype(np.float32)
    else:
        out[prefix + "|sample"] = flat[::SAMPLE_STRIDE].astype(np.float32)
        out[prefix + "|sum"] = np.float64(flat.astype(np.float64).sum())
        out[prefix + "|sumsq"] = np.float64((flat.astype(np.float64) ** 2).sum())


CASES = [
    dict(name="multimodal_full", kind="multimodal", dims=dict(A=782, B=572, S=24, L=20, E=32), n=32, seed=1,
         beta=1e-3, gamma=1.0, weights=True, present=("a", "b", "site"), train=True, steps=2),
    dict(name="rna2dna_full", kind="rna2dna", dims=dict(A=782, B=572, S=24, L=20, E=32), n=32, seed=2,
         beta=1e-3, gamma=1.0, weights=False, present=("a", "site"), train=True, steps=2),
    dict(name="dna2rna_full", kind="dna2rna", dims=dict(A=782, B=572, S=24, L=20, E=32), n=32, seed=3,
         beta=5e-4, gamma=1.0, weights=False, present=("b", "site"), train=True, steps=2),
    dict(name="multimodal_small", kind="multimodal", dims=dict(A=50, B=36, S=5, L=10, E=16), n=7, seed=4,
         beta=2e-3, gamma=2.5, weights=True, present=("a", "b", "site"), train=True, steps=2),
    dict(name="multimodal_eval_a", kind="multimodal", dims=dict(A=50, B=36, S=5, L=10, E=16), n=9, seed=5,
         beta=1e-3, gamma=1.0, weights=False, present=("a",), train=False, steps=0),
    dict(name="multimodal_eval_b_site", kind="multimodal", dims=dict(A=50, B=36, S=5, L=10, E=16), n=9, seed=6,
         beta=1e-3, gamma=1.0, weights=False, present=("b", "site"), train=False, steps=0),
    dict(name="rna2dna_train_rna_only", kind="rna2dna", dims=dict(A=50, B=36, S=5, L=12, E=64), n=6, seed=7,
         beta=1e-3, gamma=1.0, weights=False, present=("a",), train=True, steps=1),
    dict(name="dna2rna_eval_dna_only", kind="dna2rna", dims=dict(A=50, B=36, S=5, L=12, E=64), n=6, seed=8,
         beta=1e-3, gamma=1.0, weights=False, present=("b",), train=False, steps=0),
    # directional autoencoders (src/models/directional_ae.py, src/utils/ae_losses.py)
    dict(name="rna2dna_ae_full", kind="rna2dna_ae", dims=dict(A=782, B=572, S=24, L=20, E=32), n=32, seed=9,
         beta=0.0, gamma=1.0, weights=False, present=("a", "site"), train=True, steps=2),
    dict(name="dna2rna_ae_full", kind="dna2rna_ae", dims=dict(A=782, B=572, S=24, L=20, E=32), n=32, seed=10,
         beta=0.0, gamma=1.0, weights=False, present=("b", "site"), train=True, steps=2),
    dict(name="rna2dna_ae_eval_rna_only", kind="rna2dna_ae", dims=dict(A=50, B=36, S=5, L=12, E=16), n=6, seed=11,
         beta=0.0, gamma=1.0, weights=False, present=("a",), train=False, steps=0),
    dict(name="dna2rna_ae_train_site_only", kind="dna2rna_ae", dims=dict(A=50, B=36, S=5, L=12, E=16), n=7, seed=12,
         beta=0.0, gamma=1.0, weights=False, present=("site",), train=True, steps=1),
]


def build_reference_model(rm, case):
    d = case["dims"]
    if case["kind"] == "multimodal":
        m = rm.MultiModalVAE(d["A"], d["B"], d["S"], d["L"], embed_dim=d["E"])
    elif case["kind"] == "rna2dna":
        m = rm.RNA2DNAVAE(d["A"], d["B"], d["S"], d["L"], embed_dim=d["E"])
    elif case["kind"] == "dna2rna":
        m = rm.DNA2RNAVAE(d["A"], d["B"], d["S"], d["L"], embed_dim=d["E"])
    elif case["kind"] == "rna2dna_ae":
        m = rm.RNA2DNAAE(d["A"], d["B"], d["S"], d["L"], embed_dim=d["E"])
    else:
        m = rm.DNA2RNAAE(d["A"], d["B"], d["S"], d["L"], embed_dim=d["E"])
    return m


def case_inputs(case):
    """Everything the oracle test needs to rebuild the same inputs (shared with the tests)."""
    d = case["dims"]
    state = vo.init_state(case["kind"], d, seed=case["seed"])
    if not case["train"]:
        # eval cases: give BN non-trivial running statistics
        for k in state:
            if k.endswith("running_mean"):
                state[k] = (0.2 * (vo.hash_uniform(state[k].size, case["seed"], 77) - 0.5)).astype(np.float32)
            if k.endswith("running_var"):
                state[k] = (0.5 + vo.hash_uniform(state[k].size, case["seed"], 78)).astype(np.float32)
    tpm, beta, site = vo.synthetic_batch(case["n"], d, seed=case["seed"])
    eps, masks = vo.synthetic_noise(case["n"], d, case["kind"], seed=case["seed"])
    cw = vo.balanced_class_weights(site, d["S"]) if case["weights"] else None
    return state, dict(a=tpm, b=beta, site=site), eps, masks, cw


def run_case(rm, rl, rdl, case):
    state, batch, eps, masks, cw = case_inputs(case)
    model = build_reference_model(rm, case)
    sd = {k: torch.from_numpy(np.array(v)) for k, v in state.items()}
    model.load_state_dict(sd, strict=True)
    for key, keep in masks.items():
        if vo.is_ae(case["kind"]):
            prefix, _, idx = key.rpartition(".")             # the AE encoders are bare nn.Sequentials
            seq = getattr(model, prefix)
        else:
            prefix, _, idx = key.rpartition(".fc.")
            seq = getattr(model, prefix).fc
        assert isinstance(seq[int(idx)], torch.nn.Dropout) and seq[int(idx)].p == vo.DROPOUT_P
        seq[int(idx)] = ReplayDropout(torch.from_numpy(keep.astype(np.float32)), vo.DROPOUT_P)
    model.train(case["train"])
    eps_t = torch.from_numpy(eps)
    orig_randn_like = torch.randn_like
    torch.randn_like = lambda t, *a, **k: eps_t.to(t.dtype)
    out = {}
    try:
        a = torch.from_numpy(batch["a"]) if "a" in case["present"] else None
        b = torch.from_numpy(batch["b"]) if "b" in case["present"] else None
        s = torch.from_numpy(batch["site"]) if "site" in case["present"] else None
        ta, tb, ts = (torch.from_numpy(batch[k]) for k in ("a", "b", "site"))
        cwt = torch.from_numpy(cw) if cw is not None else None
        opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=1e-5)
        nsteps = max(case["steps"], 1)
        for step in range(nsteps):
            ctx = torch.enable_grad() if case["train"] else torch.no_grad()
            with ctx:
                if case["kind"] == "multimodal":
                    ra, rb, rc, mu, lv = model(a=a, b=b, site=s)
                    loss, recon, cls, kld = rl.vae_loss(ra, ta, rb, tb, rc, ts, mu, lv, beta=case["beta"],
                                                        gamma=case["gamma"], class_weights=cwt)
                    recons = {"decoder_a": ra, "decoder_b": rb, "decoder_c": rc}
                elif case["kind"] == "rna2dna":
                    rb, mu, lv = model(rna=a, site=s)
                    loss, recon, kld = rdl.rna2dna_loss(rb, tb, mu, lv, beta=case["beta"])
                    cls = 0.0
                    recons = {"decoder_dna": rb}
                elif case["kind"] == "dna2rna":
                    ra, mu, lv = model(dna=b, site=s)
                    loss, recon, kld = rdl.dna2rna_loss(ra, ta, mu, lv, beta=case["beta"])
                    cls = 0.0
                    recons = {"decoder_rna": ra}
                elif case["kind"] == "rna2dna_ae":
                    rb, mu = model(rna=a, site=s)                 # (recon, latent)
                    loss, recon = rdl.rna2dna_ae_loss(rb, tb)
                    cls, kld, lv = 0.0, 0.0, None
                    recons = {"decoder_dna": rb}
                else:
                    ra, mu = model(dna=b, site=s)
                    loss, recon = rdl.dna2rna_ae_loss(ra, ta)
                    cls, kld, lv = 0.0, 0.0, None
                    recons = {"decoder_rna": ra}
            if step == 0:
                for k, v in recons.items():
                    pack(f"out.recon.{k}", v.detach().numpy(), out)
                pack("out.mu", mu.detach().numpy(), out)
                if lv is not None:
                    pack("out.logvar", lv.detach().numpy(), out)
                out["loss"] = np.array([loss.item(), recon, cls, kld], dtype=np.float64)
            if case["steps"] > 0:
                opt.zero_grad()
                loss.backward()
                if step == 0:
                    for k, p in model.named_parameters():
                        if p.grad is not None:
                            pack(f"grad.{k}", p.grad.numpy(), out)
                        else:
                            out[f"grad.{k}|none"] = np.zeros(0, dtype=np.float32)
                opt.step()
        if case["steps"] > 0:
            out["loss_last"] = np.array([loss.item(), recon, cls, kld], dtype=np.float64)
            for k, v in model.state_dict().items():
                pack(f"final.{k}", v.numpy().astype(np.float64), out)
    finally:
        torch.randn_like = orig_randn_like
    return out


def main():
    torch.set_num_threads(1)
    rm, rl, rdl = load_reference()
    only = sys.argv[1:]                                   # optional: fixture names to (re)generate
    for case in CASES:
        if only and case["name"] not in only:
            continue
        out = run_case(rm, rl, rdl, case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(case["name"], len(out), "arrays", os.path.getsize(path) // 1024, "KiB", "loss", out["loss"])


if __name__ == "__main__":
    main()
