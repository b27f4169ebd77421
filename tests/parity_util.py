"""Shared helpers of the GPU parity tests: build the CUDA module from an oracle state, run the oracle,
compare with the tolerances BASELINE.json states."""
import os
import re

import numpy as np
import torch

from oracle import vae_oracle as vo

# BASELINE.json north_star: fp32 kernels 1e-5 relative; bf16 tensor-core kernels 2e-2 relative on losses and
# gradients; argmax site predictions exact.  "relative" is measured as ||x - ref||_2 / ||ref||_2 per tensor.
TOL_FP32 = 1e-5
TOL_BF16 = 2e-2
# Parameter gradients against the EXACT (fp64) arithmetic and against the reference fixtures: the same 2e-2.
# Round 1 needed 0.25 here: with every GEMM operand rounded to bf16 once, pre-activations within 2^-9 (relative) of zero
# flip their ReLU decision and each flip switches a whole (sample, unit) gradient contribution on or off (7-9 % at
# batch 4096, measured in numpy with no GPU involved).  The CUDA path now carries the operands of every forward GEMM
# upstream of a ReLU as hi + lo bf16 pairs (three MMA passes, DESIGN.md "Precision"), which removes the flips:
# measured <= 1.8e-2 (typically 3e-3 .. 9e-3) for all five model kinds at batch 8 .. 4096.
TOL_GRAD_VS_EXACT = 2e-2
MIN_COSINE_VS_EXACT = 0.9995
# Same algorithm at the CUDA path's declared operand precision (oracle/vae_oracle.py header); VLA_SPLIT=0 selects the
# single-rounding variant in the library and here.
SPLIT = os.environ.get("VLA_SPLIT", "1") != "0"
MATCHED_Q = vo.SPLIT_BF16 if SPLIT else vo.round_bf16
if not SPLIT:                      # documented envelope of the single-rounding variant (round 1)
    TOL_GRAD_VS_EXACT, MIN_COSINE_VS_EXACT = 0.25, 0.97


def rel_l2(x, ref):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-30))


def is_pre_bn_bias(name):
    m = re.fullmatch(r"encoder_\w+\.fc\.(\d+)\.bias", name)
    if m:
        return int(m.group(1)) % 4 == 0
    m = re.fullmatch(r"encoder_(rna|dna)\.(\d+)\.bias", name)      # autoencoders: the last Linear is the head (no BatchNorm behind it)
    if m:
        return int(m.group(2)) % 4 == 0 and int(m.group(2)) < 4 * (1 if m.group(1) == "rna" else 2)
    return False


def module_class(kind):
    from src.models import DNA2RNAVAE, MultiModalVAE, RNA2DNAVAE
    from src.models.directional_ae import DNA2RNAAE, RNA2DNAAE
    return {"multimodal": MultiModalVAE, "rna2dna": RNA2DNAVAE, "dna2rna": DNA2RNAVAE,
            "rna2dna_ae": RNA2DNAAE, "dna2rna_ae": DNA2RNAAE}[kind]


def make_module(kind, dims, state, device="cuda"):
    cls = module_class(kind)
    m = cls(dims["A"], dims["B"], dims["S"], dims["L"], embed_dim=dims.get("E", 32))
    sd = {k: torch.from_numpy(np.array(v)) for k, v in state.items()}
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def call_module(m, kind, a=None, b=None, site=None):
    """Uniform call -> dict(recon={decoder prefix: tensor}, mu, logvar)."""
    if kind == "multimodal":
        ra, rb, rc, mu, lv = m(a=a, b=b, site=site)
        recon = {"decoder_a": ra, "decoder_b": rb, "decoder_c": rc}
    elif kind == "rna2dna":
        rb, mu, lv = m(rna=a, site=site)
        recon = {"decoder_dna": rb}
    elif kind == "dna2rna":
        ra, mu, lv = m(dna=b, site=site)
        recon = {"decoder_rna": ra}
    elif kind == "rna2dna_ae":
        rb, mu = m(rna=a, site=site)                 # (reconstruction, latent)
        recon, lv = {"decoder_dna": rb}, None
    else:
        ra, mu = m(dna=b, site=site)
        recon, lv = {"decoder_rna": ra}, None
    return dict(recon=recon, mu=mu, logvar=lv)


def loss_for(kind, out, batch_t, beta, gamma, cw):
    from src.utils.directional_losses import dna2rna_loss, rna2dna_loss
    from src.utils.losses import vae_loss
    if kind == "multimodal":
        total, recon, cls, kld = vae_loss(out["recon"]["decoder_a"], batch_t["a"], out["recon"]["decoder_b"], batch_t["b"],
                                          out["recon"]["decoder_c"], batch_t["site"], out["mu"], out["logvar"],
                                          beta=beta, gamma=gamma, class_weights=cw)
    elif kind == "rna2dna":
        total, recon, kld = rna2dna_loss(out["recon"]["decoder_dna"], batch_t["b"], out["mu"], out["logvar"], beta=beta)
        cls = 0.0
    elif kind == "dna2rna":
        total, recon, kld = dna2rna_loss(out["recon"]["decoder_rna"], batch_t["a"], out["mu"], out["logvar"], beta=beta)
        cls = 0.0
    elif kind == "rna2dna_ae":
        from src.utils.ae_losses import rna2dna_ae_loss
        total, recon = rna2dna_ae_loss(out["recon"]["decoder_dna"], batch_t["b"])
        cls, kld = 0.0, 0.0
    else:
        from src.utils.ae_losses import dna2rna_ae_loss
        total, recon = dna2rna_ae_loss(out["recon"]["decoder_rna"], batch_t["a"])
        cls, kld = 0.0, 0.0
    return total, (recon, cls, kld)


def to_t(x, device="cuda"):
    return None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(device)


def oracle_step(kind, dims, state, batch, present, eps, masks, beta, gamma, cw, train=True, dtype=np.float64, q=None):
    """Oracle forward + loss + backward on a private fp64 copy of `state`.
    q=None: exact reference arithmetic; q=MATCHED_Q: same algorithm at the CUDA path's declared GEMM-operand
    precision (oracle/vae_oracle.py header)."""
    st = {k: (v.astype(dtype) if v.dtype.kind == "f" else v.copy()) for k, v in state.items()}
    bt = {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in batch.items()}
    inputs = {k: (bt[k] if k in present else None) for k in ("a", "b", "site")}
    out, cache = vo.forward(kind, dims, st, inputs, eps.astype(dtype), masks, train=train, q=q)
    scalars, og = vo.loss_and_output_grads(kind, out, bt, beta, gamma, None if cw is None else cw.astype(dtype))
    grads = vo.backward(kind, dims, st, cache, og, train=train)
    return out, scalars, grads, st


def assert_close(name, got, ref, tol, atol=0.0):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    err = np.linalg.norm(got - ref)
    bound = tol * np.linalg.norm(ref) + atol
    assert np.isfinite(got).all(), name
    assert err <= bound, f"{name}: |err|={err:.4g} > {bound:.4g} (rel {err / max(np.linalg.norm(ref), 1e-30):.3g})"


def cosine(x, ref):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    return float(x @ ref / max(np.linalg.norm(x) * np.linalg.norm(ref), 1e-300))


def grads_by_name(core, flat):
    """{state_dict name: numpy gradient} from a flat gradient arena laid out like the parameter arena."""
    from vla_b200 import _lib
    out = {}
    host = flat.detach().cpu().numpy()
    for name, kind, off, shape in core.infos:
        if kind == _lib.TENSOR_PARAM:
            n = int(np.prod(shape)) if shape else 1
            out[name] = host[off:off + n].reshape(shape)
    return out


def out_dir():
    """Where GPU tests leave their measured tables (merged back by gpurun; the summaries are copied to profiles/)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    return d
