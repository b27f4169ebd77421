"""CPU oracle for the MultiModalVAE / directional-VAE train and inference step.

TEST INFRASTRUCTURE ONLY.  This file is a plain-numpy restatement of the
reference's algorithm for the hot path.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product path (`vae-los-angeles_b200/`) never does and fails loudly when its CUDA
library is missing.

Parity status: PINNED.  The reference ships no golden vectors or tests of its own
(SURVEY.md section 4), so the pin is made here: `tests/golden/make_golden.py` imports the
unmodified reference modules from /root/reference, replays an injected epsilon and
injected dropout masks through them, and commits their outputs, losses, gradients
and AdamW-updated parameters as fixtures under `tests/golden/`;
`tests/test_oracle_golden.py` checks every function below against those fixtures.

The arithmetic of the reference lives in PyTorch (ATen), pinned by the reference
only as `torch>=2.0.0` (requirements.txt:2; 2.11.0 in this image).  The ATen
semantics restated here (and verified against it by the fixtures):
  * nn.Linear                      y = x W^T + b
  * nn.BatchNorm1d (train)         batch mean, *biased* variance for normalising,
                                   eps 1e-5; running stats use momentum 0.1 and the
                                   *unbiased* variance; num_batches_tracked += 1
  * nn.BatchNorm1d (eval)          running mean / running var
  * nn.Dropout(p) (train)          x * keep / (1 - p)
  * F.binary_cross_entropy         log terms clamped at -100; backward divides by
                                   max(y (1 - y), 1e-12)
  * F.cross_entropy(weight, sum)   sum_i w[t_i] * (-log_softmax(x_i)[t_i])
  * torch.optim.AdamW              decoupled decay, bias correction, eps outside sqrt

Operand precision: every function takes an optional `q` (default: identity = exact reference arithmetic).
The CUDA path stages the operands of its tensor-core GEMMs in bf16 (fp32 accumulation, fp32 everything
else).  `q` applies that one declared difference at exactly those staging points (GEMM input activations,
weights as GEMM operands, gradient signals as GEMM operands), which is the standard way to check a
mixed-precision kernel ("same algorithm, same operand precision"):
  * `q=round_bf16`  every operand rounded to bf16 once (the CUDA path with VLA_SPLIT=0);
  * `q=SPLIT_BF16`  the CUDA path's default: the operands of every FORWARD GEMM whose result reaches a ReLU
                    (all layers except the output layers of decoders A and B) are carried as hi + lo bf16 pairs
                    (`split_bf16`, ~16 mantissa bits, three MMA passes); the output layers and every backward
                    GEMM (data and weight gradients) use the single bf16 rounding.
DESIGN.md "Precision" has the measurement behind this split: a ReLU whose pre-activation moves by 2^-9
relative flips, and flipped units dominate the gradient error; with split operands upstream of every ReLU
the gradients stay within 2e-2 of the exact arithmetic.

Reference call sites followed (paths relative to /root/reference):
  EncoderA/B/C.forward      src/models/encoders.py:8-23, 26-46, 49-61
  DecoderA/B/C.forward      src/models/decoders.py:8-19, 22-36, 39-50
  reparameterize            src/models/vae.py:11-15
  MultiModalVAE.forward     src/models/vae.py:37-79
  RNA2DNAVAE / DNA2RNAVAE   src/models/directional_vae.py:12-60, 63-111
  vae_loss                  src/utils/losses.py:8-46
  rna2dna_loss/dna2rna_loss src/utils/directional_losses.py:8-30, 33-55
  optimizer step            train_rna2dna.py:94-96, 185-189 (torch.optim.AdamW)
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
DROPOUT_P = 0.1

# (module prefix, stack type) per model kind; order = order in which the reference
# appends to mu_list (vae.py:51-62, directional_vae.py:40-47, 91-98).
MODEL_KINDS = {
    "multimodal": dict(
        encoders=[("encoder_a", "A"), ("encoder_b", "B"), ("encoder_c", "C")],
        decoders=[("decoder_a", "A"), ("decoder_b", "B"), ("decoder_c", "C")],
    ),
    "rna2dna": dict(
        encoders=[("encoder_rna", "A"), ("encoder_site", "C")],
        decoders=[("decoder_dna", "B")],
    ),
    "dna2rna": dict(
        encoders=[("encoder_dna", "B"), ("encoder_site", "C")],
        decoders=[("decoder_rna", "A")],
    ),
    # Directional autoencoders (src/models/directional_ae.py:10-134): the same stacks with ONE head of width L per
    # encoder (the last Linear of the nn.Sequential / `site_projection`), latent = mean over the present encoders,
    # no reparameterisation, no KL term.
    "rna2dna_ae": dict(
        encoders=[("encoder_rna", "A"), ("site", "C")],
        decoders=[("decoder_dna", "B")],
        ae=True,
    ),
    "dna2rna_ae": dict(
        encoders=[("encoder_dna", "B"), ("site", "C")],
        decoders=[("decoder_rna", "A")],
        ae=True,
    ),
}

ENC_HIDDEN = {"A": [128], "B": [512, 256]}          # encoders.py:12-17, 30-39
DEC_HIDDEN = {"A": [128], "B": [256, 512], "C": [64]}  # decoders.py:12-16, 26-33, 43-47
INPUT_OF = {"A": "a", "B": "b", "C": "site"}        # which input feeds which stack type


def is_ae(kind):
    return bool(MODEL_KINDS[kind].get("ae"))


def enc_names(kind, prefix, t):
    """state_dict key stems of one encoder stack: emb (embedding weight key), fc / bn / drop (one per hidden layer; drop
    is the key of the Dropout's keep-mask in `masks`), heads ([mu, logvar] for the VAEs, [latent] for the AEs).
    VAE names: src/models/encoders.py:12-19, 30-41, 53-55; AE names: src/models/directional_ae.py:21-32, 80-95."""
    depth = 0 if t == "C" else len(ENC_HIDDEN[t])
    if not is_ae(kind):
        return dict(emb=f"{prefix}.embedding.weight", fc=[f"{prefix}.fc.{4 * i}" for i in range(depth)],
                    bn=[f"{prefix}.fc.{4 * i + 1}" for i in range(depth)], drop=[f"{prefix}.fc.{4 * i + 3}" for i in range(depth)],
                    heads=[f"{prefix}.fc_mu", f"{prefix}.fc_logvar"])
    if t == "C":
        return dict(emb="site_embedding.weight", fc=[], bn=[], drop=[], heads=["site_projection"])
    return dict(emb=None, fc=[f"{prefix}.{4 * i}" for i in range(depth)], bn=[f"{prefix}.{4 * i + 1}" for i in range(depth)],
                drop=[f"{prefix}.{4 * i + 3}" for i in range(depth)], heads=[f"{prefix}.{4 * depth}"])


def bn_affine_keys(kind):
    out = set()
    for prefix, t in MODEL_KINDS[kind]["encoders"]:
        for stem in enc_names(kind, prefix, t)["bn"]:
            out.add(stem + ".weight")
            out.add(stem + ".bias")
    return out


def feature_dim(stack_type, dims):
    return {"A": dims["A"], "B": dims["B"], "C": dims["S"]}[stack_type]


def param_shapes(kind, dims):
    """Ordered {state_dict key: shape}; trainable parameters and BN buffers.

    Order and names follow the reference's state_dict."""
    L, E = dims["L"], dims.get("E", 32)
    out = {}
    spec = MODEL_KINDS[kind]
    for prefix, t in spec["encoders"]:
        nm = enc_names(kind, prefix, t)
        if t == "C":
            out[nm["emb"]] = (dims["S"], E)
            last = E
        else:
            last = feature_dim(t, dims)
            for i, h in enumerate(ENC_HIDDEN[t]):
                out[nm["fc"][i] + ".weight"] = (h, last)
                out[nm["fc"][i] + ".bias"] = (h,)
                out[nm["bn"][i] + ".weight"] = (h,)
                out[nm["bn"][i] + ".bias"] = (h,)
                out[nm["bn"][i] + ".running_mean"] = (h,)
                out[nm["bn"][i] + ".running_var"] = (h,)
                out[nm["bn"][i] + ".num_batches_tracked"] = ()
                last = h
        for head in nm["heads"]:
            out[head + ".weight"] = (L, last)
            out[head + ".bias"] = (L,)
    for prefix, t in spec["decoders"]:
        last = L
        widths = DEC_HIDDEN[t] + [feature_dim(t, dims)]
        for i, h in enumerate(widths):
            out[f"{prefix}.fc.{2 * i}.weight"] = (h, last)
            out[f"{prefix}.fc.{2 * i}.bias"] = (h,)
            last = h
    return out


def is_buffer(key):
    return key.endswith(("running_mean", "running_var", "num_batches_tracked"))


# --------------------------------------------------------------------------------------
# Portable deterministic initialisation (no dependence on any library RNG stream)
# --------------------------------------------------------------------------------------
def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return z ^ (z >> np.uint64(31))


def hash_uniform(n, seed, stream):
    """n doubles in [0,1) from a counter hash; identical on every platform."""
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64)
        key = _splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        bits = _splitmix64(ctr ^ key)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def hash_normal(n, seed, stream):
    u1 = hash_uniform(n, seed, 2 * stream + 1000003)
    u2 = hash_uniform(n, seed, 2 * stream + 1000004)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def init_state(kind, dims, seed=0, dtype=np.float32):
    """PyTorch-style default init (Linear U(+-1/sqrt(fan_in)), Embedding N(0,1), BN 1/0/0/1),
    drawn from the portable hash stream.  BN affine parameters are perturbed slightly so that
    parity tests exercise gamma/beta gradients with non-trivial values."""
    state = {}
    for i, (key, shape) in enumerate(param_shapes(kind, dims).items()):
        n = int(np.prod(shape)) if shape else 1
        if key.endswith("num_batches_tracked"):
            state[key] = np.zeros((), dtype=np.int64)
        elif key.endswith("running_mean"):
            state[key] = np.zeros(shape, dtype=dtype)
        elif key.endswith("running_var"):
            state[key] = np.ones(shape, dtype=dtype)
        elif "embedding" in key:
            state[key] = hash_normal(n, seed, i).reshape(shape).astype(dtype)
        elif key in bn_affine_keys(kind):
            # BatchNorm affine: weight ~ 1 + 0.1 u, bias ~ 0.1 u
            u = hash_uniform(n, seed, i) * 2.0 - 1.0
            base = 1.0 if key.endswith("weight") else 0.0
            state[key] = (base + 0.1 * u).astype(dtype)
        else:
            # Linear weight [out, in] or bias [out]; the bias bound uses the layer's fan_in
            if len(shape) == 2:
                fan_in = shape[1]
            else:
                wkey = key[: -len("bias")] + "weight"
                fan_in = param_shapes(kind, dims)[wkey][1]
            bound = 1.0 / np.sqrt(fan_in)
            u = hash_uniform(n, seed, i) * 2.0 - 1.0
            state[key] = (bound * u).reshape(shape).astype(dtype)
    return state


def synthetic_batch(n, dims, seed=0, dtype=np.float32):
    """Synthetic inputs of the reference's schema (scripts/prepare_data.py:112-125):
    tpm = log1p(Gamma(1, 20)); beta in (0,1), bimodal; site uniform over S classes."""
    u = hash_uniform(n * dims["A"], seed, 11)
    tpm = np.log1p(-20.0 * np.log(1.0 - u)).reshape(n, dims["A"])       # Exp(20) == Gamma(1,20)
    v = hash_uniform(n * dims["B"], seed, 12)
    beta = np.clip(np.sin(0.5 * np.pi * v) ** 2, 1e-4, 1 - 1e-4).reshape(n, dims["B"])  # arcsine == Beta(.5,.5)
    site = (hash_uniform(n, seed, 13) * dims["S"]).astype(np.int64)
    return tpm.astype(dtype), beta.astype(dtype), site


def synthetic_noise(n, dims, kind, seed=0, dtype=np.float32):
    """Injected epsilon [n, L] and dropout keep-masks (uint8) for every Dropout in `kind`."""
    eps = hash_normal(n * dims["L"], seed, 21).reshape(n, dims["L"]).astype(dtype)
    masks = {}
    j = 0
    for prefix, t in MODEL_KINDS[kind]["encoders"]:
        if t == "C":
            continue
        nm = enc_names(kind, prefix, t)
        for i, h in enumerate(ENC_HIDDEN[t]):
            keep = hash_uniform(n * h, seed, 31 + j) >= DROPOUT_P
            masks[nm["drop"][i]] = keep.reshape(n, h).astype(np.uint8)
            j += 1
    return eps, masks


def balanced_class_weights(site, n_sites, dtype=np.float32):
    """sklearn 'balanced' weights n / (S * count_c) (optimize_hyperparameters.py:33-44)."""
    counts = np.bincount(site, minlength=n_sites).astype(np.float64)
    w = len(site) / (n_sites * np.maximum(counts, 1.0))
    return w.astype(dtype)


# --------------------------------------------------------------------------------------
# Operand rounding (declared precision of the CUDA path's GEMM operands)
# --------------------------------------------------------------------------------------
def round_bf16(x):
    """Round-to-nearest-even to bfloat16, returned in the input's dtype."""
    x = np.asarray(x)
    f = np.ascontiguousarray(x, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) >> np.uint64(16)) << np.uint64(16)
    return r.astype(np.uint32).view(np.float32).reshape(x.shape).astype(x.dtype)


def split_bf16(x):
    """hi + lo bf16 pair of x (lo = bf16 of what the first rounding dropped), returned as their sum."""
    h = round_bf16(x)
    return h + round_bf16(np.asarray(x) - h)


class OperandPrecision:
    """q(x): rounding of a GEMM operand; q.fwd(x): rounding of the operands of the forward GEMMs that feed a ReLU."""

    def __init__(self, lin, fwd=None):
        self.lin, self.fwd = lin, (fwd or lin)

    def __call__(self, x):
        return self.lin(x)


SPLIT_BF16 = OperandPrecision(round_bf16, split_bf16)


def _ident(x):
    return x


def _fwd(q):
    return getattr(q, "fwd", q)


# --------------------------------------------------------------------------------------
# Forward
# --------------------------------------------------------------------------------------
def _linear(x, w, b, q=_ident):
    return x @ q(w).T + b


def forward(kind, dims, state, inputs, eps, masks=None, train=True, update_running=True, q=None):
    """Forward pass of `kind` on `inputs` = {'a':..., 'b':..., 'site':...} (missing/None = absent).

    Returns (outputs, cache).  outputs = {'recon': {decoder prefix: array}, 'mu', 'logvar', 'z'}; for the autoencoder
    kinds 'mu' is the latent (directional_ae.py:46-56) and 'logvar' is None.
    In train mode BN running statistics in `state` are updated in place (as nn.BatchNorm1d does)
    unless update_running=False.  `eps` is the injected N(0,1) draw; reparameterize always samples,
    also in eval mode (vae.py:11-15)."""
    spec = MODEL_KINDS[kind]
    dt = eps.dtype
    q = q or _ident
    qf = _fwd(q)                  # operands of forward GEMMs upstream of a ReLU; weight-gradient GEMMs re-read them through q
    cache = {"enc": {}, "dec": {}, "present": [], "q": q}
    mus, lvs = [], []
    for prefix, t in spec["encoders"]:
        x = inputs.get(INPUT_OF[t])
        if x is None:
            continue
        c = {}
        nm = enc_names(kind, prefix, t)
        if t == "C":
            h_exact = state[nm["emb"]][x]                         # encoders.py:58
            h = qf(h_exact)
            c["site"] = x
        else:
            h_exact = x.reshape(x.shape[0], -1).astype(dt)        # encoders.py:44
            h = qf(h_exact)
            c["layers"] = []
            for i, width in enumerate(ENC_HIDDEN[t]):
                lc = {"x": q(h_exact)}
                pre = _linear(h, state[nm["fc"][i] + ".weight"], state[nm["fc"][i] + ".bias"], qf)
                bn = nm["bn"][i]
                if train:
                    n = pre.shape[0]
                    if n < 2:
                        raise ValueError("Expected more than 1 value per channel when training")
                    mean = pre.mean(0)
                    var = pre.var(0)                              # biased
                    if update_running:
                        state[bn + ".running_mean"] = ((1 - BN_MOMENTUM) * state[bn + ".running_mean"]
                                                       + BN_MOMENTUM * mean).astype(dt)
                        state[bn + ".running_var"] = ((1 - BN_MOMENTUM) * state[bn + ".running_var"]
                                                      + BN_MOMENTUM * var * n / (n - 1)).astype(dt)
                        state[bn + ".num_batches_tracked"] = state[bn + ".num_batches_tracked"] + 1
                else:
                    mean, var = state[bn + ".running_mean"], state[bn + ".running_var"]
                rstd = 1.0 / np.sqrt(var + BN_EPS)
                xhat = (pre - mean) * rstd
                y = xhat * state[bn + ".weight"] + state[bn + ".bias"]
                r = np.maximum(y, 0)
                if train:
                    keep = masks[nm["drop"][i]].astype(dt)
                    h_exact = r * keep / (1 - DROPOUT_P)
                else:
                    keep = None
                    h_exact = r
                h = qf(h_exact)
                lc.update(xhat=xhat, rstd=rstd, y=y, keep=keep)
                c["layers"].append(lc)
        c["h"] = q(h_exact)
        mu = _linear(h, state[nm["heads"][0] + ".weight"], state[nm["heads"][0] + ".bias"], qf)
        mus.append(mu)
        if len(nm["heads"]) == 2:
            lvs.append(_linear(h, state[nm["heads"][1] + ".weight"], state[nm["heads"][1] + ".bias"], qf))
        cache["enc"][prefix] = c
        cache["present"].append((prefix, t))
    if not mus:
        return None, None
    mu = mus[0] if len(mus) == 1 else np.stack(mus).mean(0)       # vae.py:70-71; directional_ae.py:53-56
    if is_ae(kind):
        logvar, std, z = None, None, mu                           # the latent itself feeds the decoder
    else:
        logvar = lvs[0] if len(lvs) == 1 else np.stack(lvs).mean(0)
        std = np.exp(0.5 * logvar)
        z = mu + eps * std                                        # vae.py:13-15
    cache.update(mu=mu, logvar=logvar, std=std, eps=eps, z=z)
    recon = {}
    for prefix, t in spec["decoders"]:
        widths = DEC_HIDDEN[t]
        k = len(widths)
        h = qf(z)
        acts = [q(z)]
        for i in range(k):
            h_exact = np.maximum(_linear(h, state[f"{prefix}.fc.{2 * i}.weight"], state[f"{prefix}.fc.{2 * i}.bias"], qf), 0)
            # the output layer (fc.{2k}) feeds no ReLU: single rounding -- except the site classifier's (argmax parity)
            h = (qf if (i + 1 < k or t == "C") else q)(h_exact)
            acts.append(q(h_exact))
        out = _linear(h, state[f"{prefix}.fc.{2 * k}.weight"], state[f"{prefix}.fc.{2 * k}.bias"], qf if t == "C" else q)
        if t == "B":
            out = 1.0 / (1.0 + np.exp(-out))                      # decoders.py:32
        recon[prefix] = out
        cache["dec"][prefix] = dict(acts=acts, out=out, type=t)
    return dict(recon=recon, mu=mu, logvar=logvar, z=z), cache


# --------------------------------------------------------------------------------------
# Losses (forward value and gradient w.r.t. the model outputs)
# --------------------------------------------------------------------------------------
def mse_sum(recon, target):
    d = recon - target
    return float(np.sum(d.astype(np.float64) ** 2)), 2.0 * d


def bce_sum(recon, target):
    """F.binary_cross_entropy(reduction='sum') incl. ATen's log clamp and backward floor."""
    logy = np.maximum(np.log(recon), -100.0)
    log1my = np.maximum(np.log1p(-recon), -100.0)
    val = -np.sum((target * logy + (1 - target) * log1my).astype(np.float64))
    grad = (recon - target) / np.maximum(recon * (1 - recon), 1e-12)
    return float(val), grad


def ce_sum(logits, site, weight=None):
    m = logits.max(1, keepdims=True)
    ex = np.exp(logits - m)
    lse = np.log(ex.sum(1, keepdims=True)) + m
    logp = logits - lse
    n = logits.shape[0]
    w = np.ones(n, dtype=logits.dtype) if weight is None else weight[site]
    val = -np.sum((w * logp[np.arange(n), site]).astype(np.float64))
    g = np.exp(logp)
    g[np.arange(n), site] -= 1.0
    return float(val), g * w[:, None]


def kld_sum(mu, logvar):
    val = -0.5 * np.sum((1 + logvar - mu ** 2 - np.exp(logvar)).astype(np.float64))
    return float(val), mu.copy(), 0.5 * (np.exp(logvar) - 1.0)


def loss_and_output_grads(kind, outputs, targets, beta=1e-3, gamma=1.0, class_weights=None):
    """total = recon + gamma*class + beta*kld with the terms each model kind uses
    (losses.py:27-44; directional_losses.py:23-28, 48-53).

    targets = {'a':..., 'b':..., 'site':...}.  Returns (scalars, grads) where
    scalars = dict(total, recon, cls, kld) and grads = {'recon': {prefix: dL/drecon}, 'mu', 'logvar'}."""
    spec = MODEL_KINDS[kind]
    recon_val, cls_val = 0.0, 0.0
    g_recon = {}
    for prefix, t in spec["decoders"]:
        r = outputs["recon"][prefix]
        if t == "A":
            v, g = mse_sum(r, targets["a"])
            recon_val += v
        elif t == "B":
            v, g = bce_sum(r, targets["b"])
            recon_val += v
        else:
            v, g = ce_sum(r, targets["site"], class_weights)
            cls_val += v
            g = gamma * g
        g_recon[prefix] = g
    if outputs["logvar"] is None:                                 # autoencoders: reconstruction only (ae_losses.py:8-39)
        return (dict(total=recon_val, recon=recon_val, cls=0.0, kld=0.0),
                dict(recon=g_recon, mu=np.zeros_like(outputs["mu"]), logvar=None))
    kld, gmu, glv = kld_sum(outputs["mu"], outputs["logvar"])
    total = recon_val + gamma * cls_val + beta * kld
    return (dict(total=total, recon=recon_val, cls=cls_val, kld=kld),
            dict(recon=g_recon, mu=beta * gmu, logvar=beta * glv))


# --------------------------------------------------------------------------------------
# Backward (explicit; what autograd does for the reference)
# --------------------------------------------------------------------------------------
def backward(kind, dims, state, cache, out_grads, train=True):
    """Gradients of every trainable parameter given dL/d(recon, mu, logvar)."""
    spec = MODEL_KINDS[kind]
    q = cache.get("q", _ident)
    grads = {}
    gz = np.zeros_like(cache["z"])
    for prefix, t in spec["decoders"]:
        g = out_grads["recon"].get(prefix)
        if g is None:
            continue
        dc = cache["dec"][prefix]
        acts = dc["acts"]
        k = len(DEC_HIDDEN[t])
        if t == "B":
            y = dc["out"]
            g = g * y * (1 - y)                                   # sigmoid backward
        g = q(g)
        for i in range(k, -1, -1):
            w = q(state[f"{prefix}.fc.{2 * i}.weight"])
            grads[f"{prefix}.fc.{2 * i}.weight"] = g.T @ acts[i]
            grads[f"{prefix}.fc.{2 * i}.bias"] = g.sum(0)
            g = g @ w
            if i > 0:
                g = q(g * (acts[i] > 0))
        gz = gz + g
    ae = is_ae(kind)
    gmu = gz + out_grads["mu"]
    m = len(cache["present"])
    gmu_e = q(gmu / m)                                            # mean fusion (identity when m == 1)
    if not ae:
        glv = gz * cache["eps"] * 0.5 * cache["std"] + out_grads["logvar"]
        glv_e = q(glv / m)
    for prefix, t in cache["present"]:
        c = cache["enc"][prefix]
        h = c["h"]
        nm = enc_names(kind, prefix, t)
        grads[nm["heads"][0] + ".weight"] = gmu_e.T @ h
        grads[nm["heads"][0] + ".bias"] = gmu_e.sum(0)
        g = gmu_e @ q(state[nm["heads"][0] + ".weight"])
        if not ae:
            grads[nm["heads"][1] + ".weight"] = glv_e.T @ h
            grads[nm["heads"][1] + ".bias"] = glv_e.sum(0)
            g = g + glv_e @ q(state[nm["heads"][1] + ".weight"])
        if t == "C":
            ge = np.zeros_like(state[nm["emb"]])
            np.add.at(ge, c["site"], q(g))                        # embedding_dense_backward
            grads[nm["emb"]] = ge
            continue
        for i in range(len(ENC_HIDDEN[t]) - 1, -1, -1):
            lc = c["layers"][i]
            bn = nm["bn"][i]
            if train:
                g = g * lc["keep"] / (1 - DROPOUT_P)
            g = g * (lc["y"] > 0)
            s1, s2 = g.sum(0), (g * lc["xhat"]).sum(0)            # column sums are taken before any rounding
            grads[bn + ".weight"] = s2
            grads[bn + ".bias"] = s1
            g = q(g)
            gam = state[bn + ".weight"]
            if train:
                n = g.shape[0]
                g = gam * lc["rstd"] * (g - s1 / n - lc["xhat"] * (s2 / n))
            else:
                g = g * gam * lc["rstd"]
            g = q(g)
            grads[nm["fc"][i] + ".weight"] = g.T @ lc["x"]
            grads[nm["fc"][i] + ".bias"] = g.sum(0)
            if i > 0:
                g = g @ q(state[nm["fc"][i] + ".weight"])
    # parameters of absent stacks get no gradient (autograd leaves .grad = None)
    return grads


# --------------------------------------------------------------------------------------
# AdamW (torch.optim.AdamW defaults used at train_rna2dna.py:185-189)
# --------------------------------------------------------------------------------------
def adamw_init(state):
    return {k: dict(m=np.zeros_like(v), v=np.zeros_like(v)) for k, v in state.items() if not is_buffer(k)}, 0


def adamw_step(state, grads, opt, step, lr=5e-4, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8):
    """One AdamW step in place; parameters without a gradient are skipped (as torch does).
    Returns the new step count."""
    step += 1
    b1, b2 = betas
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for k, g in grads.items():
        p = state[k]
        g = g.astype(p.dtype)
        p *= (1.0 - lr * weight_decay)
        o = opt[k]
        o["m"] = b1 * o["m"] + (1 - b1) * g
        o["v"] = b2 * o["v"] + (1 - b2) * g * g
        denom = np.sqrt(o["v"]) / np.sqrt(bc2) + eps
        p -= (lr / bc1) * (o["m"] / denom)
    return step


def train_step(kind, dims, state, opt, step, batch, eps, masks, beta=1e-3, gamma=1.0,
               class_weights=None, lr=5e-4, weight_decay=1e-5, q=None):
    """fwd + loss + bwd + AdamW, the loop body at train_rna2dna.py:82-99 /
    optimize_hyperparameters.py:104-113.  batch = {'a','b','site'} (the model's encoder inputs are
    selected per kind: rna2dna encodes (a, site), dna2rna encodes (b, site), multimodal all)."""
    spec = MODEL_KINDS[kind]
    enc_inputs = {INPUT_OF[t]: batch[INPUT_OF[t]] for _, t in spec["encoders"]}
    out, cache = forward(kind, dims, state, enc_inputs, eps, masks, train=True, q=q)
    scalars, og = loss_and_output_grads(kind, out, batch, beta, gamma, class_weights)
    grads = backward(kind, dims, state, cache, og, train=True)
    step = adamw_step(state, grads, opt, step, lr=lr, weight_decay=weight_decay)
    return scalars, out, grads, step
