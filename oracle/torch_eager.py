"""Stock-PyTorch restatement of the reference's train step, for the "same B200, no custom kernels" comparison line.

TEST / BENCH INFRASTRUCTURE ONLY (like the rest of oracle/): bench.py times it next to the CUDA path, tests/ check it
against the numpy oracle.  It is the reference's algorithm issued through the same ATen calls the reference's nn.Modules make
(F.linear, F.batch_norm, F.relu, F.dropout, F.embedding, F.mse_loss, F.binary_cross_entropy, F.cross_entropy,
torch.optim.AdamW), driven by a state dict with the reference's keys (oracle/vae_oracle.py param_shapes):

  encoders   src/models/encoders.py:8-61          decoders   src/models/decoders.py:8-50
  fusion     src/models/vae.py:37-79, src/models/directional_vae.py:36-60, 87-111
  losses     src/utils/losses.py:27-46, src/utils/directional_losses.py:23-30, 48-55
  loop body  train_rna2dna.py:82-99 (zero_grad -> forward -> loss -> backward -> optimizer.step)

The unmodified reference cannot travel to the GPU box (/root/reference exists only in the build container); this file can.
"""
import torch
import torch.nn.functional as F

from . import vae_oracle as vo


def params_from_state(state, device, dtype=torch.float32):
    """Trainable tensors (requires_grad) and BatchNorm buffers from an oracle / reference state dict."""
    params, buffers = {}, {}
    for k, v in state.items():
        t = torch.as_tensor(v)
        if k.endswith("num_batches_tracked"):
            buffers[k] = t.to(device)
        elif vo.is_buffer(k):
            buffers[k] = t.to(device, dtype).clone()
        else:
            params[k] = t.to(device, dtype).clone().requires_grad_(True)
    return params, buffers


def forward(kind, p, buf, inputs, train=True, eps=None, masks=None):
    """inputs = {'a', 'b', 'site'} (absent = missing).  eps / masks: injected noise (tests); None = torch's own generators."""
    spec = vo.MODEL_KINDS[kind]
    mus, lvs = [], []
    for prefix, t in spec["encoders"]:
        x = inputs.get(vo.INPUT_OF[t])
        if x is None:
            continue
        nm = vo.enc_names(kind, prefix, t)
        if t == "C":
            h = F.embedding(x, p[nm["emb"]])
        else:
            h = x.reshape(x.shape[0], -1)
            for i in range(len(vo.ENC_HIDDEN[t])):
                h = F.linear(h, p[nm["fc"][i] + ".weight"], p[nm["fc"][i] + ".bias"])
                bn = nm["bn"][i]
                h = F.batch_norm(h, buf[bn + ".running_mean"], buf[bn + ".running_var"], p[bn + ".weight"], p[bn + ".bias"],
                                 training=train, momentum=vo.BN_MOMENTUM, eps=vo.BN_EPS)
                if train:
                    buf[bn + ".num_batches_tracked"] += 1
                h = F.relu(h)
                if train:
                    if masks is not None:
                        h = h * masks[nm["drop"][i]].to(h.dtype) / (1 - vo.DROPOUT_P)
                    else:
                        h = F.dropout(h, vo.DROPOUT_P, training=True)
        mus.append(F.linear(h, p[nm["heads"][0] + ".weight"], p[nm["heads"][0] + ".bias"]))
        if len(nm["heads"]) == 2:
            lvs.append(F.linear(h, p[nm["heads"][1] + ".weight"], p[nm["heads"][1] + ".bias"]))
    mu = mus[0] if len(mus) == 1 else torch.stack(mus).mean(0)
    if vo.is_ae(kind):
        logvar, z = None, mu
    else:
        logvar = lvs[0] if len(lvs) == 1 else torch.stack(lvs).mean(0)
        std = torch.exp(0.5 * logvar)
        z = mu + (torch.randn_like(std) if eps is None else eps) * std
    recon = {}
    for prefix, t in spec["decoders"]:
        k = len(vo.DEC_HIDDEN[t])
        h = z
        for i in range(k):
            h = F.relu(F.linear(h, p[f"{prefix}.fc.{2 * i}.weight"], p[f"{prefix}.fc.{2 * i}.bias"]))
        out = F.linear(h, p[f"{prefix}.fc.{2 * k}.weight"], p[f"{prefix}.fc.{2 * k}.bias"])
        recon[prefix] = torch.sigmoid(out) if t == "B" else out
    return dict(recon=recon, mu=mu, logvar=logvar, z=z)


def loss(kind, out, targets, beta=1e-3, gamma=1.0, class_weights=None):
    recon_val, cls_val = 0.0, 0.0
    for prefix, t in vo.MODEL_KINDS[kind]["decoders"]:
        r = out["recon"][prefix]
        if t == "A":
            recon_val = recon_val + F.mse_loss(r, targets["a"], reduction="sum")
        elif t == "B":
            recon_val = recon_val + F.binary_cross_entropy(r, targets["b"], reduction="sum")
        else:
            cls_val = cls_val + F.cross_entropy(r, targets["site"], weight=class_weights, reduction="sum")
    if out["logvar"] is None:
        return recon_val, dict(total=recon_val, recon=recon_val, cls=0.0, kld=0.0)
    kld = -0.5 * torch.sum(1 + out["logvar"] - out["mu"].pow(2) - out["logvar"].exp())
    total = recon_val + gamma * cls_val + beta * kld
    return total, dict(total=total, recon=recon_val, cls=cls_val, kld=kld)


class EagerTrainer:
    """The loop body of train_rna2dna.py:82-99 on device-resident batches, plain eager PyTorch."""

    def __init__(self, kind, state, device, lr=5e-4, weight_decay=1e-5, beta=1e-3, gamma=1.0, dtype=torch.float32, autocast=None):
        self.kind, self.beta, self.gamma, self.autocast = kind, beta, gamma, autocast
        self.p, self.buf = params_from_state(state, device, dtype)
        self.opt = torch.optim.AdamW(list(self.p.values()), lr=lr, weight_decay=weight_decay)
        self.device = device

    def step(self, a, b, site):
        spec = vo.MODEL_KINDS[self.kind]
        batch = dict(a=a, b=b, site=site)
        enc_inputs = {vo.INPUT_OF[t]: batch[vo.INPUT_OF[t]] for _, t in spec["encoders"]}
        self.opt.zero_grad(set_to_none=True)
        if self.autocast is not None:
            with torch.autocast(device_type="cuda", dtype=self.autocast):
                out = forward(self.kind, self.p, self.buf, enc_inputs, train=True)
            out = dict(recon={k: v.float() for k, v in out["recon"].items()}, mu=out["mu"].float(),
                       logvar=None if out["logvar"] is None else out["logvar"].float(), z=out["z"])
        else:
            out = forward(self.kind, self.p, self.buf, enc_inputs, train=True)
        total, scal = loss(self.kind, out, batch, self.beta, self.gamma)
        total.backward()
        self.opt.step()
        return total
