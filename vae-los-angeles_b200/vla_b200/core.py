"""Host-side mirror of the reference's model interface on top of libvla_b200.

`VaeModule` is the common base of the drop-in `MultiModalVAE`, `RNA2DNAVAE` and `DNA2RNAVAE`
(`src/models/`): real `nn.Module`s with real `nn.Parameter`s under the reference's state_dict names
(SURVEY.md Appendix A), whose storage is one flat fp32 arena laid out by the library, and whose
forward / backward are two C-ABI calls wrapped in a `torch.autograd.Function`.
"""
import contextlib
import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib

# (submodule name, stack type) per model kind -- the reference's attribute names
# (src/models/vae.py:29-35; src/models/directional_vae.py:21-23, 72-74)
KINDS = {
    "multimodal": dict(encoders=[("encoder_a", "A"), ("encoder_b", "B"), ("encoder_c", "C")],
                       decoders=[("decoder_a", "A"), ("decoder_b", "B"), ("decoder_c", "C")]),
    "rna2dna": dict(encoders=[("encoder_rna", "A"), ("encoder_site", "C")], decoders=[("decoder_dna", "B")]),
    "dna2rna": dict(encoders=[("encoder_dna", "B"), ("encoder_site", "C")], decoders=[("decoder_rna", "A")]),
    # directional autoencoders (src/models/directional_ae.py:17-35, 76-98): the site encoder is two top-level modules
    "rna2dna_ae": dict(encoders=[("encoder_rna", "A"), ("site", "C")], decoders=[("decoder_dna", "B")]),
    "dna2rna_ae": dict(encoders=[("encoder_dna", "B"), ("site", "C")], decoders=[("decoder_rna", "A")]),
}
# state_dict prefixes owned by an encoder when they differ from its name
STACK_PREFIXES = {"site": ("site_embedding", "site_projection")}
ENC_HIDDEN = {"A": (128,), "B": (512, 256)}
DEC_HIDDEN = {"A": (128,), "B": (256, 512), "C": (64,)}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, name):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# Parameter containers (state_dict names only; they hold no compute)
# ------------------------------------------------------------------------------------------------
class LinearParams(nn.Module):
    """weight [out, in], bias [out]; PyTorch's nn.Linear default init: U(+-1/sqrt(fan_in)) for both."""

    def __init__(self, in_features, out_features):
        super().__init__()
        bound = 1.0 / math.sqrt(in_features)
        self.weight = nn.Parameter(torch.empty(out_features, in_features).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))


class BatchNormParams(nn.Module):
    """weight=1, bias=0, running_mean=0, running_var=1, num_batches_tracked=0 (nn.BatchNorm1d defaults)."""

    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class EmbeddingParams(nn.Module):
    """weight [n, dim] ~ N(0, 1) (nn.Embedding default)."""

    def __init__(self, n, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n, dim).normal_())


class Slots(nn.Module):
    """Numbered children `fc.0`, `fc.1`, ... at the indices the reference's nn.Sequential uses."""

    def __init__(self, children):
        super().__init__()
        for idx, mod in children:
            self.add_module(str(idx), mod)


class Stack(nn.Module):
    """One encoder or decoder stack: parameters only.  Standalone calls are not part of the hot path."""

    def __init__(self, role, stack_type, feature_dim, latent_dim, embed_dim=32):
        super().__init__()
        self.role, self.stack_type = role, stack_type
        if role == "enc":
            if stack_type == "C":
                self.embedding = EmbeddingParams(feature_dim, embed_dim)
                last = embed_dim
            else:
                kids, last = [], feature_dim
                for i, h in enumerate(ENC_HIDDEN[stack_type]):
                    kids.append((4 * i, LinearParams(last, h)))
                    kids.append((4 * i + 1, BatchNormParams(h)))
                    last = h
                self.fc = Slots(kids)
            self.fc_mu = LinearParams(last, latent_dim)
            self.fc_logvar = LinearParams(last, latent_dim)
        else:
            kids, last = [], latent_dim
            for i, h in enumerate(DEC_HIDDEN[stack_type] + (feature_dim,)):
                kids.append((2 * i, LinearParams(last, h)))
                last = h
            self.fc = Slots(kids)

    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container of the fused B200 path; call the owning VAE module "
            "(MultiModalVAE / RNA2DNAVAE / DNA2RNAVAE). There is no per-stack or CPU fallback.")


# ------------------------------------------------------------------------------------------------
# autograd bridge
# ------------------------------------------------------------------------------------------------
class _VaeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, a, b, site, *params):
        core = module._core
        outs = core.forward(a, b, site, module.training, module._injected)
        ctx.module = module
        ctx.generation = core.generation
        ctx.n_params = len(params)
        recon_b = outs[core.dec_slots.index(1)] if 1 in core.dec_slots else None
        ctx.save_for_backward(recon_b)
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        core = ctx.module._core
        if core.generation != ctx.generation:
            raise RuntimeError(
                "vla_b200: the activations of this forward were overwritten by a later forward on the same module "
                "before backward() ran; run forward -> backward back to back (one in-flight graph per module).")
        (recon_b,) = ctx.saved_tensors
        flat, active = core.backward(gouts, recon_b)
        grads = []
        for stack, (offset, shape) in zip(core.param_stacks, core.param_offsets):
            if stack in active:
                n = 1
                for d in shape:
                    n *= d
                grads.append(flat[offset:offset + n].view(shape))
            else:
                grads.append(None)      # autograd leaves .grad = None for unused stacks, as in the reference
        return (None, None, None, None) + tuple(grads)


class Core:
    """Library handle + flat arenas of one module on one CUDA device."""

    def __init__(self, module, device):
        L = _lib.lib()
        self.device = device
        cfg = _lib.Config(_lib.KIND[module.kind], module.dim_a, module.dim_b, module.n_sites, module.latent_dim,
                          module.embed_dim)
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.vla_model_create(C.byref(cfg), C.byref(handle)), "vla_model_create")
        self.handle = handle
        self.n_params = L.vla_param_count(handle)
        self.n_buffers = L.vla_buffer_count(handle)
        self.n_counters = L.vla_counter_count(handle)
        self.infos = []
        info = _lib.TensorInfo()
        for i in range(L.vla_num_tensors(handle)):
            _lib.check(L.vla_tensor_info(handle, i, C.byref(info)), "vla_tensor_info")
            shape = tuple(info.shape[: info.ndim])
            self.infos.append((info.name.decode(), info.kind, info.offset, shape))
        self.arena = torch.zeros(self.n_params, dtype=torch.float32, device=device)
        self.buffers = torch.zeros(max(self.n_buffers, 1), dtype=torch.float32, device=device)
        self.counters = torch.zeros(max(self.n_counters, 1), dtype=torch.long, device=device)
        self.kind = module.kind
        self.dims = (module.dim_a, module.dim_b, module.n_sites)
        self.latent = module.latent_dim
        # decoder output slots in the module's return order: 0 = A (RNA), 1 = B (DNA), 2 = C (site logits)
        self.dec_slots = [{"A": 0, "B": 1, "C": 2}[t] for _, t in KINDS[module.kind]["decoders"]]
        self.generation = 0
        self.calls = 0
        self.shadow_version = None
        self.last_grads = None
        self.present = ()

    def __deepcopy__(self, memo):
        return None                      # a copied module re-packs into its own arena on first use

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().vla_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- views -----------------------------------------------------------------------------------
    def view(self, kind, offset, shape):
        n = 1
        for s in shape:
            n *= s
        if kind == _lib.TENSOR_PARAM:
            return self.arena[offset:offset + n].view(shape)
        if kind == _lib.TENSOR_BUFFER:
            return self.buffers[offset:offset + n].view(shape)
        return self.counters[offset]

    # -- forward / backward ------------------------------------------------------------------------
    def _check_input(self, t, name, width, dtype):
        if t is None:
            return None
        if not t.is_cuda or t.device != self.device:
            raise RuntimeError(f"vla_b200: input `{name}` is on {t.device}, the module is on {self.device}; "
                               "there is no CPU fallback")
        if dtype == torch.long:
            if t.dtype != torch.long:
                t = t.long()
            return t.contiguous().view(-1)
        t = _f32c(t, name)
        t = t.view(t.shape[0], -1)          # EncoderB flattens: src/models/encoders.py:44
        if t.shape[1] != width:
            raise RuntimeError(f"vla_b200: `{name}` has {t.shape[1]} features, the module expects {width}")
        return t

    def forward(self, a, b, site, training, injected):
        L = _lib.lib()
        a = self._check_input(a, "a", self.dims[0], torch.float32)
        b = self._check_input(b, "b", self.dims[1], torch.float32)
        site = self._check_input(site, "site", 0, torch.long)
        batch = next(t.shape[0] for t in (a, b, site) if t is not None)
        for t in (a, b, site):
            if t is not None and t.shape[0] != batch:
                raise RuntimeError("vla_b200: modality batch sizes differ")
        dev = self.device
        outs = []
        ptrs = [None, None, None]
        for slot in self.dec_slots:
            t = torch.empty(batch, self.dims[slot], dtype=torch.float32, device=dev)
            outs.append(t)
            ptrs[slot] = t
        mu = torch.empty(batch, self.latent, dtype=torch.float32, device=dev)
        logvar = torch.empty(batch, self.latent, dtype=torch.float32, device=dev)
        eps = masks = None
        if injected is not None:
            eps = injected.get("eps")
            masks = injected.get("keep_masks")
        keep = None
        mask_arr = None
        if masks is not None:
            keep = [None if m is None else m.to(device=dev, dtype=torch.uint8).contiguous() for m in masks]
            mask_arr = (C.c_void_p * len(keep))(*[None if m is None else m.data_ptr() for m in keep])
        if eps is not None:
            eps = _f32c(eps.to(dev), "eps")
            if tuple(eps.shape) != (batch, self.latent):
                raise RuntimeError("vla_b200: injected eps has the wrong shape")
        version = self.param_version()
        by_type = {"A": a, "B": b, "C": site}
        self.present = tuple(p for name, t in KINDS[self.kind]["encoders"] if by_type[t] is not None
                             for p in STACK_PREFIXES.get(name, (name,)))
        args = _lib.ForwardArgs(
            params=_ptr(self.arena), buffers=_ptr(self.buffers), counters=_ptr(self.counters),
            x_a=_ptr(a), x_b=_ptr(b), site=_ptr(site), batch=batch, train=1 if training else 0,
            refresh_shadows=0 if self.shadow_version == version else 1,
            eps=_ptr(eps), keep_masks=mask_arr if mask_arr is not None else None,
            seed=torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, offset=self.calls,
            recon_a=_ptr(ptrs[0]), recon_b=_ptr(ptrs[1]), recon_c=_ptr(ptrs[2]), mu=_ptr(mu), logvar=_ptr(logvar))
        with torch.cuda.device(dev):
            _lib.check(L.vla_forward(self.handle, C.byref(args), _stream()), "vla_forward")
        self.shadow_version = version
        self.calls += 1
        self.generation += 1
        self._keepalive = (a, b, site, eps, keep)
        return tuple(outs) + (mu, logvar)

    def backward(self, gouts, recon_b):
        L = _lib.lib()
        dev = self.device
        g = [None, None, None]
        for slot, t in zip(self.dec_slots, gouts[: len(self.dec_slots)]):
            g[slot] = _f32c(t, "grad")
        g_mu = _f32c(gouts[len(self.dec_slots)], "grad_mu")
        g_lv = _f32c(gouts[len(self.dec_slots) + 1], "grad_logvar")
        grads = torch.empty(self.n_params, dtype=torch.float32, device=dev)
        args = _lib.BackwardArgs(params=_ptr(self.arena), g_recon_a=_ptr(g[0]), g_recon_b=_ptr(g[1]),
                                 g_recon_c=_ptr(g[2]), g_mu=_ptr(g_mu), g_logvar=_ptr(g_lv),
                                 recon_b=_ptr(recon_b), grads=_ptr(grads))
        with torch.cuda.device(dev):
            _lib.check(L.vla_backward(self.handle, C.byref(args), _stream()), "vla_backward")
        self.last_grads = grads
        active = set(self.present)
        for (name, _), t in zip(KINDS[self.kind]["decoders"], gouts):
            if t is not None:
                active.add(name)
        return grads, active

    def param_version(self):
        """Changes whenever any parameter is updated in place (optimizer step, load_state_dict)."""
        return sum(p._version for p in self.order)


class VaeModule(nn.Module):
    """Base of the drop-in VAE modules.  Subclasses define `kind` and the forward keyword names."""

    kind = None

    def __init__(self, dim_a, dim_b, n_sites, latent_dim, embed_dim=32):
        super().__init__()
        self.dim_a, self.dim_b, self.n_sites = int(dim_a), int(dim_b), int(n_sites)
        self.latent_dim, self.embed_dim = int(latent_dim), int(embed_dim)
        feat = {"A": self.dim_a, "B": self.dim_b, "C": self.n_sites}
        for name, t in KINDS[self.kind]["encoders"]:
            self.add_module(name, self._make_stack("enc", t, feat[t]))
        for name, t in KINDS[self.kind]["decoders"]:
            self.add_module(name, self._make_stack("dec", t, feat[t]))
        self.__dict__["_core"] = None
        self.__dict__["_injected"] = None

    def _make_stack(self, role, t, feature_dim):
        return Stack(role, t, feature_dim, self.latent_dim, self.embed_dim)

    # -- arena management --------------------------------------------------------------------------
    def _named_tensors(self):
        out = dict(self.named_parameters())
        out.update(dict(self.named_buffers()))
        return out

    def _pack(self, device):
        """Re-home every parameter / buffer into the library's flat arenas on `device` (values are kept)."""
        core = Core(self, device)
        tensors = self._named_tensors()
        names = {n for n, _, _, _ in core.infos}
        if names != set(tensors):
            raise RuntimeError(f"vla_b200: state_dict mismatch: {sorted(names ^ set(tensors))}")
        with torch.no_grad():
            for name, kind, offset, shape in core.infos:
                t = tensors[name]
                view = core.view(kind, offset, shape)
                view.copy_(t.detach().to(device=device, dtype=view.dtype).reshape(view.shape))
                if kind == _lib.TENSOR_PARAM:
                    t.data = view
                else:
                    owner_name, _, leaf = name.rpartition(".")
                    self.get_submodule(owner_name)._buffers[leaf] = view
        core.order = [tensors[name] for name, kind, _, _ in core.infos if kind == _lib.TENSOR_PARAM]
        core.param_offsets = [(offset, shape) for _, kind, offset, shape in core.infos if kind == _lib.TENSOR_PARAM]
        core.param_stacks = [name.split(".")[0] for name, kind, _, _ in core.infos if kind == _lib.TENSOR_PARAM]
        self.__dict__["_core"] = core
        return core

    def _is_packed(self, core):
        base = core.arena.data_ptr()
        for p, (offset, _) in zip(core.order, core.param_offsets):
            if p.data_ptr() != base + 4 * offset:
                return False
        return True

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        p = next(self.parameters())
        if p.is_cuda:
            if p.dtype != torch.float32:
                raise RuntimeError("vla_b200 keeps fp32 master parameters; .half()/.bfloat16() is not supported")
            core = self._core
            if core is None or core.device != p.device or not self._is_packed(core):
                self._pack(p.device)
        else:
            self.__dict__["_core"] = None
        return self

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_core"] = None
        state["_injected"] = None
        return state

    def _ensure_core(self):
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("vla_b200: the module is on the CPU; move it to a B200 with .to('cuda'). "
                               "There is no CPU fallback.")
        core = self._core
        if core is None or core.device != p.device or not self._is_packed(core):
            core = self._pack(p.device)
        return core

    # -- test hooks --------------------------------------------------------------------------------
    @contextlib.contextmanager
    def inject(self, eps=None, keep_masks=None):
        """Replay a recorded epsilon / dropout keep-masks (one per dropout layer, model order)."""
        self.__dict__["_injected"] = dict(eps=eps, keep_masks=keep_masks)
        try:
            yield self
        finally:
            self.__dict__["_injected"] = None

    # -- forward -----------------------------------------------------------------------------------
    def _run(self, a, b, site):
        core = self._ensure_core()
        outs = _VaeFunction.apply(self, a, b, site, *core.order)
        return outs

    def flat_parameters(self):
        """The flat fp32 parameter arena (all nn.Parameters are views of it)."""
        return self._ensure_core().arena


class AeModule(VaeModule):
    """Base of the drop-in directional autoencoders (reference src/models/directional_ae.py:10-134): the encoder is a bare
    nn.Sequential whose last Linear is the single head, the site branch is `site_embedding` + `site_projection`, the
    decoder is the VAEs' DecoderA / DecoderB.  forward returns (reconstruction, latent)."""

    def __init__(self, dim_a, dim_b, n_sites, latent_dim, embed_dim=32):
        nn.Module.__init__(self)
        self.dim_a, self.dim_b, self.n_sites = int(dim_a), int(dim_b), int(n_sites)
        self.latent_dim, self.embed_dim = int(latent_dim), int(embed_dim)
        feat = {"A": self.dim_a, "B": self.dim_b, "C": self.n_sites}
        (enc_name, enc_t), _ = KINDS[self.kind]["encoders"]
        kids, last = [], feat[enc_t]
        for i, h in enumerate(ENC_HIDDEN[enc_t]):
            kids.append((4 * i, LinearParams(last, h)))
            kids.append((4 * i + 1, BatchNormParams(h)))
            last = h
        kids.append((4 * len(ENC_HIDDEN[enc_t]), LinearParams(last, self.latent_dim)))
        self.add_module(enc_name, Slots(kids))
        self.site_embedding = EmbeddingParams(self.n_sites, self.embed_dim)
        self.site_projection = LinearParams(self.embed_dim, self.latent_dim)
        for name, t in KINDS[self.kind]["decoders"]:
            self.add_module(name, self._make_stack("dec", t, feat[t]))
        self.__dict__["_core"] = None
        self.__dict__["_injected"] = None

    def _run_ae(self, a, b, site):
        recon, latent, _ = self._run(a, b, site)
        return recon, latent
