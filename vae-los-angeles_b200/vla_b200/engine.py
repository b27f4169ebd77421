"""Fused training engine: whole step (forward + loss + backward + AdamW) as one C-ABI call,
captured into a CUDA graph, over a device-resident dataset.

`FusedAdamW` is the drop-in optimizer for scripts that want the fused multi-tensor kernel while keeping
the per-call autograd path; `Trainer` is the throughput path used by bench.py and the DP launcher:
it restates the loop body at train_rna2dna.py:82-99 / optimize_hyperparameters.py:104-113 with zero
host synchronisation per step.
"""
import ctypes as C

import torch

from . import _lib
from .core import _ptr, _stream


class _DeviceFloats:
    """Library-owned device memory seen through __cuda_array_interface__ (zero-copy torch view)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


def _wrap_device_floats(ptr, n, device):
    if not ptr:
        raise RuntimeError("null device pointer from libvla_b200")
    return torch.as_tensor(_DeviceFloats(ptr, n), device=device)


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled decay, bias correction) in one launch over the flat arena.

    Construct with the module, not a parameter list: `FusedAdamW(model, lr=5e-4, weight_decay=1e-5)`.
    LR schedulers work as usual (the learning rate is read from `param_groups[0]['lr']` every step)."""

    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.module = module
        core = module._ensure_core()
        super().__init__(list(core.order), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.exp_avg = torch.zeros_like(core.arena)
        self.exp_avg_sq = torch.zeros_like(core.arena)
        self.steps = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        core = self.module._ensure_core()
        grads = core.last_grads
        if grads is None or any(p.grad is None for p in core.order):
            raise RuntimeError("FusedAdamW.step(): every parameter needs a gradient from the fused backward "
                               "(use torch.optim.AdamW when training a subset of the stacks)")
        base = grads.data_ptr()
        for p, (offset, _) in zip(core.order, core.param_offsets):
            if p.grad.data_ptr() != base + 4 * offset:
                raise RuntimeError("FusedAdamW.step(): .grad tensors are not views of the fused gradient arena "
                                   "(gradient accumulation / clipping copies are not supported here)")
        g = self.param_groups[0]
        self.steps += 1
        args = _lib.AdamWArgs(params=_ptr(core.arena), grads=_ptr(grads), exp_avg=_ptr(self.exp_avg),
                              exp_avg_sq=_ptr(self.exp_avg_sq), lr=float(g["lr"]), beta1=float(g["betas"][0]),
                              beta2=float(g["betas"][1]), eps=float(g["eps"]), weight_decay=float(g["weight_decay"]),
                              step=self.steps)
        with torch.cuda.device(core.device):
            _lib.check(_lib.lib().vla_adamw(core.handle, C.byref(args), _stream()), "vla_adamw")
        # the kernel wrote through the arena and refreshed the bf16 copies itself
        core.shadow_version = core.param_version()
        return loss


class DeviceDataset:
    """Device-resident dataset with the reference's schema (src/data/dataset.py:19-30):
    tpm fp32 [N, A], beta fp32 [N, B], site int64 [N]."""

    def __init__(self, tpm, beta, site, device):
        self.tpm = torch.as_tensor(tpm, dtype=torch.float32).to(device).contiguous()
        self.beta = torch.as_tensor(beta, dtype=torch.float32).to(device).contiguous()
        self.site = torch.as_tensor(site, dtype=torch.long).to(device).contiguous()
        if not (len(self.tpm) == len(self.beta) == len(self.site)):
            raise ValueError("modalities differ in length")

    def __len__(self):
        return len(self.site)

    def shuffle_(self, generator=None):
        """In-place row permutation on the device (DataLoader(shuffle=True) between epochs).  The tensors keep their
        addresses, so a captured train-step graph that reads them stays valid."""
        perm = torch.randperm(len(self), device=self.site.device, generator=generator)
        self.tpm.copy_(self.tpm[perm])
        self.beta.copy_(self.beta[perm])
        self.site.copy_(self.site[perm])

    def gather_into(self, index, out):
        """Rows `index` (int64 on the device) of all three arrays into the batch slot `out` (a DeviceDataset of len(index)
        rows whose storage a captured step graph reads): what DataLoader's sampler + collate do for one batch
        (src/data/dataset.py:35-39, train_rna2dna.py:57-67), as ONE kernel on the device -- shuffled training without
        permuting the whole dataset."""
        from . import _lib
        from .core import _ptr, _stream
        index = index.to(self.site.device, torch.long).contiguous()
        n = int(index.numel())
        if n != len(out):
            raise ValueError("gather_into: the batch slot must hold exactly len(index) rows")
        with torch.cuda.device(self.site.device):
            _lib.check(_lib.lib().vla_gather_rows(_ptr(self.tpm), self.tpm.shape[1], _ptr(self.beta), self.beta.shape[1], _ptr(self.site),
                                                  len(self), _ptr(index), n, _ptr(out.tpm), _ptr(out.beta), _ptr(out.site), _stream()),
                       "vla_gather_rows")
        return out

    @staticmethod
    def synthetic(n, dim_a, dim_b, n_sites, device, seed=0):
        """Synthetic rows of the reference's schema (scripts/prepare_data.py:112-125): tpm = log1p(Gamma(1, 20)),
        beta ~ Beta(0.5, 0.5) clipped to [1e-4, 1 - 1e-4], site uniform."""
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        u = torch.rand(n, dim_a, device=device, generator=g)
        tpm = torch.log1p(-20.0 * torch.log1p(-u))
        v = torch.rand(n, dim_b, device=device, generator=g)
        beta = torch.sin(0.5 * torch.pi * v).pow(2).clamp_(1e-4, 1 - 1e-4)
        site = torch.randint(0, n_sites, (n,), device=device, generator=g)
        return DeviceDataset(tpm, beta, site, device)


class Trainer:
    """Whole-step trainer: `step()` enqueues one fused train step (a CUDA-graph replay) with no host sync.

    Data parallel (one process per GPU, one node): pass `process_group`; the step becomes
    [forward + loss + backward] -> one all-reduce(SUM) of the flat gradient arena with the 4 loss scalars
    appended -> AdamW.  SUM, not mean: the reference's losses are reduction='sum' (SURVEY D10), so the result
    equals a single-process run on the concatenated batch up to BatchNorm, which uses per-shard statistics.
    exchange="p2p" (default): the collective is this library's own kernel over NVLink peer memory (csrc/dp_exchange.cu;
    torch.distributed only carries the 64-byte IPC handles once); exchange="nccl": torch.distributed.all_reduce."""

    def __init__(self, module, dataset, batch_size, lr=5e-4, weight_decay=1e-5, betas=(0.9, 0.999), eps=1e-8,
                 beta_kl=1e-3, gamma=1.0, class_weights=None, seed=0, use_graph=True, process_group=None, exchange="p2p",
                 sync_bn=False):
        self.module = module
        if sync_bn and (process_group is None or exchange != "p2p"):
            raise ValueError("sync_bn=True needs process_group and the peer-memory exchange (exchange='p2p')")
        self.sync_bn = bool(sync_bn)
        self.core = module._ensure_core()
        self.datasets = list(dataset) if isinstance(dataset, (list, tuple)) else [dataset]
        self.batch = int(batch_size)
        if any(len(d) < self.batch for d in self.datasets):
            raise ValueError("dataset smaller than one batch")
        dev = self.core.device
        n = self.core.n_params
        self.pg = process_group
        self.dp = None
        if exchange not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        # gradient arena with the loss scalars appended: one collective moves both
        if process_group is not None and exchange == "p2p":
            self._connect_peers(n + 4)
            self.grads = self.flat[:n]
            self._loss_local = self.flat[n:]          # this rank's losses (summed into the reduced buffer with the gradients)
            self.loss_out = self.reduced           # float[4] summed over the ranks (written by the AdamW kernel)
        else:
            self.flat = torch.zeros(n + 4, dtype=torch.float32, device=dev)
            self.grads = self.flat[:n]
            self.loss_out = self._loss_local = self.flat[n:]
        self.exp_avg = torch.zeros_like(self.core.arena)
        self.exp_avg_sq = torch.zeros_like(self.core.arena)
        self.class_weights = None if class_weights is None else class_weights.to(dev, torch.float32).contiguous()
        self.betas, self.eps, self.seed = betas, eps, seed
        self.hyper = None
        self.set_hyper(lr, weight_decay, beta_kl, gamma)
        self.steps = 0
        self.use_graph = use_graph
        self.graphs = {}
        self.injected = None
        self._mask_arr = None
        _lib.check(_lib.lib().vla_model_reserve(self.core.handle, self.batch), "vla_model_reserve")
        # the graphs captured below bake workspace addresses into their kernel arguments: while this trainer lives, a call on
        # the same handle that would have to reallocate the workspace (a larger batch) fails instead of corrupting them
        _lib.check(_lib.lib().vla_model_pin(self.core.handle, 1), "vla_model_pin")
        self._pinned = True
        self.reset_counters(0, 0)

    def _connect_peers(self, n_floats):
        """Allocates this rank's exchange buffers in the library, swaps the CUDA IPC handles through the process group
        (host bytes, once) and maps every peer's buffers."""
        import torch.distributed as dist
        L = _lib.lib()
        dev = self.core.device
        world, rank = dist.get_world_size(self.pg), dist.get_rank(self.pg)
        with torch.cuda.device(dev):
            handle = C.c_void_p()
            _lib.check(L.vla_dp_create(world, rank, n_floats, C.byref(handle)), "vla_dp_create")
            self.dp = handle
            mine = C.create_string_buffer(64)
            _lib.check(L.vla_dp_ipc_handle(self.dp, mine), "vla_dp_ipc_handle")
            gathered = [None] * world
            dist.all_gather_object(gathered, bytes(mine.raw), group=self.pg)
            blob = b"".join(gathered)
            _lib.check(L.vla_dp_connect(self.dp, blob), "vla_dp_connect")
            self.flat = _wrap_device_floats(L.vla_dp_grads(self.dp), n_floats, dev)
            self.reduced = _wrap_device_floats(L.vla_dp_losses(self.dp), 4, dev)
            dist.barrier(group=self.pg)               # every rank has mapped every buffer before anyone steps

    # -- per-epoch scalars (beta warm-up train_rna2dna.py:80, ReduceLROnPlateau :216) ------------------
    def set_hyper(self, lr=None, weight_decay=None, beta_kl=None, gamma=None):
        cur = list(self.hyper) if self.hyper else [5e-4, 1e-5, 1e-3, 1.0]
        for i, v in enumerate((lr, weight_decay, beta_kl, gamma)):
            if v is not None:
                cur[i] = float(v)
        if self.hyper != tuple(cur):
            self.hyper = tuple(cur)
            with torch.cuda.device(self.core.device):
                _lib.check(_lib.lib().vla_set_hyper(self.core.handle, *self.hyper, _stream()), "vla_set_hyper")

    def reset_counters(self, completed_steps, batch_index):
        with torch.cuda.device(self.core.device):
            _lib.check(_lib.lib().vla_set_step(self.core.handle, int(completed_steps), int(batch_index),
                                               float(self.betas[0]), float(self.betas[1]), _stream()), "vla_set_step")

    def _args(self, ds, phases):
        core = self.core
        eps = masks = None
        if self.injected is not None:
            eps = self.injected.get("eps")
            keep = self.injected.get("keep_masks")
            if keep is not None:
                self._keep = [None if m is None else m.to(core.device, torch.uint8).contiguous() for m in keep]
                self._mask_arr = (C.c_void_p * len(self._keep))(*[None if m is None else m.data_ptr() for m in self._keep])
                masks = self._mask_arr
        return _lib.TrainArgs(
            params=_ptr(core.arena), grads=_ptr(self.grads), exp_avg=_ptr(self.exp_avg), exp_avg_sq=_ptr(self.exp_avg_sq),
            buffers=_ptr(core.buffers), counters=_ptr(core.counters),
            x_a=_ptr(ds.tpm), x_b=_ptr(ds.beta), site=_ptr(ds.site), class_weights=_ptr(self.class_weights),
            batch=self.batch, dataset_rows=len(ds), eps=_ptr(eps), keep_masks=masks, seed=self.seed,
            beta1=self.betas[0], beta2=self.betas[1], adam_eps=self.eps,
            recon_a=None, recon_b=None, recon_c=None, mu=None, logvar=None, loss_out=_ptr(self._loss_local), phases=phases,
            dp=self.dp, sync_bn=1 if self.sync_bn else 0)

    def _call(self, ds, phases):
        args = self._args(ds, phases)
        _lib.check(_lib.lib().vla_train_step(self.core.handle, C.byref(args), _stream()), "vla_train_step")

    def _enqueue(self, ds):
        if self.pg is None or self.dp is not None:
            self._call(ds, 0)                 # (data parallel: the peer-memory exchange is a kernel of the step)
        else:
            from .dp import allreduce_gradients
            self._call(ds, 1)
            allreduce_gradients(self.flat, self.pg)
            self._call(ds, 2)

    def step(self, which=0):
        """One optimizer step on the next resident batch of dataset `which`.  No host synchronisation."""
        core = self.core
        ds = self.datasets[which]
        with torch.cuda.device(core.device):
            if core.shadow_version != core.param_version():
                # parameters were changed from the host side (init, load_state_dict): re-derive the bf16 copies
                _lib.check(_lib.lib().vla_refresh_shadows(core.handle, _ptr(core.arena), _stream()),
                           "vla_refresh_shadows")
                core.shadow_version = core.param_version()
            if not self.use_graph:
                self._enqueue(ds)
            elif which not in self.graphs:
                self._first_step_and_capture(which, ds)
            else:
                self.graphs[which].replay()
        self.steps += 1
        core.generation += 1

    def run_epoch(self, which=0):
        """One pass over dataset `which` in storage order, as `for batch in DataLoader(ds, batch_size)` walks it
        (train_rna2dna.py:82-99, vae_cross_modality_cv.py:121-158): every full batch by the captured step, then the ragged
        last batch (len(ds) % batch rows, the reference's loaders keep it: no drop_last) by one eager call on the same handle.
        Shuffle between epochs with `dataset.shuffle_()`.  Returns the number of optimizer steps taken."""
        ds = self.datasets[which]
        n_full, tail = divmod(len(ds), self.batch)
        self.reset_counters(self.steps, 0)                  # the epoch starts at the dataset's first row
        for _ in range(n_full):
            self.step(which)
        if tail:
            self._step_tail(which, n_full * self.batch, tail)
        return n_full + (1 if tail else 0)

    def _step_tail(self, which, first_row, rows):
        if self.dp is not None or self.pg is not None:
            raise RuntimeError("run_epoch(): ragged last batches are not supported under data parallelism (shards must agree on the step count)")
        if rows < 2:
            raise ValueError("Expected more than 1 value per channel when training (BatchNorm1d): the last batch has one row")
        core = self.core
        ds = self.datasets[which]
        with torch.cuda.device(core.device):
            self._refresh_shadows_if_needed()
            args = self._args(ds, 0)
            args.x_a = C.c_void_p(ds.tpm.data_ptr() + first_row * ds.tpm.shape[1] * 4)
            args.x_b = C.c_void_p(ds.beta.data_ptr() + first_row * ds.beta.shape[1] * 4)
            args.site = C.c_void_p(ds.site.data_ptr() + first_row * 8)
            args.batch = int(rows)
            args.dataset_rows = int(rows)
            _lib.check(_lib.lib().vla_train_step(core.handle, C.byref(args), _stream()), "vla_train_step")
        self.steps += 1
        core.generation += 1

    def _refresh_shadows_if_needed(self):
        core = self.core
        if core.shadow_version != core.param_version():
            _lib.check(_lib.lib().vla_refresh_shadows(core.handle, _ptr(core.arena), _stream()), "vla_refresh_shadows")
            core.shadow_version = core.param_version()

    def forward_backward(self, which=0):
        """First half of a step (vla_train_step phases = 1): forward + loss + backward; the gradients stay in `self.grads`
        (loss.backward() at train_rna2dna.py:95) until `apply_optimizer()`.  Not available with the peer-memory exchange."""
        if self.dp is not None:
            raise RuntimeError("forward_backward(): the peer-memory exchange runs the whole step")
        with torch.cuda.device(self.core.device):
            self._refresh_shadows_if_needed()
            self._call(self.datasets[which], 1)
        self.core.generation += 1

    def apply_optimizer(self, which=0):
        """Second half (phases = 2): fused AdamW on `self.grads`, which it clears (optimizer.step() at train_rna2dna.py:96)."""
        with torch.cuda.device(self.core.device):
            self._call(self.datasets[which], 2)
        self.steps += 1

    def _first_step_and_capture(self, which, ds):
        # The first step runs eagerly on a side stream (it also loads the kernels and sizes the workspace);
        # the same call sequence is then captured, without executing, for every later step.
        dev = self.core.device
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            self._enqueue(ds)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._enqueue(ds)
        self.graphs[which] = g

    def close(self):
        """Release the captured graphs (do this before destroying a process group whose collectives they contain)."""
        torch.cuda.synchronize(self.core.device)
        self.graphs.clear()
        if getattr(self, "_pinned", False):
            _lib.lib().vla_model_pin(self.core.handle, -1)
            self._pinned = False
        if self.dp is not None:
            import torch.distributed as dist
            dist.barrier(group=self.pg)       # no peer may still be reading or writing this rank's buffers
            torch.cuda.synchronize(self.core.device)
            self.flat = self.reduced = self.grads = self.loss_out = self._loss_local = None
            _lib.check(_lib.lib().vla_dp_disconnect(self.dp), "vla_dp_disconnect")   # unmap the peers' buffers ...
            dist.barrier(group=self.pg)       # ... every rank has, so nobody imports this rank's buffer any more ...
            _lib.lib().vla_dp_destroy(self.dp)                                        # ... and it can be freed
            self.dp = None

    def __del__(self):
        try:
            if getattr(self, "_pinned", False) and self.core.handle:
                _lib.lib().vla_model_pin(self.core.handle, -1)
                self._pinned = False
        except Exception:
            pass

    def losses(self):
        """(total, recon, class, kld) of the last completed step -- one 16-byte device->host read.
        Under data parallelism these are the sums over all ranks."""
        return tuple(self.loss_out.tolist())

    def exchange_trace(self):
        """Data parallel, peer-memory exchange: spans (us) of block 0 in the last exchange kernel on this rank, from its
        %globaltimer stamps -- push to the peers, reduce of the own shard (waits for the peers' pushes) and push of the sums."""
        if self.dp is None:
            return None
        buf = (C.c_ulonglong * 8)()
        with torch.cuda.device(self.core.device):
            _lib.check(_lib.lib().vla_dp_trace(self.dp, buf), "vla_dp_trace")
        t = [int(x) for x in buf]
        out = {}
        for part, name in ((1, "decoder_part_side_stream"), (0, "encoder_part_main_stream")):
            o = 4 * part
            out[name] = dict(push_us=(t[o + 1] - t[o]) / 1e3, reduce_us=(t[o + 2] - t[o + 1]) / 1e3,
                             total_us=(t[o + 2] - t[o]) / 1e3)
            if part == 0 and t[o + 3] > t[o + 1]:
                out[name]["own_gradient_loaded_us"] = (t[o + 3] - t[o + 1]) / 1e3
        out["decoder_part_done_before_encoder_part_starts_us"] = (t[0] - t[6]) / 1e3
        return out

    def timeline(self, which=0):
        """Per-phase timing inside the chain launches of ONE replayed step, from the %globaltimer stamps every CTA writes
        (vla_chain_timeline).  Returns a list with one entry per chain launch: dict(name, ctas, span_us, flops, bytes,
        phases=[dict(name, start_us, work_us (mean over CTAs that had work), wait_us (mean time a CTA then waited at the
        cluster barrier), span_us (first start to last end))]).  Per phase, relative to the CTA's phase start and averaged over
        the CTAs that had work: operands_us (first operands in shared memory), mma_issued_us, acc_ready_us, epi_first_us /
        epi_last_us (first / last epilogue warp done), barrier_in_us (thread 0 reaches the cluster barrier), barrier_out_us."""
        import numpy as np
        L = _lib.lib()
        dev = self.core.device
        out = []
        with torch.cuda.device(dev):
            self.step(which)
            torch.cuda.synchronize(dev)
            _lib.check(L.vla_chain_timeline(self.core.handle, 1), "vla_chain_timeline")
            self.step(which)
            torch.cuda.synchronize(dev)
            if which in self.graphs:
                # a graph replay does not pass through the library: ask it which plans an eager step launches
                self._enqueue(self.datasets[which])
                self.steps += 1
                torch.cuda.synchronize(dev)
            n = L.vla_chain_count(self.core.handle)
            for w in range(n):
                name = C.create_string_buffer(48)
                nph, nct, fl, by = C.c_int(), C.c_int(), C.c_double(), C.c_double()
                _lib.check(L.vla_chain_info(self.core.handle, w, name, C.byref(nph), C.byref(nct), C.byref(fl), C.byref(by)), "vla_chain_info")
                buf = (C.c_ulonglong * (nct.value * 24 * 8))()
                _lib.check(min(L.vla_chain_timeline_read(self.core.handle, w, buf), 0), "vla_chain_timeline_read")
                t = np.frombuffer(buf, dtype=np.uint64).reshape(nct.value, 24, 8)[:, :nph.value].astype(np.int64)
                t0 = int(t[:, 0, 0].min())
                phases = []
                for p in range(nph.value):
                    pn = C.create_string_buffer(48)
                    _lib.check(L.vla_chain_phase_name(self.core.handle, w, p, pn), "vla_chain_phase_name")
                    s, e, b = t[:, p, 0], t[:, p, 1], t[:, p, 2]
                    busy = t[:, p, 7] > s                      # CTAs whose epilogue / element-wise warps ran in this phase

                    def rel(slot):                             # mean over the busy CTAs of (stamp - phase start), us
                        v = (t[:, p, slot] - s)[busy & (t[:, p, slot] > s)]
                        return float(v.mean()) / 1e3 if v.size else 0.0
                    phases.append(dict(name=pn.value.decode(), start_us=(int(s.min()) - t0) / 1e3, busy_ctas=int(busy.sum()),
                                       operands_us=rel(3), mma_issued_us=rel(4), acc_ready_us=rel(5), epi_first_us=rel(6),
                                       epi_last_us=rel(7), barrier_in_us=float((e - s).mean()) / 1e3,
                                       barrier_out_us=float((b - s).mean()) / 1e3, span_us=(int(b.max()) - int(s.min())) / 1e3))
                out.append(dict(name=name.value.decode(), ctas=nct.value, span_us=(int(t[:, :, 2].max()) - t0) / 1e3,
                                flops=fl.value, bytes=by.value, phases=phases))
            _lib.check(L.vla_chain_timeline(self.core.handle, 0), "vla_chain_timeline")
        return out

    def profile(self, steps=3, which=0):
        """Per-launch device times inside a replayed CUDA graph: the step is captured once with an event pair around every
        launch (event-record nodes), the graph is replayed `steps` times and the pairs are read after each replay.
        Returns a list of (name, ms, flops, bytes), `steps` entries per launch."""
        L = _lib.lib()
        dev = self.core.device
        ds = self.datasets[which]
        out = []
        buf = (_lib.ProfEntry * 512)()
        with torch.cuda.device(dev):
            torch.cuda.synchronize(dev)
            _lib.check(L.vla_profile_begin(self.core.handle), "vla_profile_begin")
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(ds)
            _lib.check(L.vla_profile_pause(self.core.handle), "vla_profile_pause")
            g.replay()                                   # warm
            torch.cuda.synchronize(dev)
            self.steps += 1
            for _ in range(steps):
                g.replay()
                torch.cuda.synchronize(dev)
                self.steps += 1
                n = L.vla_profile_read(self.core.handle, buf, 512)
                for i in range(max(n, 0)):
                    out.append((buf[i].name.decode(), float(buf[i].ms), float(buf[i].flops), float(buf[i].bytes)))
            L.vla_profile_collect(self.core.handle, buf, 512)   # releases the events
            del g
        return out


class _DeviceInt16:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i2", "data": (int(ptr), False), "version": 3}


def workspace_view(core, what, rows, i=0, j=0):
    """Test hook (vla_test_workspace): zero-copy view of a workspace buffer left by the last forward / train step.
    what = "eps": fp32 [rows, latent]; what = "act": bf16 [rows, row pitch] of encoder i, BatchNorm layer j."""
    ptr, ld = C.c_void_p(), C.c_int()
    code = {"eps": 0, "act": 1}[what]
    _lib.check(_lib.lib().vla_test_workspace(core.handle, code, i, j, C.byref(ptr), C.byref(ld)), "vla_test_workspace")
    n = int(rows) * ld.value
    if what == "eps":
        return _wrap_device_floats(ptr.value, n, core.device).view(rows, ld.value)
    return torch.as_tensor(_DeviceInt16(ptr.value, n), device=core.device).view(torch.bfloat16).view(rows, ld.value)
