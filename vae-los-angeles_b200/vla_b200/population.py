"""Populations of independent models on one GPU (BASELINE configs[4]).

The reference trains its hyper-parameter trials (`optimize_hyperparameters.py:68-133`) and cross-validation folds
(`vae_cross_modality_cv.py:113-196, 198-283`) one after the other.  The models are independent -- no data-path
exchange, "replicas only" -- and each one's train step is a chain of short dependent launches that leaves most SMs idle
most of the time.  `Population` steps its members in LOCK-STEP (`grouped=True`, the default): launch j of
the train step is issued once for every live member (`vla_train_step_group`: grouped tcgen05 GEMMs whose grid holds all
members' tiles, the element-wise launches likewise), captured as one CUDA graph per set of live members -- a population's
step costs the launch latency of one model's.  `grouped=False` keeps the earlier scheme: one fused `Trainer` graph per model
on its own stream, stepped round-robin, overlapping on the device.  Across GPUs the population is sharded by index
(`models[rank::world]`), one process per GPU, no collective.

The per-epoch control flow of the reference loops is restated on the host, per model, from ONE device->host read per epoch
for the whole population: beta warm-up (`train_rna2dna.py:80`), `ReduceLROnPlateau` (torch's own scheduler object on a
dummy optimizer, so the semantics are exactly the reference's, `vae_cross_modality_cv.py:127`), early stopping with a
device-side snapshot of the best state (`vae_cross_modality_cv.py:129-196`).
"""
import ctypes as C

import torch

from . import _lib
from .engine import Trainer, _stream

GROUP_MAX = 256      # MULTI_MAX_MEMBERS of the library: members per merged launch


class Member:
    """One model of the population with its trainer, stream and the host-side training-control state."""

    def __init__(self, model, trainer, stream, hyper):
        self.model, self.trainer, self.stream, self.hyper = model, trainer, stream, hyper
        self.best_val = float("inf")
        self.best_state = None
        self.bad_epochs = 0
        self.stopped = False
        self.history = []
        self._dummy = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=hyper["lr"])
        self.scheduler = None


class Population:
    """`specs`: list of dicts with `model` (an un-trained drop-in module on the CPU or the device) and optional `lr`,
    `weight_decay`, `beta_start`, `gamma`, `seed`, `class_weights`; `datasets`: one DeviceDataset shared by all members or
    one per member (folds); every member trains at `batch_size`."""

    def __init__(self, specs, datasets, batch_size, device="cuda", beta_warmup_epochs=50, lr_factor=0.5, lr_patience=5,
                 patience=15, use_graph=True, grouped=True):
        self.device = torch.device(device)
        self.grouped = bool(grouped)
        self.use_graph = use_graph
        self._group_graphs = {}
        self.batch = int(batch_size)
        self.beta_warmup_epochs = beta_warmup_epochs
        self.patience = patience
        self.members = []
        shared = not isinstance(datasets, (list, tuple))
        main = torch.cuda.current_stream(self.device)
        for i, spec in enumerate(specs):
            hyper = dict(lr=spec.get("lr", 5e-4), weight_decay=spec.get("weight_decay", 1e-5),
                         beta_start=spec.get("beta_start", 1e-3), gamma=spec.get("gamma", 1.0), seed=spec.get("seed", i))
            model = spec["model"].to(self.device).train()
            ds = datasets if shared else datasets[i]
            stream = main if self.grouped else torch.cuda.Stream(device=self.device)
            if not self.grouped:
                stream.wait_stream(main)
            with torch.cuda.stream(stream):
                tr = Trainer(model, ds, self.batch, lr=hyper["lr"], weight_decay=hyper["weight_decay"],
                             beta_kl=hyper["beta_start"], gamma=hyper["gamma"], class_weights=spec.get("class_weights"),
                             seed=hyper["seed"], use_graph=use_graph)
            mem = Member(model, tr, stream, hyper)
            mem.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(mem._dummy, mode="min", factor=lr_factor,
                                                                        patience=lr_patience)
            self.members.append(mem)

    def __len__(self):
        return len(self.members)

    # -- lock-step stepping (vla_train_step_group) -----------------------------------------------------------------------
    def _group_call(self, members, args_list):
        L = _lib.lib()
        for lo in range(0, len(members), GROUP_MAX):
            ms, ar = members[lo:lo + GROUP_MAX], args_list[lo:lo + GROUP_MAX]
            n = len(ms)
            handles = (C.c_void_p * n)(*[getattr(m.trainer.core.handle, "value", m.trainer.core.handle) for m in ms])
            ptrs = (C.POINTER(_lib.TrainArgs) * n)(*[C.pointer(a) for a in ar])
            _lib.check(L.vla_train_step_group(handles, ptrs, n, _stream()), "vla_train_step_group")

    def _group_step(self, members, tails=None):
        """One optimizer step of every member of `members` as merged launches.  tails: per member (first_row, rows) of a
        ragged last batch (eager call), else the next resident batch of each member's dataset (graph replay)."""
        if not members:
            return
        with torch.cuda.device(self.device):
            for mem in members:
                mem.trainer._refresh_shadows_if_needed()
            if tails is not None:
                args = []
                for mem, (first_row, rows) in zip(members, tails):
                    if rows < 2:
                        raise ValueError("Expected more than 1 value per channel when training (BatchNorm1d): the last batch has one row")
                    tr = mem.trainer
                    ds = tr.datasets[0]
                    a = tr._args(ds, 0)
                    a.x_a = C.c_void_p(ds.tpm.data_ptr() + first_row * ds.tpm.shape[1] * 4)
                    a.x_b = C.c_void_p(ds.beta.data_ptr() + first_row * ds.beta.shape[1] * 4)
                    a.site = C.c_void_p(ds.site.data_ptr() + first_row * 8)
                    a.batch = int(rows)
                    a.dataset_rows = int(rows)
                    args.append(a)
                self._group_call(members, args)
            else:
                key = tuple(id(m) for m in members)
                if not self.use_graph:
                    self._group_call(members, [m.trainer._args(m.trainer.datasets[0], 0) for m in members])
                elif key in self._group_graphs:
                    self._group_graphs[key].replay()
                else:
                    # first step eagerly on a side stream (builds and caches the merged launch tables), then the capture
                    s = torch.cuda.Stream(device=self.device)
                    s.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(s):
                        self._group_call(members, [m.trainer._args(m.trainer.datasets[0], 0) for m in members])
                    torch.cuda.current_stream(self.device).wait_stream(s)
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._group_call(members, [m.trainer._args(m.trainer.datasets[0], 0) for m in members])
                    self._group_graphs[key] = g
        for mem in members:
            mem.trainer.steps += 1
            mem.trainer.core.generation += 1

    # -- stepping ----------------------------------------------------------------------------------------------------
    def step(self, n=1):
        """`n` optimizer steps of every active member (lock-step merged launches, or round-robin over the members'
        streams with grouped=False); no host synchronisation."""
        for _ in range(n):
            if self.grouped:
                self._group_step([m for m in self.members if not m.stopped])
                continue
            for mem in self.members:
                if not mem.stopped:
                    with torch.cuda.stream(mem.stream):
                        mem.trainer.step()

    def run_epoch(self):
        """One pass of every active member over its own dataset, as the reference's per-model `for batch in train_loader`
        (vae_cross_modality_cv.py:121-158, optimize_hyperparameters.py:104-113): full batches round-robin over the members'
        streams, then each member's ragged last batch (the loaders keep it).  No host synchronisation."""
        live = [m for m in self.members if not m.stopped]
        plan = []
        for mem in live:
            n_full, tail = divmod(len(mem.trainer.datasets[0]), self.batch)
            plan.append((mem, n_full, tail))
            with torch.cuda.stream(mem.stream):
                mem.trainer.reset_counters(mem.trainer.steps, 0)
        if self.grouped:
            for i in range(max((n for _, n, _ in plan), default=0)):
                self._group_step([mem for mem, n_full, _ in plan if i < n_full])
            with_tail = [(mem, n_full, tail) for mem, n_full, tail in plan if tail]
            self._group_step([mem for mem, _, _ in with_tail], tails=[(n_full * self.batch, tail) for _, n_full, tail in with_tail])
            return
        for i in range(max((n for _, n, _ in plan), default=0)):
            for mem, n_full, _ in plan:
                if i < n_full:
                    with torch.cuda.stream(mem.stream):
                        mem.trainer.step()
        for mem, n_full, tail in plan:
            if tail:
                with torch.cuda.stream(mem.stream):
                    mem.trainer._step_tail(0, n_full * self.batch, tail)

    def synchronize(self):
        main = torch.cuda.current_stream(self.device)
        for mem in self.members:
            main.wait_stream(mem.stream)
        torch.cuda.synchronize(self.device)

    def losses(self):
        """[(total, recon, class, kld)] of every member's last step (synchronises once)."""
        self.synchronize()
        stacked = torch.stack([mem.trainer.loss_out for mem in self.members]).tolist()
        return [tuple(x) for x in stacked]

    # -- the reference's per-epoch control flow ------------------------------------------------------------------------
    def begin_epoch(self, epoch):
        """beta warm-up (train_rna2dna.py:80): beta = min(1, epoch / warmup) * beta_start, per member."""
        w = min(1.0, epoch / self.beta_warmup_epochs) if self.beta_warmup_epochs > 0 else 1.0
        for mem in self.members:
            if not mem.stopped:
                with torch.cuda.stream(mem.stream):
                    mem.trainer.set_hyper(beta_kl=w * mem.hyper["beta_start"])

    def end_epoch(self, val_losses):
        """`val_losses`: one validation loss per member (floats).  Steps every member's ReduceLROnPlateau, keeps a device
        snapshot of the best state and applies early stopping (vae_cross_modality_cv.py:177-190).  Returns the number of
        members still training."""
        for mem, v in zip(self.members, val_losses):
            if mem.stopped:
                continue
            v = float(v)
            mem.history.append(v)
            mem.scheduler.step(v)
            lr = mem._dummy.param_groups[0]["lr"]
            with torch.cuda.stream(mem.stream):
                mem.trainer.set_hyper(lr=lr)
                if v < mem.best_val:
                    mem.best_val, mem.bad_epochs = v, 0
                    core = mem.trainer.core
                    mem.best_state = (core.arena.clone(), core.buffers.clone(), core.counters.clone())
                else:
                    mem.bad_epochs += 1
                    if mem.bad_epochs >= self.patience:
                        mem.stopped = True
        return sum(not m.stopped for m in self.members)

    def restore_best(self):
        """Load every member's best snapshot back (vae_cross_modality_cv.py:192-194)."""
        for mem in self.members:
            if mem.best_state is not None:
                with torch.cuda.stream(mem.stream), torch.no_grad():
                    core = mem.trainer.core
                    core.arena.copy_(mem.best_state[0])
                    core.buffers.copy_(mem.best_state[1])
                    core.counters.copy_(mem.best_state[2])
                    core.shadow_version = None     # the bf16 operand copies are re-derived before the next step
        self.synchronize()

    def validate(self, batches, loss_fn):
        """Validation loss of every member on `batches` (a list of input tuples already on the device) in eval mode, the
        loops at optimize_hyperparameters.py:113-125 / vae_cross_modality_cv.py:160-173.  `loss_fn(member, model, batch)`
        returns the loss tensor of one batch; the per-member sums are read back with one synchronisation."""
        for batch in batches:
            if batch[0].shape[0] > self.batch:
                raise ValueError("validation batches must not exceed the training batch size (the captured train-step graphs "
                                 "hold pointers into a workspace sized for it)")
        sums = []
        for mem in self.members:
            with torch.cuda.stream(mem.stream), torch.no_grad():
                mem.model.eval()
                tot = torch.zeros((), device=self.device)
                for batch in batches:
                    tot = tot + loss_fn(mem, mem.model, batch).detach()
                mem.model.train()
                sums.append(tot / max(len(batches), 1))
        self.synchronize()
        return torch.stack(sums).tolist()

    def close(self):
        self.synchronize()
        self._group_graphs.clear()
        for mem in self.members:
            mem.trainer.close()


def shard(items, rank, world):
    """Population members of process `rank` out of `world` (one process per GPU; no communication between shards)."""
    return list(items)[rank::world]
