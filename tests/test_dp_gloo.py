"""Host-side data-parallel logic on CPU, world_size 2 over gloo: gradient packing into the library's flat arena layout,
the single all-reduce(SUM) of [gradients | loss scalars], and identical AdamW results on every rank.
The per-shard arithmetic is the oracle's; the DP oracle is 'reference run once per shard, gradients summed'
(BatchNorm uses per-shard statistics)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vae_oracle as vo

DIMS = dict(A=50, B=36, S=5, L=10, E=16)
KIND = "multimodal"
N = 24


def _shard_grads(rank, world):
    state = vo.init_state(KIND, DIMS, seed=3)
    st = {k: (v.astype(np.float64) if v.dtype.kind == "f" else v.copy()) for k, v in state.items()}
    tpm, beta, site = vo.synthetic_batch(N, DIMS, seed=3)
    eps, masks = vo.synthetic_noise(N, DIMS, KIND, seed=3)
    lo, hi = rank * N // world, (rank + 1) * N // world
    batch = dict(a=tpm[lo:hi].astype(np.float64), b=beta[lo:hi].astype(np.float64), site=site[lo:hi])
    m = {k: v[lo:hi] for k, v in masks.items()}
    out, cache = vo.forward(KIND, DIMS, st, batch, eps[lo:hi].astype(np.float64), m, train=True)
    scal, og = vo.loss_and_output_grads(KIND, out, batch, 1e-3, 1.0, None)
    grads = vo.backward(KIND, DIMS, st, cache, og, train=True)
    return st, grads, scal


def _worker(rank, world, port, ret):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "vae-los-angeles_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from vla_b200 import Layout, allreduce_gradients
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = Layout(KIND, DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"], DIMS["E"])
        st, grads, scal = _shard_grads(rank, world)
        flat = lay.pack(grads, extra=4, dtype=torch.float64)
        flat[lay.n_params:] = torch.tensor([scal["total"], scal["recon"], scal["cls"], scal["kld"]], dtype=torch.float64)
        allreduce_gradients(flat)
        summed = {k: v.numpy().copy() for k, v in lay.unpack(flat[: lay.n_params]).items()}
        opt, step = vo.adamw_init(st)
        vo.adamw_step(st, {k: summed[k] for k in grads}, opt, step)
        ret[rank] = dict(flat=flat.numpy().copy(), params={k: st[k].copy() for k in grads})
    finally:
        dist.destroy_process_group()


def test_dp_allreduce_sum_matches_summed_shard_gradients():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    r0, r1 = ret[0], ret[1]
    np.testing.assert_array_equal(r0["flat"], r1["flat"])                      # every rank holds the same reduced buffer
    for k in r0["params"]:
        np.testing.assert_array_equal(r0["params"][k], r1["params"][k])        # -> identical replicas after AdamW
    # reference: per-shard oracle gradients summed in this process
    from vla_b200 import Layout
    lay = Layout(KIND, DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"], DIMS["E"])
    total = None
    loss = np.zeros(4)
    for rank in range(world):
        _, grads, scal = _shard_grads(rank, world)
        flat = lay.pack(grads, dtype=torch.float64).numpy()
        total = flat if total is None else total + flat
        loss += [scal["total"], scal["recon"], scal["cls"], scal["kld"]]
    np.testing.assert_allclose(r0["flat"][: lay.n_params], total, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(r0["flat"][lay.n_params:], loss, rtol=1e-12)
    # SUM (not mean) reproduces the single-process gradient on the concatenated batch for everything downstream of the
    # BatchNorm layers' statistics, e.g. the decoder biases given the same activations: checked here through the loss sums
    _, _, scal_full = _shard_grads(0, 1)
    assert abs(loss[3] - scal_full["kld"]) / abs(scal_full["kld"]) < 0.2       # same order; BN statistics differ per shard
