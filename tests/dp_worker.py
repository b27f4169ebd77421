"""torchrun worker of tests/test_gpu_dp.py: 2+ ranks, fused DP train steps vs the per-shard oracle with summed gradients.

DP_EXCHANGE = p2p (the library's peer-memory kernel, default) | nccl (torch.distributed.all_reduce).
One GPU per rank, always: kernels of different ranks wait on one another's stores, and nothing guarantees that two
processes sharing one GPU run at the same time (B200_PROFILING.md: such runs raised Xid 109).  The index arithmetic of
the protocol is covered on the CPU by tests/test_dp_protocol_model.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-los-angeles_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oracle import vae_oracle as vo  # noqa: E402
from parity_util import MATCHED_Q, is_pre_bn_bias, make_module, rel_l2, to_t  # noqa: E402
from vla_b200 import DeviceDataset, Trainer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    exchange = os.environ.get("DP_EXCHANGE", "p2p")
    if torch.cuda.device_count() < world:
        raise SystemExit("dp_worker needs one GPU per rank")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sync_bn = os.environ.get("DP_SYNCBN", "0") == "1"
    kind, dims, per_rank, steps = os.environ.get("DP_KIND", "rna2dna"), dict(A=782, B=572, S=24, L=20, E=32), 64, 3
    state = vo.init_state(kind, dims, seed=31)
    tpm, beta, site = vo.synthetic_batch(per_rank * world, dims, seed=31)
    eps, masks = vo.synthetic_noise(per_rank * world, dims, kind, seed=31)
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    m = make_module(kind, dims, state, device=f"cuda:{local}").train()
    ds = DeviceDataset(tpm[sl], beta[sl], site[sl], f"cuda:{local}")
    tr = Trainer(m, ds, per_rank, beta_kl=1e-3, process_group=dist.group.WORLD, use_graph=os.environ.get("DP_GRAPH", "1") == "1",
                 exchange=exchange, sync_bn=sync_bn and os.environ.get("DP_SYNCBN_NEGATIVE_CONTROL") != "1")
    tr.injected = dict(eps=to_t(eps[sl], f"cuda:{local}"), keep_masks=[to_t(v[sl], f"cuda:{local}") for v in masks.values()])
    losses = []
    for _ in range(steps):
        tr.step()
        losses.append(tr.losses())
    torch.cuda.synchronize()
    # replicas must stay bit-identical
    flat = tr.core.arena.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(flat, ref), "replicas diverged"
    if rank == 0 and sync_bn:
        # SyncBN oracle: the reference on the CONCATENATED batch (global BatchNorm statistics), one process
        st = {k: (v.astype(np.float64) if v.dtype.kind == "f" else v.copy()) for k, v in state.items()}
        opt, step = vo.adamw_init(st)
        ref_losses = []
        full = dict(a=tpm.astype(np.float64), b=beta.astype(np.float64), site=site)
        for _ in range(steps):
            scal, _, _, step = vo.train_step(kind, dims, st, opt, step, full, eps.astype(np.float64), masks, beta=1e-3, gamma=1.0,
                                             q=MATCHED_Q)
            ref_losses.append([scal["total"], scal["recon"], scal["cls"], scal["kld"]])
        got = np.array(losses)
        np.testing.assert_allclose(got[:, 0], np.array(ref_losses)[:, 0], rtol=2e-2)
        np.testing.assert_allclose(got[:, 3], np.array(ref_losses)[:, 3], rtol=2e-2)
        sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        worst = 0.0
        for name, refv in st.items():
            if name.endswith("num_batches_tracked"):
                assert int(sd[name]) == steps, name
                continue
            if vo.is_buffer(name):        # running statistics of the GLOBAL batch on every rank
                np.testing.assert_allclose(sd[name], refv, rtol=2e-2, atol=4 * 5e-4 * steps * np.sqrt(refv.size), err_msg=name)
                continue
            if is_pre_bn_bias(name):
                continue
            d_ref = refv - state[name].astype(np.float64)
            d_got = sd[name].astype(np.float64) - state[name].astype(np.float64)
            worst = max(worst, rel_l2(d_got, d_ref))
        assert worst <= 0.2, worst
        print(f"DP_OK world={world} exchange={exchange} syncbn kind={kind} worst displacement rel err {worst:.3f} losses {got[-1].tolist()}", flush=True)
    elif rank == 0:
        # oracle: every shard forward/backward separately (per-shard BatchNorm statistics), gradients summed, one AdamW
        st = {k: (v.astype(np.float64) if v.dtype.kind == "f" else v.copy()) for k, v in state.items()}
        opt, step = vo.adamw_init(st)
        ref_losses = []
        for _ in range(steps):
            total, tot_loss = None, np.zeros(4)
            for r in range(world):
                s2 = slice(r * per_rank, (r + 1) * per_rank)
                stc = {k: v.copy() for k, v in st.items()}
                batch = dict(a=tpm[s2].astype(np.float64), b=beta[s2].astype(np.float64), site=site[s2])
                out, cache = vo.forward(kind, dims, stc, dict(a=batch["a"], site=batch["site"]), eps[s2].astype(np.float64),
                                        {k: v[s2] for k, v in masks.items()}, train=True, q=MATCHED_Q)
                scal, og = vo.loss_and_output_grads(kind, out, batch, 1e-3, 1.0, None)
                g = vo.backward(kind, dims, stc, cache, og, train=True)
                total = g if total is None else {k: total[k] + g[k] for k in g}
                tot_loss += [scal["total"], scal["recon"], scal["cls"], scal["kld"]]
                if r == 0:
                    bn_state = {k: stc[k] for k in stc if vo.is_buffer(k)}      # rank 0's running statistics
            st.update(bn_state)
            step = vo.adamw_step(st, total, opt, step)
            ref_losses.append(tot_loss)
        got = np.array(losses)
        np.testing.assert_allclose(got[:, 0], np.array(ref_losses)[:, 0], rtol=2e-2)
        sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        worst = 0.0
        for name, refv in st.items():
            if vo.is_buffer(name) or is_pre_bn_bias(name):
                continue
            d_ref = refv - state[name].astype(np.float64)
            d_got = sd[name].astype(np.float64) - state[name].astype(np.float64)
            worst = max(worst, rel_l2(d_got, d_ref))
        assert worst <= 0.2, worst
        print(f"DP_OK world={world} exchange={exchange} worst displacement rel err {worst:.3f} losses {got[-1].tolist()}", flush=True)
    import threading
    threading.Timer(20.0, lambda: os._exit(0)).start()       # never hang at exit: the verdict is already printed
    tr.close()
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
