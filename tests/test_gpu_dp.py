"""Data-parallel fused training: the peer-memory gradient exchange (csrc/dp_exchange.cu) and the NCCL variant against the
per-shard oracle with summed gradients.  On a one-GPU box the peer-memory protocol still runs between two processes that
share the GPU; the two-GPU cases are skipped there."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_worker(env_extra, port, world=2):
    env = dict(os.environ, **env_extra)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(HERE, "dp_worker.py")], capture_output=True, text=True, env=env,
                         timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert f"DP_OK world={world}" in res.stdout
    return res.stdout


@pytest.mark.parametrize("world,graph", [(2, "1"), (2, "0"), (5, "1")])
def test_multi_process_p2p_exchange_on_one_gpu(world, graph):
    """world = 5: shards of unequal length (the flat buffer does not divide evenly) and more than two contributions."""
    out = run_worker(dict(DP_GRAPH=graph, DP_SAME_GPU="1", DP_EXCHANGE="p2p"), 29611, world)
    assert "exchange=p2p" in out


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("graph", ["1", "0"])
def test_two_rank_dp_matches_summed_shard_oracle(graph, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = run_worker(dict(DP_GRAPH=graph, DP_EXCHANGE=exchange), 29617)
    assert f"exchange={exchange}" in out
