"""Data-parallel fused training over NCCL (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("graph", ["1", "0"])
def test_two_rank_dp_matches_summed_shard_oracle(graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DP_GRAPH=graph)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", os.path.join(HERE, "dp_worker.py")], capture_output=True, text=True, env=env,
                         timeout=180)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DP_OK world=2" in res.stdout
