"""Data-parallel fused training: the peer-memory gradient exchange (csrc/dp_exchange.cu) and the NCCL variant against the
per-shard oracle with summed gradients.  One GPU per rank (skipped on a one-GPU box: ranks whose kernels wait on one another
must never share a GPU); the protocol's index arithmetic is tested on the CPU in tests/test_dp_protocol_model.py."""
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


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("graph", ["1", "0"])
def test_two_rank_dp_matches_summed_shard_oracle(graph, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = run_worker(dict(DP_GRAPH=graph, DP_EXCHANGE=exchange), 29617)
    assert f"exchange={exchange}" in out


@pytest.mark.parametrize("kind", ["rna2dna", "dna2rna"])
def test_two_rank_sync_bn_equals_reference_on_concatenated_batch(kind):
    """Opt-in SyncBN: with the BatchNorm column sums all-reduced forward and backward, the 2-rank step
    equals the reference run ONCE on the concatenated batch (encoders.py:14,32,36 see the global statistics)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = run_worker(dict(DP_GRAPH="1", DP_EXCHANGE="p2p", DP_SYNCBN="1", DP_KIND=kind), 29631)
    assert f"syncbn kind={kind}" in out
