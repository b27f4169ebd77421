"""The stock-PyTorch restatement bench.py times on the GPU (oracle/torch_eager.py) against the numpy oracle -- which is
itself pinned by the reference fixtures (tests/test_oracle_golden.py) -- on CPU in fp64: same losses, same gradients."""
import numpy as np
import pytest
import torch

from oracle import torch_eager as te
from oracle import vae_oracle as vo

DIMS = dict(A=30, B=22, S=5, L=6, E=32)


@pytest.mark.parametrize("kind", ["rna2dna", "dna2rna", "multimodal", "rna2dna_ae"])
def test_eager_restatement_matches_oracle(kind):
    n = 12
    state = vo.init_state(kind, DIMS, seed=2, dtype=np.float64)
    tpm, beta, site = vo.synthetic_batch(n, DIMS, seed=2, dtype=np.float64)
    eps, masks = vo.synthetic_noise(n, DIMS, kind, seed=2, dtype=np.float64)
    batch = dict(a=tpm, b=beta, site=site)
    spec = vo.MODEL_KINDS[kind]
    enc = {vo.INPUT_OF[t]: batch[vo.INPUT_OF[t]] for _, t in spec["encoders"]}
    st = {k: v.copy() for k, v in state.items()}
    out, cache = vo.forward(kind, DIMS, st, enc, eps, masks, train=True)
    scal, og = vo.loss_and_output_grads(kind, out, batch, beta=2e-3, gamma=1.5)
    grads = vo.backward(kind, DIMS, st, cache, og, train=True)

    p, buf = te.params_from_state(state, "cpu", torch.float64)
    tb = {k: torch.as_tensor(v) for k, v in batch.items()}
    tenc = {k: tb[k] for k in enc}
    tout = te.forward(kind, p, buf, tenc, train=True, eps=torch.as_tensor(eps), masks={k: torch.as_tensor(v) for k, v in masks.items()})
    total, tscal = te.loss(kind, tout, tb, beta=2e-3, gamma=1.5)
    total.backward()
    assert abs(float(total) - scal["total"]) <= 1e-9 * abs(scal["total"])
    for k, g in grads.items():
        tg = p[k].grad
        ref = np.asarray(g)
        if tg is None:
            assert np.abs(ref).max() == 0.0, k
            continue
        assert np.abs(tg.numpy() - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max()), k
    for k in buf:
        if k.endswith("running_mean") or k.endswith("running_var"):
            np.testing.assert_allclose(buf[k].numpy(), st[k], rtol=1e-10, atol=1e-12)
