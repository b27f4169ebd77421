"""tcgen05 GEMM kernel in isolation (through the C ABI test hook) against torch.matmul on the same
bf16-rounded operands.  Tolerance 2e-3 relative: identical products, only the fp32 accumulation order differs."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, A, B, M, N, K, bn, splits, want_bias=False):
    from vla_b200 import _lib
    L = _lib.lib()
    Cout = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    bias = torch.zeros(M, dtype=torch.float32, device="cuda") if want_bias else None
    rc = L.vla_test_gemm(mode, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cout.data_ptr(), M, N, K, bn, splits,
                         None if bias is None else bias.data_ptr(), None)
    _lib.check(rc, "vla_test_gemm")
    torch.cuda.synchronize()
    return Cout, bias


def _padded(rows, cols, seed):
    ld = (cols + 7) // 8 * 8
    g = torch.Generator(device="cuda").manual_seed(seed)
    buf = torch.randn(rows, ld, device="cuda", generator=g).to(torch.bfloat16)
    return buf, buf[:, :cols]


NT_CASES = [(128, 128, 64, 128), (256, 64, 782, 32), (4096, 572, 512, 144), (33, 40, 128, 32), (4096, 128, 782, 32),
            (100, 448, 20, 64), (4096, 20, 448, 32), (515, 160, 200, 144), (4096, 782, 128, 96), (4096, 572, 512, 128)]


@pytest.mark.parametrize("M,N,K,bn", NT_CASES)
def test_gemm_nt(M, N, K, bn):
    Ab, A = _padded(M, K, 1)
    Bb, B = _padded(N, K, 2)
    out, _ = _run(0, Ab, Bb, M, N, K, bn, 1)
    ref = A.float() @ B.float().t()
    err = (out - ref).norm() / ref.norm()
    assert err < 2e-3, float(err)
    assert (out - ref).abs().max() < 2e-3 * ref.abs().max() + 1e-3


TN_CASES = [(128, 128, 64, 128, 1), (572, 512, 4096, 128, 8), (40, 128, 4096, 128, 4), (448, 20, 4096, 64, 8),
            (24, 32, 100, 64, 1), (128, 782, 4096, 128, 7), (512, 257, 1000, 128, 3)]


@pytest.mark.parametrize("M,N,Kb,bn,splits", TN_CASES)
def test_gemm_tn_with_bias_grad(M, N, Kb, bn, splits):
    Gb, G = _padded(Kb, M, 3)
    Xb, X = _padded(Kb, N, 4)
    out, bias = _run(1, Gb, Xb, M, N, Kb, bn, splits, want_bias=True)
    ref = G.float().t() @ X.float()
    err = (out - ref).norm() / ref.norm()
    assert err < 2e-3, float(err)
    bref = G.float().sum(0)
    berr = (bias - bref).norm() / bref.norm()
    assert berr < 2e-3, float(berr)


NN_CASES = [(128, 64, 64, 64), (4096, 128, 40, 128), (4096, 512, 572, 128), (4096, 20, 448, 64), (300, 256, 512, 128),
            (4096, 32, 40, 64)]


@pytest.mark.parametrize("M,N,K,bn", NN_CASES)
def test_gemm_nn_mixed_major(M, N, K, bn):
    """C[M,N] = A[M,K] * B[K,N]: A K-major, B MN-major (data gradients read the forward weight copy)."""
    Ab, A = _padded(M, K, 5)
    Bb, B = _padded(K, N, 6)
    out, _ = _run(2, Ab, Bb, M, N, K, bn, 1)
    ref = A.float() @ B.float()
    err = (out - ref).norm() / ref.norm()
    assert err < 2e-3, float(err)


# Launches of more than 1.5 waves with short main loops take the PERSISTENT form (gemm_tc_persist_kernel: a CTA walks a
# contiguous range of tiles, two TMEM accumulators, two-stage ring); same arithmetic, same tolerance.  Ragged M / N / K on purpose.
@pytest.mark.parametrize("M,N,K,bn", [(40000, 572, 512, 144), (33001, 160, 200, 32), (70000, 40, 128, 64), (20011, 782, 128, 96)])
def test_gemm_nt_persistent_form(M, N, K, bn):
    Ab, A = _padded(M, K, 11)
    Bb, B = _padded(N, K, 12)
    out, _ = _run(0, Ab, Bb, M, N, K, bn, 1)
    ref = A.float() @ B.float().t()
    assert (out - ref).norm() / ref.norm() < 2e-3
    assert (out - ref).abs().max() < 2e-3 * ref.abs().max() + 1e-3


@pytest.mark.parametrize("M,N,K,bn", [(40000, 256, 512, 128), (25003, 128, 40, 64)])
def test_gemm_nn_persistent_form(M, N, K, bn):
    Ab, A = _padded(M, K, 13)
    Bb, B = _padded(K, N, 14)
    out, _ = _run(2, Ab, Bb, M, N, K, bn, 1)
    ref = A.float() @ B.float()
    assert (out - ref).norm() / ref.norm() < 2e-3


def test_gemm_tn_persistent_form():
    """Weight-gradient tiles with one k-block per split: 7 x 8 x 8 = 448 tiles, red.add partial sums."""
    M, N, Kb, bn, splits = 782, 512, 512, 64, 8
    Gb, G = _padded(Kb, M, 15)
    Xb, X = _padded(Kb, N, 16)
    out, bias = _run(1, Gb, Xb, M, N, Kb, bn, splits, want_bias=True)
    ref = G.float().t() @ X.float()
    assert (out - ref).norm() / ref.norm() < 2e-3
    bref = G.float().sum(0)
    assert (bias - bref).norm() / bref.norm() < 2e-3
