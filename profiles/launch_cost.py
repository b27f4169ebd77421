"""Steady-state cost per launch inside a CUDA graph (20 identical launches per graph, 20 replays)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-los-angeles_b200"))
import torch  # noqa: E402
from vla_b200 import _lib  # noqa: E402

L = _lib.lib()


def graph_time(fn, n_in_graph=20, replays=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n_in_graph):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n_in_graph * replays)


def gemm(mode, M, N, K, bn, splits=1):
    if mode == 0:
        A = torch.randn(M, (K + 7) // 8 * 8, device="cuda").bfloat16(); B = torch.randn(N, (K + 7) // 8 * 8, device="cuda").bfloat16()
    elif mode == 2:
        A = torch.randn(M, (K + 7) // 8 * 8, device="cuda").bfloat16(); B = torch.randn(K, (N + 7) // 8 * 8, device="cuda").bfloat16()
    else:
        A = torch.randn(K, (M + 7) // 8 * 8, device="cuda").bfloat16(); B = torch.randn(K, (N + 7) // 8 * 8, device="cuda").bfloat16()
    Cc = torch.zeros(M, N, device="cuda")
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lambda: _lib.check(L.vla_test_gemm(mode, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cc.data_ptr(), M, N, K, bn, splits, None, st()), "g")
    t = graph_time(fn)
    fl = 2.0 * M * N * K
    print(f"gemm mode {mode} M={M} N={N} K={K} bn={bn} splits={splits}: {t:7.2f} us/launch  {fl / t / 1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    x = torch.zeros(1 << 20, device="cuda")
    print(f"torch x.add_(1) on 4 MB: {graph_time(lambda: x.add_(1.0)):.2f} us/launch")
    y = torch.zeros(64, device="cuda")
    print(f"torch tiny add_: {graph_time(lambda: y.add_(1.0)):.2f} us/launch")
    gemm(0, 128, 32, 64, 32)          # one CTA: fixed cost
    gemm(0, 4096, 32, 64, 32)         # 32 CTAs, trivial work
    gemm(0, 4096, 512, 256, 128)
    gemm(0, 4096, 572, 512, 144)
    gemm(0, 4096, 128, 782, 32)
    gemm(0, 4096, 128, 782, 128)
    gemm(2, 4096, 512, 572, 128)
    gemm(1, 572, 512, 4096, 128, 7)
    gemm(0, 16384, 572, 512, 144)
    gemm(0, 65536, 572, 512, 144)
    a = torch.randn(8192, 8192, device="cuda").bfloat16()
    print(f"torch bf16 matmul 4096x512x572-ish: {graph_time(lambda: torch.matmul(a[:4096, :512], a[:512, :576])):.2f} us/launch")
