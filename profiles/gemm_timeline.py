"""In-kernel timeline of the tcgen05 GEMM (globaltimer stamps per CTA) for the step's layer shapes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-los-angeles_b200"))
import torch  # noqa: E402
from vla_b200 import _lib  # noqa: E402

L = _lib.lib()
NAMES = ["entry", "dep", "setup", "first_ops", "mma_issued", "acc_ready", "epi_done"]


def run(mode, M, N, K, bn, splits=1, reps=5):
    if mode == 0:
        A = torch.randn(M, (K + 7) // 8 * 8, device="cuda").bfloat16(); B = torch.randn(N, (K + 7) // 8 * 8, device="cuda").bfloat16()
    elif mode == 2:
        A = torch.randn(M, (K + 7) // 8 * 8, device="cuda").bfloat16(); B = torch.randn(K, (N + 7) // 8 * 8, device="cuda").bfloat16()
    else:
        A = torch.randn(K, (M + 7) // 8 * 8, device="cuda").bfloat16(); B = torch.randn(K, (N + 7) // 8 * 8, device="cuda").bfloat16()
    C = torch.zeros(M, N, device="cuda")
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn) * splits
    dbg = torch.zeros(tiles * 8, dtype=torch.int64, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for r in range(reps):
        L.vla_test_set_timeline(dbg.data_ptr() if r == reps - 1 else None)
        if r == reps - 1:
            ev0.record()
        _lib.check(L.vla_test_gemm(mode, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), M, N, K, bn, splits, None, None), "gemm")
        if r == reps - 1:
            ev1.record()
    torch.cuda.synchronize()
    L.vla_test_set_timeline(None)
    t = dbg.view(tiles, 8).cpu().double()
    t0 = t[:, 0].min()
    rel = (t[:, :7] - t0) / 1e3
    print(f"mode {mode} M={M} N={N} K={K} bn={bn} splits={splits} tiles={tiles}  event time {ev0.elapsed_time(ev1) * 1e3:.1f} us")
    print("   stamp        min      median      max   (us since first CTA entry)")
    for i, n in enumerate(NAMES):
        col = rel[:, i]
        print(f"   {n:10s} {col.min():8.2f} {col.median():10.2f} {col.max():8.2f}")


if __name__ == "__main__":
    run(0, 4096, 512, 256, 128)
    run(0, 4096, 572, 512, 144)
    run(0, 4096, 128, 782, 32)
    run(0, 4096, 40, 128, 64)
    run(2, 4096, 512, 572, 128)
    run(1, 572, 512, 4096, 128, 7)
