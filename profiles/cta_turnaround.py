"""Per-SM CTA turnaround of the tcgen05 GEMM: gap between one CTA's last stamp and the next CTA's entry on the same SM."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-los-angeles_b200"))
import torch  # noqa: E402
from vla_b200 import _lib  # noqa: E402

L = _lib.lib()


def run(M, N, K, bn, flags=0, label=""):
    L.vla_test_set_flags(flags)
    A = torch.randn(M, (K + 7) // 8 * 8, device="cuda").bfloat16(); B = torch.randn(N, (K + 7) // 8 * 8, device="cuda").bfloat16()
    C = torch.zeros(M, N, device="cuda")
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    dbg = torch.zeros(tiles * 8, dtype=torch.int64, device="cuda")
    for r in range(3):
        L.vla_test_set_timeline(dbg.data_ptr() if r == 2 else None)
        _lib.check(L.vla_test_gemm(0, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), M, N, K, bn, 1, None, None), "gemm")
    torch.cuda.synchronize()
    L.vla_test_set_timeline(None)
    t = dbg.view(tiles, 8).cpu()
    t0 = int(t[:, 0].min())
    per_sm = {}
    for i in range(tiles):
        per_sm.setdefault(int(t[i, 7]), []).append([(int(t[i, k]) - t0) / 1e3 for k in range(7)])
    L.vla_test_set_flags(0)
    print(f"[{label}] M={M} N={N} K={K} bn={bn}: tiles={tiles}, SMs used={len(per_sm)}, kernel span {(int(t[:, :7].max()) - t0) / 1e3:.1f} us")
    sm = sorted(per_sm)[3]
    rows = sorted(per_sm[sm])
    print(f"  SM {sm}: CTAs in order (entry, dep, setup, first_ops, mma_issued, acc_ready, epi_done) us")
    for r in rows[:6]:
        print("   ", " ".join(f"{x:7.2f}" for x in r))
    gaps = []
    for rows in per_sm.values():
        rows = sorted(rows)
        for a, b in zip(rows, rows[1:]):
            gaps.append(b[0] - max(a))
    if gaps:
        gaps.sort()
        print(f"  gap between a CTA's last stamp and the next CTA's entry on the same SM: median {gaps[len(gaps) // 2]:.2f} us, "
              f"p90 {gaps[int(len(gaps) * 0.9)]:.2f} us; CTA lifetime median {sorted(max(r) - r[0] for rs in per_sm.values() for r in rs)[tiles // 2]:.2f} us")


if __name__ == "__main__":
    run(16384, 572, 512, 144, 0, "full kernel")
    run(16384, 572, 512, 144, 8, "no patch/stores")
    run(16384, 572, 512, 144, 8 | 16, "no stores, no TMEM loads")
    run(16384, 572, 512, 144, 1, "no epilogue at all")
    run(16384, 576, 512, 144, 0, "full kernel, N=576 (no partial chunk)")
    run(16384, 512, 512, 128, 0, "full kernel, N=512 bn=128")
