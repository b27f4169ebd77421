"""In-tree build of libvla_b200.so: `nvcc -gencode arch=compute_100a,code=sm_100a` only (no other arch, no JIT)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
SOURCES = ["gemm_tc.cu", "elementwise.cu", "chain_kernel.cu", "rowchain.cu", "headblock.cu", "dp_exchange.cu", "metrics.cu", "vla_api.cu"]
OUT = os.path.join(HERE, "libvla_b200.so")


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in ("tc_ptx.cuh", "vla_internal.h", "gemm_tile.cuh", "elementwise_dev.cuh", "dp_frame.cuh", "loss_math.cuh")] + \
        [os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "vla_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "--threads", "0", "-o", OUT] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libvla_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
