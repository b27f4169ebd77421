"""Static checks of the built library (CPU only, `cuobjdump -sass` on libvla_b200.so): the GEMM kernels really are
tcgen05 / TMA / TMEM code, the library is built for sm_100a only, and no kernel outgrows the instruction-cache budget that
DESIGN.md section 4 found to matter (83 KB of SASS in the GEMM epilogue cost 30-40 % per launch, 190 KB in AdamW likewise)."""
import collections
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    from vla_b200 import _lib
    try:
        res = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300)
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    assert res.returncode == 0, res.stderr[-500:]
    kernels, cur = collections.OrderedDict(), None
    archs = set(re.findall(r"arch = (sm_\w+)", res.stdout))
    for line in res.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur].append(m.group(1))
    return archs, kernels


def _find(kernels, needle):
    return {k: v for k, v in kernels.items() if needle in k}


def test_built_for_sm_100a_only(sass):
    archs, _ = sass
    assert archs == {"sm_100a"}, archs


def test_gemm_kernels_use_tcgen05_tma_tmem(sass):
    _, kernels = sass
    gemms = _find(kernels, "gemm_tc_kernel")
    assert len(gemms) >= 5
    for name, ops in gemms.items():
        have = collections.Counter(op.split(".")[0] for op in ops)
        assert have["UTCHMMA"] > 0, name          # tcgen05.mma
        assert have["UTMALDG"] > 0, name          # cp.async.bulk.tensor (TMA)
        assert have["LDTM"] > 0, name             # tcgen05.ld (TMEM -> registers)
        assert have["UTCBAR"] > 0, name           # tcgen05.commit -> mbarrier
        assert not any(op.startswith(("HMMA", "WGMMA")) for op in ops), name      # no mma.sync / wgmma path
    # the lock-step population form and the persistent form run the same tile body
    for needle, at_least in (("gemm_tc_multi_kernel", 5), ("gemm_tc_persist_kernel", 4)):
        more = _find(kernels, needle)
        assert len(more) >= at_least, (needle, len(more))
        for name, ops in more.items():
            have = {op.split(".")[0] for op in ops}
            assert {"UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"} <= have, name
    # the 4-CTA cluster form: A tiles by multicast TMA, stage release by multicast tcgen05.commit, cluster barriers
    cl = _find(kernels, "gemm_tc_cluster_kernel")
    assert len(cl) >= 4, len(cl)
    for name, ops in cl.items():
        assert {"UTCHMMA", "UTMALDG.2D.MULTICAST", "UTCBAR.MULTICAST", "UCGABAR_ARV", "UCGABAR_WAIT"} <= set(ops) | {op.split(".")[0] for op in ops}, name
    # the MMAs of a k-block issue back to back: no vote loop (ELECT ... BRA.U.ANY) around tcgen05.mma any more
    for name, ops in gemms.items():
        idx = [i for i, op in enumerate(ops) if op.startswith("UTCHMMA")]
        gaps = [b - a for a, b in zip(idx, idx[1:])]
        assert gaps and sorted(gaps)[len(gaps) // 2] <= 4, (name, sorted(gaps)[len(gaps) // 2])
    rowchain = _find(kernels, "15rowchain_kernel")
    assert rowchain and all({"UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"} <= {op.split(".")[0] for op in ops} for ops in rowchain.values())
    chain = _find(kernels, "12chain_kernel")
    assert chain and all({"UTCHMMA", "UTMALDG", "LDTM", "UCGABAR_ARV", "UCGABAR_WAIT"} <= {op.split(".")[0] for op in ops}
                         for ops in chain.values()), {k: sorted({op.split(".")[0] for op in v} & {"UTCHMMA", "UTMALDG", "LDTM", "UCGABAR_ARV", "UCGABAR_WAIT"}) for k, v in chain.items()}


# (needles are fragments of the mangled names, with their length prefixes)
# Current sizes plus a small margin: the test exists to stop silent growth.  The chain kernel contains every row-local phase
# body (seven GEMM tile instantiations as separate functions plus the element-wise bodies); only one of them runs at a time.
@pytest.mark.parametrize("needle,limit_kb", [("gemm_tc_kernelILi0ELi30919", 92), ("gemm_tc_kernelILi0ELi10373", 56), ("gemm_tc_kernelILi0ELi6273", 48), ("gemm_tc_kernelILi0ELi207", 52), ("gemm_tc_kernelILi0ELi195", 30),
                                              ("gemm_tc_kernelILi2ELi240", 52), ("gemm_tc_kernelILi2ELi208", 30), ("gemm_tc_kernelILi1ELi768", 24),
                                              ("12adamw_kernel", 28), ("15dp_adamw_kernel", 40), ("18dp_exchange_kernel", 28),
                                              ("13ingest_kernel", 20), ("13bn_act_kernel", 36), ("13bn_bwd_kernel", 12), ("17latent_fwd_kernel", 16),
                                              ("17latent_bwd_kernel", 8), ("14metrics_kernel", 26), ("11loss_kernel", 56), ("12chain_kernel", 420)])
def test_kernel_code_size_budget(sass, needle, limit_kb):
    _, kernels = sass
    found = _find(kernels, needle)
    assert found, needle
    for name, ops in found.items():
        kb = len(ops) * 16 / 1024
        assert kb <= limit_kb, f"{name}: {kb:.1f} KB of SASS > {limit_kb} KB"


def test_latent_backward_epilogue_is_compiled_in(sass):
    """The opt-in GF_LATBWD instantiation really contains its epilogue (exponentials, 16-bit stores): the tile body works on a
    local snapshot of the problem descriptor, and a field the snapshot does not copy is undefined -- the compiler then drops
    the whole epilogue without a warning (this happened once while the path was written)."""
    _, kernels = sass
    found = _find(kernels, "gemm_tc_kernelILi2ELi32768")
    assert found
    for name, ops in found.items():
        assert sum(op.startswith("MUFU.EX2") for op in ops) >= 8, name
        assert sum(op.startswith("STG.E.U16") for op in ops) >= 8, name
        assert any(op.startswith("LDTM") for op in ops) and any(op.startswith("UTCHMMA") for op in ops), name
