"""Static checks of the built sm_100a library (no GPU needed): the hot kernels exist, do not spill, and the SASS contains the
tensor-core / TMA / NVLink-reduction instructions the design relies on (mnemonics per B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
cp.async.bulk.tensor -> UTMALDG / UTMASTG)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "marius_b200", "lib", "libmarius_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)), reason="needs the built library and cuobjdump")


@pytest.fixture(scope="module")
def resources():
    out = subprocess.run([CUOBJDUMP, "--dump-resource-usage", LIB], capture_output=True, text=True, timeout=300).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        res[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)), local=int(m.group(5)))
    return res


def test_built_for_sm_100a():
    out = subprocess.run([CUOBJDUMP, "--list-elf", LIB], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out


def test_hot_kernels_exist_and_do_not_spill(resources):
    hot = ["gemm_tc_group_kernelILb0E", "neg_rows_kernelILi4E", "edge_rows_kernelILi2ELi2E", "edge_backward_kernelILi2ELi2E", "loss_kernelILi8E",
           "segment_reduce_kernelILi2ELi4E", "fetch_remote_rows_kernelILi4ELb0E", "gather_rows_kernel", "rank_kernel", "sample_negatives_kernel",
           "loss_merge_kernel", "inbox_apply_kernelILi4E", "neg_rows_bulk_kernelILi4E"]
    for name in hot:
        found = [(k, v) for k, v in resources.items() if name in k]
        assert found, f"kernel {name} not in the library"
        for k, v in found:
            assert v["local"] == 0 and v["stack"] == 0, f"{k} spills: {v}"
    for name in ("shard_barrier_kernel", "owner_bounds_kernel"):  # one-warp control kernels of the sharded step (dynamic indexing of their
        assert any(name in k for k in resources), name                # parameter arrays costs them a few bytes of stack: not hot)
    # the persistent contractions: plain (192 threads), and the two backward kernels with converter warps (576 threads: A operand through
    # shared memory / through tensor memory); one CTA per SM.  The 576-thread kernels are capped at 96 registers and keep one table entry
    # (16 bytes, touched once per tile, not per k-block) on the stack.
    gemms = {k: v for k, v in resources.items() if "gemm_tc_group_kernel" in k or "gemm_tc_ts_kernel" in k}
    assert len(gemms) == 3
    for k, v in gemms.items():
        conv = "ILb0E" not in k
        assert v["reg"] * (640 if conv else 192) <= 65536
        assert v["local"] == 0 and v["stack"] <= (16 if conv else 0), f"{k} spills: {v}"


def test_sass_has_tcgen05_tma_and_system_reductions():
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    assert re.search(r"UTC[A-Z]*MMA", sass), "no tcgen05.mma (UTC*MMA) in the SASS"
    assert "UTCHMMA.2CTA" in sass, "the grouped contraction is a cta_group::2 kernel"
    assert "UTMALDG" in sass and "UTMASTG" in sass, "no TMA tensor loads / stores in the SASS"
    assert re.search(r"STG\.E\.64\.STRONG\.SYS", sass) and re.search(r"LDG\.E\.64\.STRONG\.SYS", sass), \
        "the cross-rank barrier publishes / polls its epoch counters with system-scope release / acquire accesses"
    assert "UBLKCP" in sass, "the negative-row kernel stages table rows with bulk asynchronous copies (cp.async.bulk)"
    assert "STTM" in sass and re.search(r"UTCHMMA\.2CTA tmem\[\w+\], gdesc", sass), "the backward contraction takes its A operand from tensor memory (tcgen05.st + [a_tmem])"
    assert "WGMMA" not in sass and "HMMA.16" not in sass  # neither Hopper wgmma nor legacy mma.sync anywhere
