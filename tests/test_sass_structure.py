"""Structural checks of the compiled K1 kernels (no GPU needed: cuobjdump reads the sm_100a SASS of the built
library).  They pin the two properties the kernel's throughput rests on: the unrolled sweeps
are branch-free (a branch per stage splits the sweep into basic blocks and ptxas stops scheduling across stages)
and nothing spills."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "incentive-design-mpc_b200", "csrc", "liblompc_b200.so")
sys.path.insert(0, os.path.join(ROOT, "tools"))

KERNELS = {  # the default shapes (lompc_api.cu: launch_solve_reg_variant)
    "small": "_ZN5lompc22lompc_solve_reg_kernelILi24ELi1ELi64ELi4ELb1EEEvNS_6ConstsENS_9SolveArgsE",
    "large": "_ZN5lompc22lompc_solve_reg_kernelILi24ELi4ELi64ELi4ELb1EEEvNS_6ConstsENS_9SolveArgsE",
    "large_saturating": "_ZN5lompc22lompc_solve_reg_kernelILi24ELi4ELi256ELi1ELb1EEEvNS_6ConstsENS_9SolveArgsE",
}

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB),
                                reason="needs cuobjdump and the built library")


@pytest.mark.parametrize("name", sorted(KERNELS))
def test_sweeps_are_branch_free_and_spill_free(name, tmp_path):
    from sass_sched import opclass, parse
    sass = tmp_path / "k.sass"
    out = subprocess.run(["cuobjdump", "-sass", "-fun", KERNELS[name], LIB], capture_output=True, text=True).stdout
    sass.write_text(out)
    prog = parse(str(sass))
    assert len(prog) > 2000, "kernel not found in the library"
    ops = [opclass(p["text"]) for p in prog]
    # two loop bodies (optimistic + safeguarded) of 2 x 24 unrolled stages each: a branch per stage would be
    # >= 96 branches; the loops, the phase switch and the epilogue need a few dozen
    assert ops.count("BRA") <= 60, ops.count("BRA")
    # every backward stage forms one reciprocal: MUFU.RCP64H appears once per unrolled stage of each copy
    assert sum(1 for p in prog if p["text"].startswith("MUFU.RCP64H")) >= 48
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    m = re.search(re.escape(KERNELS[name]) + r":\s*\n\s*REG:(\d+) STACK:(\d+)", res)
    assert m, "resource usage not found"
    assert int(m.group(1)) <= 255 and int(m.group(2)) <= 64, m.groups()


TMA_KERNELS = {  # automatic mode at N = 24 from 65,536 QPs (lompc_api.cu: launch_solve_reg_variant)
    "small": "_ZN5lompc26lompc_solve_reg_tma_kernelILi24ELi1ELi64ELi4ELb0EEEvNS_6ConstsENS_9SolveArgsE",
    "large": "_ZN5lompc26lompc_solve_reg_tma_kernelILi24ELi4ELi128ELi2ELb1EEEvNS_6ConstsENS_9SolveArgsE",
}


@pytest.mark.parametrize("name", sorted(TMA_KERNELS))
def test_bulk_copy_kernel_moves_rows_with_the_copy_engine(name):
    """The saturating K1 kernel loads every price row with ONE bulk copy into shared memory (UBLKCP.S.G, completion
    through an mbarrier: SYNCS...TRANS64), reads it with 36 LDS.128, and stores the result row with one bulk store
    (UBLKCP.G.S): no wide per-thread global loads or stores are left (those touched 32 lines per instruction)."""
    out = subprocess.run(["cuobjdump", "-sass", "-fun", TMA_KERNELS[name], LIB], capture_output=True, text=True).stdout
    ops = [m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", out)]
    assert len(ops) > 2000, "kernel not found in the library"
    assert sum(o.startswith("UBLKCP.S.G") for o in ops) == 1 and sum(o.startswith("UBLKCP.G.S") for o in ops) == 1
    assert any(o.startswith("SYNCS.ARRIVE.TRANS64") for o in ops) and any("TRYWAIT" in o for o in ops)
    assert sum(o.startswith("LDS.128") for o in ops) >= 36
    assert not any(o.startswith("LDG.E.128") or o.startswith("STG.E.128") for o in ops)
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    m = re.search(re.escape(TMA_KERNELS[name]) + r":\s*\n\s*REG:(\d+) STACK:(\d+)", res)
    assert m and int(m.group(1)) <= 255 and int(m.group(2)) <= 64, m and m.groups()
