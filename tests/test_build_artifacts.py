"""What the compiler made of the kernels (CPU-only checks of the in-tree build): the claims of
DESIGN.md section 4 about registers, spills and the instructions that prove TMA streaming and the
warp-uniform header are read back from `ptxas -v` and `cuobjdump -sass`."""
import os
import re
import shutil
import subprocess

import pytest

from xline_b200 import build as B


def _ptxas(unit):
    path = os.path.join(B.OBJ, unit + ".ptxas.txt")
    if not os.path.exists(path):
        B.build()
    if not os.path.exists(path):  # objects came prebuilt and up to date: nothing was recompiled
        pytest.skip("no ptxas report next to the prebuilt objects")
    kernels = {}
    cur = None
    for line in open(path):
        m = re.search(r"Compiling entry function '(\w+)'", line)
        if m:
            cur = m.group(1)
            kernels[cur] = {}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and "stack" not in kernels[cur]:
            kernels[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            kernels[cur]["regs"] = int(m.group(1))
    return kernels


def _find(kernels, ppt, threads):
    key = "track_kernelILi%dELi%dE" % (ppt, threads)
    hits = [v for k, v in kernels.items() if key in k and k.endswith("Lb0EEEvNS_5KArgsE")]
    assert len(hits) == 1, (key, list(kernels))
    return hits[0]


def test_default_thin_lens_kernel_fits_three_ctas_per_sm_without_spills():
    """3 particles/thread x 128 threads x 3 CTAs/SM needs <= 168 registers (65536 / 384)."""
    k = _find(_ptxas("track_fast"), 3, 128)
    assert k["regs"] <= 168
    assert k["spill_st"] == 0 and k["spill_ld"] == 0


def test_two_particle_kernels_fit_128_registers():
    k = _find(_ptxas("track_fast"), 2, 256)
    assert k["regs"] <= 128 and k["spill_st"] == 0
    k = _find(_ptxas("track_fast_bf"), 2, 256)  # BeamBeam4D / space charge: the default for C5
    assert k["regs"] <= 128


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_has_tma_bulk_copies_and_uniform_header():
    obj = os.path.join(B.OBJ, "track_fast.o")
    if not os.path.exists(obj):
        B.build()
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    n_kernels = sass.count("Function :")
    assert n_kernels >= 8
    assert sass.count("UBLKCP") >= n_kernels      # cp.async.bulk: the lattice ring (prologue + refill)
    assert sass.count("SYNCS") >= n_kernels       # mbarrier arrive / try_wait
    assert sass.count("REDUX.OR") == n_kernels    # one warp-uniform header decode per kernel
    assert "HMMA" not in sass and "UTCHMMA" not in sass  # no tensor cores on this path
    # FP64 is where the arithmetic is: DFMA dominates the FP64-pipe instructions
    assert sass.count("DFMA") > 10 * sass.count("FFMA")
