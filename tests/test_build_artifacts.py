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


def test_default_thin_lens_kernels_fit_three_ctas_per_sm():
    """128 threads x 3 CTAs/SM allow 168 registers (65536 / 384).  The general kernel (chi column,
    3 particles/thread) fits without spills; the one-species kernels (no chi register, 4 particles
    per thread -- the default) fit with a handful of spilled words, all of them per-chunk loop
    state outside the record loop (checked in SASS below)."""
    k = _find(_ptxas("track_fast"), 3, 128)
    assert k["regs"] <= 168
    assert k["spill_st"] == 0 and k["spill_ld"] == 0
    for unit in ("track_fast_nc", "track_fast_nc_lo"):
        k = _find(_ptxas(unit), 4, 128)
        assert k["regs"] <= 168
        assert k["spill_st"] <= 64, (unit, k)


def test_two_particle_kernels_fit_128_registers():
    k = _find(_ptxas("track_fast"), 2, 256)
    assert k["regs"] <= 128 and k["spill_st"] <= 16
    k = _find(_ptxas("track_fast_bf"), 2, 256)  # BeamBeam4D / space charge: the default for C5
    assert k["regs"] <= 128


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_four_particle_kernel_does_not_spill_inside_the_record_loop():
    """The spilled words of the default kernel are chunk-loop state: no STL/LDL between the
    warp-uniform header decode (REDUX.OR, top of the record loop) and the END_CHUNK return."""
    obj = os.path.join(B.OBJ, "track_fast_nc.o")
    if not os.path.exists(obj):
        B.build()
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    body = sass.split("track_kernelILi4ELi128ELi3ELb0E")[1].split("Function :")[0]
    lines = [ln for ln in body.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln)]
    redux = [i for i, ln in enumerate(lines) if "REDUX.OR" in ln]
    assert len(redux) == 1
    spills = [i for i, ln in enumerate(lines) if re.search(r"\b(STL|LDL)\b", ln)]
    # the record loop spans from the header decode to the last of its back-edges; everything
    # the chunk loop spills sits before the decode or after the loop
    addr = lambda ln: int(re.search(r"/\*([0-9a-f]+)\*/", ln).group(1), 16)
    top_lo, top_hi = addr(lines[redux[0] - 12]), addr(lines[redux[0]])
    back = [i for i, ln in enumerate(lines) if i > redux[0] and re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", ln)
            and top_lo <= int(re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", ln).group(1), 16) <= top_hi]
    assert back
    loop_end = max(back)
    # top of the record loop = the earliest back-edge target (a few instructions above the decode)
    first_target = min(int(re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", lines[i]).group(1), 16) for i in back)
    loop_top = min(i for i, ln in enumerate(lines) if addr(ln) >= first_target)
    assert redux[0] - 12 <= loop_top <= redux[0]
    inside = [i for i in spills if loop_top <= i <= loop_end]
    assert not inside, [lines[i] for i in inside[:5]]


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


def _kernel_sass(obj_name, kernel_key):
    obj = os.path.join(B.OBJ, obj_name)
    if not os.path.exists(obj):
        B.build()
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    return sass, sass.split(kernel_key)[1].split("Function :")[0]


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_beam_field_kernels_keep_their_warps_in_step():
    """DESIGN.md section 4.2, "Warps in step": the beam-field families carry one CTA barrier per
    lattice chunk (item fetch, publish, chunk loop: three BAR.SYNC), the thin-lens families do
    without it (two)."""
    _, bf = _kernel_sass("track_fast_bf_nc_lo.o", "track_kernelILi2ELi256ELi2ELb0E")
    _, lean = _kernel_sass("track_fast_nc.o", "track_kernelILi4ELi128ELi3ELb0E")
    n_bf, n_lean = bf.count("BAR.SYNC"), lean.count("BAR.SYNC")
    assert n_bf == n_lean + 1, (n_bf, n_lean)


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_faddeeva_loop_is_straight_line_for_two_particles_per_thread():
    """Up to four chains the Weideman recurrence is unrolled completely: every pair of
    coefficients is read ONCE (LDCU.128 from the constant bank into uniform registers, the DFMAs
    take them as operands) -- (N - 2) / 2 loads per copy of the field code, no loop; the body is
    inlined at its two call sites (BeamBeam4D, space charge).  Three and four particles per
    thread (six, eight chains) keep one out-of-line copy with the rolled loop."""
    n_coeff = int(re.search(r"#define XLB_WEID_N (\d+)",
                            open(os.path.join(B.CSRC, "faddeeva_coeffs.inc")).read()).group(1))
    loads = {}
    for ppt, key in ((1, "ILi1ELi256ELi2ELb0E"), (2, "ILi2ELi256ELi2ELb0E"), (3, "ILi3ELi128ELi3ELb0E")):
        _, body = _kernel_sass("track_fast_bf_nc_lo.o", "track_kernel" + key)
        loads[ppt] = len(re.findall(r"LDCU\.128 UR\d+, c\[0x3\]", body))
    assert loads[1] == loads[2] == 2 * ((n_coeff - 2) // 2), loads
    assert loads[3] < loads[2], loads
