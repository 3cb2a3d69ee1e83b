"""Bit-level comparison of the kernel variants (particles per thread x threads) on C2."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from xline_b200 import _cabi

if len(sys.argv) > 1 and sys.argv[1] != "base":
    _cabi.LIB_PATH = os.path.join("xline_b200", "exp", "lib_%s.so" % sys.argv[1])
import xline_b200 as xl
from xline_b200 import configs

n = 200_000
line, cols, p0c, m0 = configs.config_lhc(n)
ref = None
for shape in ("3x128", "1x128", "1x256", "1x512", "2x128", "2x256", "4x128"):
    ppt, thr = (int(v) for v in shape.split("x"))
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=2, particles_per_thread=ppt, threads_per_block=thr)
    cur = {k: getattr(p, k).cpu().numpy() for k in ("x", "px", "y", "py", "zeta", "delta", "state", "at_element", "at_turn")}
    if ref is None:
        ref = cur
        print(shape, "reference; lost", int((cur["state"] == 0).sum()))
        continue
    out = []
    for k in cur:
        a, b = ref[k], cur[k]
        neq = ~((a == b) | (np.isnan(a.astype(float)) & np.isnan(b.astype(float))))
        if neq.any():
            i = np.flatnonzero(neq)
            out.append("%s: %d differ, max |d| %.3e (first idx %d, state %d/%d at_el %d/%d)" % (
                k, neq.sum(), np.nanmax(np.abs(a[i].astype(float) - b[i].astype(float))), i[0],
                ref["state"][i[0]], cur["state"][i[0]], ref["at_element"][i[0]], cur["at_element"][i[0]]))
    print(shape, "regs", line.last_stats["regs_per_thread"], "IDENTICAL" if not out else "; ".join(out))
