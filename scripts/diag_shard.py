"""Replays tests/test_gpu_parity.py::test_sharding_invariance_full_size with a report of what differs."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np

from xline_b200 import _cabi

if len(sys.argv) > 1 and sys.argv[1] != "base":
    _cabi.LIB_PATH = os.path.join("xline_b200", "exp", "lib_%s.so" % sys.argv[1])
import xline_b200 as xl
from xline_b200 import configs

n = 300_000
line, cols, p0c, m0 = configs.config_lhc(n)
KEYS = ("x", "px", "y", "py", "zeta", "delta", "rpp", "rvv", "s", "state", "at_element", "at_turn")


def run(sl, **kw):
    p = xl.Particles(p0c=p0c, mass0=m0, **{k: v[sl] for k, v in cols.items()})
    line.track(p, num_turns=2, **kw)
    return {k: getattr(p, k).cpu().numpy() for k in KEYS}


def cmp(name, a, b):
    out = []
    for k in KEYS:
        neq = ~((a[k] == b[k]) | (np.isnan(a[k].astype(float)) & np.isnan(b[k].astype(float))))
        if neq.any():
            i = np.flatnonzero(neq)
            out.append("%s: %d differ (lost among them %d), max|d| %.3e, idx %s, at_element %s" % (
                k, neq.sum(), int((a["state"][i] == 0).sum()), np.nanmax(np.abs(a[k][i].astype(float) - b[k][i].astype(float))),
                i[:5].tolist(), a["at_element"][i[:5]].tolist()))
    print(name, "IDENTICAL" if not out else "\n   ".join([""] + out), flush=True)


whole = run(slice(0, n))
cmp("whole again", whole, run(slice(0, n)))
cmp("whole ppt1", whole, run(slice(0, n), particles_per_thread=1))
cut = 123_457
pa, pb = run(slice(0, cut), particles_per_thread=1), run(slice(cut, n), particles_per_thread=1)
cmp("parts ppt1", whole, {k: np.concatenate([pa[k], pb[k]]) for k in KEYS})
pa, pb = run(slice(0, cut)), run(slice(cut, n))
cmp("parts ppt3", whole, {k: np.concatenate([pa[k], pb[k]]) for k in KEYS})
