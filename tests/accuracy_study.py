"""Accuracy of the fast kernel on the C2 lattice (LHC + apertures), measured on the GPU:

* one turn, 2000 particles: fast and strict kernels against the CPU oracle;
* error growth: fast vs strict (the strict kernel keeps the reference's operation order and
  is bit-identical to the NumPy path except for sin() in the 12 cavities) after 1, 10, 100
  and 1000 turns on 20 000 particles, per amplitude bin.

Writes profiles/accuracy_<tag>.json (tag = first argument, default r2).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xline_b200 as xl  # noqa: E402
from xline_b200 import configs  # noqa: E402
from tests import helpers as H  # noqa: E402

COORDS = ("x", "px", "y", "py", "zeta", "delta")


def stats(a, b, mask):
    out = {}
    for k in COORDS:
        rms = float(np.sqrt(np.mean(b[k][mask] ** 2)))
        e = np.abs(a[k][mask] - b[k][mask]) / rms
        out[k] = dict(median=float(np.median(e)), p99=float(np.quantile(e, 0.99)), max=float(e.max()))
    return out


def main():
    res = {}
    line, cols, p0c, m0 = configs.config_lhc(2000)
    ref = H.run_oracle(line.to_specs(), cols, p0c, m0, num_turns=1)
    for name, strict in (("fast", False), ("strict", True)):
        p = xl.Particles(p0c=p0c, mass0=m0, **cols)
        line.track(p, num_turns=1, strict=strict)
        got = p.to_numpy()
        alive = ref["state"] == 1
        res["one_turn_%s_vs_oracle" % name] = dict(
            n=2000, lost=int((~alive).sum()),
            state_mismatch=int((got["state"] != ref["state"]).sum()),
            at_element_mismatch=int((got["at_element"] != ref["at_element"]).sum()),
            err_rel_to_beam_rms=stats(got, ref, alive))
    # yardstick: the same arithmetic in 80-bit extended precision (x87 long double).  How far
    # is each float64 implementation from it after one turn?
    from oracle import xline_oracle as xo

    nl = 400
    line, cols, p0c, m0 = configs.config_lhc(nl)
    c64 = {k: v for k, v in cols.items() if k != "particle_id"}
    pl = xo.OracleParticles(nl, p0c=p0c, mass0=m0, dtype=np.longdouble, **c64)
    with np.errstate(all="ignore"):
        xo.line_track(line.to_specs(), pl, 1)
    truth = xo.gather_full(pl, nl)
    ref64 = H.run_oracle(line.to_specs(), c64, p0c, m0, num_turns=1)
    ok = (truth["state"] == 1) & (ref64["state"] == 1)
    yard = {}
    for name, arr in (("numpy_float64_oracle", ref64),):
        yard[name] = stats({k: arr[k].astype(np.float64) for k in COORDS},
                           {k: truth[k].astype(np.float64) for k in COORDS}, ok)
    for name, strict in (("gpu_fast", False), ("gpu_strict", True)):
        p = xl.Particles(p0c=p0c, mass0=m0, **cols)
        line.track(p, num_turns=1, strict=strict)
        got = p.to_numpy()
        # subtract in extended precision, then report relative to the beam r.m.s.
        yard[name] = {}
        for k in COORDS:
            rms = float(np.sqrt(np.mean(truth[k][ok].astype(np.float64) ** 2)))
            e = np.abs(got[k][ok].astype(np.longdouble) - truth[k][ok]).astype(np.float64) / rms
            yard[name][k] = dict(median=float(np.median(e)), p99=float(np.quantile(e, 0.99)), max=float(e.max()))
    e = yard["numpy_float64_oracle"]
    for k in COORDS:
        rms = float(np.sqrt(np.mean(truth[k][ok].astype(np.float64) ** 2)))
        d = np.abs(ref64[k][ok].astype(np.longdouble) - truth[k][ok]).astype(np.float64) / rms
        e[k] = dict(median=float(np.median(d)), p99=float(np.quantile(d, 0.99)), max=float(d.max()))
    res["one_turn_distance_to_extended_precision"] = dict(n=nl, **yard)
    print("vs extended precision, x:", {k: v["x"] for k, v in yard.items()})

    n = 20000
    line, cols, p0c, m0 = configs.config_lhc(n)
    amp = np.sqrt(cols["x"] ** 2 + cols["y"] ** 2) / 1e-4
    pf = xl.Particles(p0c=p0c, mass0=m0, **cols)
    ps = xl.Particles(p0c=p0c, mass0=m0, **cols)
    done = 0
    growth = {}
    for target in (1, 10, 100, 1000):
        line.track(pf, num_turns=target - done, turns_per_launch=50)
        line.track(ps, num_turns=target - done, strict=True, turns_per_launch=50)
        done = target
        a, b = pf.to_numpy(), ps.to_numpy()
        both = (a["state"] == 1) & (b["state"] == 1)
        entry = dict(
            survivors_fast=int((a["state"] == 1).sum()), survivors_strict=int((b["state"] == 1).sum()),
            state_mismatch=int((a["state"] != b["state"]).sum()),
            loss_place_mismatch=int(((a["state"] == 0) & (b["state"] == 0) &
                                     ((a["at_element"] != b["at_element"]) | (a["at_turn"] != b["at_turn"]))).sum()),
            all=stats(a, b, both))
        for lo, hi in ((0, 2), (2, 4), (4, 8), (8, 100)):
            m = both & (amp >= lo) & (amp < hi)
            if m.sum() > 10:
                entry["amp_%g_%g_sigma" % (lo, hi)] = dict(n=int(m.sum()), x=stats(a, b, m)["x"])
        growth[str(target)] = entry
        print(target, entry["state_mismatch"], entry["all"]["x"])
    res["fast_vs_strict_growth"] = growth
    res["notes"] = ("errors are |a-b| / rms(b) per coordinate over particles alive in both runs; "
                    "amplitude = sqrt(x0^2+y0^2)/1e-4 m")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for d in ("profiles", "gpurun_out"):  # gpurun_out/ is what travels back from the GPU box
        os.makedirs(os.path.join(ROOT, d), exist_ok=True)
        with open(os.path.join(ROOT, d, "accuracy_%s.json" % (sys.argv[1] if len(sys.argv) > 1 else "r2")), "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
