"""Heavy-loss C2 beam (sigma 3e-4 m, A ~ U(0, 12)) -- the bench leg `c2_heavy_loss`
by itself, for schedule experiments: survivor-weighted particle-turns/s of 2 x 100 turns,
launches of at most 50 turns.

    python scripts/probe_heavy_loss.py out.json [thr=0.0078 ...]

thr = compact_threshold of Line.track.  (profiles/r2d_heavy_loss_schedule.json was taken with a
library that also read the two-particles-per-thread threshold -- `fill`, as a fraction of one wave
of the four-particle kernel -- and the launch-length constant -- `ramp`, 0.06 / loss rate -- from the
environment: none of the three moves the figure by more than 1.5 %, the leg runs at the
throughput of a 200 k-particle beam.)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from xline_b200 import _cabi  # noqa: E402

LIBNAME = os.environ.get("XLB_LIB", "base")
if LIBNAME != "base":
    _cabi.LIB_PATH = os.path.join(ROOT, "xline_b200", "exp", "lib_%s.so" % LIBNAME)

import xline_b200 as xl  # noqa: E402
from xline_b200 import configs  # noqa: E402


def main():
    out_path = sys.argv[1]
    settings = [dict(kv.split("=") for kv in a.split(",")) for a in (sys.argv[2:] or ["thr=0"])]
    line, _, p0c, m0 = configs.config_lhc(1000)
    cols = configs.gaussian_beam(1_000_000, 2, 0, sx=3e-4, spx=3e-6, amp_max=12.0)
    rows = []
    w = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(w, num_turns=3, turns_per_launch=50)
    del w
    for st in settings:
        thr = float(st.get("thr", 0.0))
        os.environ["XLB_PPT4_FILL"] = st.get("fill", "1.0")
        os.environ["XLB_RAMP_C"] = st.get("ramp", "0.06")
        best = None
        for rep in range(2):
            p = xl.Particles(p0c=p0c, mass0=m0, **cols)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            launches = 0
            for _ in range(2):
                line.track(p, num_turns=100, turns_per_launch=50, compact_threshold=thr, timed=True)
                launches += line.last_stats["kernel_launches"]
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            v = int(p.at_turn.sum()) / (ms * 1e-3)
            if best is None or v > best["ptps"]:
                best = {"lib": LIBNAME, "env": {k: os.environ[k] for k in ("XLB_PPT4_FILL", "XLB_RAMP_C") if k in os.environ},
                        "compact_threshold": thr, "ptps": v, "ms": ms, "launches": launches,
                        "survivors": int((p.state == 1).sum()), "lost_tally": int(line.loss_tally.sum())}
        rows.append(best)
        print(json.dumps(best), flush=True)
    old = []
    if os.path.exists(out_path):
        with open(out_path) as fh:
            old = json.load(fh)["rows"]
    with open(out_path, "w") as fh:
        json.dump({"gpu": torch.cuda.get_device_name(0), "rows": old + rows}, fh, indent=1)


if __name__ == "__main__":
    main()
