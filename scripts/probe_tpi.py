"""Work-item length (turns per item) against beam size on the C2 lattice: one 50-turn launch.
    python scripts/probe_tpi.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs

for n in (125_000, 200_000, 250_000, 1_000_000):
    line, cols, p0c, m0 = configs.config_lhc(n)
    for tpi in (5, 3, 2, 1):
        best = 0.0
        for rep in range(2):
            p = xl.Particles(p0c=p0c, mass0=m0, **cols)
            if rep == 0:
                line.track(p, num_turns=2)
                p = xl.Particles(p0c=p0c, mass0=m0, **cols)
            line.track(p, num_turns=50, turns_per_item=tpi, turns_per_launch=-1, timed=True)
            st = line.last_stats
            best = max(best, int(p.at_turn.sum()) / (st["kernel_ms"] * 1e-3))
        print(json.dumps({"n": n, "tpi": tpi, "blocks": st["blocks"], "threads": st["threads"],
                          "regs": st["regs_per_thread"], "ptps": best}), flush=True)
