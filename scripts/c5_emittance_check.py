"""C5 (PS Booster + 120 frozen space-charge kicks): distribution-level comparison of the fast
kernel with the strict one (the bit-faithful stand-in for the NumPy path) over many turns.
Particle by particle the two part company after a few hundred turns (the map is chaotic for a
good part of the beam); emittances and loss counts must not.

    python scripts/c5_emittance_check.py [n] [turns] [out.json]
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
line, cols, p0c, m0 = configs.config_psb(n)


def emit(p):
    ok = (p.state == 1)
    out = {}
    for a, b in (("x", "px"), ("y", "py")):
        u, v = getattr(p, a)[ok].double(), getattr(p, b)[ok].double()
        u, v = u - u.mean(), v - v.mean()
        out[a] = float(torch.sqrt((u * u).mean() * (v * v).mean() - (u * v).mean() ** 2))
    return out, int(ok.sum())


rows = []
parts = {name: xl.Particles(p0c=p0c, mass0=m0, **cols) for name in ("fast", "strict")}
step = max(turns // 8, 1)
for done in range(0, turns + 1, step):
    row = {"turn": done}
    for name, p in parts.items():
        if done:
            line.track(p, num_turns=step, strict=(name == "strict"))
        e, alive = emit(p)
        row[name] = {"emit_x": e["x"], "emit_y": e["y"], "alive": alive}
    row["rel_diff_emit_x"] = abs(row["fast"]["emit_x"] / row["strict"]["emit_x"] - 1)
    row["rel_diff_emit_y"] = abs(row["fast"]["emit_y"] / row["strict"]["emit_y"] - 1)
    dx = (parts["fast"].x - parts["strict"].x).abs()
    row["median_abs_dx_m"] = float(dx.median())
    rows.append(row)
    print(json.dumps(row), flush=True)
res = {"n": n, "turns": turns, "gpu": torch.cuda.get_device_name(0), "rows": rows,
       "statistical_error_of_an_emittance": float(1 / np.sqrt(n))}
if len(sys.argv) > 3:
    json.dump(res, open(sys.argv[3], "w"), indent=1)
