"""Small-N regime of the C2 lattice on one GPU: throughput against the number of particles for
every (particles per thread, threads) shape and for the library's own choice (shape 0x0).

    python scripts/sweep_n.py [out.json] [turns]
"""
import json
import sys

import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs

out_path = sys.argv[1] if len(sys.argv) > 1 else None
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = []
for n in (10_000, 30_000, 62_500, 100_000, 125_000, 150_000, 200_000, 300_000, 500_000, 1_000_000):
    line, cols, p0c, m0 = configs.config_lhc(n)
    for ppt, thr in ((0, 0), (4, 128), (3, 128), (2, 128), (2, 256), (1, 128), (1, 256)):
        best = 0.0
        for rep in range(2):
            p = xl.Particles(p0c=p0c, mass0=m0, **cols)
            if rep == 0:
                line.track(p, num_turns=1, particles_per_thread=ppt, threads_per_block=thr)
                p = xl.Particles(p0c=p0c, mass0=m0, **cols)
            line.track(p, num_turns=turns, particles_per_thread=ppt, threads_per_block=thr, timed=True,
                       turns_per_launch=-1)
            st = line.last_stats
            best = max(best, int(p.at_turn.sum()) / (st["kernel_ms"] * 1e-3))
        rows.append({"n": n, "ppt": ppt, "threads": thr, "chosen_threads": st["threads"], "blocks": st["blocks"],
                     "regs": st["regs_per_thread"], "ptps": best})
        print(json.dumps(rows[-1]), flush=True)
if out_path:
    json.dump({"gpu": torch.cuda.get_device_name(0), "turns": turns, "rows": rows}, open(out_path, "w"), indent=1)
