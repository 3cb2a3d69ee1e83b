"""BASELINE configuration C5 at full size on one GPU: the PS Booster lattice of tests/psb
(xline_b200.configs.config_psb: 120 SCQGaussProfile kicks, 264 apertures, RF), 1e6 particles x
1e4 turns, one BeamMonitor storing every 100th turn for the first 1e5 particle ids.
Writes gpurun_out/c5_psb_full.json.

    python scripts/run_c5_psb.py [n_particles] [n_turns]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xline_b200 as xl  # noqa: E402
from xline_b200 import configs  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
turns = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000
skip = 100
line, cols, p0c, m0 = configs.config_psb(n, monitor_stores=turns // skip, monitor_ids=min(n, 100_000),
                                         monitor_skip=skip)
p = xl.Particles(p0c=p0c, mass0=m0, **cols)
warm = p.copy()
line.track(warm, num_turns=2)          # warm-up on a copy (module load, lattice upload)
line.reset_monitors()
torch.cuda.synchronize()
eps0 = {u: float(torch.sqrt(torch.var(getattr(p, u)) * torch.var(getattr(p, "p" + u))
                            - torch.mean((getattr(p, u) - getattr(p, u).mean())
                                         * (getattr(p, "p" + u) - getattr(p, "p" + u).mean())) ** 2))
        for u in ("x", "y")}
t0 = time.perf_counter()
kernel_ms, done = 0.0, 0
seg = 1000
for start in range(0, turns, seg):     # host loop only to report progress; each call is segmented inside
    before = int(p.at_turn.sum())
    line.track(p, num_turns=min(seg, turns - start), turns_per_launch=100, timed=True)
    kernel_ms += line.last_stats["kernel_ms"]
    done += int(p.at_turn.sum()) - before
    print("turn %d: alive %d, %.3g particle-turns/s so far" % (start + seg, int((p.state == 1).sum()),
                                                            done / (kernel_ms * 1e-3)), flush=True)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
alive = p.state == 1
eps1 = {}
for u in ("x", "y"):
    a, b = getattr(p, u)[alive], getattr(p, "p" + u)[alive]
    eps1[u] = float(torch.sqrt(torch.var(a) * torch.var(b) - torch.mean((a - a.mean()) * (b - b.mean())) ** 2))
mon = [el for el in line.elements if type(el).__name__ == "BeamMonitor"][0]
out = dict(config="C5 PS Booster (tests/psb), 120 SCQGaussProfile kicks, BeamMonitor", particles=n, turns=turns,
           elements_per_turn=len(line), algorithmic_ops_per_turn=line.algorithmic_ops_per_turn(),
           particle_turns_done=done, kernel_ms=kernel_ms, wall_s=wall,
           particle_turns_per_s=done / (kernel_ms * 1e-3), particle_turns_per_s_wall=done / wall,
           survivors=int(alive.sum()), emittance_start=eps0, emittance_end=eps1,
           zeta_rms_end=float(p.zeta[alive].std()), delta_rms_end=float(p.delta[alive].std()),
           monitor_slots_written=int((~torch.isnan(mon.data["x"])).sum()),
           monitor_bytes=int(sum(v.numel() * v.element_size() for v in mon.data.values())),
           regs=line.last_stats["regs_per_thread"])
print(json.dumps(out), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/c5_psb_full.json", "w"), indent=1)
