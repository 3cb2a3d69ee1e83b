"""Throughput of the BASELINE configurations C3 and C5 (stand-in) at their particle counts.
Writes gpurun_out/r1_configs.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xline_b200 as xl  # noqa: E402
from xline_b200 import configs  # noqa: E402

out = {}


def run(name, line, cols, p0c, m0, turns, **kw):
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=1)
    torch.cuda.synchronize()
    before = int(p.at_turn.sum())
    t0 = time.perf_counter()
    line.track(p, num_turns=turns, timed=True, **kw)
    torch.cuda.synchronize()
    st = line.last_stats
    done = int(p.at_turn.sum()) - before
    out[name] = dict(particles=len(p), turns=turns, elements=len(line), kernel_ms=st["kernel_ms"],
                     particle_turns_done=done, particle_turns_per_s=done / (st["kernel_ms"] * 1e-3),
                     survivors=int((p.state == 1).sum()), regs=st["regs_per_thread"], wall_s=time.perf_counter() - t0,
                     algorithmic_ops_per_turn=line.algorithmic_ops_per_turn())
    print(name, out[name], flush=True)
    return p


line, cols, p0c, m0 = configs.config_lhc_beambeam(10_000_000)
run("C3_lhc_beambeam_10M", line, cols, p0c, m0, 20, turns_per_launch=10)
del cols
line, cols, p0c, m0 = configs.config_psb_like(1_000_000, monitor_stores=50, monitor_ids=100_000)
p = run("C5_psb_like_1M_monitor", line, cols, p0c, m0, 200, turns_per_launch=100)
mon = [el for el in line.elements if type(el).__name__ == "BeamMonitor"][0]
out["C5_psb_like_1M_monitor"]["monitor_slots_written"] = int((~torch.isnan(mon.data["x"])).sum())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r1_configs.json", "w"), indent=1)
