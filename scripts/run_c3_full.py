"""BASELINE configuration C3 at full size on one GPU: LHC lattice with 72 BeamBeam4D and
2 BeamBeam6D (15 slices) lenses (examples/beambeam), 1e7 particles x 1e3 turns.
Writes gpurun_out/c3_full.json.

    python scripts/run_c3_full.py [n_particles] [n_turns]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xline_b200 as xl  # noqa: E402
from xline_b200 import configs  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
turns = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000
line, cols, p0c, m0 = configs.config_lhc_beambeam(n)
p = xl.Particles(p0c=p0c, mass0=m0, **cols)
del cols
warm = xl.Particles(p0c=p0c, mass0=m0, x=[0.0] * 1000)
line.track(warm, num_turns=1)          # module load, lattice upload
torch.cuda.synchronize()
t0 = time.perf_counter()
kernel_ms, done, launches = 0.0, 0, 0
seg = 100
for start in range(0, turns, seg):
    before = int(p.at_turn.sum())
    line.track(p, num_turns=min(seg, turns - start), turns_per_launch=50, timed=True)
    kernel_ms += line.last_stats["kernel_ms"]
    launches += line.last_stats["kernel_launches"]
    done += int(p.at_turn.sum()) - before
    print("turn %d: alive %d, %.3g particle-turns/s so far" % (start + seg, int((p.state == 1).sum()),
                                                            done / (kernel_ms * 1e-3)), flush=True)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
out = dict(config="C3 LHC + 72 BeamBeam4D + 2 BeamBeam6D x 15 slices (examples/beambeam)", particles=n, turns=turns,
           elements_per_turn=len(line), algorithmic_ops_per_turn=line.algorithmic_ops_per_turn(),
           particle_turns_done=done, kernel_ms=kernel_ms, wall_s=wall, kernel_launches=launches,
           particle_turns_per_s=done / (kernel_ms * 1e-3), particle_turns_per_s_wall=done / wall,
           survivors=int((p.state == 1).sum()), regs=line.last_stats["regs_per_thread"],
           segments=line.pack().segments.tolist())
print(json.dumps(out), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/c3_full.json", "w"), indent=1)
