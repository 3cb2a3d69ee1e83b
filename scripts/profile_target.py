"""Small fixed job for ncu: one wave of particles, few turns.
    python scripts/profile_target.py [n] [turns] [particles_per_thread] [c2|c3|c4|c5]"""
import sys

import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 170496
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ppt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
which = sys.argv[4] if len(sys.argv) > 4 else "c2"
line, cols, p0c, m0 = {"c2": configs.config_lhc, "c3": configs.config_lhc_beambeam, "c4": configs.config_petra4,
                       "c5": configs.config_psb}[which](n)
p = xl.Particles(p0c=p0c, mass0=m0, **cols)
line.track(p, num_turns=turns, particles_per_thread=ppt, timed=True)
torch.cuda.synchronize()
print(line.last_stats, int((p.state == 1).sum()))
