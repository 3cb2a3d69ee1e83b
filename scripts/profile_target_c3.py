"""Small fixed job for ncu on the beam-field kernel: C3 lattice, one wave, 2 turns."""
import sys

import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 75776
line, cols, p0c, m0 = configs.config_lhc_beambeam(n)
p = xl.Particles(p0c=p0c, mass0=m0, **cols)
line.track(p, num_turns=2, timed=True)
torch.cuda.synchronize()
print(line.last_stats)
