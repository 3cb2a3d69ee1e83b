import sys, torch
sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs
line, cols, p0c, m0 = configs.config_lhc_beambeam(1000000)
p = xl.Particles(p0c=p0c, mass0=m0, **cols)
line.track(p, num_turns=2, timed=True)
torch.cuda.synchronize()
print(line.last_stats)
