import sys
import torch
sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs
n = 303104
line, cols, p0c, m0 = configs.config_lhc(n)
for strict, ppt in ((True, 1), (True, 2)):
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=1, strict=strict, particles_per_thread=ppt)
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=3, strict=strict, particles_per_thread=ppt, timed=True)
    st = line.last_stats
    print("strict ppt=%d regs=%d %.3e p-t/s" % (ppt, st["regs_per_thread"], int(p.at_turn.sum()) / (st["kernel_ms"] * 1e-3)))
