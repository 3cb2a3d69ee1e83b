"""Survival vs amplitude on the LHC config: how many particles of the synthetic beam
survive N turns, binned by the amplitude factor A."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs

n = 200_000
turns = int(sys.argv[1]) if len(sys.argv) > 1 else 300
amp_max = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
line, _, p0c, m0 = configs.config_lhc(10)
cols = configs.gaussian_beam(n, 2, 0, amp_max=amp_max, sx=1e-4, spx=1e-6)
rng = np.random.default_rng(configs.SEED0 + 2000)
a = rng.uniform(0.0, amp_max, n)  # same stream as gaussian_beam's first draw
p = xl.Particles(p0c=p0c, mass0=m0, **cols)
line.track(p, num_turns=turns, turns_per_launch=25, timed=True)
st = p.state.cpu().numpy()
at = p.at_turn.cpu().numpy()
print("stats", line.last_stats)
print("survivors %.4f" % (st == 1).mean(), "particle-turns done %.4e of %.4e" % (at.sum(), n * turns))
for lo in np.arange(0, amp_max, amp_max / 12):
    m = (a >= lo) & (a < lo + amp_max / 12)
    print("A in [%.2f,%.2f): survive %.3f   median turn of loss %s" % (
        lo, lo + amp_max / 12, (st[m] == 1).mean(), np.median(at[m][st[m] == 0]) if (st[m] == 0).any() else None))
