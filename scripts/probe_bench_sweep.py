"""Sweep (ppt, threads, turns_per_item) at the bench size (1M particles, C2)."""
import sys

import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs

n, turns = 1_000_000, 40
line, cols, p0c, m0 = configs.config_lhc(n)
ALL = ((2, 256, 5), (2, 256, 2), (2, 256, 10), (2, 256, -1), (3, 128, 5), (4, 128, 5), (2, 128, 5), (1, 512, 5),
       (3, 128, -1), (1, 512, -1), (1, 256, 5), (1, 128, 5), (3, 128, 3), (3, 128, 10), (3, 96, 5), (4, 128, 5), (3, 64, 5), (3, 192, 5), (3, 160, 5))
sel = [ALL[int(a)] for a in sys.argv[1:] if not a.startswith("cw=")] or ALL
for a in sys.argv[1:]:
    if a.startswith("cw="):
        line.chunk_words = int(a[3:])
for ppt, thr, tpi in sel:
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=2, particles_per_thread=ppt, threads_per_block=thr, turns_per_item=tpi)
    torch.cuda.synchronize()
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=turns, particles_per_thread=ppt, threads_per_block=thr, turns_per_item=tpi, timed=True)
    st = line.last_stats
    done = int(p.at_turn.sum())
    print("ppt=%d thr=%d tpi=%d regs=%d blocks=%d ms=%.1f  %.3e p-t/s" % (
        ppt, thr, tpi, st["regs_per_thread"], st["blocks"], st["kernel_ms"], done / (st["kernel_ms"] * 1e-3)), flush=True)
