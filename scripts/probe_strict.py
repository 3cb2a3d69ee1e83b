"""Strict (bit-exact) kernel on the C2 lattice: throughput per (particles per thread, threads) shape
at 1 M particles, with a checksum of the survivors' coordinates (all shapes must print the same)."""
import hashlib
import json
import sys

import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import configs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 20
line, cols, p0c, m0 = configs.config_lhc(n)
for ppt, thr in ((2, 128), (3, 128), (4, 128)):
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=1, strict=True, particles_per_thread=ppt, threads_per_block=thr)
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=turns, strict=True, particles_per_thread=ppt, threads_per_block=thr, timed=True)
    st = line.last_stats
    h = hashlib.sha1()
    for k in ("x", "px", "y", "py", "zeta", "delta", "state", "at_element", "at_turn"):
        h.update(getattr(p, k).cpu().numpy().tobytes())
    print(json.dumps({"strict": True, "ppt": ppt, "threads": thr, "regs": st["regs_per_thread"],
                      "blocks": st["blocks"], "ptps": int(p.at_turn.sum()) / (st["kernel_ms"] * 1e-3),
                      "sha1": h.hexdigest()[:16]}), flush=True)
