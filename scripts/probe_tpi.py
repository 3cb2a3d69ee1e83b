"""Work-item length sweep: throughput of the regular library against `turns_per_item` (0 = the
library's own choice, at least 16 lattice chunks per item) on the probe configurations, same seeded
beams, with a checksum of the result (the item length never changes a bit).

    python scripts/probe_tpi.py out.json [c5:0,12,25,50 c3:0,2,4 c4r:0,2,4 c2@250000:0,2]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from xline_b200 import _cabi  # noqa: E402

LIBNAME = os.environ.get("XLB_LIB", "base")  # an experiment library of scripts/build_exp.py
if LIBNAME != "base":
    _cabi.LIB_PATH = os.path.join(ROOT, "xline_b200", "exp", "lib_%s.so" % LIBNAME)
SHAPE = os.environ.get("XLB_SHAPE", "")  # "2x512": particles per thread x threads per block
KW = dict(zip(("particles_per_thread", "threads_per_block"), (int(v) for v in SHAPE.split("x")))) if SHAPE else {}
if os.environ.get("XLB_STRICT"):  # the bit-exact kernels
    KW["strict"] = True

import xline_b200 as xl  # noqa: E402
from xline_b200 import configs  # noqa: E402

TABLE = {"c2": (configs.config_lhc, 1_000_000, 40), "c3": (configs.config_lhc_beambeam, 1_000_000, 10),
         "c4r": (configs.config_petra4, 1_000_000, 40), "c5": (configs.config_psb, 1_000_000, 200)}


def main():
    out_path = sys.argv[1]
    jobs = sys.argv[2:] or ["c5:0,12,25,50", "c3:0,2,4", "c4r:0,2,4", "c2@250000:0,2"]
    rows = []
    for job in jobs:
        name, tpis = job.split(":")
        cfg, _, n = name.partition("@")
        fn, n_def, turns = TABLE[cfg]
        n = int(n) if n else n_def
        line, cols, p0c, m0 = fn(n)
        if os.environ.get("XLB_CHUNK_WORDS"):
            line.chunk_words = int(os.environ["XLB_CHUNK_WORDS"])
        for tpi in (int(v) for v in tpis.split(",")):
            best = 0.0
            for rep in range(3):
                p = xl.Particles(p0c=p0c, mass0=m0, **cols)
                if rep == 0:
                    line.track(p, num_turns=2, turns_per_item=tpi, **KW)
                    torch.cuda.synchronize()
                    continue
                line.track(p, num_turns=turns, timed=True, turns_per_item=tpi, **KW)
                st = line.last_stats
                best = max(best, int(p.at_turn.sum()) / (st["kernel_ms"] * 1e-3))
            ok = p.state == 1
            chk = float(p.x[ok].double().sum() + p.py[ok].double().sum() + p.zeta[ok].double().sum())
            row = {"lib": LIBNAME, "shape": SHAPE, "strict": bool(KW.get("strict")), "config": cfg, "n": n, "turns": turns, "turns_per_item": tpi, "ptps": best,
                   "n_chunks": line.pack().n_chunks, "chunk_words": line.pack().chunk_words,
                   "blocks": st["blocks"], "threads": st["threads"], "regs": st["regs_per_thread"],
                   "alive": int(ok.sum()), "checksum": repr(chk)}
            rows.append(row)
            print(json.dumps(row), flush=True)
    old = []
    if os.path.exists(out_path):
        with open(out_path) as fh:
            old = json.load(fh)["rows"]
    with open(out_path, "w") as fh:
        json.dump({"gpu": torch.cuda.get_device_name(0), "rows": old + rows}, fh, indent=1)


if __name__ == "__main__":
    main()
