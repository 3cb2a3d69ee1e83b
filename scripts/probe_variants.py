"""Time experiment libraries (scripts/build_exp.py) against the regular one: each library in a
fresh process, same seeded beams, throughput + a checksum of the final coordinates.

    python scripts/probe_variants.py [--configs c2,c4r,c3,c5] name1 name2 ...   ("base" = regular library)
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def child(libname, cfgs):
    sys.path.insert(0, ROOT)
    import torch

    from xline_b200 import _cabi

    if libname != "base":
        _cabi.LIB_PATH = os.path.join(ROOT, "xline_b200", "exp", "lib_%s.so" % libname)
    import xline_b200 as xl
    from xline_b200 import configs

    table = {"c2": (configs.config_lhc, 1_000_000, 40), "c3": (configs.config_lhc_beambeam, 1_000_000, 10),
             "c4r": (configs.config_petra4, 1_000_000, 40), "c5": (configs.config_psb, 1_000_000, 200)}
    for c in cfgs:
        fn, n, turns = table[c]
        n = int(os.environ.get("XLB_PROBE_N", n))  # particles (default: 1 M)
        line, cols, p0c, m0 = fn(n)
        # XLB_PROBE_SHAPES="3x128,2x256": particles per thread x threads per block (default: library's choice)
        for shape in os.environ.get("XLB_PROBE_SHAPES", "0x0").split(","):
            ppt, thr = (int(v) for v in shape.split("x"))
            kw = dict(particles_per_thread=ppt, threads_per_block=thr) if ppt else {}
            best = 0.0
            for rep in range(2):
                p = xl.Particles(p0c=p0c, mass0=m0, **cols)
                if rep == 0:
                    line.track(p, num_turns=2, **kw)
                    torch.cuda.synchronize()
                    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
                line.track(p, num_turns=turns, timed=True, **kw)
                st = line.last_stats
                best = max(best, int(p.at_turn.sum()) / (st["kernel_ms"] * 1e-3))
            ok = p.state == 1
            chk = float(p.x[ok].double().sum() + p.py[ok].double().sum() + p.zeta[ok].double().sum())
            print(json.dumps({"lib": libname, "config": c, "shape": shape, "ptps": best, "regs": st["regs_per_thread"],
                              "alive": int(ok.sum()), "nan_alive": int(torch.isnan(p.x[ok]).sum()),
                              "checksum": repr(chk)}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3].split(","))
    else:
        args = sys.argv[1:]
        cfgs = "c2"
        if args[0] == "--configs":
            cfgs = args[1]
            args = args[2:]
        for name in args:
            r = subprocess.run([sys.executable, __file__, "--child", name, cfgs], capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode:
                sys.stdout.write("FAILED %s: %s\n" % (name, r.stderr[-800:]))
            sys.stdout.flush()
