"""Build experiment variants of libxline_b200.so: the fast translation units recompiled with
extra -D flags, everything else taken from the regular build.  One library per variant under
xline_b200/exp/ (git-ignored, travels to the GPU box); scripts/probe_variants.py times them.

    python scripts/build_exp.py name1:-DXLB_FOO=1,-DXLB_BAR=1 name2:...
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, ".")
from xline_b200 import build as B

B.build()
OUT = os.path.join(B.HERE, "exp")
os.makedirs(OUT, exist_ok=True)
FAST = [u for u in B.UNITS if u[1] == "track_fast.cu"]
REST = [os.path.join(B.OBJ, u[0] + ".o") for u in B.UNITS if u[1] != "track_fast.cu"]


def one(job):
    name, flags, unit = job
    obj = os.path.join(OUT, "%s_%s.o" % (name, unit[0]))
    cmd = [B._nvcc()] + B.ARCH + B.COMMON + unit[2] + flags + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, unit[1]), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        raise RuntimeError(res.stderr)
    open(obj[:-2] + ".ptxas.txt", "w").write(res.stderr)
    return obj


specs = []
for a in sys.argv[1:]:
    name, _, fl = a.partition(":")
    specs.append((name, [f for f in fl.split(",") if f]))
jobs = [(n, f, u) for n, f in specs for u in FAST]
with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
    objs = list(ex.map(one, jobs))
for n, f in specs:
    mine = [o for (jn, _, _), o in zip(jobs, objs) if jn == n]
    lib = os.path.join(OUT, "lib_%s.so" % n)
    subprocess.run([B._nvcc()] + B.ARCH + ["-shared", "-o", lib] + mine + REST, check=True)
    print(lib, f)
