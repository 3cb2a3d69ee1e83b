"""Per-instruction residence time from an ncu source page: cycles a warp spends at each SASS
instruction per execution (stall samples scaled by warp-cycles per sample).
    python scripts/ncu_path.py src.csv warps_per_sm n_sm cycles [lo_addr hi_addr]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
data = rows[2:]
warp_cycles = float(sys.argv[2]) * float(sys.argv[3]) * float(sys.argv[4])
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[isamp]) for r in data)
per = warp_cycles / tot
base = int(data[0][ia], 16)
lo = int(sys.argv[5], 16) if len(sys.argv) > 5 else 0
hi = int(sys.argv[6], 16) if len(sys.argv) > 6 else 1 << 30
for r in data:
    a = int(r[ia], 16) - base
    if a < lo or a > hi:
        continue
    e, s = int(r[iex]), int(r[isamp])
    st = sorted(((c[6:], int(r[h.index(c)])) for c in stalls), key=lambda kv: -kv[1])[:2]
    print("%05x %-46s ex=%10d cyc/ex=%6.1f  %s" % (a, r[isrc].strip()[:46], e, s * per / e if e else 0,
                                                   " ".join("%s:%d" % kv for kv in st if kv[1])))
