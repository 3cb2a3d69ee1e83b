"""Per-region view of an ncu source page (SASS): consecutive instructions with the same
execution count are one region; prints executed count, instruction mix and stall samples.
    ncu -i rep --page source --csv --print-source sass > src.csv ; python scripts/ncu_regions.py src.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
data = rows[2:]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[isamp]) for r in data)
totex = sum(int(r[iex]) for r in data)
base = int(data[0][ia], 16)
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
regions = []
cur = None
for r in data:
    ex = int(r[iex])
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    isbr = op in ("BRA", "BRX", "EXIT", "BSYNC", "CALL", "RET")
    if cur is None or abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"], 1):
        cur = {"ex": ex, "rows": []}
        regions.append(cur)
    cur["rows"].append(r)
    if isbr:
        cur = None
print("total samples %d, executed %d" % (tot, totex))
FP = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")
for reg in regions:
    s = sum(int(r[isamp]) for r in reg["rows"])
    e = sum(int(r[iex]) for r in reg["rows"])
    if 100.0 * s / tot < minshare:
        continue
    nfp = 0
    for r in reg["rows"]:
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        nfp += op in FP
    a0 = int(reg["rows"][0][ia], 16) - base
    st = {}
    for c in stalls:
        st[c] = sum(int(r[h.index(c)]) for r in reg["rows"])
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print("@%05x n=%3d fp64=%3d exec/instr=%11d  samples %5.2f%%  exec %5.2f%%  samples/exec ratio %.2f  %s" % (
        a0, len(reg["rows"]), nfp, reg["ex"], 100.0 * s / tot, 100.0 * e / totex, (s / tot) / (e / totex) if e else 0,
        " ".join("%s=%.0f%%" % (k[6:], 100.0 * v / max(s, 1)) for k, v in top)))
