"""Summarise an .ncu-rep (read with the ncu CLI, no GPU needed): headline metrics, SASS opcode
mix and where the warp-stall samples sit.  Usage: python scripts/ncu_summary.py rep [out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
for row in raw[2:]:
    print("kernel:", row[hdr.index("Kernel Name")], file=out)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("  %-66s %s %s" % (w, row[i], units[i]), file=out)
    st = [(h, float(row[i])) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    print("  stall cycles per issued instruction:", file=out)
    for h, v in sorted(st, key=lambda kv: -kv[1])[:8]:
        print("    %-28s %.3f" % (h.split("stalled_")[1].split("_per_issue")[0], v), file=out)

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
h2 = src[1]
data = src[2:]
ia, isrc, isamp, iex = h2.index("Address"), h2.index("Source"), h2.index("# Samples"), h2.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data) or 1
totex = sum(int(r[iex]) for r in data) or 1
agg = {}
for r in data:
    t = r[isrc].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    a = agg.setdefault(op, [0, 0])
    a[0] += int(r[isamp])
    a[1] += int(r[iex])
fp64 = sum(v[1] for k, v in agg.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
print("  SASS mix (warp-level instructions executed: %d, FP64-pipe share %.1f%%):" % (totex, 100.0 * fp64 / totex), file=out)
for op, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("    %-10s executed %5.1f%%   stall samples %5.1f%%" % (op, 100.0 * e / totex, 100.0 * s / tot), file=out)
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
print("  hottest instructions by stall samples:", file=out)
for r in sorted(data, key=lambda r: -int(r[isamp]))[:14]:
    st = sorted(((h, int(r[h2.index(h)])) for h in stalls), key=lambda kv: -kv[1])[:2]
    print("    %5.1f%%  %-52s %s" % (100.0 * int(r[isamp]) / tot, r[isrc][:52], st), file=out)
