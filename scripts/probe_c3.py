"""Throughput probe of BASELINE config C3 (LHC + 72 BeamBeam4D + 2 BeamBeam6D lenses)."""
import sys

import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import _cabi, configs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
turns = int(sys.argv[2]) if len(sys.argv) > 2 else 5
which = sys.argv[3] if len(sys.argv) > 3 else "c3"
line, cols, p0c, m0 = {"c3": configs.config_lhc_beambeam, "c4": configs.config_petra_like,
                       "c5": configs.config_psb, "c5like": configs.config_psb_like,
                       "c4r": configs.config_petra4}[which](n)
ops = line.algorithmic_ops_per_turn()
fl, _ = _cabi.measure_fp64_peak(3)
print("elements", len(line), "alg ops/turn", ops, "records", line.pack().record_counts)
for ppt, thr in ((1, 256), (2, 256), (3, 128)):
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=1, particles_per_thread=ppt, threads_per_block=thr)
    torch.cuda.synchronize()
    p = xl.Particles(p0c=p0c, mass0=m0, **cols)
    line.track(p, num_turns=turns, particles_per_thread=ppt, threads_per_block=thr, timed=True)
    st = line.last_stats
    ptps = int(p.at_turn.sum()) / (st["kernel_ms"] * 1e-3)
    print("ppt=%d thr=%d regs=%d ms=%.1f  %.3e p-t/s  %.2f TFLOP/s(alg) frac=%.3f alive=%d" % (
        ppt, thr, st["regs_per_thread"], st["kernel_ms"], ptps, ptps * ops / 1e12, ptps * ops / fl,
        int((p.state == 1).sum())))
