"""Quick GPU probe: FP64 peak + particle-turns/s of the LHC config across kernel variants."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import xline_b200 as xl
from xline_b200 import _cabi, configs


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 303104
    turns = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    print("variants", json.dumps(_cabi.kernel_variants()))
    fl, ms = _cabi.measure_fp64_peak(5)
    print("fp64 peak TFLOP/s %.2f (%.3f ms)" % (fl / 1e12, ms))
    line, cols, p0c, m0 = configs.config_lhc(n)
    ops = line.algorithmic_ops_per_turn()
    print("elements", len(line), "alg ops/turn", ops)
    for strict in (False,):
        for ppt, thr in ((1, 256), (1, 512), (2, 128), (2, 256), (3, 128), (4, 128), (4, 160)):
            p = xl.Particles(p0c=p0c, mass0=m0, **cols)
            try:
                line.track(p, num_turns=1, particles_per_thread=ppt, threads_per_block=thr, strict=strict)
                torch.cuda.synchronize()
                p = xl.Particles(p0c=p0c, mass0=m0, **cols)
                line.track(p, num_turns=turns, particles_per_thread=ppt, threads_per_block=thr,
                           strict=strict, timed=True)
            except Exception as e:  # noqa: BLE001
                print("ppt", ppt, "thr", thr, "failed:", e)
                continue
            st = line.last_stats
            ptps = n * turns / (st["kernel_ms"] * 1e-3)
            print("strict=%d ppt=%d thr=%d regs=%d blocks=%d ms=%.1f  %.3e p-t/s  %.2f TFLOP/s(alg) frac=%.3f alive=%d"
                  % (strict, ppt, thr, st["regs_per_thread"], st["blocks"], st["kernel_ms"], ptps,
                     ptps * ops / 1e12, ptps * ops / fl, int((p.state == 1).sum())))


if __name__ == "__main__":
    main()
