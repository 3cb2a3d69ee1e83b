#!/usr/bin/env python
"""Benchmark of the ``Line.track`` particle push on B200 (BASELINE.json metric:
particle-turns/s on the LHC lattice, fp64).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (``config.workload``): BASELINE config C2 -- the LHC lattice of the reference's
``examples/lhc`` (18 657 elements + 7 640 synthetic apertures), 1 000 000 Gaussian particles
per GPU with amplitude scale A ~ U(0, 4), tracked ``--turns-per-step`` turns per step
(defaults: 10 steps x 100 turns = the 1000 turns of C2).  A "step" is one
``Line.track(p, num_turns=T)`` over the resident particle set.  ``value`` = particle-turns
actually tracked (sum over particles of the turns they survived) / device time, whole job.

Printed JSON (one line, rank 0): the contract keys plus ``roofline`` (FP64 pipe: the path is
register-resident arithmetic, neither HBM- nor tensor-bound), ``cpu_baseline`` (the NumPy
oracle port on this box's host cores), ``e2e`` (the same metric through the C-ABI host
entry point ``xlb_track_host`` with pinned host buffers, copies inside the timed region),
``clocks`` and ``gpu_launches``.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-turns/s on LHC lattice (fp64)"
UNIT = "particle-turns/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU")
    ap.add_argument("--turns-per-step", type=int, default=100)
    ap.add_argument("--turns-per-launch", type=int, default=50)
    ap.add_argument("--ppt", type=int, default=3)
    ap.add_argument("--threads", type=int, default=128)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-particles", type=int, default=5000, help="CPU sample: particles per process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# CPU leg: the oracle port (NumPy restatement of the reference's element code) on host cores
# ------------------------------------------------------------------------------------------
def _cpu_worker(job):
    """One process: track `n` particles through the C2 lattice for `turns` turns on the oracle."""
    rank, n, turns = job
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np

    from oracle import xline_oracle as xo
    from xline_b200 import configs

    line, cols, p0c, m0 = configs.config_lhc(n, rank=1000 + rank)
    cols.pop("particle_id", None)
    specs = line.to_specs()
    p = xo.OracleParticles(n, p0c=p0c, mass0=m0, **cols)
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        xo.line_track(specs, p, turns)
    dt = time.perf_counter() - t0
    full = xo.gather_full(p, n)
    return int(full["at_turn"].sum()), dt


_POOL = None


def cpu_step(n_per_proc, turns, procs):
    """One bounded CPU sample: `procs` processes x `n_per_proc` particles x `turns` turns.
    Returns (particle-turns done, wall seconds of the tracking part)."""
    global _POOL
    import multiprocessing as mp

    if _POOL is None:
        import atexit

        _POOL = mp.get_context("spawn").Pool(procs)
        atexit.register(lambda: (_POOL.close(), _POOL.join()))
        _POOL.map(_cpu_worker, [(r, 16, 1) for r in range(procs)])  # import + lattice load
    t0 = time.perf_counter()
    res = _POOL.map(_cpu_worker, [(r, n_per_proc, turns) for r in range(procs)])
    wall = time.perf_counter() - t0
    # lattice loading/packing is setup, not tracking: charge the slowest worker's tracking time
    track = max(r[1] for r in res)
    return sum(r[0] for r in res), min(wall, track) if track > 0 else wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = host_cores()
    n = args.cpu_particles
    for _ in range(max(args.warmup, 0)):
        cpu_step(max(n // 8, 64), 1, procs)
    done, secs = 0, 0.0
    for _ in range(args.steps):
        d, s = cpu_step(n, 1, procs)
        done += d
        secs += s
    value = done / secs
    sample = "%d processes x %d particles x 1 turn per step, C2 lattice (LHC + apertures)" % (procs, n)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(args, n_gpus=args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python/NumPy and un-importable (xline/__init__.py:3 raises; xpart absent): "
                "timed here is oracle/xline_oracle.py, the NumPy restatement pinned against the reference's "
                "own element code (tests/golden), one process per host core",
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n_gpus):
    return {
        "workload": "C2: LHC lattice (examples/lhc, 18657 elements + 7640 apertures), "
                    "%d particles/GPU x %d turns/step, Gaussian beam A~U(0,4)" % (args.particles, args.turns_per_step),
        "particles_per_gpu": args.particles, "turns_per_step": args.turns_per_step,
        "turns_per_launch": args.turns_per_launch, "parallelism": "particle-index shards x%d" % n_gpus,
        "l2_flush_between_steps": True,
    }


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi while the timed region runs)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:  # pragma: no cover
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.rows:
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0]))
                mx.append(float(t[1]))
                pw.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(names, t[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    import xline_b200 as xl
    from xline_b200 import _cabi, configs, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.particles
    T = args.turns_per_step
    line, cols, p0c, m0 = configs.config_lhc(n, rank=rank, first_id=rank * n)
    ops_per_turn = line.algorithmic_ops_per_turn()
    p = xl.Particles(p0c=p0c, mass0=m0, device=dev, **cols)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    launches = {"track": 0, "compact": 0}
    kernel_ms = []

    def step(timed=False):
        flush.zero_()  # > L2 (126 MB): nothing of the previous step stays cached
        line.track(p, num_turns=T, turns_per_launch=args.turns_per_launch,
                   particles_per_thread=args.ppt, threads_per_block=args.threads, timed=timed)
        st = line.last_stats
        launches["track"] += st["kernel_launches"]
        launches["compact"] += 3 * st["compactions"]
        if timed:
            kernel_ms.append(st["kernel_ms"])
        if world > 1:
            sharding.allreduce_loss_tally(line.loss_tally)

    for _ in range(args.warmup):
        step()
    barrier()
    turns_before = p.at_turn.sum().item()
    launches["track"] = launches["compact"] = 0
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step(timed=True)
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = ev0.elapsed_time(ev1)
    done = p.at_turn.sum().item() - turns_before
    alive = int((p.state == 1).sum().item())
    stats = torch.tensor([elapsed_ms, float(done), float(alive), float(sum(kernel_ms))],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, done, alive = float(mx[0]), float(sm[1]), int(sm[2])
    value = done / (elapsed_ms * 1e-3)

    # ---- end-to-end through the C-ABI host entry point (host buffers, copies timed)
    e2e = None
    if args.e2e_steps > 0:
        hp = xl.Particles(p0c=p0c, mass0=m0, device="cpu", pinned=True, **cols)
        packed = line.pack()
        lat = _cabi.Lattice(packed.words.ctypes.data, packed.words.size, packed.chunk_words,
                            packed.n_chunks, packed.n_elements, packed.flags)
        cp = _cabi.Particles()
        cp.n = len(hp)
        ncols_in = 0
        for k, t in hp._columns():
            setattr(cp, k, t.data_ptr())
            ncols_in += 1
        cp.q0, cp.mass0, cp.p0c = hp.q0, hp.mass0, hp.p0c
        cp.beta0, cp.gamma0, cp.energy0 = hp.beta0, hp.gamma0, hp.energy0
        tally = torch.zeros(packed.n_elements, dtype=torch.int64).pin_memory()
        opts = _cabi.TrackOptions()
        opts.num_turns, opts.turns_per_launch = T, args.turns_per_launch
        opts.particles_per_thread, opts.threads_per_block = args.ppt, args.threads
        opts.loss_tally = tally.data_ptr()
        lib = _cabi.lib()
        _cabi.check(lib.xlb_track_host(C.byref(lat), C.byref(cp), C.byref(opts)))  # warm-up (arena alloc)
        barrier()
        before = int(hp.at_turn.sum())
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            _cabi.check(lib.xlb_track_host(C.byref(lat), C.byref(cp), C.byref(opts)))
        barrier()
        dt = time.perf_counter() - t0
        e_done = float(int(hp.at_turn.sum()) - before)
        e_stats = torch.tensor([dt, e_done], dtype=torch.float64, device=dev)
        if world > 1:
            a = e_stats.clone()
            dist.all_reduce(a, op=dist.ReduceOp.MAX)
            b = e_stats.clone()
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
            dt, e_done = float(a[0]), float(b[1])
        h2d = packed.nbytes + ncols_in * 8 * n + packed.n_elements * 8
        d2h = 12 * 8 * n + packed.n_elements * 8
        e2e = {"value": e_done / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
               "api": "xlb_track_host (C ABI, pinned host SoA buffers)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline: FP64 pipe, measured peak (register-resident DFMA chains, this GPU, now)
    peak_flops, _ = _cabi.measure_fp64_peak(5)
    per_launch_ms = sum(kernel_ms) / max(launches["track"], 1)
    local_done = p.at_turn.sum().item() - turns_before
    achieved = ops_per_turn * local_done / (sum(kernel_ms) * 1e-3) if kernel_ms else None
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    ncu_note = None
    if os.path.exists(prof):
        with open(prof) as fh:
            pj = json.load(fh)
        traffic, ncu_note = pj.get("dram_bytes_per_launch"), pj
    roofline = {
        "bound": "fp64", "achieved": achieved / 1e12 if achieved else None, "peak": peak_flops / 1e12,
        "unit": "TFLOP/s", "frac": (achieved / peak_flops) if achieved else None, "traffic": traffic,
        "peak_source": "measured: xlb_measure_fp64_peak (8 independent DFMA chains/thread), same process; "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "algorithmic_fp64_ops_per_particle_turn": ops_per_turn,
        "kernel": "track_kernel<ppt=%d>" % args.ppt, "avg_launch_ms": per_launch_ms,
        "particle_turns_per_launch": local_done / max(launches["track"], 1),
        "hbm_view": {"bytes_per_launch_algorithmic": 196 * n * max(1, -(-args.turns_per_launch // 5)),
                     "note": "particle state is register-resident inside a work item; per particle and 5-turn "
                             "item 92 B are loaded and 104 B stored"},
        "ncu": ncu_note,
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        procs = host_cores()
        cpu_step(64, 1, procs)
        d, s = cpu_step(args.cpu_particles, 1, procs)
        d1, s1 = _cpu_worker((0, args.cpu_particles, 1))  # one process, one core
        cpu_baseline = {"value": d / s, "unit": UNIT, "cores": procs, "kind": "port",
                        "sample": "%d processes x %d particles x 1 turn, same C2 lattice and beam recipe"
                                  % (procs, args.cpu_particles),
                        "single_core_value": d1 / s1}

    line_out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world), "impl": "b200",
        "survivors": alive, "particle_turns_done": done,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
        "gpu_launches": launches["track"] + launches["compact"],
        "gpu_launches_detail": launches,
    }
    print(json.dumps(line_out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    # Exactly one JSON line on stdout: libraries (NCCL's version banner, ...) write to fd 1
    # behind Python's back, so fd 1 is pointed at stderr for the run and the line goes to a
    # duplicate of the original stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if args.impl == "reference":
            return run_reference(args)
        return run_b200(args)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    sys.exit(main())
