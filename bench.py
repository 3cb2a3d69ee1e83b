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
``--scaling strong`` divides the 1 000 000 particles over the ranks instead.

Printed JSON (one line, rank 0): the contract keys plus
  ``roofline``      FP64 pipe (the path is register-resident arithmetic, neither HBM- nor
                    tensor-bound), against the peak measured in this process AND the nominal one;
  ``strict``        the same workload through the bit-exact kernel (reference operation order,
                    results identical to the NumPy path bit for bit): value + roofline fraction.
                    ``value`` above is the fast kernel (FMA contraction, folded constants: equal
                    to the reference within its own rounding noise, 1e-12 per element);
  ``cpu_baseline``  the NumPy oracle port on this box's host cores;
  ``e2e``           the same metric through the C-ABI host entry point ``xlb_track_host`` with
                    pinned host buffers, copies inside the timed region;
  ``configs``       short measurements of the other BASELINE configurations with their own
                    roofline fraction and clocks (c1 FODO incl. the host entry point, c2_heavy_loss =
                    the SURVEY 8(d) beam, c2_125k = one rank's share of C2 under strong scaling at 8
                    GPUs, c3 LHC + beam-beam, c4 PETRA IV, c5 PS Booster + space charge);
  ``clocks``, ``gpu_launches``.
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
NOMINAL_FP64_PEAK = 148 * 64 * 2 * 1.965e9  # SMs x DFMA/clk/SM x 2 flops x boost clock


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000, help="particles per GPU (total with --scaling strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--turns-per-step", type=int, default=100)
    ap.add_argument("--turns-per-launch", type=int, default=50)
    ap.add_argument("--ppt", type=int, default=0, help="particles per thread (0 = library default for N)")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-particles", type=int, default=5000, help="CPU sample: particles per process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the strict leg and the other configurations")
    ap.add_argument("--extras", default="strict,c1,c2_heavy_loss,c2_125k,c3,c4,c5")
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
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": workload_config(args, n_gpus=args.gpus, n_local=args.particles),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python/NumPy and un-importable (xline/__init__.py:3 raises; xpart absent): "
                "timed here is oracle/xline_oracle.py, the NumPy restatement pinned against the reference's "
                "own element code (tests/golden), one process per host core",
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n_gpus, n_local):
    return {
        "workload": "C2: LHC lattice (examples/lhc, 18657 elements + 7640 apertures), "
                    "%d particles/GPU x %d turns/step, Gaussian beam A~U(0,4)" % (n_local, args.turns_per_step),
        "particles_per_gpu": n_local, "turns_per_step": args.turns_per_step,
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

    def __init__(self, index, period_ms=200):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", str(period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def mark(self):
        return time.perf_counter()

    def summary(self, t0=None, t1=None):
        """Median SM clock, maximum clock and throttle reasons of the samples taken in [t0, t1]."""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in list(self.rows):
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1 + 0.25):
                continue
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

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:  # pragma: no cover
            self.proc.kill()


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of the GPU arm: device, ranks, clock sampler, FP64 peak."""


def _barrier(ctx):
    import torch
    import torch.distributed as dist

    if ctx.world > 1:
        dist.barrier()
    torch.cuda.synchronize(ctx.dev)


def _reduce(ctx, values, op):
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return [float(v) for v in t.tolist()]


def roofline_of(ctx, ops_per_turn, particle_turns, kernel_ms, launches, kernel_name, extra=None):
    achieved = ops_per_turn * particle_turns / (kernel_ms * 1e-3) if kernel_ms else None
    out = {
        "bound": "fp64", "achieved": achieved / 1e12 if achieved else None, "peak": ctx.peak / 1e12,
        "unit": "TFLOP/s", "frac": achieved / ctx.peak if achieved else None,
        "frac_of_nominal": achieved / NOMINAL_FP64_PEAK if achieved else None,
        "peak_nominal": NOMINAL_FP64_PEAK / 1e12,
        "algorithmic_fp64_ops_per_particle_turn": ops_per_turn, "kernel": kernel_name,
        "avg_launch_ms": kernel_ms / max(launches, 1), "particle_turns_per_launch": particle_turns / max(launches, 1),
    }
    if extra:
        out.update(extra)
    return out


def measure(ctx, line, p_factory, turns, steps, warm_turns, name, track_kw=None, all_ranks=False, flush=True):
    """Times `steps` x Line.track(p, num_turns=turns) on fresh particles after a warm-up of
    `warm_turns` turns on a throw-away copy.  Device time by CUDA events (max over ranks when
    `all_ranks`), clocks sampled over the timed region.  Returns a dict for the JSON line."""
    import torch

    track_kw = dict(track_kw or {})
    p = p_factory()
    if warm_turns:
        w = p.copy()
        line.track(w, num_turns=warm_turns, **track_kw)
        del w
    if all_ranks:
        _barrier(ctx)
    else:
        torch.cuda.synchronize(ctx.dev)
    before = int(p.at_turn.sum())
    n0 = int((p.state == 1).sum())
    launches = kernel_ms = 0.0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = ctx.sampler.mark() if ctx.sampler else None
    ev0.record()
    for _ in range(steps):
        if flush:
            ctx.flush.zero_()
        line.track(p, num_turns=turns, timed=True, **track_kw)
        st = line.last_stats
        launches += st["kernel_launches"] + 3 * st["compactions"]
        kernel_ms += st["kernel_ms"]
    ev1.record()
    if all_ranks:
        _barrier(ctx)
    else:
        torch.cuda.synchronize(ctx.dev)
    t_b = ctx.sampler.mark() if ctx.sampler else None
    ms = ev0.elapsed_time(ev1)
    done = int(p.at_turn.sum()) - before
    alive = int((p.state == 1).sum())
    local_done = done
    if all_ranks and ctx.world > 1:
        ms, = _reduce(ctx, [ms], "MAX")
        done, alive, n0 = (int(v) for v in _reduce(ctx, [done, alive, n0], "SUM"))
    ops = line.algorithmic_ops_per_turn()
    st = line.last_stats
    out = {
        "workload": name, "value": done / (ms * 1e-3), "unit": UNIT, "ms": ms, "particles": n0, "survivors": alive,
        "turns": turns * steps, "particle_turns_done": done, "gpu_launches": int(launches),
        "kernel_shape": "%d threads x %d CTAs, %d registers" % (st["threads"], st["blocks"], st["regs_per_thread"]),
        "roofline": roofline_of(ctx, ops, local_done, kernel_ms, launches, "track_kernel"),
        "clocks": ctx.sampler.summary(t_a, t_b) if ctx.sampler else None,
    }
    return out, p


def run_extras(ctx, args, which):
    """The bit-exact leg and the other BASELINE configurations, a few seconds each."""
    import ctypes as C

    import numpy as np
    import torch

    import xline_b200 as xl
    from xline_b200 import _cabi, configs

    out = {}
    rank, world = ctx.rank, ctx.world

    def factory(cols, p0c, m0):
        return lambda: xl.Particles(p0c=p0c, mass0=m0, device=ctx.dev, **cols)

    if "c1" in which and world == 1:
        # C1: FODO cell, 10 k particles x 100 turns -- device entry point and the host entry point
        # (host buffers, copies inside) on the same call size
        line, cols, p0c, m0 = configs.config_fodo(10_000)
        r, _ = measure(ctx, line, factory(cols, p0c, m0), 100, 20, 100, "C1: FODO cell, 10k particles x 100 turns",
                       flush=False)
        hp = xl.Particles(p0c=p0c, mass0=m0, device="cpu", pinned=True, **cols)
        packed = line.pack()
        lat = packed.c_lattice()
        cp = _cabi.Particles()
        cp.n = len(hp)
        for k, t in hp._columns():
            setattr(cp, k, t.data_ptr())
        cp.q0, cp.mass0, cp.p0c = hp.q0, hp.mass0, hp.p0c
        cp.beta0, cp.gamma0, cp.energy0 = hp.beta0, hp.gamma0, hp.energy0
        opts = _cabi.TrackOptions()
        opts.num_turns = 100
        lib = _cabi.lib()
        for _ in range(3):
            _cabi.check(lib.xlb_track_host(C.byref(lat), C.byref(cp), C.byref(opts)))
        before = int(hp.at_turn.sum())
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            _cabi.check(lib.xlb_track_host(C.byref(lat), C.byref(cp), C.byref(opts)))
        dt = time.perf_counter() - t0
        r["e2e"] = {"value": (int(hp.at_turn.sum()) - before) / dt, "unit": UNIT, "ms_per_call": 1e3 * dt / reps,
                    "device_ms_per_call": r["ms"] / 20, "api": "xlb_track_host (pinned host SoA buffers)"}
        r["e2e"]["host_over_device"] = r["e2e"]["ms_per_call"] / r["e2e"]["device_ms_per_call"]
        out["c1"] = r

    if "strict" in which and world == 1:
        line, cols, p0c, m0 = ctx.c2
        r, _ = measure(ctx, line, factory(cols, p0c, m0), args.turns_per_step, 1, 5,
                       "C2 through the bit-exact (strict) kernel: reference operation order, no FMA contraction, "
                       "exactly rounded divisions; results identical to the NumPy path bit for bit",
                       track_kw=dict(strict=True, turns_per_launch=args.turns_per_launch))
        r["frac"] = r["roofline"]["frac"]
        out["strict"] = r

    if "c2_heavy_loss" in which and world == 1:
        # the beam SURVEY.md 8(d) proposes for C2: sigma 3e-4 m, A ~ U(0, 12) -- most of it is outside
        # the aperture and is lost within the first turns
        line = ctx.c2[0]
        cols = configs.gaussian_beam(args.particles, 2, rank, sx=3e-4, spx=3e-6, amp_max=12.0)
        r, _ = measure(ctx, line, factory(cols, ctx.c2[2], ctx.c2[3]), args.turns_per_step, 2, 0,
                       "C2 lattice, SURVEY 8(d) beam (sigma 3e-4 m, A~U(0,12): heavy early losses), "
                       "survivor-weighted", track_kw=dict(turns_per_launch=args.turns_per_launch))
        out["c2_heavy_loss"] = r

    if "c2_125k" in which and world == 1:
        # strong-scaling regime on one GPU: the share of one rank when the 1 M particles of C2 are
        # divided over 8 GPUs (no data-path collective: 8 x this value is what 8 ranks deliver)
        line, cols, p0c, m0 = configs.config_lhc(125_000)
        r, _ = measure(ctx, line, factory(cols, p0c, m0), args.turns_per_step, 3, 5,
                       "C2 lattice, 125 000 particles (1/8 of C2: one rank's share under strong scaling at 8 GPUs)",
                       track_kw=dict(turns_per_launch=args.turns_per_launch))
        out["c2_125k"] = r

    if "c3" in which and world == 1:
        line, cols, p0c, m0 = configs.config_lhc_beambeam(4_000_000)
        r, _ = measure(ctx, line, factory(cols, p0c, m0), 20, 1, 1,
                       "C3: LHC + 72 BeamBeam4D + 2 BeamBeam6D (15 slices), 4M particles x 20 turns "
                       "(full size 1e7 x 1e3: scripts/run_c3_full.py)")
        out["c3"] = r

    if "c4" in which:
        n4 = 12_500_000
        line, _, p0c, m0 = configs.config_petra4(4)
        line.append_element(xl.LimitEllipse(a=8e-3, b=4e-3), "scraper")
        n_total = n4 * world
        side = int(np.ceil(np.sqrt(n_total)))
        ids = np.arange(rank, n_total, world, dtype=np.int64)  # interleaved shard of the amplitude grid
        cols = dict(x=(ids % side) * (6e-3 / (side - 1)), y=(ids // side) * (3e-3 / (side - 1)), particle_id=ids)
        r, _ = measure(ctx, line, factory(cols, p0c, m0), 6, 1, 1,
                       "C4: PETRA IV dynamic-aperture scan, 12.5M grid points per GPU x 6 turns "
                       "(100M on 8 GPUs; full length: scripts/run_c4_sharded.py)", all_ranks=True,
                       track_kw=dict(turns_per_launch=3))
        out["c4"] = r

    if "c5" in which and world == 1:
        line, cols, p0c, m0 = configs.config_psb(1_000_000, monitor_stores=4, monitor_ids=100_000, monitor_skip=100)
        r, _ = measure(ctx, line, factory(cols, p0c, m0), 400, 1, 5,
                       "C5: PS Booster + 120 SCQGaussProfile kicks + BeamMonitor, 1M particles x 400 turns "
                       "(full size 1e6 x 1e4: scripts/run_c5_psb.py)")
        # the convention counts a wofz call as 100 operations; the Weideman evaluation (synthetic
        # division by a real quadratic: 34 x 2 FMA + set-up and remainder) spends ~170
        nsc = sum(1 for e in line.elements if type(e).__name__.startswith("SC"))
        r["roofline"]["executed_flops_estimate_per_particle_turn"] = r["roofline"][
            "algorithmic_fp64_ops_per_particle_turn"] + nsc * 2 * 70
        r["roofline"]["frac_executed_flops_estimate"] = (
            r["roofline"]["frac"] * r["roofline"]["executed_flops_estimate_per_particle_turn"]
            / r["roofline"]["algorithmic_fp64_ops_per_particle_turn"])
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            with open(prof) as fh:
                r["roofline"]["ncu"] = json.load(fh).get("c5")
        out["c5"] = r
    return out


def run_b200(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import xline_b200 as xl
    from xline_b200 import _cabi, configs, sharding

    ctx = Ctx()
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    ctx.dev = dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.scaling == "strong":
        lo, hi = sharding.shard_bounds(args.particles, rank, world)
        n, first_id = hi - lo, lo
    else:
        n, first_id = args.particles, rank * args.particles
    T = args.turns_per_step
    line, cols, p0c, m0 = configs.config_lhc(n, rank=rank, first_id=first_id)
    ctx.c2 = (line, cols, p0c, m0)
    ops_per_turn = line.algorithmic_ops_per_turn()
    p = xl.Particles(p0c=p0c, mass0=m0, device=dev, **cols)
    ctx.flush = flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    ctx.sampler = ClockSampler(local_rank) if rank == 0 else None
    ctx.peak, _ = _cabi.measure_fp64_peak(5)

    launches = {"track": 0, "compact": 0}
    kernel_ms = []
    shape = {}

    def step(timed=False):
        flush.zero_()  # > L2 (126 MB): nothing of the previous step stays cached
        line.track(p, num_turns=T, turns_per_launch=args.turns_per_launch,
                   particles_per_thread=args.ppt, threads_per_block=args.threads, timed=timed)
        st = line.last_stats
        launches["track"] += st["kernel_launches"]
        launches["compact"] += 3 * st["compactions"]
        shape.update(threads=st["threads"], blocks=st["blocks"], regs=st["regs_per_thread"])
        if timed:
            kernel_ms.append(st["kernel_ms"])

    for _ in range(args.warmup):
        step()
    _barrier(ctx)
    turns_before = p.at_turn.sum().item()
    launches["track"] = launches["compact"] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(ctx)
    t_a = ctx.sampler.mark() if ctx.sampler else None
    ev0.record()
    for _ in range(args.steps):
        step(timed=True)
    ev1.record()
    _barrier(ctx)
    t_b = ctx.sampler.mark() if ctx.sampler else None
    clocks = ctx.sampler.summary(t_a, t_b) if ctx.sampler else None
    elapsed_ms = ev0.elapsed_time(ev1)
    local_done = p.at_turn.sum().item() - turns_before
    done = local_done
    alive = int((p.state == 1).sum().item())
    n_lost_local = int((p.state != 1).sum().item())
    if world > 1:
        elapsed_ms, = _reduce(ctx, [elapsed_ms], "MAX")
        done, alive, n_lost = (int(v) for v in _reduce(ctx, [done, alive, n_lost_local], "SUM"))
    else:
        n_lost = n_lost_local
    value = done / (elapsed_ms * 1e-3)
    # loss tallies: one SUM all-reduce of a COPY after the loop (the line's own tally keeps
    # accumulating this rank's losses); every lost particle is in exactly one bin
    tally = sharding.allreduce_loss_tally(line.loss_tally.clone())
    tally_sum = int(tally.sum().item())
    if tally_sum != n_lost:
        raise RuntimeError("loss tallies (%d) do not add up to the lost particles (%d)" % (tally_sum, n_lost))

    # ---- end-to-end through the C-ABI host entry point (host buffers, copies timed)
    e2e = None
    if args.e2e_steps > 0:
        hp = xl.Particles(p0c=p0c, mass0=m0, device="cpu", pinned=True, **cols)
        packed = line.pack()
        lat = packed.c_lattice()
        cp = _cabi.Particles()
        cp.n = len(hp)
        ncols_in = 0
        for k, t in hp._columns():
            setattr(cp, k, t.data_ptr())
            ncols_in += 1
        cp.q0, cp.mass0, cp.p0c = hp.q0, hp.mass0, hp.p0c
        cp.beta0, cp.gamma0, cp.energy0 = hp.beta0, hp.gamma0, hp.energy0
        tally_h = torch.zeros(packed.n_elements, dtype=torch.int64).pin_memory()
        opts = _cabi.TrackOptions()
        opts.num_turns, opts.turns_per_launch = T, args.turns_per_launch
        opts.particles_per_thread, opts.threads_per_block = args.ppt, args.threads
        opts.loss_tally = tally_h.data_ptr()
        lib = _cabi.lib()
        _cabi.check(lib.xlb_track_host(C.byref(lat), C.byref(cp), C.byref(opts)))  # warm-up (arena alloc)
        _barrier(ctx)
        before = int(hp.at_turn.sum())
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            _cabi.check(lib.xlb_track_host(C.byref(lat), C.byref(cp), C.byref(opts)))
        _barrier(ctx)
        dt = time.perf_counter() - t0
        e_done = float(int(hp.at_turn.sum()) - before)
        if world > 1:
            dt, = _reduce(ctx, [dt], "MAX")
            e_done, = _reduce(ctx, [e_done], "SUM")
        h2d = packed.nbytes + ncols_in * 8 * n + packed.n_elements * 8
        d2h = 12 * 8 * n + packed.n_elements * 8
        e2e = {"value": e_done / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
               "api": "xlb_track_host (C ABI, pinned host SoA buffers)"}
        del hp

    # ---- the bit-exact leg and the other configurations
    del p
    extras = {}
    if not args.no_extras:
        which = [w for w in args.extras.split(",") if w]
        extras = run_extras(ctx, args, which)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline: FP64 pipe, measured peak (register-resident DFMA chains, this GPU, now)
    traffic = None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    ncu_note = None
    if os.path.exists(prof):
        with open(prof) as fh:
            pj = json.load(fh)
        traffic, ncu_note = pj.get("dram_bytes_per_launch"), pj
    roofline = roofline_of(
        ctx, ops_per_turn, local_done, sum(kernel_ms), launches["track"],
        "track_kernel (%s threads x %s CTAs, %s registers)" % (shape.get("threads"), shape.get("blocks"),
                                                               shape.get("regs")),
        extra={
            "traffic": traffic,
            "peak_source": "measured: xlb_measure_fp64_peak (8 independent DFMA chains/thread), same process; "
                           "MEASURED_PEAKS.json has no FP64 entry; peak_nominal = 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz",
            "hbm_view": {"bytes_per_launch_algorithmic": 188 * n * args.turns_per_launch,
                         "note": "particle state is register-resident inside a work item (one turn of the LHC "
                                 "lattice: 121 chunks); per particle and item 84 B are loaded and 104 B stored "
                                 "(one species: no chi column) -- through L2, where the beam (1 M x 120 B) stays "
                                 "from one item to the next: see ncu.dram_bytes_per_launch"},
            "ncu": ncu_note,
        })

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
    if ctx.sampler:
        ctx.sampler.stop()

    strict = extras.pop("strict", None)
    line_out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / max(args.steps, 1), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world, n), "impl": "b200",
        "kernel_mode": "fast (FMA contraction, folded constants; 1e-12 per element, rounding-noise level over a "
                       "turn); the bit-exact kernel is reported under 'strict'",
        "survivors": alive, "particle_turns_done": done, "loss_tally_sum": tally_sum,
        "roofline": roofline, "strict": strict, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
        "configs": extras,
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
