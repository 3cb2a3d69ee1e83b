"""BASELINE config C4: 8-GPU sharded dynamic-aperture scan on the PETRA IV lattice.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29520 scripts/run_c4_sharded.py [n_total] [turns]

100 M start points on a 2-D (x, y) amplitude grid, interleaved over the ranks by index (so
every GPU sees every amplitude band and the shards lose particles at the same rate); one
process per GPU, no data-path collective; NCCL only for the loss tallies and the survivor
counts at the end.  Rank 0 prints one JSON line (also written to gpurun_out/c4_sharded.json).
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xline_b200 as xl  # noqa: E402
from xline_b200 import configs, sharding  # noqa: E402


def main():
    n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    turns = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line, _, p0c, m0 = configs.config_petra4(4)
    line.append_element(xl.LimitEllipse(a=8e-3, b=4e-3), "scraper")
    side = int(np.ceil(np.sqrt(n_total)))
    ids = np.arange(rank, n_total, world, dtype=np.int64)  # interleaved shard
    x = (ids % side) * (6e-3 / (side - 1))
    y = (ids // side) * (3e-3 / (side - 1))
    p = xl.Particles(p0c=p0c, mass0=m0, device=dev, x=x, y=y, particle_id=ids)
    line.track(p, num_turns=1)  # warm-up turn (first touch of the lattice, scratch allocation)
    before = sharding.global_counts(p)[2]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    line.track(p, num_turns=turns, turns_per_launch=10)
    sharding.allreduce_loss_tally(line.loss_tally)
    alive, lost, done = sharding.global_counts(p)
    ev1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    # per-rank device time and work (shard imbalance, if any, shows here)
    per_rank = torch.zeros(world, 3, dtype=torch.float64, device=dev)
    per_rank[rank, 0] = ms[0]
    per_rank[rank, 1] = float(int(p.at_turn.sum()))
    per_rank[rank, 2] = float(int((p.state == 1).sum()))
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank)
    # survival per amplitude band (radial, in units of the grid extent), reduced over ranks
    rr = np.sqrt((x / 6e-3) ** 2 + (y / 3e-3) ** 2)
    band = torch.from_numpy(np.minimum((rr * 10).astype(np.int64), 14)).to(dev)
    tot = torch.bincount(band, minlength=15).to(torch.float64)
    surv = torch.bincount(band, weights=(p.state == 1).to(torch.float64), minlength=15)
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(surv)
    if rank == 0:
        out = {
            "config": "C4: PETRA IV (examples/petra4/h7ba_n8.seq, 31026 elements), %d particles on a 2-D "
                      "amplitude grid, %d turns, %d GPUs, interleaved particle-index shards" % (n_total, turns, world),
            "n_gpus": world, "particles": n_total, "turns": turns,
            "particle_turns_done": int(done - before), "alive": alive, "lost": lost,
            "device_ms_max_over_ranks": float(ms.item()), "wall_s": time.perf_counter() - t0,
            "particle_turns_per_s": (done - before) / (float(ms.item()) * 1e-3),
            "loss_tally_total": int(line.loss_tally.sum().item()),
            "per_rank": [{"rank": r, "device_ms": float(per_rank[r, 0]), "particle_turns_incl_warmup": int(per_rank[r, 1]),
                          "alive": int(per_rank[r, 2])} for r in range(world)],
            "gpu": torch.cuda.get_device_name(dev),
            "survival_by_amplitude_band": [float(s / t) if t > 0 else None for s, t in zip(surv.tolist(), tot.tolist())],
        }
        line_json = json.dumps(out)
        print(line_json)
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/c4_sharded.json", "w") as fh:
            fh.write(line_json + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
