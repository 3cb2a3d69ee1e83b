"""N>1 host logic on CPU: two gloo ranks exercise the sharding helpers (index blocks,
loss-tally all-reduce, monitor merge, ragged column gather) that the NCCL path uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xline_b200 import sharding


def test_shard_bounds_cover_and_balance():
    for n, w in ((10, 3), (1_000_000, 8), (7, 8), (0, 2)):
        blocks = [sharding.shard_bounds(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import xline_b200 as xl

        rng = np.random.default_rng(3)
        cols = dict(x=rng.normal(0, 1, n_total), y=rng.normal(0, 1, n_total))
        mine = sharding.shard_columns(cols, rank, world)
        lo, hi = sharding.shard_bounds(n_total, rank, world)
        assert np.array_equal(mine["particle_id"], np.arange(lo, hi))
        p = xl.Particles(p0c=1e9, device="cpu", **mine)
        # pretend a kernel ran: lose the particles with |x| > 1 at element 5, others did 3 turns
        lost = p.x.abs() > 1
        p.state[lost] = 0
        p.at_element[lost] = 5
        p.at_turn[~lost] = 3
        tally = torch.zeros(8, dtype=torch.int64)
        tally[5] = int(lost.sum())
        sharding.allreduce_loss_tally(tally)
        alive, dead, turns = sharding.global_counts(p)
        full = sharding.gather_columns(p, names=("x", "state", "particle_id"))
        # monitor slab: each rank fills only its own particle columns
        slab = torch.full((2, n_total), float("nan"), dtype=torch.float64)
        slab[:, lo:hi] = p.x.unsqueeze(0) * torch.tensor([[1.0], [2.0]], dtype=torch.float64)
        merged = sharding.merge_monitor(slab)
        if rank == 0:
            want_lost = int((np.abs(cols["x"]) > 1).sum())
            assert int(tally[5]) == want_lost and int(tally.sum()) == want_lost
            assert alive == n_total - want_lost and dead == want_lost and turns == 3 * alive
            assert np.array_equal(full["x"].numpy(), cols["x"])
            assert np.array_equal(full["particle_id"].numpy(), np.arange(n_total))
            assert np.array_equal(merged[0].numpy(), cols["x"]) and np.array_equal(merged[1].numpy(), 2 * cols["x"])
            out.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(150)
        assert pr.exitcode == 0
    assert out.get() == "ok"
