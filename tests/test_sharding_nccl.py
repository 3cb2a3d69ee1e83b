"""N>1 path on hardware: two NCCL ranks (one per GPU) track their index shards of one beam through
a line with apertures and a BeamMonitor, then all-reduce the loss tallies, merge the monitor slabs
and gather the columns -- and the result equals the single-GPU run of the whole beam bit for bit.
Skipped with fewer than two devices (run it with `gpurun --gpus 2`); the same helpers are covered
on CPU by tests/test_sharding_gloo.py."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _line_and_beam(n):
    import xline_b200 as xl
    from xline_b200 import configs

    base, cols, p0c, m0 = configs.config_fodo(n)
    mon = xl.BeamMonitor(num_stores=4, start=0, skip=1, min_particle_id=0, max_particle_id=n - 1)
    line = xl.Line(list(base.elements) + [xl.LimitEllipse(a=2.5e-3, b=2.5e-3), mon,
                                           xl.LimitRect(min_x=-2e-3, max_x=2e-3, min_y=-2e-3, max_y=2e-3)])
    return line, mon, cols, p0c, m0


def _worker(rank, world, port, n, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import xline_b200 as xl
        from xline_b200 import sharding

        line, mon, cols, p0c, m0 = _line_and_beam(n)
        mine = sharding.shard_columns(cols, rank, world)
        p = xl.Particles(p0c=p0c, mass0=m0, device=dev, **mine)
        line.track(p, num_turns=4)
        tally = sharding.allreduce_loss_tally(line.loss_tally.clone())
        alive, lost, turns = sharding.global_counts(p)
        merged = {k: sharding.merge_monitor(v) for k, v in mon.data.items()}
        full = sharding.gather_columns(p)
        if rank == 0:
            out.put({"tally": tally.cpu().numpy(), "counts": (alive, lost, turns),
                     "monitor": {k: v.cpu().numpy() for k, v in merged.items()},
                     "full": {k: v.cpu().numpy() for k, v in full.items()}})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_nccl_tally_monitor_and_gather_equal_the_single_gpu_run():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp

    import xline_b200 as xl

    n = 20_001  # ragged shards
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = out.get()
    for pr in procs:
        pr.join(200)
        assert pr.exitcode == 0
    line, mon, cols, p0c, m0 = _line_and_beam(n)
    p = xl.Particles(p0c=p0c, mass0=m0, device="cuda:0", particle_id=np.arange(n), **cols)
    line.track(p, num_turns=4)
    want = p.to_numpy()
    n_lost = int((want["state"] == 0).sum())
    assert 0 < n_lost < n
    assert np.array_equal(got["tally"], line.loss_tally.cpu().numpy()) and int(got["tally"].sum()) == n_lost
    assert got["counts"] == (n - n_lost, n_lost, int(want["at_turn"].sum()))
    for k in ("x", "px", "y", "py", "zeta", "delta", "state", "at_element", "at_turn", "particle_id"):
        assert np.array_equal(got["full"][k], want[k], equal_nan=True), k
    for k, v in mon.data.items():
        assert np.array_equal(got["monitor"][k], v.cpu().numpy(), equal_nan=True), k
    assert int(np.isfinite(got["monitor"]["x"]).sum()) > n
