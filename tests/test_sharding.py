"""Sharding of batches of independent problems: partition arithmetic, and the final gather on a
world_size-2 gloo group (the N > 1 path of bench.py / the batch drivers, minus the GPUs)."""
import os
import socket

import numpy as np
import pytest

from fastest_lap_b200.sharding import shard_bounds, shard_range, owner_of, gather_results


@pytest.mark.parametrize("count,world", [(4096, 1), (4096, 2), (4096, 8), (10, 3), (3, 8), (0, 4), (1, 2)])
def test_contiguous_partition_matches_floor_rule(count, world):
    b = shard_bounds(count, world)
    assert b[0] == 0 and b[-1] == count and all(b[r] <= b[r + 1] for r in range(world))
    for p in range(count):
        r = owner_of(p, count, world)
        assert b[r] <= p < b[r + 1]
    sizes = [b[r + 1] - b[r] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_bad_arguments():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    with pytest.raises(ValueError):
        shard_bounds(10, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, count, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(count, world, rank)
    # per-problem "results": (lap time, iterations) as a function of the global problem index
    local = np.stack([70.0 + np.arange(lo, hi) * 0.25, np.arange(lo, hi) % 7], axis=1)
    full = gather_results(local, count)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("count", [7, 8])
def test_gather_on_two_gloo_ranks(count):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, count, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([70.0 + np.arange(count) * 0.25, np.arange(count) % 7], axis=1)
    for rank, full in got:
        assert full.shape == expect.shape and np.array_equal(full, expect), rank
