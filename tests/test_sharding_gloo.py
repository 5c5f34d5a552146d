"""Multi-rank host logic on CPU (gloo, world_size 2): rotation-list sharding and the single end-of-build
collective.  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffsims_b200.library import active_quaternions, gather_counts, shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 300001, 1 << 20):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_active_quaternions_conjugate():
    from diffsims_b200.crystal import Rotation
    r = Rotation.random(5, rng=1)
    q = active_quaternions(r)
    np.testing.assert_allclose(q, (~r).data)
    np.testing.assert_allclose(active_quaternions(r.data), (~r).data)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 11
        lo, hi = shard_bounds(n, rank, world)
        local = torch.arange(lo, hi, dtype=torch.int32) * 10  # pretend per-template spot counts
        allc = gather_counts(local)
        q.put((rank, allc.tolist()))
    finally:
        dist.destroy_process_group()


def test_gather_counts_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [i * 10 for i in range(11)]
    assert res[0] == expect and res[1] == expect
