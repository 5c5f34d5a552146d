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


# ------------------------------------------------------------------------------------------------------------
# sharded multi-phase library: split by cost, gather of the packed spot lists
# ------------------------------------------------------------------------------------------------------------
def test_split_work_covers_every_unit_once_and_balances():
    from diffsims_b200.library import split_work
    rng = np.random.default_rng(0)
    for _ in range(300):
        n_ph, world = int(rng.integers(1, 5)), int(rng.integers(1, 9))
        counts = [int(rng.integers(0, 60)) for _ in range(n_ph)]
        costs = [int(rng.integers(1, 6000)) for _ in range(n_ph)]
        plan = split_work(counts, costs, world)
        assert len(plan) == world
        seen = [np.zeros(c, int) for c in counts]
        flat = [seg for segs in plan for seg in segs]
        assert flat == sorted(flat)                                   # contiguous slices in (phase, rotation) order
        for p, lo, hi in flat:
            seen[p][lo:hi] += 1
        assert all((s == 1).all() for s in seen)
        loads = [sum((hi - lo) * costs[p] for p, lo, hi in segs) for segs in plan]
        total = sum(c * w for c, w in zip(counts, costs))
        assert max(loads) <= total / world + max(costs) + 1
    # BASELINE configs[4]: Fe bcc + Fe fcc + Fe3C, rendered -> nearly equal template counts per rank
    plan = split_work([333333, 333333, 333334], [40000 + 55, 40000 + 103, 40000 + 391], 8)
    sizes = [sum(hi - lo for _, lo, hi in segs) for segs in plan]
    assert max(sizes) - min(sizes) < 0.01 * sizes[0]


def _oracle_packed(phase_name, quats, rr=1.0, s_max=0.02):
    """CSR spot lists of a rotation list from the CPU oracle (stand-in for K2 + ds_pack_csr on a rank)."""
    from diffsims_b200.library import PackedSpots
    from oracle import kinematical as K
    from tests.golden import cases
    from tests.helpers import oracle_G_from_active_quat
    phase = cases.phase(phase_name)
    gs = K.GSet(phase.structure, rr, True)
    wl = K.get_electron_wavelength(200)
    off, g, x, i = [0], [], [], []
    for q in quats:
        r = K.simulate_rotation(phase.structure, gs, oracle_G_from_active_quat(q), wl, s_max)
        off.append(off[-1] + len(r["intensity"]))
        g.append(np.asarray(r["g_index"], np.int32))
        x.append(np.asarray(r["xyz"], float).reshape(-1, 3))
        i.append(np.asarray(r["intensity"], float))
    return PackedSpots(torch.tensor(off, dtype=torch.int64), torch.from_numpy(np.concatenate(g)),
                       torch.from_numpy(np.concatenate(x)), torch.from_numpy(np.concatenate(i)))


_PHASES = [("fe_bcc", 9, 0), ("si", 7, 1), ("fe3c", 5, 2)]     # name, orientations, seed
_COSTS = [92, 690, 650]


def _library_worker(rank, world, port, q):
    from diffsims_b200.library import gather_shards, split_work
    from tests.helpers import random_quats
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        quats = [random_quats(n, seed) for _, n, seed in _PHASES]
        plan = split_work([n for _, n, _ in _PHASES], _COSTS, world)
        packed = [_oracle_packed(_PHASES[p][0], quats[p][lo:hi]) for p, lo, hi in plan[rank]]
        lib, nbytes = gather_shards(plan, packed, rank, world, len(_PHASES))
        q.put((rank, [(t.offsets.numpy(), t.g_index.numpy(), t.xyz.numpy(), t.intensity.numpy()) for t in lib], nbytes))
    finally:
        dist.destroy_process_group()


def test_sharded_library_gather_equals_single_rank_world2():
    """Two ranks build their slices of a 3-phase library and gather; every rank must hold exactly the library a
    single rank builds (reflection order, indices and every float64 bit)."""
    from tests.helpers import random_quats
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_library_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        rank, lib, nbytes = q.get(timeout=300)
        res[rank] = lib
        assert nbytes > 0
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = [_oracle_packed(name, random_quats(n, seed)) for name, n, seed in _PHASES]
    for rank in (0, 1):
        for got, ref in zip(res[rank], single):
            for a, b in zip(got, (ref.offsets.numpy(), ref.g_index.numpy(), ref.xyz.numpy(), ref.intensity.numpy())):
                np.testing.assert_array_equal(a, b)
