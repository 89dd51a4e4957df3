"""Host logic of the multi-GPU path (coffe_b200/sharding.py) on CPU: block partition of the (z, r) items and the
single gather of the result slices, world size 2 and 3 over gloo.  The per-rank compute call is replaced by a
stand-in with the layout of coffe_b200_multipoles_range (the kernels need a GPU; their sharded result is
compared bit for bit with the unsharded one in tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest

from coffe_b200 import sharding


def test_block_partition_covers_the_grid_once():
    for n in (0, 1, 7, 32000, 32001):
        for world in (1, 2, 3, 8):
            seen = np.zeros(n, dtype=int)
            for rank in range(world):
                b, e = sharding.block_partition(n, world, rank)
                assert 0 <= b <= e <= n
                seen[b:e] += 1
            assert np.all(seen == 1)
    with pytest.raises(ValueError):
        sharding.block_partition(10, 2, 2)


def test_weighted_bounds_balance_and_cover():
    rng = np.random.default_rng(3)
    for n in (1, 7, 1000):
        for world in (1, 2, 3, 8):
            w = rng.uniform(0.5, 2.0, n) * np.linspace(1.0, 3.0, n)
            b = sharding.weighted_bounds(w, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            if n >= 100 * world:
                parts = np.array([w[b[r]:b[r + 1]].sum() for r in range(world)])
                assert parts.max() / parts.mean() < 1.02
    assert sharding.weighted_bounds(np.ones(8), 4) == [0, 2, 4, 6, 8]
    with pytest.raises(ValueError):
        sharding.weighted_bounds([1.0, -1.0], 2)


def test_ranges_to_grid_is_the_reference_loop_order():
    n_z, n_sep, n_l = 3, 5, 2
    flat = np.empty((n_l, n_z * n_sep))
    for il in range(n_l):
        for iz in range(n_z):
            for ir in range(n_sep):
                flat[il, iz * n_sep + ir] = 100 * iz + 10 * il + ir
    grid = sharding.ranges_to_grid(flat, n_z, n_sep)
    assert grid.shape == (n_z, n_l, n_sep)
    assert grid[2, 1, 4] == 214 and grid[1, 0, 3] == 103


class FakeEngine:
    """stand-in for coffe_b200.Engine.multipoles_range: value = f(z, r, l), layout [n_l, end - begin]"""
    device = 0

    def multipoles_range(self, z, sep, l, begin, end):
        out = np.empty((len(l), end - begin))
        for k, item in enumerate(range(begin, end)):
            iz, ir = divmod(item, len(sep))
            for il, ll in enumerate(l):
                out[il, k] = np.sin(z[iz]) * sep[ir] + ll
        return out


class CountingPlan:
    def __init__(self, z, sep):
        self.z, self.sep = z, sep

    def count(self):
        c = sum((1.0 + zz) * (10.0 + ss) for zz in self.z for ss in self.sep)
        return (0.0, 0.0, 0.0, c)

    def close(self):
        pass


class FakeEngineWithPlans(FakeEngine):
    def plan_multipoles(self, z, sep, l, begin=0, end=None):
        return CountingPlan(list(z), list(sep))


def _worker(rank, world, port, n_z, n_sep, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z, sep, l = np.linspace(0.5, 2.0, n_z), np.linspace(10, 300, n_sep), [0, 2, 4]
        got = sharding.multipoles_sharded(FakeEngine(), z, sep, l, dist=dist)
        bal = sharding.multipoles_sharded(FakeEngineWithPlans(), z, sep, l, dist=dist, balanced=True)
        if rank == 0:
            assert np.array_equal(bal, got)
        else:
            assert bal is None
        if rank == 0:
            want = sharding.ranges_to_grid(FakeEngine().multipoles_range(z, sep, l, 0, n_z * n_sep), n_z, n_sep)
            q.put(bool(np.array_equal(got, want)) and got.shape == (n_z, 3, n_sep))
        else:
            q.put(got is None)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_z,n_sep", [(2, 4, 25), (2, 3, 7), (3, 2, 5)])
def test_sharded_gather_equals_unsharded(world, n_z, n_sep):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_z, n_sep, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)
