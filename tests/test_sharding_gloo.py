"""Host logic of the multi-GPU path (coffe_b200/sharding.py + coffe_b200_balance_bounds in the library) on CPU: block partition of the
(z, r) items and the gather of the result blocks, world size 2 and 3 over a gloo process group this test sets up (the product package
itself is torch-free: on the GPUs the gather is the library's own NCCL call).  The per-rank compute call is replaced by a stand-in
with the layout of coffe_b200_multipoles_range (the kernels need a GPU; their sharded result is compared bit for bit with the
unsharded one in tests/test_gpu_parity.py and tests/test_multigpu.py)."""
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
    """stand-in for coffe_b200.Engine: value = f(z, r, l), layout [n_l, end - begin]; cost model (1 + z)(10 + r)"""
    device = 0

    def multipoles_range(self, z, sep, l, begin, end):
        out = np.empty((len(l), end - begin))
        for k, item in enumerate(range(begin, end)):
            iz, ir = divmod(item, len(sep))
            for il, ll in enumerate(l):
                out[il, k] = np.sin(z[iz]) * sep[ir] + ll
        return out

    def multipoles_cost(self, z, sep, l):
        return np.outer(1.0 + np.asarray(z), 10.0 + np.asarray(sep))


class GlooComm:
    """the communicator protocol of coffe_b200.sharding over torch.distributed (gloo): variable-size blocks padded to the largest"""

    def __init__(self, dist):
        self.dist, self.rank, self.world = dist, dist.get_rank(), dist.get_world_size()

    def allgather_blocks(self, local, bounds):
        import torch
        n_l = local.shape[0]
        per = max(max(bounds[r + 1] - bounds[r] for r in range(self.world)), 1)
        pad = torch.zeros((n_l, per), dtype=torch.float64)
        pad[:, :local.shape[1]] = torch.from_numpy(local)
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(parts, pad)
        return torch.cat([parts[r][:, :bounds[r + 1] - bounds[r]] for r in range(self.world)], dim=1).numpy()


def _worker(rank, world, port, n_z, n_sep, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z, sep, l = np.linspace(0.5, 2.0, n_z), np.linspace(10, 300, n_sep), [0, 2, 4]
        comm = GlooComm(dist)
        want = sharding.ranges_to_grid(FakeEngine().multipoles_range(z, sep, l, 0, n_z * n_sep), n_z, n_sep)
        eq = sharding.multipoles_sharded(FakeEngine(), z, sep, l, comm=comm, balanced=False)
        bal = sharding.multipoles_sharded(FakeEngine(), z, sep, l, comm=comm, balanced=True)
        # every rank derives the same cost-balanced boundaries, and they differ from the equal-count ones on this grid
        bounds = sharding.weighted_bounds(sharding.item_cost(FakeEngine(), z, sep, l), world)
        allb = [None] * world
        dist.all_gather_object(allb, bounds)
        ok = all(b == bounds for b in allb) and np.array_equal(eq, want) and np.array_equal(bal, want) and eq.shape == (n_z, 3, n_sep)
        q.put(bool(ok))
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


def test_product_package_is_torch_free():
    """north_star: "no PyTorch" in the product; torch is plumbing of bench.py and of the tests only"""
    from pathlib import Path
    import coffe_b200
    for f in Path(coffe_b200.__file__).resolve().parent.rglob("*.py"):
        text = f.read_text()
        assert "import torch" not in text and "from torch" not in text, f
