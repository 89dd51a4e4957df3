"""Multi-GPU sharding of the output grid: host logic only.

Every (z_mean, separation) work item is independent given the tables -- the reference itself runs a flat
`omp parallel for` over them (src/multipoles.c:88-125) -- so the flattened items iz*n_sep + ir are cut into
contiguous blocks of equal estimated cost, one per GPU, with the tables replicated.  Two ways to drive the GPUs:

  * one process, all devices: `coffe_b200.Group` (coffe_b200_group_* in the library: one host thread per device, every block
    copied device -> host straight into the caller's array; no collective);
  * one process per GPU (torchrun / mpirun): `multipoles_sharded` below -- each rank computes its block and the blocks are
    gathered device to device by the library's own NCCL calls (`NcclComm`).

The arithmetic of the block boundaries lives in the library (coffe_b200_balance_bounds, coffe_b200_multipoles_cost); this
module only arranges arrays.  The collective is reached through a small communicator protocol (`rank`, `world`,
`allgather_blocks`) so that the CPU tests can drive the same code over a gloo process group they set up themselves.
"""
import numpy as np

from . import engine as _engine


def block_partition(n_items, world, rank):
    """[begin, end) of rank `rank`: equal blocks of ceil(n_items / world), the last ones possibly short or empty"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    per = (n_items + world - 1) // world
    return min(rank * per, n_items), min((rank + 1) * per, n_items)


def block_size(n_items, world):
    return (n_items + world - 1) // world


def weighted_bounds(weights, world):
    """world + 1 boundaries of contiguous blocks of the flattened items with (nearly) equal total weight: the cost of a
    (z, r) item grows with the separation (more mu and outer nodes) and with the distance (more tail panels), so equal
    item counts leave the ranks that hold the high-redshift rows with ~10 % more work.  Deterministic: every rank
    computes the same boundaries from the same weights (coffe_b200_balance_bounds, host arithmetic of the library)."""
    w = np.asarray(weights, dtype=np.float64).ravel()
    if world < 1:
        raise ValueError("need world >= 1")
    if w.size == 0:
        return [0] * (world + 1)
    return _engine.balance_bounds(w, world)


def item_cost(engine, z_mean, sep, l):
    """[n_z, n_sep] estimated cost of every (z, r) item: integrand evaluations counted by the library on the device
    (coffe_b200_multipoles_cost: exact on a few rows of the grid, every separation; linear in z between them)"""
    return engine.multipoles_cost(z_mean, sep, l)


def ranges_to_grid(flat, n_z, n_sep):
    """[n_l, n_z*n_sep] (item = iz*n_sep + ir) -> [n_z, n_l, n_sep], the order of coffe_multipoles_init
    (src/multipoles.c:75-86) and of coffe_b200_multipoles"""
    flat = np.asarray(flat)
    n_l = flat.shape[0]
    return np.ascontiguousarray(flat.reshape(n_l, n_z, n_sep).transpose(1, 0, 2))


class NcclComm:
    """The library's own communicator (coffe_b200_comm_init): the gather runs device to device inside
    coffe_b200_multipoles_sharded, so `multipoles_sharded` hands the whole call to the engine."""

    def __init__(self, engine, rank, world, unique_id):
        engine.comm_init(rank, world, unique_id)
        self.engine, self.rank, self.world = engine, rank, world
        self.device_gather = True


def multipoles_sharded(engine, z_mean, sep, l, comm=None, bounds=None, balanced=True):
    """Sharded equivalent of Engine.multipoles, one process per GPU: returns the full [n_z, n_l, n_sep] array on every rank.
    comm: None (single process) | NcclComm | any object with `rank`, `world` and
    `allgather_blocks(local [n_l, n_local], bounds) -> [n_l, n_items]` (the CPU tests pass a gloo-backed one)."""
    n_z, n_sep = len(z_mean), len(sep)
    n_items = n_z * n_sep
    if comm is None or comm.world == 1:
        return ranges_to_grid(engine.multipoles_range(z_mean, sep, l, 0, n_items), n_z, n_sep)
    if bounds is None:
        if balanced:
            bounds = weighted_bounds(item_cost(engine, z_mean, sep, l), comm.world)
        else:
            bounds = [block_partition(n_items, comm.world, r)[0] for r in range(comm.world)] + [n_items]
    if len(bounds) != comm.world + 1 or bounds[0] != 0 or bounds[-1] != n_items:
        raise ValueError("bounds must run from 0 to n_items in world + 1 steps")
    if getattr(comm, "device_gather", False):
        out, _, _ = engine.multipoles_sharded(z_mean, sep, l, bounds=bounds)
        return out
    local = engine.multipoles_range(z_mean, sep, l, bounds[comm.rank], bounds[comm.rank + 1])
    flat = comm.allgather_blocks(np.ascontiguousarray(local, dtype=np.float64), bounds)
    return ranges_to_grid(flat, n_z, n_sep)
