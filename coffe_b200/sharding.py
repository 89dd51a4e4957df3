"""Multi-GPU sharding of the output grid.

Every (z_mean, separation) work item is independent given the tables -- the reference itself runs a flat
`omp parallel for` over them (src/multipoles.c:88-125) -- so the flattened items iz*n_sep + ir are cut into
contiguous blocks, one per rank, the tables are replicated, and the only communication is one gather of the
FP64 result slices.  This module holds that host logic; it works with any torch.distributed backend (NCCL
on the GPUs, gloo in the CPU tests) and does no arithmetic on the results.
"""
import numpy as np


def block_partition(n_items, world, rank):
    """[begin, end) of rank `rank`: equal blocks of ceil(n_items / world), the last ones possibly short or empty"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    per = (n_items + world - 1) // world
    return min(rank * per, n_items), min((rank + 1) * per, n_items)


def block_size(n_items, world):
    return (n_items + world - 1) // world


def gather_ranges(local, n_items, dist, device="cpu", dst=0):
    """Gather the per-rank result slices of coffe_b200_multipoles_range.

    local : [n_l, end - begin] array of this rank (layout of coffe_b200_multipoles_range)
    Returns [n_l, n_items] on rank `dst`, None elsewhere.  One collective, padded to equal blocks.
    """
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    per = block_size(n_items, world)
    local = np.ascontiguousarray(local, dtype=np.float64)
    n_l = local.shape[0]
    begin, end = block_partition(n_items, world, rank)
    if local.shape != (n_l, end - begin):
        raise ValueError(f"rank {rank}: expected a slice of shape {(n_l, end - begin)}, got {local.shape}")
    pad = torch.zeros((n_l, per), dtype=torch.float64, device=device)
    if end > begin:
        pad[:, :end - begin] = torch.from_numpy(local).to(device)
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat(parts, dim=1)[:, :n_items].cpu().numpy()


def ranges_to_grid(flat, n_z, n_sep):
    """[n_l, n_z*n_sep] (item = iz*n_sep + ir) -> [n_z, n_l, n_sep], the order of coffe_multipoles_init
    (src/multipoles.c:75-86) and of coffe_b200_multipoles"""
    flat = np.asarray(flat)
    n_l = flat.shape[0]
    return np.ascontiguousarray(flat.reshape(n_l, n_z, n_sep).transpose(1, 0, 2))


def multipoles_sharded(engine, z_mean, sep, l, dist=None, device=None):
    """Sharded equivalent of Engine.multipoles: every rank computes its block on its own GPU, rank 0 returns
    the full [n_z, n_l, n_sep] array (None on the other ranks).  Without a process group: single GPU."""
    n_z, n_sep = len(z_mean), len(sep)
    n_items = n_z * n_sep
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ranges_to_grid(engine.multipoles_range(z_mean, sep, l, 0, n_items), n_z, n_sep)
    begin, end = block_partition(n_items, dist.get_world_size(), dist.get_rank())
    local = engine.multipoles_range(z_mean, sep, l, begin, end)
    if device is None:
        device = f"cuda:{engine.device}" if dist.get_backend() == "nccl" else "cpu"
    flat = gather_ranges(local, n_items, dist, device=device)
    return None if flat is None else ranges_to_grid(flat, n_z, n_sep)
