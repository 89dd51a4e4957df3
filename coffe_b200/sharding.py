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


def weighted_bounds(weights, world):
    """world + 1 boundaries of contiguous blocks of the flattened items with (nearly) equal total weight: the cost of a
    (z, r) item grows with the separation (more mu and outer nodes) and with the distance (more tail panels), so equal
    item counts leave the ranks that hold the high-redshift rows with ~10 % more work.  Deterministic: every rank
    computes the same boundaries from the same weights."""
    w = np.asarray(weights, dtype=np.float64).ravel()
    if world < 1:
        raise ValueError("need world >= 1")
    if w.size == 0:
        return [0] * (world + 1)
    if np.any(w < 0) or not np.all(np.isfinite(w)) or w.sum() <= 0:
        raise ValueError("weights must be finite, non-negative and not all zero")
    c = np.concatenate([[0.0], np.cumsum(w)])
    targets = c[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(c, targets, side="left")
    # the boundary closest to the target of the two neighbours
    cuts = np.array([k if k == 0 or abs(c[k] - t) <= abs(c[k - 1] - t) else k - 1 for k, t in zip(cuts, targets)], dtype=int)
    bounds = np.concatenate([[0], np.clip(cuts, 0, w.size), [w.size]])
    return [int(b) for b in np.maximum.accumulate(bounds)]


def gather_ranges(local, n_items, dist, device="cpu", dst=0, bounds=None):
    """Gather the per-rank result slices of coffe_b200_multipoles_range.

    local : [n_l, end - begin] array of this rank (layout of coffe_b200_multipoles_range)
    bounds: world + 1 block boundaries (weighted_bounds); default: equal blocks (block_partition)
    Returns [n_l, n_items] on rank `dst`, None elsewhere.  One collective, padded to the largest block.
    """
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    if bounds is None:
        bounds = [block_partition(n_items, world, r)[0] for r in range(world)] + [n_items]
    if len(bounds) != world + 1 or bounds[0] != 0 or bounds[-1] != n_items:
        raise ValueError("bounds must run from 0 to n_items in world + 1 steps")
    per = max(max(bounds[r + 1] - bounds[r] for r in range(world)), 1)
    local = np.ascontiguousarray(local, dtype=np.float64)
    n_l = local.shape[0]
    begin, end = bounds[rank], bounds[rank + 1]
    if local.shape != (n_l, end - begin):
        raise ValueError(f"rank {rank}: expected a slice of shape {(n_l, end - begin)}, got {local.shape}")
    pad = torch.zeros((n_l, per), dtype=torch.float64, device=device)
    if end > begin:
        pad[:, :end - begin] = torch.from_numpy(local).to(device)
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=dst)
    if rank != dst:
        return None
    out = [parts[r][:, :bounds[r + 1] - bounds[r]] for r in range(world)]
    return torch.cat(out, dim=1).cpu().numpy()


def separable_cost(engine, z_mean, sep, l, n_z_samples=5, n_sep_samples=16):
    """Estimated relative cost of every (z, r) item, [n_z, n_sep]: integrand evaluations counted by the library itself
    (coffe_b200_plan_count) on a few rows and columns of the grid, combined as f(z) g(r).  Costs a few ms of GPU time."""
    z, s = np.asarray(z_mean, float), np.asarray(sep, float)

    def cost(zs, ss):
        plan = engine.plan_multipoles(zs, ss, l)
        _, _, n_si, n_di = plan.count()
        plan.close()
        return n_di + 2.0 * n_si          # a local x integrated evaluation costs about two integrated x integrated ones

    iz = np.unique(np.linspace(0, len(z) - 1, min(n_z_samples, len(z))).round().astype(int))
    js = np.unique(np.linspace(0, len(s) - 1, min(n_sep_samples, len(s))).round().astype(int))
    f = np.array([cost(z[i:i + 1], s[js]) for i in iz])
    g = np.array([cost(z[iz[len(iz) // 2]:iz[len(iz) // 2] + 1], s[j:j + 1]) for j in js])
    fz = np.interp(z, z[iz], f) if len(iz) > 1 else np.full(len(z), f[0])
    gs = np.interp(s, s[js], g) if len(js) > 1 else np.full(len(s), g[0])
    return np.outer(fz, gs)


def ranges_to_grid(flat, n_z, n_sep):
    """[n_l, n_z*n_sep] (item = iz*n_sep + ir) -> [n_z, n_l, n_sep], the order of coffe_multipoles_init
    (src/multipoles.c:75-86) and of coffe_b200_multipoles"""
    flat = np.asarray(flat)
    n_l = flat.shape[0]
    return np.ascontiguousarray(flat.reshape(n_l, n_z, n_sep).transpose(1, 0, 2))


def multipoles_sharded(engine, z_mean, sep, l, dist=None, device=None, balanced=False):
    """Sharded equivalent of Engine.multipoles: every rank computes its block on its own GPU, rank 0 returns
    the full [n_z, n_l, n_sep] array (None on the other ranks).  Without a process group: single GPU."""
    n_z, n_sep = len(z_mean), len(sep)
    n_items = n_z * n_sep
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ranges_to_grid(engine.multipoles_range(z_mean, sep, l, 0, n_items), n_z, n_sep)
    bounds = None
    if balanced:
        bounds = weighted_bounds(separable_cost(engine, z_mean, sep, l), dist.get_world_size())
        begin, end = bounds[dist.get_rank()], bounds[dist.get_rank() + 1]
    else:
        begin, end = block_partition(n_items, dist.get_world_size(), dist.get_rank())
    local = engine.multipoles_range(z_mean, sep, l, begin, end)
    if device is None:
        device = f"cuda:{engine.device}" if dist.get_backend() == "nccl" else "cpu"
    flat = gather_ranges(local, n_items, dist, device=device, bounds=bounds)
    return None if flat is None else ranges_to_grid(flat, n_z, n_sep)
