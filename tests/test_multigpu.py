"""Multi-GPU drivers of the library (group.inc): one call -> all devices (coffe_b200_group_*), the range / row variants they are
built from, and the one-process-per-GPU path with the library's own NCCL gather.  Everything is compared BIT FOR BIT with the
single-device calls: a result must not depend on how the grid was cut or on which device computed it."""
import multiprocessing as mp

import numpy as np
import pytest

from conftest import ALL, H

pytestmark = pytest.mark.gpu


def _device_count():
    import ctypes as C
    try:
        cuda = C.CDLL("libcudart.so")
    except OSError:
        import torch
        return torch.cuda.device_count()
    n = C.c_int()
    return n.value if cuda.cudaGetDeviceCount(C.byref(n)) == 0 else 0


@pytest.fixture(scope="module")
def single():
    import coffe_b200
    from coffe_b200.default_tables import load_default_tableset
    eng = coffe_b200.Engine(0)
    eng.upload_tables(load_default_tableset())
    eng.set_options(ALL)
    return eng


@pytest.fixture(scope="module")
def group():
    import coffe_b200
    from coffe_b200.default_tables import load_default_tableset
    g = coffe_b200.Group(-1)
    g.upload_tables(load_default_tableset())
    g.set_options(ALL)
    return g


GRID = (np.linspace(0.6, 1.9, 5), np.linspace(15.0, 420.0, 37), [0, 2, 4])
COV = dict(sep=np.linspace(20.0, 218.0, 40) / H, l=[0, 2, 4], z_mean=[0.6, 1.4], deltaz=[0.1, 0.1])


def _cov_args():
    import coffe_b200
    from pathlib import Path
    k, pk = np.loadtxt(Path(coffe_b200.__file__).resolve().parent / "data" / "PkL_CLASS.dat", unpack=True)
    return ((k * H, pk / H ** 3), COV["sep"], COV["l"], [0.2] * 2, [1e-3 * H ** 3] * 2, [1e-3 * H ** 3] * 2, [2.0 / H] * 2)


def test_group_calls_match_single_device_bit_for_bit(single, group):
    z, sep, ls = GRID
    assert group.size >= 1
    a = group.multipoles(z, sep, ls)
    assert np.array_equal(a, single.multipoles(z, sep, ls))
    bounds, ms = group.last_blocks()
    assert bounds[0] == 0 and bounds[-1] == len(z) * len(sep) and all(x <= y for x, y in zip(bounds, bounds[1:]))
    if group.size > 1:
        assert all(m > 0 for m in ms), "every device must have run kernels"
        sizes = np.diff(bounds)
        assert sizes.min() > 0 and sizes[0] > sizes[-1], "cost-balanced blocks: the cheap low-redshift rows make the first block the longest"
    mu = [-0.9, -0.2, 0.4, 0.97]
    assert np.array_equal(group.corrfunc(z[:3], mu, sep[:11]), single.corrfunc(z[:3], mu, sep[:11]))
    zmin, zmax = [0.5, 1.2, 1.8125], [0.6875, 1.4, 2.0]
    assert np.array_equal(group.average_multipoles(zmin, zmax, sep[:7], ls), single.average_multipoles(zmin, zmax, sep[:7], ls))
    single.set_options(["den", "rsd"]); group.set_options(["den", "rsd"])
    try:
        for window in (False, True):
            kw = dict(z_mean=COV["z_mean"], deltaz=COV["deltaz"], window=window)
            assert np.array_equal(group.covariance(*_cov_args(), **kw), single.covariance(*_cov_args(), **kw))
    finally:
        single.set_options(ALL); group.set_options(ALL)


def test_range_and_row_variants_tile_the_full_results(single):
    z, sep, ls = GRID
    mu = [-0.9, -0.2, 0.4, 0.97]
    single.set_quadrature(accuracy_target=1)        # pieces are always served at piece accuracy: pin the totals to the same rules
    full = single.corrfunc(z[:3], mu, sep[:11])
    n = full.size
    cuts = [0, 5, 44, 45, 100, n]
    parts = [single.corrfunc_pieces(z[:3], mu, sep[:11], a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    tiled = np.concatenate(parts, axis=1)
    assert np.array_equal(tiled.sum(0).reshape(full.shape), full) or np.array_equal((tiled[0] + tiled[1] + tiled[2]).reshape(full.shape), full)
    zmin, zmax = [0.5, 1.8125], [0.6875, 2.0]
    fa = single.average_multipoles(zmin, zmax, sep[:7], ls)
    single.set_quadrature()
    pa_auto = single.average_multipoles(zmin, zmax, sep[:7], ls)
    assert np.all(np.abs(pa_auto - fa) < 1e-5 * np.max(np.abs(fa), axis=1, keepdims=True))      # automatic target: within the gate of the piece rules
    single.set_quadrature(accuracy_target=1)
    pa = np.concatenate([single.average_multipoles_pieces(zmin, zmax, sep[:7], ls, a, b) for a, b in ((0, 3), (3, 4), (4, 14))], axis=2)
    assert np.array_equal((pa[0] + pa[1] + pa[2]).reshape(3, 2, 7).transpose(1, 0, 2), fa)
    single.set_quadrature()
    single.set_options(["den", "rsd"])
    try:
        kw = dict(z_mean=COV["z_mean"], deltaz=COV["deltaz"])
        fc = single.covariance(*_cov_args(), **kw)
        rows = 3 * 40
        tiled = np.full_like(fc, np.nan)
        for a, b in ((0, 17), (17, 64), (64, 65), (65, rows)):
            part = single.covariance(*_cov_args(), rows=(a, b), **kw).reshape(2, rows, rows)
            assert np.all(np.isnan(part[:, :a])) and np.all(np.isnan(part[:, b:]))
            tiled.reshape(2, rows, rows)[:, a:b] = part[:, a:b]
        assert np.array_equal(tiled, fc)
    finally:
        single.set_options(ALL)


def test_cost_model_follows_the_counted_evaluations(single):
    """coffe_b200_multipoles_cost (exact on a few rows, interpolated in z) against the evaluations the kernels count themselves"""
    z, sep, ls = np.linspace(0.5, 2.0, 12), np.linspace(15.0, 440.0, 24), [0, 2, 4]
    cost = single.multipoles_cost(z, sep, ls)
    assert cost.shape == (12, 24) and np.all(cost > 0)
    for iz in (0, 5, 11):
        plan = single.plan_multipoles(z[iz:iz + 1], sep, ls)
        _, _, n_si, n_di = plan.count()
        plan.close()
        assert abs(cost[iz].sum() / (n_di + 2 * n_si) - 1) < 0.01
    import coffe_b200.engine as E
    b = E.balance_bounds(cost, 8)
    parts = np.array([cost.ravel()[b[r]:b[r + 1]].sum() for r in range(8)])
    assert parts.max() / parts.mean() < 1.05


def test_single_rank_communicator_gathers_in_place(single):
    """world = 1 NCCL communicator: the sharded call (kernels + device-side gather + D2H) equals the plain one"""
    import coffe_b200.engine as E
    z, sep, ls = GRID
    single.comm_init(0, 1, E.comm_unique_id())
    try:
        out, kms, gms = single.multipoles_sharded(z, sep, ls)
        assert np.array_equal(out, single.multipoles(z, sep, ls)) and kms > 0 and gms >= 0
        plan = single.plan_multipoles(z, sep, ls)
        plan.execute_gather([0, len(z) * len(sep)])
        assert np.array_equal(plan.fetch_gathered(), plan.fetch())
        plan.close()
    finally:
        single.comm_destroy()


def _rank_main(rank, world, uid, q):
    import coffe_b200
    import coffe_b200.engine as E
    from coffe_b200 import sharding
    from coffe_b200.default_tables import load_default_tableset
    eng = coffe_b200.Engine(rank)
    eng.upload_tables(load_default_tableset())
    eng.set_options(ALL)
    comm = sharding.NcclComm(eng, rank, world, uid)
    z, sep, ls = GRID
    out = sharding.multipoles_sharded(eng, z, sep, ls, comm=comm)
    plan_bounds = sharding.weighted_bounds(sharding.item_cost(eng, z, sep, ls), world)
    plan = eng.plan_multipoles(z, sep, ls, plan_bounds[rank], plan_bounds[rank + 1])
    kms, gms = plan.execute_gather(plan_bounds)
    flat = plan.fetch_gathered()
    plan.close()
    eng.comm_destroy()
    q.put((rank, out, sharding.ranges_to_grid(flat, len(z), len(sep)), plan_bounds, kms, gms))


def test_two_processes_gather_with_nccl_equals_one_gpu(single):
    """one process per GPU, blocks gathered device to device by the library's NCCL calls: every rank ends up with the full grid,
    bit-identical to the single-GPU result"""
    if _device_count() < 2:
        pytest.skip("needs two GPUs")
    import coffe_b200.engine as E
    world = 2
    uid = E.comm_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, world, uid, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    z, sep, ls = GRID
    want = single.multipoles(z, sep, ls)
    for rank, out, planned, bounds, kms, gms in res:
        assert np.array_equal(out, want) and np.array_equal(planned, want), rank
        assert bounds[0] == 0 and bounds[-1] == want.shape[0] * want.shape[2] and kms > 0
