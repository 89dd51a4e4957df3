"""The compiled drop-in (dropin/): COFFE's own driver symbols -- coffe_multipoles_init(par*, bg*, integral*, mp*) and friends,
with the reference's struct layouts -- exported by a library that keeps the reference's host sources (common.c, parser.c,
background.c, errors.c) and forwards the arithmetic to libcoffe_b200.  The harness is oracle/ref_driver.c (the same ctypes driver
the parity oracle uses: real coffe_parameters_t structs, the reference's setters) compiled against the drop-in instead of the
reference's integrators, so `RefCoffe(lib_path=harness)` runs the reference's call sequence over the GPU kernels."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ALL, H, ROOT

BUILD = ROOT / "dropin" / "_build"
LIB, HARNESS = BUILD / "libcoffe_dropin.so", BUILD / "libcoffe_dropin_harness.so"
REF = Path("/root/reference")


def _build():
    if (REF / "src" / "common.c").exists():
        subprocess.run(["make", "-C", str(ROOT / "dropin"), f"REF={REF}", "all", "harness"], check=True, stdout=subprocess.DEVNULL)
    if not LIB.exists() or not HARNESS.exists():
        pytest.skip("dropin/_build not built (run __graft_entry__.build() where /root/reference exists)")


def test_dropin_exports_the_symbols_the_reference_binds_and_nothing_of_the_replaced_sources():
    """python/coffe/coffe.pxd:399-558 binds these; nm must show them defined, and no functions_*, Cuhre, twofast_*, twobessel_*"""
    _build()
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True, check=True).stdout
    defined = {line.split()[-1] for line in out.splitlines() if line.strip()}
    for sym in ("coffe_integrals_init", "coffe_integrals_free", "coffe_find_integral", "coffe_integrate", "coffe_corrfunc_init", "coffe_corrfunc_free",
                "coffe_multipoles_init", "coffe_multipoles_free", "coffe_average_multipoles_init", "coffe_average_multipoles_free",
                "coffe_covariance_init", "coffe_covariance_free", "coffe_average_covariance_free", "coffe_covariance_find",
                "coffe_parse_default_parameters", "coffe_background_init", "coffe_init_spline", "coffe_interp_spline"):
        assert sym in defined, sym
    every = subprocess.run(["nm", "-D", str(LIB)], capture_output=True, text=True, check=True).stdout
    for banned in ("functions_nonintegrated", "functions_single_integrated", "functions_double_integrated", "Cuhre", "twofast_", "twobessel"):
        assert banned not in every, banned
    undefined = subprocess.run(["nm", "-D", "--undefined-only", str(LIB)], capture_output=True, text=True, check=True).stdout
    assert "coffe_b200_group_multipoles" in undefined and "coffe_b200_integrals" in undefined


def test_dropin_without_a_gpu_fails_loudly_not_silently():
    """no CPU fallback behind COFFE's API either: the driver returns EXIT_FAILURE and leaves size == 0"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _build()
    from oracle import coffe_oracle as O
    r = O.RefCoffe(lib_path=HARNESS)
    r.set_power_spectrum(O.load_pk())
    r.set_contributions(["den", "rsd"]); r.set_sep([50.0]); r.set_z_mean([1.0]); r.set_deltaz([0.1]); r.set_l([0])
    r.lib.ref_background(r.h_)
    assert r.lib.ref_background_size(r.h_) == 10000              # the kept host sources (background.c) run as they are
    out = np.zeros(1)
    # coffe_integrals_init -> EXIT_FAILURE -> integral->size stays 0 -> the driver's own multipoles call fails the same way
    n = r.lib.ref_multipoles(r.h_, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert n == 0


@pytest.fixture(scope="module")
def drop():
    _build()
    from oracle import coffe_oracle as O
    r = O.RefCoffe(lib_path=HARNESS)
    r.set_power_spectrum(O.load_pk())
    r.set_sep(np.array([10.0, 20.0, 40.0, 100.0, 150.0]) / H)
    r.set_z_mean([0.5, 1.0, 1.5, 2.0]); r.set_deltaz([0.3] * 4); r.set_l([0, 2, 4])
    r.set_contributions(ALL)
    return r


@pytest.mark.gpu
def test_dropin_integrals_init_matches_the_reference_tables(drop, ref):
    """coffe_integrals_init of the drop-in (coffe_b200_integrals underneath) against the reference's own driver: same knots, values
    within the reference's QAG tolerance; the 2-FAST part to 1e-7 of each table's scale"""
    for (n, l) in [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (2, 0), (2, 2), (3, 1), (4, 0)]:
        xa, ya = drop.integral_table(n, l)
        xb, yb = ref.integral_table(n, l)
        assert np.allclose(xa, xb, rtol=1e-13, atol=0)
        scale = np.max(np.abs(yb))
        tol = 3e-5 if (n, l) == (4, 0) else 1e-7
        big = xb > 1.0
        assert np.max(np.abs(ya - yb)[big]) < tol * scale, (n, l)
        assert np.max(np.abs(ya - yb)[~big]) < 3e-5 * scale, (n, l)
    ca, za = drop.renorm_table(); cb, zb = ref.renorm_table()
    assert np.allclose(ca, cb, rtol=1e-12) and np.max(np.abs(za - zb)) < 3e-5 * np.max(np.abs(zb))


@pytest.mark.gpu
def test_dropin_multipoles_init_on_real_parameter_structs(drop, ref, engine):
    """coffe_multipoles_init(&par, &bg, &integral, &mp) of the drop-in, AoS out in the reference's order: local terms against the
    reference's own driver to 1e-5; all effects against the kernels fed the same tables directly (bit for bit) -- the drop-in adds
    plumbing, not arithmetic"""
    drop.set_contributions(["den", "rsd", "d1"]); ref.set_contributions(["den", "rsd", "d1"])
    try:
        got, want = drop.multipoles(), ref.multipoles()
        assert got.shape == want.shape == (4, 3, 5)
        scale = np.max(np.abs(want), axis=2, keepdims=True)
        assert np.all(np.abs(got - want) < 1e-5 * scale), np.max(np.abs(got - want) / scale)
    finally:
        drop.set_contributions(ALL); ref.set_contributions(ALL)
    got = drop.multipoles()
    import coffe_b200
    from oracle_tables import tableset_from_oracle
    eng = coffe_b200.Engine(0)
    eng.upload_tables(tableset_from_oracle(drop))           # the drop-in's own tables (built by its coffe_integrals_init)
    eng.set_options(ALL)
    direct = eng.multipoles([0.5, 1.0, 1.5, 2.0], np.array([10.0, 20.0, 40.0, 100.0, 150.0]) / H, [0, 2, 4])
    assert np.array_equal(got, direct)
    # and against the kernels on the REFERENCE's tables: what the drop-in's own table construction contributes
    engine.set_options(ALL)
    on_ref_tables = engine.multipoles([0.5, 1.0, 1.5, 2.0], np.array([10.0, 20.0, 40.0, 100.0, 150.0]) / H, [0, 2, 4])
    assert np.max(np.abs(got / on_ref_tables - 1)) < 5e-5


@pytest.mark.gpu
def test_dropin_corrfunc_average_covariance_and_single_points(drop, ref):
    drop.set_contributions(["den", "rsd"]); ref.set_contributions(["den", "rsd"])
    try:
        for r in (drop, ref):
            r.set_mu([0.0, 0.5, 0.9]); r.set_z_min([0.9]); r.set_z_max([1.1])
        a, b = drop.corrfunc(), ref.corrfunc()
        assert a.shape == b.shape == (4, 3, 5) and np.all(np.abs(a - b) < 1e-5 * np.max(np.abs(b), axis=2, keepdims=True))
        a, b = drop.average_multipoles(), ref.average_multipoles()
        assert a.shape == b.shape == (1, 3, 5) and np.allclose(a, b, rtol=5e-4)          # the reference's 2-D Cuhre at its own tolerance
        # single points through coffe_integrate(flag_integral, flag_output): src/signal.c:472-771
        for piece, out in ((0, 2), (0, 1)):
            va = drop.integrate(1.0, 60.0, 0.3, 2, piece, out)
            vb = ref.integrate(1.0, 60.0, 0.3, 2, piece, out)
            assert abs(va / vb - 1) < 1e-5, (piece, out, va, vb)
        # covariance: the reference's test set-up (tests/test_coffe.py:382-399)
        for r in (drop, ref):
            r.set_z_mean([1.0]); r.set_deltaz([0.1]); r.set_sep(np.arange(50.0, 350.0, 50.0) / H)
            r.set_fsky([0.2]); r.set_density1([1e-3 * H ** 3]); r.set_density2([1e-3 * H ** 3]); r.set_pixelsize([50.0 / H])
            r.set_covariance(window=0)
        ref.set_integration(prec=1e-8)
        a, b = drop.covariance(), ref.covariance()
        ref.set_integration(prec=1e-5)
        d = np.sqrt(np.abs(np.einsum("ziaia->zia", b)))
        scale = d[:, :, :, None, None] * d[:, None, None, :, :]
        assert a.shape == b.shape == (1, 3, 6, 3, 6) and np.all(np.abs(a - b) <= 2e-6 * np.abs(b) + 1e-8 * scale)
    finally:
        for r in (drop, ref):
            r.set_contributions(ALL); r.set_z_mean([0.5, 1.0, 1.5, 2.0]); r.set_deltaz([0.3] * 4)
            r.set_sep(np.array([10.0, 20.0, 40.0, 100.0, 150.0]) / H)
    # all effects, one point, the three pieces of coffe_integrate against the converged oracle
    from oracle.converged import ConvergedOracle
    want = ConvergedOracle(ref, n_outer=64, n_panel=8, mu_kw=dict(n=6, W=24.0)).multipoles_pieces(1.0, 60.0, [2])[:, 0]
    got = np.array([drop.integrate(1.0, 60.0, 0.0, 2, piece, 2) for piece in (0, 1, 2)])
    assert np.all(np.abs(got - want) < 1e-5 * abs(want.sum()) + 5e-5 * np.abs(want)), (got, want)
