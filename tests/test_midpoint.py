"""the CLASS-midpoint layer (pk_type != linear with midpoint_approximation; src/functions.c:136-155 and the same
#ifdef HAVE_CLASS pattern in every standard term, tables src/integrals.c:663-767).  The oracle is the reference compiled WITH CLASS
(oracle/Makefile: class) and run with halofit; its inputs (P_lin(k), the rows P_nl(k, z_i) CLASS produced) and its results travel in
tests/golden/class_midpoint.npz (tests/golden/make_class_midpoint.py).  The product builds the 2-D tables itself on the GPU
(coffe_b200_integrals2d), uploads them and evaluates den + rsd + len with the midpoint option."""
import numpy as np
import pytest

from conftest import H, ROOT

FIXTURE = ROOT / "tests" / "golden" / "class_midpoint.npz"
FAMILY = [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (2, 0), (2, 2), (3, 1)]
FINE_RULES = dict(n_mu_panel=12, n_mu_panel_di=10, mu_panel_width=16.0, n_outer=96, n_panel_si=12, n_panel_di=12, panel_width=8.0,
                  panel_uniform_extent=640.0, panel_width_di=16.0, panel_uniform_extent_di=640.0, mu_panel_width_di=16.0,
                  tail_ratio=1.3, tail_ratio_di=1.3, geometric_shift_di=1, geometric_growth=1.0, geometric_growth_di=1.0, n_renorm=64, tail_map=1, tail_map_di=1, mu_geometric_growth=1.0, mu_geometric_growth_di=1.0, mu_switch_break_di=1, mu_tail_ratio=1.3, n_mu_ni=64)


def test_fixture_holds_a_nonlinear_spectrum_and_the_layer_matters():
    """CPU: the fixture is what it says -- halofit rows above the linear spectrum at high k and low z, equal to D1^2 P_lin at low k --
    and the midpoint results differ from the plain linear ones by far more than the parity tolerance (so the GPU test can tell)"""
    fx = np.load(FIXTURE)
    k, lin, z2, rows = fx["k"], fx["pk_lin"], fx["z2d"], fx["pk2d"]
    assert rows.shape == (100, len(k)) and np.allclose(z2, 3.0 * np.arange(100) / 100)
    hi, lo = k > 1.0, k < 1e-3
    assert np.all(rows[0][hi][:10] > 1.5 * lin[hi][:10])                  # non-linear boost at z = 0
    assert np.allclose(rows[0][lo], lin[lo], rtol=2e-3)                   # linear regime untouched (window aside)
    assert fx["converged_pieces"].shape == (2, 3, 3, 3) and np.all(np.isfinite(fx["converged_pieces"]))
    stock, conv = fx["stock_multipoles"], fx["converged_pieces"].sum(2).transpose(0, 2, 1)
    assert np.all(np.abs(stock / conv - 1) < 5e-3)                        # the stock Cuhre path around the converged integrals


@pytest.mark.gpu
def test_midpoint_multipoles_match_the_class_reference():
    import coffe_b200
    fx = np.load(FIXTURE)
    cosmo = coffe_b200.Coffe(sep=fx["sep"], z_mean=fx["z_mean"], deltaz=[0.2, 0.2], l=[int(v) for v in fx["l"]],
                             has_density=True, has_rsd=True, has_lensing=True)
    cosmo.set_power_spectrum_linear(fx["k"], fx["pk_lin"])
    cosmo.set_power_spectrum_2d(fx["k"], fx["z2d"], fx["pk2d"], k_min=float(fx["k_min"]), k_max=float(fx["k_max"]))
    linear = cosmo.compute_multipoles_array().copy()
    cosmo.midpoint_approximation = True
    got = cosmo.compute_multipoles_array()
    # (i) the builder: the product's I^n_l(z, r) against samples of the reference's 2-D tables (bicubic in between: compare at knots' neighbourhood)
    res = cosmo._integrals
    from coffe_b200.background import akima_eval
    for t, key in enumerate(FAMILY):
        tab = res["tables2d"][key]                               # [n_r, n_z]
        for iz, zq in enumerate(fx["sample_z"]):
            i = int(round(zq * 100 / 3.0))
            if abs(fx["z2d"][i] - zq) > 1e-12:
                continue                                          # only samples that sit on a redshift knot (no z interpolation involved)
            col = tab[:, i]
            for ir, rq in enumerate(fx["sample_r"]):
                want = fx["sample_i2d"][t, iz, ir]
                j = np.searchsorted(res["r"], rq)
                lo, hi = max(j - 4, 0), min(j + 4, len(res["r"]))
                have = np.interp(rq, res["r"][lo:hi], col[lo:hi])
                scale = np.max(np.abs(col[lo:hi]))
                assert abs(have - want) < 2e-3 * scale, (key, zq, rq, have, want)      # linear interpolation between knots vs bicubic
    # (ii) the signal: every multipole within 1e-5 of the l = 0..4 scale of the converged integral of the reference's integrands
    want = fx["converged_pieces"].sum(2).transpose(0, 2, 1)       # [z, l, sep]
    scale = np.max(np.abs(want), axis=1, keepdims=True)
    err = np.abs(got - want) / scale
    assert err.max() < 1e-5, err
    # (iii) and the layer is really on: the plain linear evaluation of the same inputs is off by per cents at small separations
    assert np.max(np.abs(linear - want) / scale) > 1e-2
    # pieces: local x integrated and integrated x integrated each within 1e-5 of their own scale.  The library's default rules were
    # sized on the linear spectrum; halofit puts more power below 10 Mpc, and the PIECES then need more nodes (default rules: 2e-5 /
    # 7e-5 of the piece, 4e-7 of the total; tools/midpoint_diag.py) -- node counts are per call, like the reference's integration_bins
    eng = cosmo._get_engine()
    eng.set_quadrature(**FINE_RULES)
    try:
        pieces = eng.multipoles_pieces(fx["z_mean"], fx["sep"], [int(v) for v in fx["l"]])      # [3, z, l, sep]
    finally:
        eng.set_quadrature()
    wp = fx["converged_pieces"].transpose(2, 0, 3, 1)
    for piece in (0, 1, 2):
        ps = np.max(np.abs(wp[piece]), axis=1, keepdims=True)
        assert np.max(np.abs(pieces[piece] - wp[piece]) / ps) < 1e-5, (piece, np.abs(pieces[piece] - wp[piece]) / ps)
    # (iv) the builder exactly at knots: the 2-FAST part of every table to 1e-10 of its scale, the small-r knots to the reference's QAG tolerance
    kz, kr = fx["knot_z_idx"], fx["knot_r_idx"]
    for t, key in enumerate(FAMILY):
        mine, ref_v = res["tables2d"][key][np.ix_(kr, kz)], fx["knot_values"][t]
        sc = np.max(np.abs(ref_v), axis=0, keepdims=True)
        small = fx["knot_r"] <= 1.0
        assert np.max(np.abs(mine - ref_v)[~small] / np.broadcast_to(sc, ref_v.shape)[~small]) < 1e-10, key
        assert np.max(np.abs(mine - ref_v)[small] / np.broadcast_to(sc, ref_v.shape)[small]) < 1e-4, key


@pytest.mark.gpu
def test_midpoint_needs_its_tables_and_refuses_flat_sky(engine):
    import coffe_b200
    engine.set_options(["den", "rsd"], midpoint=True)
    try:
        with pytest.raises(coffe_b200.CoffeB200Error) as e:
            engine.multipoles([1.0], [50.0], [0])
        assert e.value.code == 3                                    # ENOTABLES: no 2-D tables uploaded
    finally:
        engine.set_options(["den", "rsd"])
