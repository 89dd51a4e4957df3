#!/usr/bin/env python
"""Diagnostics of the CLASS-midpoint layer on the GPU: (1) the product's 2-D tables against the reference's knot values in the
fixture, (2) the pieces with the default and with the doubled fixed-node rules against the converged fixture."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import coffe_b200  # noqa: E402

FAMILY = [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (2, 0), (2, 2), (3, 1)]
FINE = dict(n_mu_panel=12, n_mu_panel_di=10, mu_panel_width=16.0, n_outer=96, n_panel_si=12, n_panel_di=12, panel_width=8.0,
            panel_uniform_extent=640.0, panel_width_di=16.0, panel_uniform_extent_di=640.0, mu_panel_width_di=16.0,
            tail_ratio=1.3, tail_ratio_di=1.3, geometric_shift_di=1, geometric_growth=1.0, geometric_growth_di=1.0, n_renorm=64, tail_map=1, tail_map_di=1, mu_geometric_growth=1.0, mu_geometric_growth_di=1.0, mu_switch_break_di=1, mu_tail_ratio=1.3, n_mu_ni=64)
fx = np.load(ROOT / "tests" / "golden" / "class_midpoint.npz")
ls = [int(v) for v in fx["l"]]
cosmo = coffe_b200.Coffe(sep=fx["sep"], z_mean=fx["z_mean"], deltaz=[0.2, 0.2], l=ls, has_density=True, has_rsd=True, has_lensing=True)
cosmo.set_power_spectrum_linear(fx["k"], fx["pk_lin"])
cosmo.set_power_spectrum_2d(fx["k"], fx["z2d"], fx["pk2d"], k_min=float(fx["k_min"]), k_max=float(fx["k_max"]))
cosmo.midpoint_approximation = True
got = cosmo.compute_multipoles_array()
res = cosmo._integrals
if "knot_values" in fx:
    kz, kr = fx["knot_z_idx"], fx["knot_r_idx"]
    print("r knots equal:", np.max(np.abs(res["r"][kr] / np.where(fx["knot_r"] == 0, 1, fx["knot_r"]) - 1)[1:]))
    for t, key in enumerate(FAMILY):
        mine = res["tables2d"][key][np.ix_(kr, kz)]
        want = fx["knot_values"][t]
        scale = np.max(np.abs(want), axis=0, keepdims=True)
        d = np.abs(mine - want) / scale
        small = fx["knot_r"] <= 1.0
        print(key, "small-r max", d[small].max(), "FFTLog part max", d[~small].max())
eng = cosmo._get_engine()
wp = fx["converged_pieces"].transpose(2, 0, 3, 1)          # [piece, z, l, sep]
for name, kw in (("default", {}), ("fine", FINE)):
    eng.set_quadrature(**kw)
    pieces = eng.multipoles_pieces(fx["z_mean"], fx["sep"], ls)
    for piece in (0, 1, 2):
        ps = np.max(np.abs(wp[piece]), axis=1, keepdims=True)
        print(name, "piece", piece, "max err / piece scale", np.max(np.abs(pieces[piece] - wp[piece]) / ps), "per (z, sep):", np.max(np.abs(pieces[piece] - wp[piece]) / ps, axis=1).round(8).tolist())
    tot = pieces.sum(0); want = wp.sum(0)
    print(name, "total", np.max(np.abs(tot - want) / np.max(np.abs(want), axis=1, keepdims=True)))
eng.set_quadrature()
