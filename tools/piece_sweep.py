#!/usr/bin/env python
"""Piece-level accuracy of candidate rules (GPU): lensing only (integrated x integrated piece), density x lensing cross only (local x
integrated piece) and the potential terms alone, on 4 z x 24 log-spaced separations, against the FINE rules of tools/quad_sweep.py.
Errors relative to the l = 0..4 scale of the piece at the point (the gate of the piece target) and, for lensing only, of each l to itself."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import coffe_b200  # noqa: E402
from coffe_b200.default_tables import load_default_tableset  # noqa: E402

H = 0.67
FINE = dict(n_mu_panel=12, n_mu_panel_di=10, mu_panel_width=16.0, n_outer=96, n_panel_si=12, n_panel_di=12, panel_width=8.0,
            panel_uniform_extent=640.0, panel_width_di=16.0, panel_uniform_extent_di=640.0, mu_panel_width_di=16.0,
            tail_ratio=1.3, tail_ratio_di=1.3, geometric_shift_di=1, geometric_growth=1.0, geometric_growth_di=1.0, n_renorm=64,
            tail_map=1, tail_map_di=1, mu_geometric_growth=1.0, mu_geometric_growth_di=1.0, mu_switch_break_di=1, mu_tail_ratio=1.3, n_mu_ni=64)
z, sep, ls = np.array([0.5, 1.0, 1.5, 2.0]), np.geomspace(10.0, 300.0, 24) / H, [0, 2, 4]
eng = coffe_b200.Engine(0)
eng.upload_tables(load_default_tableset())
CONFIGS = {"len": (["len"], {}, 2), "den x len": (["den", "len"], {"only_cross": True}, 1), "g4 g5": (["g4", "g5"], {}, 2),
           "d1 g1 x len g4": (["d1", "g1", "len", "g4"], {"only_cross": True}, 1)}
cands = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {"default": {}}
for cname, (contrib, flags, piece) in CONFIGS.items():
    eng.set_options(contrib, **flags)
    eng.set_quadrature(**FINE)
    fine = eng.multipoles_pieces(z, sep, ls)[piece]          # [nz, nl, ns]
    scale = np.max(np.abs(fine), axis=1, keepdims=True)
    for name, kw in cands.items():
        eng.set_quadrature(accuracy_target=1, **kw)
        got = eng.multipoles_pieces(z, sep, ls)[piece]
        err = np.abs(got - fine) / scale
        self_err = np.abs(got / fine - 1)
        k = np.unravel_index(np.argmax(err), err.shape)
        print(f"{cname:16s} {name:24s} max {err.max():.2e} at z={z[k[0]]} r={sep[k[2]] * H:.0f} l={ls[k[1]]} | each l to itself: l0 {self_err[:, 0].max():.2e} "
              f"l2 {self_err[:, 1].max():.2e} l4 {self_err[:, 2].max():.2e}", flush=True)
