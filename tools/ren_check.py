#!/usr/bin/env python
"""n_renorm convergence: the integrated x integrated piece with the potential terms only (where renormalization(l1, l2) is half of a
near-cancelling difference) at n_renorm = 8..64 against 128, everything else at the library defaults."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import coffe_b200
from coffe_b200.default_tables import load_default_tableset
H = 0.67
z, sep, ls = np.linspace(0.5, 2.0, 8), np.linspace(10.0, 300.0, 60) / H, [0, 2, 4]
eng = coffe_b200.Engine(0)
eng.upload_tables(load_default_tableset())
for contrib in (["g4", "g5"], ["g4"], list(coffe_b200.capi.CONTRIBUTIONS)):
    eng.set_options(contrib)
    eng.set_quadrature(n_renorm=128)
    ref = eng.multipoles_pieces(z, sep, ls)[2]
    scale = np.max(np.abs(ref), axis=1, keepdims=True)
    for n in (8, 12, 16, 24, 32, 64):
        eng.set_quadrature(n_renorm=n)
        got = eng.multipoles_pieces(z, sep, ls)[2]
        print(contrib if len(contrib) < 4 else "all", n, f"max err rel. to the piece's l=0..4 scale {np.max(np.abs(got - ref) / scale):.2e}", flush=True)
