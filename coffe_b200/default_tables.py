"""Loader of a table file in the layout of tests/golden/default_tables.npz (the tables the REFERENCE builds for its
DEFAULT_SETTINGS cosmology with PkL_CLASS.dat, written by tests/golden/make_default_tables.py): a ready-made example of what a
reference-side driver passes to coffe_b200_upload_tables.  The data file is reference output and lives with the test fixtures,
not in the product package; the package's own table pipeline is Coffe._tableset() (background.py + coffe_b200_integrals)."""
from pathlib import Path

import numpy as np

from . import capi

PATH = Path(__file__).resolve().parents[1] / "tests" / "golden" / "default_tables.npz"
FAMILY = [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (2, 0), (2, 2), (3, 1)]
BIAS_Z = np.arange(0.0, 11.0)                       # src/parser.c:42-48
BIAS_DEFAULT = {"galaxy_bias1": 1.0, "galaxy_bias2": 1.0, "galaxy_bias3": 1.0, "galaxy_bias4": 1.0,
                "magnification_bias1": 0.0, "magnification_bias2": 0.0, "evolution_bias1": 0.0, "evolution_bias2": 0.0}


def load_default_tableset(bias=None, path=None):
    d = np.load(path or PATH)
    bg = {name: (d["bg_z"], d["bg_" + name]) for name in capi.BG_NAMES if name != "z_as_chi"}
    bg["z_as_chi"] = (d["bg_z_as_chi_x"], d["bg_z_as_chi_y"])
    b = {name: (BIAS_Z, np.full_like(BIAS_Z, v)) for name, v in BIAS_DEFAULT.items()}
    if bias:
        b.update(bias)
    integrals = [(n, l, 0, d["ifam_x"], d[f"ifam_{n}{l}"]) for (n, l) in FAMILY]
    integrals.append((4, 0, 0, d["i40_x"], d["i40_y"]))
    return capi.TableSet(bg, b, integrals, float(d["omega_matter"]), float(d["h"]), interp_method=int(d["interp_method"]),
                         renorm=(d["renorm_chi"], d["renorm_z"]), renorm0=(d["renorm0_x"], d["renorm0_y"]))
