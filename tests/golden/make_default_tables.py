"""Generate tests/golden/default_tables.npz: the tables the reference builds for its DEFAULT_SETTINGS
cosmology with PkL_CLASS.dat (background splines, the I^n_l family, the divergent I^4_0 extras with the
renormalisation grid out to chi(z = 2.3)), i.e. exactly what a patched reference driver hands to
coffe_b200_upload_tables.  bench.py uses them as its synthetic input (the benchmark times signal
evaluation, SURVEY.md 8d: "tables already resident"), so that the bench itself never touches the oracle.

Run here (container with /root/reference and a built oracle):  python tests/golden/make_default_tables.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import coffe_oracle as O  # noqa: E402

ALL = ("den", "rsd", "len", "d1", "d2", "g1", "g2", "g3", "g4", "g5")
FAMILY = [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (2, 0), (2, 2), (3, 1)]
OUT = ROOT / "tests" / "golden" / "default_tables.npz"

r = O.RefCoffe()
r.set_power_spectrum(O.load_pk())
r.set_z_mean([2.0]); r.set_deltaz([0.3])          # renormalisation grid to chi(2.3)
r.set_contributions(ALL)
store = {"omega_matter": np.array(0.3), "h": np.array(0.67), "interp_method": np.array(5)}
z, _ = r.background_table("a")
store["bg_z"] = z
for name in ("a", "Hz", "conformal_Hz", "conformal_Hz_prime", "D1", "f", "G1", "G2", "comoving_distance"):
    store["bg_" + name] = r.background_table(name)[1]
x, y = r.background_table("z_as_chi")
store["bg_z_as_chi_x"], store["bg_z_as_chi_y"] = x, y
x, _ = r.integral_table(0, 0)
# only r <= 2e4 Mpc is ever needed (r <= 2 chi(z = 15)); keeps the file small
keep = x <= 2.0e4
store["ifam_x"] = x[keep]
for (n, l) in FAMILY:
    store[f"ifam_{n}{l}"] = r.integral_table(n, l)[1][keep]
x, y = r.integral_table(4, 0)
store["i40_x"], store["i40_y"] = x, y
chi, zz = r.renorm_table()
store["renorm_chi"], store["renorm_z"] = chi, zz
x, y = r.renorm0_table()
store["renorm0_x"], store["renorm0_y"] = x, y
np.savez_compressed(OUT, **store)
print(f"wrote {OUT}: {OUT.stat().st_size / 1e6:.2f} MB")
