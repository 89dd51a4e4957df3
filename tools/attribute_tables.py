#!/usr/bin/env python
"""Where does the 'same inputs' deviation come from?  (task 3)

Kernels fed the reference's own tables are the baseline; then ONE table group at a time is replaced by the repo's own
(background.py / coffe_b200_integrals) and the all-effects multipoles of a small grid are compared.  Needs a GPU and the oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import coffe_b200  # noqa: E402
from coffe_b200 import capi  # noqa: E402
from oracle import coffe_oracle as O  # noqa: E402
from oracle_tables import FAMILY, DEFAULT_BIAS, DEFAULT_BIAS_Z  # noqa: E402

H = 0.67
ALL = list(capi.CONTRIBUTIONS)
z = [0.5, 1.0, 1.5, 2.0]
sep = np.array([10.0, 20.0, 40.0, 100.0, 150.0, 300.0]) / H
ls = [0, 2, 4]

ref = O.RefCoffe(); ref.set_power_spectrum(O.load_pk())
ref.set_sep(sep); ref.set_z_mean(z); ref.set_deltaz([0.3] * 4); ref.set_l(ls); ref.set_contributions(ALL)
cosmo = coffe_b200.Coffe(sep=sep, z_mean=z, deltaz=[0.3] * 4, l=ls, **{f"has_{v}": True for v in ("density", "rsd", "lensing", "d1", "d2", "g1", "g2", "g3", "g4", "g5")})
cosmo.set_power_spectrum_linear(*O.load_pk())
cosmo._background_init(); cosmo._integrals_init()
own_bg, own = cosmo._background, cosmo._integrals

ref_bg = {name: ref.background_table(name) for name in capi.BG_NAMES}
ref_int = {(n, l): ref.integral_table(n, l) for (n, l) in FAMILY}
ref_i40 = ref.integral_table(4, 0); ref_ren = ref.renorm_table(); ref_ren0 = ref.renorm0_table()
bias = {name: (DEFAULT_BIAS_Z, np.full_like(DEFAULT_BIAS_Z, DEFAULT_BIAS[name])) for name in capi.BIAS_NAMES}


def tableset(bg=None, fam_big=False, fam_small=False, i40=False, ren=False, ren0=False, bg_only=None):
    b = dict(ref_bg)
    if bg:
        for name in (bg_only or capi.BG_NAMES):
            b[name] = own_bg[name]
    ints = []
    for i, (n, l) in enumerate(FAMILY):
        x, y = ref_int[(n, l)]
        y = y.copy()
        oy = own["tables"][own["keys"].index((n, l, 0))]
        assert np.allclose(own["r"], x, rtol=1e-12)
        if fam_big:
            y[x > 1.0] = oy[x > 1.0]
        if fam_small:
            y[x <= 1.0] = oy[x <= 1.0]
        ints.append((n, l, 0, x, y))
    ints.append((4, 0, 0, ref_i40[0], own["i40"] if i40 else ref_i40[1]))
    renorm = (ref_ren[0], own["renorm_z"] if ren else ref_ren[1])
    x0 = ref_ren0[0]
    renorm0 = (x0, own["renorm0"] if ren0 else ref_ren0[1])
    return capi.TableSet(b, bias, ints, 0.3, H, interp_method=5, renorm=renorm, renorm0=renorm0)


eng = coffe_b200.Engine(0)
eng.set_options(ALL)


def run(**kw):
    eng.upload_tables(tableset(**kw))
    return eng.multipoles(z, sep, ls)


base = run()
scale = np.max(np.abs(base), axis=1, keepdims=True)
print(f"{'replaced by the repo own table':42s} max |dev|/scale   max rel dev")
cases = [("background (all)", dict(bg=True)), ("  background: D1 only", dict(bg=True, bg_only=["D1"])), ("  background: f only", dict(bg=True, bg_only=["f"])),
         ("  background: chi, z(chi)", dict(bg=True, bg_only=["comoving_distance", "z_as_chi"])), ("  background: G1, G2", dict(bg=True, bg_only=["G1", "G2"])),
         ("  background: a, H, Hc, Hc'", dict(bg=True, bg_only=["a", "Hz", "conformal_Hz", "conformal_Hz_prime"])),
         ("I^n_l family, r > 1 Mpc (2-FAST)", dict(fam_big=True)), ("I^n_l family, r <= 1 Mpc (direct)", dict(fam_small=True)),
         ("I^4_0", dict(i40=True)), ("renormalisation grid", dict(ren=True)), ("zero-separation renormalisation", dict(ren0=True)),
         ("everything", dict(bg=True, fam_big=True, fam_small=True, i40=True, ren=True, ren0=True))]
for name, kw in cases:
    got = run(**kw)
    print(f"{name:42s} {np.max(np.abs(got - base) / scale):.2e}        {np.max(np.abs(got / base - 1)):.2e}")
for name in ("D1", "f", "comoving_distance", "G1", "conformal_Hz"):
    x, y = ref_bg[name]; yo = own_bg[name][1]
    sel = (x > 0.3) & (x < 2.6)
    print(f"background table {name}: max rel diff on 0.3 < z < 2.6: {np.max(np.abs(yo[sel] / y[sel] - 1)):.2e}")
print("tables: i40 rel", np.max(np.abs(own['i40'] / ref_i40[1] - 1)[1:]), "renorm grid rel-to-scale", np.max(np.abs(own['renorm_z'] - ref_ren[1])) / np.max(np.abs(ref_ren[1])),
      "renorm0", np.max(np.abs(own['renorm0'] - ref_ren0[1])) / np.max(np.abs(ref_ren0[1])))
