"""Generate tests/golden/class_midpoint.npz: the reference built WITH CLASS (oracle/_ref/libcoffe_ref_class.so, `make -C oracle class`:
/root/reference/src with -DHAVE_CLASS + the vendored external/class_public compiled in place) run with a non-linear power spectrum
(halofit) and midpoint_approximation = 1 -- SURVEY.md 8f rank 3, src/functions.c:136-155 ff., src/integrals.c:663-767.

What travels to the GPU box (CLASS itself cannot: it reads bbn/ files under /root/reference):
  * the inputs CLASS produced inside the reference: P_lin(k, z = 0) (par->power_spectrum) and the rows P_nl(k, z_i) of the reference's
    2-D spline on its 100 redshifts z_i = 3 i/100 (src/parser.c:655-700) -- 124 KB;
  * the reference's own results for den + rsd + len: multipoles l = 0, 2, 4 (stock Cuhre path, so the integrated pieces carry its
    5e-4) and, per piece, the CONVERGED integrals of its integrands under oracle/converged.py (Cuhre directly for the 2-D piece);
  * samples of its 2-D tables I^n_l(z, r) to check the product's builder against.
Run here:  make -C oracle class && python tests/golden/make_class_midpoint.py      (CLASS: ~20 s, tables: ~15 s, oracle: minutes)"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import coffe_oracle as O  # noqa: E402
from oracle.converged import ConvergedOracle  # noqa: E402

H = 0.67
LIB = ROOT / "oracle" / "_ref" / "libcoffe_ref_class.so"
FAMILY = [(0, 0), (0, 2), (0, 4), (1, 1), (1, 3), (2, 0), (2, 2), (3, 1)]
CONTRIB = ["den", "rsd", "len"]
Z_MEAN, SEP, LS = [0.8, 1.2], np.array([20.0, 60.0, 120.0]) / H, [0, 2, 4]

r = O.RefCoffe(lib_path=LIB)
for f in ("ref_k_min", "ref_k_max", "ref_integral2d_eval"):
    getattr(r.lib, f).restype = C.c_double
r.set_contributions(CONTRIB)
r.set_z_mean(Z_MEAN); r.set_deltaz([0.2, 0.2]); r.set_sep(SEP); r.set_l(LS)
assert r.lib.ref_run_class(r.h_, C.c_int(2), C.c_int(1)) == 0          # COFFE_PK_NONLINEAR_HALOFIT, midpoint_approximation
n = r.lib.ref_power_spectrum_size(r.h_)
k, pk = np.empty(n), np.empty(n)
r.lib.ref_power_spectrum_table(r.h_, k.ctypes.data_as(O.dp), pk.ctypes.data_as(O.dp))
store = {"k": k, "pk_lin": pk, "k_min": np.array(r.lib.ref_k_min(r.h_)), "k_max": np.array(r.lib.ref_k_max(r.h_)),
         "z_mean": np.array(Z_MEAN), "sep": SEP, "l": np.array(LS), "contributions": np.array(CONTRIB)}
# P(k, z_i) rows of the reference's own 2-D spline (its knots in k are CLASS's, the same as the 1-D spline's)
nz = r.lib.ref_pk2d_nz(r.h_)
z2 = np.empty(nz); rows = np.empty((nz, n))
r.lib.ref_pk2d_table(r.h_, z2.ctypes.data_as(O.dp), rows.ctypes.data_as(O.dp))
store["z2d"], store["pk2d"] = z2, rows
r.lib.ref_integrals(r.h_)
# background of the run (unchanged cosmology) is rebuilt by the product; a few samples of the 2-D tables pin its builder
zs = np.array([0.0, 0.45, 1.0, 1.77, 2.9]); rs = np.array([0.3, 5.0, 33.0, 150.0, 1200.0, 6000.0])
store["sample_z"], store["sample_r"] = zs, rs
store["sample_i2d"] = np.array([[[r.lib.ref_integral2d_eval(r.h_, C.c_int(nn), C.c_int(ll), C.c_double(a), C.c_double(b)) for b in rs] for a in zs]
                                for (nn, ll) in FAMILY])
store["stock_multipoles"] = r.multipoles()
# exact knot values of the reference's 2-D tables on a sub-grid (no interpolation involved on either side): pins the product's builder
nzr, nrr = C.c_int(), C.c_int()
kz = np.array([0, 27, 40, 60, 99]); kr = None
knots = []
for (nn, ll) in FAMILY:
    r.lib.ref_integral2d_sizes(r.h_, C.c_int(nn), C.c_int(ll), C.byref(nzr), C.byref(nrr))
    zz_, rr_, vv_ = np.empty(nzr.value), np.empty(nrr.value), np.empty(nzr.value * nrr.value)
    r.lib.ref_integral2d_table(r.h_, C.c_int(nn), C.c_int(ll), zz_.ctypes.data_as(O.dp), rr_.ctypes.data_as(O.dp), vv_.ctypes.data_as(O.dp))
    if kr is None:
        kr = np.unique(np.concatenate([np.arange(0, 44), np.arange(44, nrr.value, 37)]))
    knots.append(vv_.reshape(nrr.value, nzr.value)[np.ix_(kr, kz)])
store["knot_z_idx"], store["knot_r_idx"], store["knot_r"], store["knot_values"] = kz, kr, rr_[kr], np.array(knots)
# converged pieces: local x integrated by Cuhre called directly on the reference's integrand (scheme-independent, 2e-7), integrated x
# integrated by oracle/converged.py at its larger rule size (the fixed rule's own local x integrated value is kept for comparison)
co = ConvergedOracle(r, n_outer=96, n_panel=10, mu_kw=dict(n=8, W=16.0))
conv = np.zeros((len(Z_MEAN), len(SEP), 3, len(LS)))
si_rule = np.zeros((len(Z_MEAN), len(SEP), len(LS)))
si_err = np.zeros((len(Z_MEAN), len(SEP), len(LS)))
for iz, zz in enumerate(Z_MEAN):
    for ir, ss in enumerate(SEP):
        conv[iz, ir] = co.multipoles_pieces(zz, float(ss), LS)
        si_rule[iz, ir] = conv[iz, ir, 1]
        conv[iz, ir, 1], si_err[iz, ir], _ = co.si_multipoles_cuhre(zz, float(ss), LS, epsrel=2e-7, maxeval=10_000_000)
        print(zz, ss * H, conv[iz, ir].sum(0), store["stock_multipoles"][iz, :, ir], "SI rule vs Cuhre", np.abs(si_rule[iz, ir] - conv[iz, ir, 1]) / np.abs(conv[iz, ir, 1]).max(), flush=True)
store["converged_pieces"], store["si_rule"], store["si_cuhre_rel_error_estimate"] = conv, si_rule, si_err
# the same integrands WITHOUT the midpoint approximation (linear D1 D1 I(r) with CLASS's linear spectrum): how much the layer moves the result
out = ROOT / "tests" / "golden" / "class_midpoint.npz"
np.savez_compressed(out, **store)
print(f"wrote {out}: {out.stat().st_size / 1e3:.0f} KB")
