"""
TEST INFRASTRUCTURE ONLY -- ctypes front-end of the parity oracle (oracle/_ref/libcoffe_ref.so).

The shared object is the reference's own C sources (compiled in place from /root/reference/src by
oracle/Makefile against the GSL/FFTW API shims in oracle/shim) plus oracle/ref_driver.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product package `coffe_b200` never does.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIBS = {}

NONINTEGRATED, SINGLE_INTEGRATED, DOUBLE_INTEGRATED = 0, 1, 2
CORRFUNC, MULTIPOLES, AVERAGE_MULTIPOLES = 1, 2, 3
CONTRIBUTIONS = ("den", "rsd", "len", "d1", "d2", "g1", "g2", "g3", "g4", "g5")
BG_NAMES = ("z_as_chi", "a", "Hz", "conformal_Hz", "conformal_Hz_prime", "D1", "f", "G1", "G2", "comoving_distance")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(dp)


def available(fast=False):
    return (HERE / "_ref" / ("libcoffe_ref_fast.so" if fast else "libcoffe_ref.so")).exists()


def load(fast=False, path=None):
    """path: another build of oracle/ref_driver.c with the same entry points -- the drop-in harness
    (dropin/_build/libcoffe_dropin_harness.so: the same driver over COFFE's symbols as exported by the drop-in library)"""
    name = str(path) if path else ("libcoffe_ref_fast.so" if fast else "libcoffe_ref.so")
    if name in _LIBS:
        return _LIBS[name]
    os.environ.setdefault("CUBACORES", "0")  # coffe.pyx:507
    lib = C.CDLL(name if path else str(HERE / "_ref" / name))
    lib.ref_new.restype = C.c_void_p
    for f in ("ref_background_eval", "ref_integral_eval", "ref_renorm_eval", "ref_integrate",
              "ref_f_nonintegrated", "ref_f_single", "ref_f_double", "ref_cuhre_direct"):
        if hasattr(lib, f):
            getattr(lib, f).restype = C.c_double
    _LIBS[name] = lib
    return lib


def gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1), 0.5 * w


def composite_gl_01(panels, n):
    """composite Gauss-Legendre rule on [0,1] with `panels` equal panels of n nodes."""
    x, w = gauss_legendre_01(n)
    xs = np.concatenate([(x + p) / panels for p in range(panels)])
    ws = np.concatenate([w / panels for _ in range(panels)])
    return xs, ws


class RefCoffe:
    """Thin mirror of the subset of `coffe.Coffe` (python/coffe/coffe.pyx) the parity tests need."""

    def __init__(self, fast=False, lib_path=None, **kw):
        self.lib = load(fast, lib_path)
        self.h_ = C.c_void_p(self.lib.ref_new())
        self._n = {"sep": 100, "mu": 1, "z_mean": 1, "l": 3, "z_min": 1}
        for k, v in kw.items():
            getattr(self, "set_" + k)(v)

    def __del__(self):
        try:
            self.lib.ref_free(self.h_)
        except Exception:
            pass

    # --- setters -------------------------------------------------------------------------
    def set_power_spectrum(self, kpk):
        k, pk = _d(kpk[0]), _d(kpk[1])
        self.lib.ref_set_power_spectrum(self.h_, _p(k), _p(pk), C.c_int(len(k)))

    def set_cosmology(self, d):
        self.lib.ref_set_cosmology(self.h_, *(C.c_double(d[k]) for k in ("omega_m", "omega_baryon", "omega_gamma", "h", "w0", "wa")))

    def set_contributions(self, names):
        flags = (C.c_int * 10)(*[1 if n in names else 0 for n in CONTRIBUTIONS])
        self.lib.ref_set_contributions(self.h_, flags)

    def set_divergent(self, v):
        self.lib.ref_set_divergent(self.h_, C.c_int(int(v)))

    def set_flags(self, d):
        self.lib.ref_set_flags(self.h_, C.c_int(int(d.get("flatsky_local", 0))), C.c_int(int(d.get("flatsky_local_nonlocal", 0))),
                               C.c_int(int(d.get("flatsky_nonlocal", 0))), C.c_int(int(d.get("only_cross", 0))))

    def _arr(self, name, v):
        a = _d(v)
        getattr(self.lib, "ref_set_" + name)(self.h_, _p(a), C.c_int(len(a)))
        self._n[name] = len(a)

    def set_sep(self, v): self._arr("sep", v)
    def set_mu(self, v): self._arr("mu", v)
    def set_z_mean(self, v): self._arr("z_mean", v)
    def set_deltaz(self, v): self._arr("deltaz", v)
    def set_z_min(self, v): self._arr("z_min", v)
    def set_z_max(self, v): self._arr("z_max", v)
    def set_fsky(self, v): self._arr("fsky", v)
    def set_pixelsize(self, v): self._arr("pixelsize", v)
    def set_density1(self, v): self._arr("density1", v)
    def set_density2(self, v): self._arr("density2", v)

    def set_l(self, v):
        a = np.ascontiguousarray(v, dtype=np.int32)
        self.lib.ref_set_l(self.h_, a.ctypes.data_as(ip), C.c_int(len(a)))
        self._n["l"] = len(a)

    def set_bias(self, which, z, v):
        z, v = _d(z), _d(v)
        self.lib.ref_set_bias(self.h_, C.c_int(which), _p(z), _p(v), C.c_int(len(z)))

    def set_integration(self, bins=0, prec=0.0, type_1d=0):
        self.lib.ref_set_integration(self.h_, C.c_int(bins), C.c_double(prec), C.c_int(type_1d))

    def set_sampling(self, background=0, bessel=0):
        self.lib.ref_set_sampling(self.h_, C.c_int(background), C.c_int(bessel))

    def set_covariance(self, pops=(1, 1, 1, 1), cosmic=1, mixed=1, poisson=1, window=0, method=0):
        self.lib.ref_set_covariance(self.h_, (C.c_int * 4)(*pops), C.c_int(cosmic), C.c_int(mixed), C.c_int(poisson),
                                    C.c_int(window), C.c_int(method))

    def set_covariance_bins(self, bins):
        self.lib.ref_set_covariance_bins(self.h_, C.c_int(int(bins)))

    def set_nthreads(self, n):
        self.lib.ref_set_nthreads(self.h_, C.c_int(n))

    # --- tables --------------------------------------------------------------------------
    def background_table(self, name):
        which = BG_NAMES.index(name)
        n = self.lib.ref_background_size(self.h_)
        x, y = np.empty(n), np.empty(n)
        self.lib.ref_background_table(self.h_, C.c_int(which), _p(x), _p(y))
        return x, y

    def background_eval(self, name, x):
        which = BG_NAMES.index(name)
        return np.array([self.lib.ref_background_eval(self.h_, C.c_int(which), C.c_double(v)) for v in np.atleast_1d(x)])

    def integral_table(self, n, l, half=False):
        size = self.lib.ref_integral_size(self.h_, C.c_int(n), C.c_int(l), C.c_int(int(half)))
        if size == 0:
            return None
        x, y = np.empty(size), np.empty(size)
        self.lib.ref_integral_table(self.h_, C.c_int(n), C.c_int(l), C.c_int(int(half)), _p(x), _p(y))
        return x, y

    def integral_eval(self, n, l, r, half=False):
        return np.array([self.lib.ref_integral_eval(self.h_, C.c_int(n), C.c_int(l), C.c_int(int(half)), C.c_double(v))
                         for v in np.atleast_1d(r)])

    def renorm_table(self):
        n = self.lib.ref_renorm_size(self.h_)
        if n == 0:
            return None
        chi, z = np.empty(n), np.empty(n * n)
        self.lib.ref_renorm_table(self.h_, _p(chi), _p(z))
        return chi, z.reshape(n, n)

    def renorm0_table(self):
        n = self.lib.ref_renorm0_size(self.h_)
        if n == 0:
            return None
        x, y = np.empty(n), np.empty(n)
        self.lib.ref_renorm0_table(self.h_, _p(x), _p(y))
        return x, y

    # --- stock outputs -------------------------------------------------------------------
    def multipoles(self):
        out = np.empty(self._n["z_mean"] * self._n["l"] * self._n["sep"])
        n = self.lib.ref_multipoles(self.h_, _p(out))
        assert n == out.size
        return out.reshape(self._n["z_mean"], self._n["l"], self._n["sep"])

    def corrfunc(self):
        out = np.empty(self._n["z_mean"] * self._n["mu"] * self._n["sep"])
        n = self.lib.ref_corrfunc(self.h_, _p(out))
        assert n == out.size
        return out.reshape(self._n["z_mean"], self._n["mu"], self._n["sep"])

    def average_multipoles(self):
        out = np.empty(self._n["z_min"] * self._n["l"] * self._n["sep"])
        n = self.lib.ref_average_multipoles(self.h_, _p(out))
        assert n == out.size
        return out.reshape(self._n["z_min"], self._n["l"], self._n["sep"])

    def covariance(self, average=False):
        nz = self._n["z_min"] if average else self._n["z_mean"]
        nl, ns = self._n["l"], self._n["sep"]
        out = np.empty(nz * nl * ns * nl * ns)
        n = self.lib.ref_covariance(self.h_, C.c_int(int(average)), _p(out))
        assert n == out.size
        return out.reshape(nz, nl, ns, nl, ns)

    def integrate(self, z_mean, sep, mu, l, piece, output, z_min=0.0, z_max=0.0):
        return self.lib.ref_integrate(self.h_, C.c_double(z_mean), C.c_double(z_min), C.c_double(z_max), C.c_double(sep),
                                      C.c_double(mu), C.c_int(l), C.c_int(piece), C.c_int(output))

    def cuhre_direct(self, which, z, sep, mu, l, epsrel=1e-7, maxeval=50_000_000):
        """the vendored Cuhre called directly on the reference's integrands with caller-chosen accuracy (ref_driver.c:
        ref_cuhre_direct).  which: 0 = local x integrated multipole l (2-D), 1 = integrated x integrated 2PCF at mu (2-D),
        2 = integrated x integrated multipole l (3-D).  Returns (value, absolute error estimate, evaluations, fail)."""
        err, ne, fail = C.c_double(), C.c_int(), C.c_int()
        v = self.lib.ref_cuhre_direct(self.h_, C.c_int(which), C.c_double(z), C.c_double(sep), C.c_double(mu), C.c_int(l),
                                      C.c_double(epsrel), C.c_int(int(maxeval)), C.byref(err), C.byref(ne), C.byref(fail))
        return v, err.value, ne.value, fail.value

    # --- raw integrands ------------------------------------------------------------------
    def f_nonintegrated(self, z, mu, sep):
        z, mu, sep = np.broadcast_arrays(_d(z), _d(mu), _d(sep))
        z, mu, sep = _d(z), _d(mu), _d(sep)
        out = np.empty(z.size)
        self.lib.ref_f_nonintegrated_v(self.h_, C.c_int(z.size), _p(z), _p(mu), _p(sep), _p(out))
        return out.reshape(z.shape)

    def f_single(self, z, mu, sep, x):
        z, mu, sep, x = (_d(a) for a in np.broadcast_arrays(_d(z), _d(mu), _d(sep), _d(x)))
        out = np.empty(z.size)
        self.lib.ref_f_single_v(self.h_, C.c_int(z.size), _p(z), _p(mu), _p(sep), _p(x), _p(out))
        return out.reshape(z.shape)

    def f_double(self, z, mu, sep, x1, x2):
        z, mu, sep, x1, x2 = (_d(a) for a in np.broadcast_arrays(_d(z), _d(mu), _d(sep), _d(x1), _d(x2)))
        out = np.empty(z.size)
        self.lib.ref_f_double_v(self.h_, C.c_int(z.size), _p(z), _p(mu), _p(sep), _p(x1), _p(x2), _p(out))
        return out.reshape(z.shape)

    # --- converged oracle ----------------------------------------------------------------
    def multipoles_converged(self, z_mean, sep, ls, mu_rule, x_rule):
        """returns array [3 pieces, n_l]; rules are (nodes, weights) on [0, 1]."""
        ls = np.ascontiguousarray(ls, dtype=np.int32)
        um, wm = _d(mu_rule[0]), _d(mu_rule[1])
        ux, wx = _d(x_rule[0]), _d(x_rule[1])
        out = np.empty(3 * len(ls))
        self.lib.ref_multipoles_converged(self.h_, C.c_double(z_mean), C.c_double(sep), ls.ctypes.data_as(ip), C.c_int(len(ls)),
                                          _p(um), _p(wm), C.c_int(len(um)), _p(ux), _p(wx), C.c_int(len(ux)), _p(out))
        return out.reshape(3, len(ls))

    def corrfunc_converged(self, z_mean, sep, mu, x_rule):
        ux, wx = _d(x_rule[0]), _d(x_rule[1])
        out = np.empty(3)
        self.lib.ref_corrfunc_converged(self.h_, C.c_double(z_mean), C.c_double(sep), C.c_double(mu), _p(ux), _p(wx),
                                        C.c_int(len(ux)), _p(out))
        return out


def load_pk(path=None):
    """PkL_CLASS.dat (k in h/Mpc, P in (Mpc/h)^3) -> (k [1/Mpc], P [Mpc^3]) with h = 0.67 (test_coffe.py:120-124)."""
    if path is None:
        path = HERE.parent / "coffe_b200" / "data" / "PkL_CLASS.dat"
    k, pk = np.loadtxt(path, unpack=True)
    h = 0.67
    return k * h, pk / h ** 3
