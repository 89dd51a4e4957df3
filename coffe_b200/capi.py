"""ctypes mirror of include/coffe_b200.h (the C ABI of libcoffe_b200.so).

Only struct layouts, enum values and the loader live here; no arithmetic.  The loader raises if the
CUDA extension has not been built -- there is deliberately no CPU fallback.
"""
import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "csrc" / "libcoffe_b200.so"

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

BG_NAMES = ("z_as_chi", "a", "Hz", "conformal_Hz", "conformal_Hz_prime", "D1", "f", "G1", "G2", "comoving_distance")
BIAS_NAMES = ("galaxy_bias1", "galaxy_bias2", "galaxy_bias3", "galaxy_bias4",
              "magnification_bias1", "magnification_bias2", "evolution_bias1", "evolution_bias2")
CONTRIBUTIONS = ("den", "rsd", "len", "d1", "d2", "g1", "g2", "g3", "g4", "g5")
ERRORS = {0: "OK", 1: "EINVAL", 2: "ECUDA", 3: "ENOTABLES", 4: "ENONFINITE", 5: "ERANGE"}


class Spline(C.Structure):
    _fields_ = [("x", dp), ("y", dp), ("n", C.c_int)]


class Integral(C.Structure):
    _fields_ = [("n", C.c_int), ("l", C.c_int), ("half_integer", C.c_int), ("result", Spline),
                ("z2d", dp), ("values2d", dp), ("n_z2d", C.c_int)]


class Tables(C.Structure):
    _fields_ = [
        ("background", Spline * 10),
        ("bias", Spline * 8),
        ("interp_method", C.c_int),
        ("integrals", C.POINTER(Integral)),
        ("n_integrals", C.c_int),
        ("renorm_chi", dp),
        ("renorm_z", dp),
        ("n_renorm", C.c_int),
        ("renorm_zero_separation", Spline),
        ("omega_matter", C.c_double),
        ("h", C.c_double),
    ]


class Options(C.Structure):
    _fields_ = [(n, C.c_int) for n in CONTRIBUTIONS] + [
        ("only_cross_correlations", C.c_int),
        ("flatsky_local", C.c_int), ("flatsky_local_nonlocal", C.c_int), ("flatsky_nonlocal", C.c_int), ("midpoint", C.c_int)]


class Quadrature(C.Structure):
    _fields_ = [("n_mu_ni", C.c_int), ("n_mu_panel", C.c_int),
                ("mu_rho0", C.c_double), ("mu_panel_width", C.c_double), ("mu_uniform_extent", C.c_double),
                ("n_outer", C.c_int), ("n_panel_si", C.c_int), ("n_panel_di", C.c_int), ("n_z", C.c_int),
                ("panel_width", C.c_double), ("panel_uniform_extent", C.c_double), ("eps_min", C.c_double),
                ("tail_ratio", C.c_double),
                ("n_mu_panel_di", C.c_int), ("panel_width_di", C.c_double), ("panel_uniform_extent_di", C.c_double),
                ("mu_panel_width_di", C.c_double), ("tail_ratio_di", C.c_double), ("geometric_shift_di", C.c_int),
                ("accuracy_target", C.c_int), ("geometric_growth", C.c_double), ("geometric_growth_di", C.c_double), ("n_renorm", C.c_int), ("tail_map", C.c_int), ("tail_map_di", C.c_int),
                ("mu_geometric_growth", C.c_double), ("mu_geometric_growth_di", C.c_double), ("mu_switch_break_di", C.c_int),
                ("n_outer_min", C.c_int), ("n_outer_step", C.c_double), ("mu_tail_ratio", C.c_double)]


class CovarianceParams(C.Structure):
    _fields_ = [("pk_k", dp), ("pk_p", dp), ("n_pk", C.c_int), ("k_min", C.c_double), ("k_max", C.c_double), ("n_bins", C.c_int),
                ("z_mean", dp), ("deltaz", dp), ("z_min", dp), ("z_max", dp),
                ("fsky", dp), ("density1", dp), ("density2", dp), ("pixelsize", dp),
                ("sep", dp), ("n_sep", C.c_int), ("l", ip), ("n_l", C.c_int), ("pop", C.c_int * 4),
                ("cosmic", C.c_int), ("mixed", C.c_int), ("poisson", C.c_int), ("window", C.c_int), ("average", C.c_int),
                ("integration_method", C.c_int), ("integration_bins", C.c_int)]


class IntegralsParams(C.Structure):
    _fields_ = [("pk_k", dp), ("pk_p", dp), ("n_pk", C.c_int), ("interp_method", C.c_int), ("k_min", C.c_double), ("k_max", C.c_double),
                ("bessel_bins", C.c_int), ("n", ip), ("l", ip), ("half", ip), ("n_tables", C.c_int), ("divergent", C.c_int),
                ("chi_max", C.c_double)]


def darr(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def iarr(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a):
    return a.ctypes.data_as(dp)


def iptr(a):
    return a.ctypes.data_as(ip)


class TableSet:
    """Owns the numpy buffers behind a `Tables` struct (keeps them alive while C reads them)."""

    def __init__(self, background, bias, integrals, omega_matter, h, interp_method=5, renorm=None, renorm0=None, integrals2d=None):
        """background: dict name -> (x, y); bias: dict name -> (x, y); integrals: list of (n, l, half, x, y);
        renorm: (chi, z[n, n]) with z[j, i] <-> (chi_i, chi_j); renorm0: (x, y);
        integrals2d: optional (z [n_z], {(n, l): values [n_r, n_z]}) -- the I^n_l(z, r) tables of the midpoint approximation."""
        self._keep = []
        t = Tables()
        for i, name in enumerate(BG_NAMES):
            if name in background:
                t.background[i] = self._spline(*background[name])
        for i, name in enumerate(BIAS_NAMES):
            if name in bias:
                t.bias[i] = self._spline(*bias[name])
        t.interp_method = int(interp_method)
        arr = (Integral * max(len(integrals), 1))()
        for i, (n, l, half, x, y) in enumerate(integrals):
            arr[i].n, arr[i].l, arr[i].half_integer = int(n), int(l), int(bool(half))
            arr[i].result = self._spline(x, y)
            if integrals2d is not None and not half and (int(n), int(l)) in integrals2d[1]:
                z2, v2 = darr(integrals2d[0]), darr(integrals2d[1][(int(n), int(l))])
                assert v2.shape == (len(x), len(z2)), "values2d must be [n_r, n_z]"
                self._keep += [z2, v2]
                arr[i].z2d, arr[i].values2d, arr[i].n_z2d = ptr(z2), ptr(v2), len(z2)
        self._keep.append(arr)
        t.integrals = C.cast(arr, C.POINTER(Integral))
        t.n_integrals = len(integrals)
        if renorm is not None:
            chi, z = darr(renorm[0]), darr(renorm[1])
            self._keep += [chi, z]
            t.renorm_chi, t.renorm_z, t.n_renorm = ptr(chi), ptr(z), len(chi)
            t.renorm_zero_separation = self._spline(*renorm0)
        else:
            t.n_renorm = 0
        t.omega_matter = float(omega_matter)
        t.h = float(h)
        self.struct = t

    def _spline(self, x, y):
        x, y = darr(x), darr(y)
        assert x.shape == y.shape and x.ndim == 1
        self._keep += [x, y]
        s = Spline()
        s.x, s.y, s.n = ptr(x), ptr(y), len(x)
        return s


def make_options(contributions, only_cross=False, flatsky_local=False, flatsky_local_nonlocal=False, flatsky_nonlocal=False, midpoint=False):
    o = Options()
    for n in CONTRIBUTIONS:
        setattr(o, n, 1 if n in contributions else 0)
    o.only_cross_correlations = int(only_cross)
    o.flatsky_local = int(flatsky_local)
    o.flatsky_local_nonlocal = int(flatsky_local_nonlocal)
    o.flatsky_nonlocal = int(flatsky_nonlocal)
    o.midpoint = int(midpoint)
    return o


def make_quadrature(**kw):
    """keyword arguments = fields of coffe_b200_quadrature; anything omitted keeps the library default."""
    q = Quadrature()
    for k, v in kw.items():
        if not hasattr(q, k):
            raise TypeError(f"unknown quadrature field {k!r}")
        setattr(q, k, v)
    return q


_lib = None


def load():
    """Load libcoffe_b200.so; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ.get("COFFE_B200_LIB", LIB_PATH))      # override: kernel-variant experiments only
    if not path.exists():
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(coffe_b200 has no CPU fallback)")
    lib = C.CDLL(str(path))
    lib.coffe_b200_last_error.restype = C.c_char_p
    lib.coffe_b200_last_error.argtypes = [C.c_void_p]
    _lib = lib
    return lib
