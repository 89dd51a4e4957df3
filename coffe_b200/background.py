"""Background functions of the cosmology: host-side set-up, the counterpart of coffe_background_init
(src/background.c:301-696).  One-off, sequential, O(10^4) work per cosmology -- it stays on the host (DESIGN.md
section 8); the tables it returns are what `coffe_b200_upload_tables` takes as `background[]`.

What is tabulated follows the reference (10 000 uniform knots on z in [0, 15], `background_sampling`):
    a, H(z), conformal H, its conformal-time derivative, D1 (normalised to 1 at z = 0), f, G1, G2, chi(z) [Mpc]
and the inverse z(chi).  w(z) = w0 + wa z/(1+z) is integrated in closed form (the reference integrates linear interpolants of
it numerically) and chi(z) is a cumulative 8-point Gauss-Legendre sum (the reference calls QAG per knot at 1e-5): both agree
with the reference's tables to 4e-15.  The growth ODE

    D'' = -3/2 (1 - w/(1+x)) D'/a + 3/2 x/(1+x) D/a^2,    x(a) = Omega_m/(1-Omega_m) a^(3(w0+wa)) e^(3 wa (1-a)),

(src/background.c:192-211) is different: the reference's adaptive RK2 at eps_rel = 1e-6, restarted for every knot, leaves a
scheme-specific 2e-6 in D1 and f, so it is integrated by the same scheme (coffe_b200_background_growth, host code of the library);
what remains is the reproducibility of that scheme under rounding: 8e-8 in D1, 6e-7 in f (tests/test_coffe_class.py).
"""
import numpy as np

COFFE_H0 = 1.0 / 2997.92458          # h / Mpc      src/common.h:75-77


def akima_eval(x, y, xq):
    """GSL's non-periodic Akima interpolant through (x, y) at xq; constant for n < 5 equal values"""
    x, y, xq = np.asarray(x, float), np.asarray(y, float), np.atleast_1d(np.asarray(xq, float))
    n = len(x)
    if n < 5:
        return np.interp(xq, x, y)
    m = np.empty(n + 3)
    m[2:n + 1] = np.diff(y) / np.diff(x)
    m[1] = 2 * m[2] - m[3]
    m[0] = 3 * m[2] - 2 * m[3]
    m[n + 1] = 2 * m[n] - m[n - 1]
    m[n + 2] = 3 * m[n] - 2 * m[n - 1]
    i = np.clip(np.searchsorted(x, xq, side="right") - 1, 0, n - 2)
    out = np.empty_like(xq)
    for q, (ii, v) in enumerate(zip(i, xq)):
        k = ii + 2                       # m[k] = slope of interval ii

        def slope_at(kk):                # Akima derivative at the left knot of the interval with slope m[kk]
            ne = abs(m[kk + 1] - m[kk]) + abs(m[kk - 1] - m[kk - 2])
            if ne == 0.0:
                return m[kk]
            al = abs(m[kk - 1] - m[kk - 2]) / ne
            return (1 - al) * m[kk - 1] + al * m[kk]

        h = x[ii + 1] - x[ii]
        b, t = slope_at(k), slope_at(k + 1)
        c = (3 * m[k] - 2 * b - t) / h
        d = (b + t - 2 * m[k]) / h ** 2
        dx = v - x[ii]
        out[q] = y[ii] + dx * (b + dx * (c + dx * d))
    return out


def _spline_on_grid(knots, z):
    """bias spline (11 knots by default, src/parser.c:42-48) on the background grid; vectorised for constants"""
    xk, yk = knots
    yk = np.asarray(yk, float)
    if np.all(yk == yk[0]):
        return np.full_like(z, yk[0])
    return akima_eval(xk, yk, z)


def compute_background(omega_m, omega_gamma, omega_de, h, w0=-1.0, wa=0.0, n=10000, z_max=15.0,
                       magnification_bias1=None, magnification_bias2=None, evolution_bias1=None, evolution_bias2=None):
    """dict name -> (x, y) in the naming of coffe_b200.capi.BG_NAMES.  omega_m = Omega0_cdm + Omega0_baryon."""
    z = z_max * np.arange(n) / (n - 1.0)
    a = 1.0 / (1.0 + z)
    w = w0 + wa * z / (1.0 + z)
    # exp(3 int_0^z (1 + w)/(1 + z') dz')
    wint = (1.0 + z) ** (3.0 * (1.0 + w0 + wa)) * np.exp(-3.0 * wa * z / (1.0 + z))
    e2 = omega_m * (1 + z) ** 3 + omega_gamma * (1 + z) ** 4 + omega_de * wint
    hz = np.sqrt(e2) * h * COFFE_H0
    conf = a * hz
    conf_prime = -((1 + z) ** 3 * (2 * (1 + z) * omega_gamma + omega_m) + (1 + 3 * w) * omega_de * wint) / (1 + z) ** 2 / 2.0 * (h * COFFE_H0) ** 2

    # comoving distance: cumulative Gauss-Legendre over the knot intervals
    gx, gw = np.polynomial.legendre.leggauss(8)
    zl, zr = z[:-1], z[1:]
    zz = 0.5 * (zl + zr)[:, None] + 0.5 * (zr - zl)[:, None] * gx[None, :]
    ww = (1.0 + zz) ** (3.0 * (1.0 + w0 + wa)) * np.exp(-3.0 * wa * zz / (1.0 + zz))
    ee = np.sqrt(omega_m * (1 + zz) ** 3 + omega_gamma * (1 + zz) ** 4 + omega_de * ww)
    seg = (0.5 * (zr - zl))[:, None] * gw[None, :] / ee
    chi = np.concatenate([[0.0], np.cumsum(seg.sum(1))]) / (h * COFFE_H0)

    # growth factor and rate: the reference's own scheme (adaptive RK2(3) at 1e-6, restarted per knot), restated as host code of the
    # library (coffe_b200/csrc/growth_host.cpp).  A converged solution of the same ODE differs from the reference's tables by 2e-6
    # in D1 and up to 2.4e-6 in f -- the whole "same inputs" deviation of the multipoles -- so the scheme is part of the result.
    import ctypes as C
    from . import capi
    lib = capi.load()
    d1, f = np.empty(n), np.empty(n)
    if abs(z_max - 15.0) > 0:
        raise ValueError("the background grid of coffe_background_init ends at z = 15 (src/background.c:446)")
    rc = lib.coffe_b200_background_growth(C.c_double(omega_m), C.c_double(w0), C.c_double(wa), C.c_int(n), capi.ptr(d1), capi.ptr(f))
    if rc:
        raise ValueError("coffe_b200_background_growth: need 0 < omega_m < 1 and at least two knots")

    bias_z = np.arange(0.0, 11.0)
    default0 = (bias_z, np.zeros_like(bias_z))
    g = {}
    for name, sb, eb in (("G1", magnification_bias1, evolution_bias1), ("G2", magnification_bias2, evolution_bias2)):
        s_ = _spline_on_grid(sb if sb is not None else default0, z)
        e_ = _spline_on_grid(eb if eb is not None else default0, z)
        with np.errstate(divide="ignore", invalid="ignore"):
            val = conf_prime / conf ** 2 + (2 - 5 * s_) / (chi * conf) + 5 * s_ - e_
        val[z <= 1e-10] = 0.0            # src/background.c:505-535
        g[name] = val
    return {
        "a": (z, a), "Hz": (z, hz), "conformal_Hz": (z, conf), "conformal_Hz_prime": (z, conf_prime),
        "D1": (z, d1), "f": (z, f), "G1": (z, g["G1"]), "G2": (z, g["G2"]), "comoving_distance": (z, chi),
        "z_as_chi": (chi, z),
    }
