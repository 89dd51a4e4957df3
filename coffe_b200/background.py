"""Background functions of the cosmology: host-side set-up, the counterpart of coffe_background_init
(src/background.c:301-696).  One-off, sequential, O(10^4) work per cosmology -- it stays on the host (DESIGN.md
section 8); the tables it returns are what `coffe_b200_upload_tables` takes as `background[]`.

What is tabulated follows the reference (10 000 uniform knots on z in [0, 15], `background_sampling`):
    a, H(z), conformal H, its conformal-time derivative, D1 (normalised to 1 at z = 0), f, G1, G2, chi(z) [Mpc]
and the inverse z(chi).  How it is computed does not: w(z) = w0 + wa z/(1+z) is integrated in closed form (the
reference integrates linear interpolants of it numerically), chi(z) is a cumulative 8-point Gauss-Legendre sum
(the reference calls QAG per knot at 1e-5), and the growth ODE

    D'' = -3/2 (1 - w/(1+x)) D'/a + 3/2 x/(1+x) D/a^2,    x(a) = Omega_m/(1-Omega_m) a^(3(w0+wa)) e^(3 wa (1-a)),

(src/background.c:192-211) is integrated ONCE from a = 0.05, (D, D') = (0.05, 1) with a fixed-step classical
Runge-Kutta on a fine grid in ln a instead of an adaptive RK2 restarted for every knot (its 1e-6 step control is the
largest difference between the two: ~1e-6 relative in D1 and f, pinned by tests/test_coffe_class.py against the
reference's own golden background file and its compiled sources).
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

    # growth: RK4 in s = ln a from a = 0.05 to a = 1 on a grid fine enough for ~1e-10, sampled at the knots
    def x_of_a(av):
        return omega_m / (1.0 - omega_m) * av ** (3.0 * (w0 + wa)) * np.exp(3.0 * wa * (1.0 - av))

    def rhs(s, y):
        av = np.exp(s)
        wv = w0 + wa * (1.0 - av)
        xv = x_of_a(av)
        # y = (D, dD/da); d/ds = a d/da
        return np.array([av * y[1], av * (-1.5 * (1 - wv / (1 + xv)) * y[1] / av + 1.5 * xv / (1 + xv) * y[0] / av ** 2)])

    s_knots = np.log(a[::-1])                        # increasing, from ln(1/16) to 0
    n_sub = 4000
    s_grid = np.linspace(np.log(0.05), 0.0, n_sub + 1)
    ys = np.empty((n_sub + 1, 2))
    ys[0] = (0.05, 1.0)
    ds = s_grid[1] - s_grid[0]
    for i in range(n_sub):
        s0, y0 = s_grid[i], ys[i]
        k1 = rhs(s0, y0); k2 = rhs(s0 + ds / 2, y0 + ds / 2 * k1); k3 = rhs(s0 + ds / 2, y0 + ds / 2 * k2); k4 = rhs(s0 + ds, y0 + ds * k3)
        ys[i + 1] = y0 + ds / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    # cubic Hermite through the RK4 states (value and derivative known at every grid point)
    dys = np.array([rhs(s, y) for s, y in zip(s_grid, ys)])
    idx = np.clip(np.searchsorted(s_grid, s_knots, side="right") - 1, 0, n_sub - 1)
    u = (s_knots - s_grid[idx]) / ds
    h00, h10, h01, h11 = 2 * u ** 3 - 3 * u ** 2 + 1, u ** 3 - 2 * u ** 2 + u, -2 * u ** 3 + 3 * u ** 2, u ** 3 - u ** 2
    d_at = h00 * ys[idx, 0] + h10 * ds * dys[idx, 0] + h01 * ys[idx + 1, 0] + h11 * ds * dys[idx + 1, 0]
    dp_at = h00 * ys[idx, 1] + h10 * ds * dys[idx, 1] + h01 * ys[idx + 1, 1] + h11 * ds * dys[idx + 1, 1]
    d1 = d_at[::-1]
    f = (dp_at[::-1]) * a / d1
    d1 = d1 / d1[0]

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
