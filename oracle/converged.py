"""
TEST INFRASTRUCTURE ONLY -- the "converged oracle": the REFERENCE's own integrand functions
(functions_nonintegrated / functions_single_integrated / functions_double_integrated, src/functions.c,
compiled unmodified into oracle/_ref/libcoffe_ref.so) integrated with a high-order fixed rule written
independently in numpy.  It defines truth for the 1e-5 gate: the stock reference integrates the same
functions with Cuhre under a 750 000-evaluation budget that the 3-D lensing integrals exhaust at an
estimated error of 1e-3..4e-2, so it cannot arbitrate at 1e-5 itself.

The node placement (panels graded towards the point of closest approach of the two lines of sight and
towards |mu| = 1) restates the idea of coffe_b200/csrc/quadrature.cuh but shares no code with it, and the
integrand values come from the reference, not from the product.
"""
import numpy as np

MU_SWITCH = 0.999      # src/functions.c:2199


def gl01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1), 0.5 * w


def breakpoints(extent, eps, W, Du, ratio, eps_min=1e-4):
    """0 = B_0 < B_1 < ... >= extent: geometric from eps until the step reaches W, uniform W to Du, then geometric"""
    eps = max(eps, eps_min)
    bps = [0.0]
    b = eps
    if eps < W:
        while b < W:
            bps.append(b); b *= 2
    else:
        b = 0.0
    start = b
    k = 0
    while b < Du:
        if b > 0:
            bps.append(b)
        k += 1
        b = start + k * W
    while b < extent:
        bps.append(b); b *= ratio
    bps.append(b)
    return np.array(sorted(set(bps)))


def ridge_breaks(hi, center, eps, W, Du, ratio):
    """break points in [0, hi] of an integration variable with closest approach at `center`, width eps"""
    out = [0.0, hi]
    if center > 0:
        b = breakpoints(center, eps, W, Du, ratio)
        out += list(center - b)
    if hi > center:
        b = breakpoints(hi - center, eps, W, Du, ratio)
        out += list(center + b)
    out = np.array(out)
    out = out[(out >= 0) & (out <= hi)]
    return np.unique(out)


def panel_nodes(breaks, n):
    x, w = gl01(n)
    a, b = breaks[:-1], breaks[1:]
    keep = (b - a) > 0
    a, b = a[keep], b[keep]
    nodes = (a[:, None] + (b - a)[:, None] * x[None, :]).ravel()
    weights = ((b - a)[:, None] * w[None, :]).ravel()
    return nodes, weights


def mu_rule(sep, n=8, rho0=1.0, W=16.0, Du=400.0, ratio=1.4):
    """nodes and weights for int_{-1}^{1} dmu/2, graded in r_perp = sep sqrt(1-mu^2), split at |mu| = 0.999"""
    bp = list(breakpoints(sep, rho0, W, Du, ratio))
    bp = [b for b in bp if b < sep] + [sep, sep * np.sqrt(1 - MU_SWITCH ** 2)]
    phi = np.arcsin(np.clip(np.unique(bp) / sep, 0, 1))
    nodes, weights = panel_nodes(phi, n)
    mu = np.cos(nodes)
    w = weights * np.sin(nodes) / 2
    return np.concatenate([mu, -mu]), np.concatenate([w, w])


class ConvergedOracle:
    """ref: oracle.coffe_oracle.RefCoffe (contributions and tables already configured)"""

    def __init__(self, ref, n_panel=10, W=8.0, Du=640.0, ratio=1.3, n_outer=96, n_mu_ni=64,
                 mu_kw=None):
        self.ref = ref
        self.n_panel, self.W, self.Du, self.ratio, self.n_outer, self.n_mu_ni = n_panel, W, Du, ratio, n_outer, n_mu_ni
        self.mu_kw = mu_kw or {}

    def geometry(self, z, sep, mu):
        chi = float(self.ref.background_eval("comoving_distance", z)[0])
        chi1, chi2 = chi - sep * mu / 2, chi + sep * mu / 2
        den = 2 * chi * chi - mu * mu * sep * sep / 2
        c = (2 * chi * chi - sep * sep + mu * mu * sep * sep / 2) / den
        s = np.sqrt(max(sep * sep * (1 - mu * mu) / den * (1 + c), 0.0))
        return chi1, chi2, c, s

    def si(self, z, sep, mu):
        """int_0^1 dx functions_single_integrated(z, mu, sep, x)          (src/signal.c:584-606)"""
        chi1, chi2, c, s = self.geometry(z, sep, mu)
        # the two summands have their closest approach at lambda2 = chi1 c (x = chi1 c/chi2) and lambda1 = chi2 c
        ba = ridge_breaks(chi2, chi1 * c, chi1 * s, self.W, self.Du, self.ratio) / chi2
        bb = ridge_breaks(chi1, chi2 * c, chi2 * s, self.W, self.Du, self.ratio) / chi1
        br = np.unique(np.clip(np.concatenate([ba, bb]), 0, 1))
        x, w = panel_nodes(br, self.n_panel)
        f = self.ref.f_single(z, mu, sep, x)
        return float(np.sum(w * f))

    def di(self, z, sep, mu):
        """int_0^1 dx1 int_0^1 dx2 functions_double_integrated(z, mu, sep, x1, x2)   (src/signal.c:695-703)"""
        chi1, chi2, c, s = self.geometry(z, sep, mu)
        xo, wo = gl01(self.n_outer)
        X1, X2, WW = [], [], []
        for x1, w1 in zip(xo, wo):
            l1 = chi1 * x1
            br = ridge_breaks(chi2, l1 * c, l1 * s, self.W, self.Du, self.ratio) / chi2
            x2, w2 = panel_nodes(br, self.n_panel)
            X1.append(np.full_like(x2, x1)); X2.append(x2); WW.append(w1 * w2)
        X1, X2, WW = np.concatenate(X1), np.concatenate(X2), np.concatenate(WW)
        f = self.ref.f_double(z, mu, sep, X1, X2)
        return float(np.sum(WW * f))

    def ni(self, z, sep, mu):
        return float(np.ravel(self.ref.f_nonintegrated(z, mu, sep))[0])

    def corrfunc_pieces(self, z, sep, mu, want=(True, True, True)):
        return np.array([self.ni(z, sep, mu) if want[0] else 0.0,
                         self.si(z, sep, mu) if want[1] else 0.0,
                         self.di(z, sep, mu) if want[2] else 0.0])

    def multipoles_pieces(self, z, sep, ls, want=(True, True, True)):
        """[3 pieces, n_l]:  (2l+1) int_0^1 du P_l(mu) piece(mu), mu = 2u-1      (src/signal.c:83-237, 521-752)"""
        out = np.zeros((3, len(ls)))
        P = [np.polynomial.legendre.Legendre.basis(l) for l in ls]
        if want[0]:
            u, w = gl01(self.n_mu_ni)
            mu = 2 * u - 1
            f = self.ref.f_nonintegrated(z, mu, sep)
            for il, l in enumerate(ls):
                out[0, il] = (2 * l + 1) * np.sum(w * P[il](mu) * f)
        if want[1] or want[2]:
            mu, w = mu_rule(sep, **self.mu_kw)
            si = np.array([self.si(z, sep, m) for m in mu]) if want[1] else np.zeros_like(mu)
            di = np.array([self.di(z, sep, m) for m in mu]) if want[2] else np.zeros_like(mu)
            for il, l in enumerate(ls):
                out[1, il] = (2 * l + 1) * np.sum(w * P[il](mu) * si)
                out[2, il] = (2 * l + 1) * np.sum(w * P[il](mu) * di)
        return out

    # ---- redshift average (src/signal.c:243-463) --------------------------------------------------------------------
    def average_nodes(self, z_min, z_max, sep, n_z=16):
        """nodes and weights of the reference's redshift average of one (bin, separation): the bin limits are moved inwards by r/2 in
        comoving distance, z = z1 + (z2 - z1) v with v in [0, 1], weight H0 h / (Hc(z) (1 + z)) and NO Jacobian (z2 - z1)
        (src/signal.c:262-283: the same four lines open every average_multipoles_*_integrand)."""
        H0h = 0.67 / 2997.92458 if not hasattr(self, "H0h") else self.H0h       # COFFE_H0 * par->h (src/common.h:75-77)
        c1 = float(self.ref.background_eval("comoving_distance", z_min)[0]) + sep / 2
        c2 = float(self.ref.background_eval("comoving_distance", z_max)[0]) - sep / 2
        z1 = float(self.ref.background_eval("z_as_chi", c1)[0]); z2 = float(self.ref.background_eval("z_as_chi", c2)[0])
        v, w = gl01(n_z)
        z = z1 + (z2 - z1) * v
        return z, w * H0h / (self.ref.background_eval("conformal_Hz", z) * (1 + z))

    def average_multipoles_pieces(self, z_min, z_max, sep, ls, n_z=16, want=(True, True, True)):
        """[3 pieces, n_l] of the redshift-averaged multipoles of one bin: Gauss-Legendre in v over multipoles_pieces(z(v))"""
        z, w = self.average_nodes(z_min, z_max, sep, n_z)
        return sum(wk * self.multipoles_pieces(float(zk), sep, ls, want) for zk, wk in zip(z, w))

    # ---- pieces with the scheme-independent integrators where they exist --------------------------------------------
    def si_multipoles_cuhre(self, z, sep, ls, epsrel=1e-7, maxeval=50_000_000):
        """local x integrated multipoles by the vendored Cuhre called directly on the reference's integrand (2-D, converges):
        returns (values[n_l], estimated relative errors[n_l], fail flags[n_l])"""
        out = [self.ref.cuhre_direct(0, z, sep, 0.0, int(l), epsrel, maxeval) for l in ls]
        val = np.array([o[0] for o in out])
        return val, np.array([o[1] for o in out]) / np.max(np.abs(val)), np.array([o[3] for o in out])

    def di_corrfunc_cuhre(self, z, sep, mu, epsrel=1e-7, maxeval=200_000_000):
        """integrated x integrated 2PCF at fixed mu by Cuhre directly (2-D): (value, estimated relative error, fail)"""
        v, e, n, f = self.ref.cuhre_direct(1, z, sep, mu, 0, epsrel, maxeval)
        return v, e / abs(v) if v else e, f
