// Integrands of the relativistic 2-point function: every auto/cross term of
// {den, rsd, d1, d2, g1, g2, g3} (local) and {len, g4, g5} (line-of-sight integrated).
//
// What is computed follows the reference term by term (src/functions.c):
//   local x local        functions_nonintegrated      src/functions.c:89-1484
//   local x integrated   functions_single_integrated  src/functions.c:1486-3457
//   integrated^2         functions_double_integrated  src/functions.c:3460-4016
// but it is organised for a GPU inner loop instead of as 55 independent branches:
//   * each term is (amplitude that depends only on the two end points) x (geometric bracket of
//     I^n_l(r) values), so the amplitudes are folded once per (z, r, mu) work item into a handful of
//     coefficients and the per-node work is one interleaved table record plus polynomial brackets;
//   * the three potential terms g1, g2, g3 enter every formula only through
//     Phi_i = [g1](1+G_i) + [g2](5 s_i - 2) + [g3](f_i - 1), so their 6 autos/crosses and 12 crosses
//     with den/rsd/d1/d2/len/g4/g5 collapse to bilinear forms in Phi_1, Phi_2;
//   * the two halves of every local x integrated term are the same function with the roles of the two
//     galaxies swapped, so one evaluator (si_half_eval) serves both, each half on its own nodes.
// Reference quirks reproduced on purpose: d1-len second half uses bg.G1 and s2
// (functions.c:2776-2779), den-g5 multiplies by a(z) (functions.c:3178-3187), g5-len uses Hc(Hc-1)
// (functions.c:3930-3959), rsd-len switches its I^0_4 bracket at |mu| >= 0.999 (functions.c:2199),
// len-len falls back to 4 + 24 l1 l2 I^1_1(0)/15 for r^2 <= 1e-20 (functions.c:3814-3833).
#pragma once
#include "compact.cuh"

namespace coffe {

struct Options {
    int den, rsd, len, d1, d2, g1, g2, g3, g4, g5;
    int only_cross;
    int flatsky_local, flatsky_local_nonlocal, flatsky_nonlocal;
    int divergent;
    int midpoint;         /* pk_type != linear with midpoint_approximation: the standard terms read I^n_l((z_a + z_b)/2, r) (tables.cuh: Mid2DView) */
    int out_multipoles;   /* 1 while serving coffe_multipoles_init: the flat-sky closed forms of src/signal.c:610-631, 707-735
                             exist only for flag_output == MULTIPOLES, not for the 2PCF or the redshift average */
    double Om;     /* Omega0_cdm + Omega0_baryon */
    double H0h;    /* COFFE_H0 * h  [1/Mpc] */
};

COFFE_HD bool opt_has_ni(const Options &o) { return o.den || o.rsd || o.d1 || o.d2 || o.g1 || o.g2 || o.g3; }
COFFE_HD bool opt_has_integrated(const Options &o) { return o.len || o.g4 || o.g5; }
COFFE_HD bool opt_has_si(const Options &o) { return opt_has_ni(o) && opt_has_integrated(o); }

/* everything the formulas need at one end point (galaxy i = 1 or 2 at comoving distance chi) */
struct EndVals {
    double chi, z;
    double a, Hc, D1, f;
    double G;      /* bg.G1 for end 1, bg.G2 for end 2 */
    double G1raw;  /* bg.G1 regardless of the end (needed by the d1-len quirk) */
    double b, s, fevo;
    double phi;    /* [g1](1+G) + [g2](5s-2) + [g3](f-1) */
    double phi_auto[3];
};

COFFE_HD EndVals end_eval_at(const Options &o, const Tables &t, const double chi, const double z, const int which);

COFFE_HD EndVals end_eval(const Options &o, const Tables &t, const double chi, const int which /* 1 or 2 */)
{
    return end_eval_at(o, t, chi, spline_eval(t.bg.z_of_chi, chi), which);
}

/* the same with the redshift of the end point supplied by the caller (kernels: from the line-of-sight table in shared memory) */
COFFE_HD EndVals end_eval_at(const Options &o, const Tables &t, const double chi, const double z, const int which /* 1 or 2 */)
{
    EndVals e;
    e.chi = chi;
    e.z = z;
    const BackgroundAtZ b = bg_eval_all(t.bg, e.z);
    e.a = b.a; e.Hc = b.Hc; e.D1 = b.D1; e.f = b.f;
    e.G = (which == 1) ? b.G1 : b.G2;
    e.G1raw = b.G1;
    e.b = spline_eval(t.bias[which == 1 ? B_GAL1 : B_GAL2], e.z);
    e.s = spline_eval(t.bias[which == 1 ? B_MAG1 : B_MAG2], e.z);
    e.fevo = spline_eval(t.bias[which == 1 ? B_EVO1 : B_EVO2], e.z);
    e.phi_auto[0] = o.g1 ? (1.0 + e.G) : 0.0;
    e.phi_auto[1] = o.g2 ? (5.0 * e.s - 2.0) : 0.0;
    e.phi_auto[2] = o.g3 ? (e.f - 1.0) : 0.0;
    e.phi = e.phi_auto[0] + e.phi_auto[1] + e.phi_auto[2];
    return e;
}

struct Geometry {
    double chi_mean, sep, mu;
    double chi1, chi2;
    double c;      /* cos(theta)            src/functions.c:101-103 */
    double omc;    /* 1 - cos(theta), evaluated without cancellation */
    double s2;     /* 1 - cos^2(theta) */
};

COFFE_HD Geometry geometry(const double chi_mean, const double sep, const double mu)
{
    Geometry g;
    g.chi_mean = chi_mean; g.sep = sep; g.mu = mu;
    g.chi1 = chi_mean - sep * mu / 2.;
    g.chi2 = chi_mean + sep * mu / 2.;
    const double den = 2. * chi_mean * chi_mean - mu * mu * sep * sep / 2.;
    g.c = (2. * chi_mean * chi_mean - sep * sep + mu * mu * sep * sep / 2.) / den;
    g.omc = sep * sep * (1.0 - mu * mu) / den;
    g.s2 = g.omc * (1.0 + g.c);
    return g;
}

/* ---- values at the separation / mean redshift, constant over mu ------------------------------ */
struct SepConst {
    double i[N_ITAB];   /* I^n_l(sep) */
    double imid[N_ITAB];/* midpoint approximation: I^n_l(z_mean, sep) from the 2-D tables (src/functions.c:136-155) */
    double i40;         /* I^4_0(sep) */
    /* mean-redshift values (flat-sky branches and the final growth factor) */
    double z_mean, chi_mean, fm, Hm, bm1, bm2, Gm1, Gm2, D1m, sm1, sm2;
};

COFFE_HD SepConst sep_const(const Options &o, const Tables &t, const double z_mean, const double sep)
{
    SepConst s;
    ifam_eval_all(t.ifam, sep, s.i);
    if (o.midpoint) mid2d_eval_all(t.mid, t.ifam, z_mean, sep, s.imid);
    s.i40 = o.divergent ? hspline_eval(t.i40, sep) : 0.0;
    s.z_mean = z_mean;
    s.chi_mean = bg_chi(t.bg, z_mean);
    const BackgroundAtZ b = bg_eval_all(t.bg, z_mean);
    s.fm = b.f; s.Hm = b.Hc; s.Gm1 = b.G1; s.Gm2 = b.G2; s.D1m = b.D1;
    s.bm1 = spline_eval(t.bias[B_GAL1], z_mean);
    s.bm2 = spline_eval(t.bias[B_GAL2], z_mean);
    s.sm1 = spline_eval(t.bias[B_MAG1], z_mean);
    s.sm2 = spline_eval(t.bias[B_MAG2], z_mean);
    return s;
}

COFFE_HD double legendre_p(const int l, const double x)
{
    if (l == 0) return 1.0;
    double p0 = 1.0, p1 = x;
    for (int k = 1; k < l; ++k) {
        const double p2 = ((2.0 * k + 1.0) * x * p1 - k * p0) / (k + 1.0);
        p0 = p1; p1 = p2;
    }
    return p1;
}

/* ---- local x local ------------------------------------------------------------------------------ */
COFFE_HD double ni_eval(const Options &o, const Tables &t, const SepConst &sc, const Geometry &g,
                        const EndVals &e1, const EndVals &e2)
{
    const double c = g.c, c2 = c * c, s2 = g.s2;
    const double chi1 = g.chi1, chi2 = g.chi2, r = g.sep, r2 = r * r, mu = g.mu;
    const double Om = o.Om, h2 = o.H0h * o.H0h, h4 = h2 * h2;
    const double f1 = e1.f, f2 = e2.f, H1 = e1.Hc, H2 = e2.Hc, b1 = e1.b, b2 = e2.b;
    const double G1 = e1.G, G2 = e2.G, a1 = e1.a, a2 = e2.a;
    const double q1 = 3.0 - e1.fevo, q2 = 3.0 - e2.fevo;
    const bool autos = !o.only_cross;
    double res = 0.0;
    /* Midpoint approximation (HAVE_CLASS, pk_type != linear; src/functions.c:136-155, 188-199, 237-297, 370-571, 677-786, 778-1022):
       the six standard terms den-den, rsd-rsd, d1-d1, den-rsd, den-d1, rsd-d1 take I^n_l(z_mean, r) from the 2-D tables, growth
       included, instead of D1 D1 I^n_l(r).  They are summed apart (res_std) so that the growth factor at the end leaves them alone. */
    double res_std = 0.0;
    {
    const double *I = o.midpoint ? sc.imid : sc.i;
    double &res = res_std;
    if (o.flatsky_local) {
        const double p2 = 0.5 * (3.0 * mu * mu - 1.0);
        const double fm = sc.fm, Hm = sc.Hm;
        if (o.den && autos) res += sc.bm1 * sc.bm2 * I[I00];                                  /* :133-199 */
        if (o.rsd && autos)                                                                    /* :236-297 */
            res += fm * fm * I[I00] / 5. - 4 * fm * fm * I[I02] / 7. * p2 + 8 * fm * fm * I[I04] / 35. * legendre_p(4, mu);
        if (o.d1 && autos)                                                                     /* :369-440 */
            res += Hm * Hm * fm * fm * sc.Gm1 * sc.Gm2 * I[I20] / 3.
                 + 2 * Hm * Hm * fm * fm * sc.Gm1 * sc.Gm2 * r2 * I[I22] * p2 / 3.;
        if (o.den && o.rsd)                                                                    /* :676-738 */
            res += (sc.bm1 * fm / 3. + sc.bm2 * fm / 3.) * I[I00] - (2 * sc.bm1 * fm / 3. + 2 * sc.bm2 * fm / 3.) * I[I02] * p2;
        if (o.den && o.d1)                                                                     /* :789-809 */
            res += (sc.bm1 * fm * Hm * sc.Gm2 * mu + sc.bm2 * fm * Hm * sc.Gm1 * (-mu)) * r * I[I11];
        if (o.rsd && o.d1)                                                                     /* :893-931 */
            res += (-3. / 5. * (fm * fm * Hm * sc.Gm2 * (-mu) + fm * fm * Hm * sc.Gm1 * mu) * I[I11] * r
                    + (fm * fm * Hm * sc.Gm2 * 2. / 5. * legendre_p(3, -mu) + fm * fm * Hm * sc.Gm1 * 2. / 5. * legendre_p(3, mu)) * I[I13] * r);
    }
    else {
        if (o.den && autos) res += b1 * b2 * I[I00];                                          /* :200-228 */
        if (o.rsd && autos) {                                                                  /* :299-360 */
            /* The reference's I^0_4 coefficient
                   [4(3c^2-1)(chi1^4+chi2^4) + chi1 chi2 (3+c^2)(3(3+c^2) chi1 chi2 - 8(chi1^2+chi2^2) c)] / (35 r^4)
               cancels from O((chi/r)^4) to O(1) (rounding noise ~1e-6 at r = 15 Mpc in the reference).  With
               c = (chi1^2+chi2^2-r^2)/(2 chi1 chi2) and chi2^2-chi1^2 = 2 chi_mean r mu it is identically
                   [560 X^2 mu^4 - 240 X mu^2 S + 12 S^2 + 24 X mu^2 r^2 + 4 r^2 S + 3 r^4] / (16 chi1^2 chi2^2 35),
               X = chi_mean^2, S = chi1^2 + chi2^2, which has no cancellation. */
            const double X = g.chi_mean * g.chi_mean, S = chi1 * chi1 + chi2 * chi2, m2 = mu * mu;
            const double k4 = (560 * X * X * m2 * m2 - 240 * X * m2 * S + 12 * S * S + 24 * X * m2 * r2 + 4 * r2 * S + 3 * r2 * r2)
                            / (16. * chi1 * chi1 * chi2 * chi2 * 35.);
            res += f1 * f2 * (1 + 2 * c2) / 15 * I[I00]
                 - f1 * f2 / 21. * ((1 + 11. * c2) - 18 * c * s2 * chi1 * chi2 / r2) * I[I02]
                 + f1 * f2 * k4 * I[I04];
        }
        if (o.d1 && autos)                                                                     /* :441-571 */
            res += H1 * H2 * f1 * f2 * G1 * G2 * (c * I[I20] / 3. - ((chi2 - chi1 * c) * (chi1 - chi2 * c) + r2 * c / 3.) * I[I22]);
        if (o.den && o.rsd)                                                                    /* :739-786 */
            res += (b1 * f2 / 3. + b2 * f1 / 3.) * I[I00]
                 - (b1 * f2 * (2. / 3. - s2 * (chi1 / r) * (chi1 / r)) + b2 * f1 * (2. / 3. - s2 * (chi2 / r) * (chi2 / r))) * I[I02];
        if (o.den && o.d1)                                                                     /* :810-827 */
            res += -(b1 * f2 * H2 * G2 * (chi1 * c - chi2) + b2 * f1 * H1 * G1 * (chi2 * c - chi1)) * I[I11];
        if (o.rsd && o.d1) {                                                                   /* :932-1022 */
            const double pa = (1. - 3 * c2) * chi2 * chi2 * chi2 + c * (5. + c2) * chi2 * chi2 * chi1
                            - 2 * (2. + c2) * chi2 * chi1 * chi1 + 2 * chi1 * chi1 * chi1 * c;
            const double pb = (1. - 3 * c2) * chi1 * chi1 * chi1 + c * (5. + c2) * chi1 * chi1 * chi2
                            - 2 * (2. + c2) * chi1 * chi2 * chi2 + 2 * chi2 * chi2 * chi2 * c;
            res += (f1 * f2 * H2 * G2 * ((1. + 2 * c2) * chi2 - 3 * chi1 * c) / 5.
                    + f2 * f1 * H1 * G1 * ((1. + 2 * c2) * chi1 - 3 * chi2 * c) / 5.) * I[I11]
                 + (f1 * f2 * H2 * G2 * pa / 5 + f2 * f1 * H1 * G1 * pb / 5) * I[I13] / r2;
        }
    }
    }
    const double *I = sc.i;
    /* divergent bracket I^4_0(r) - renormalisation(chi1, chi2)   src/functions.c:574-597 */
    double ren = 0.0;
    if (o.divergent && (o.d2 || o.g1 || o.g2 || o.g3))
        ren = sc.i40 - renorm2d_eval(t.ren2d, chi1, chi2);
    /* the remaining terms have no flat-sky variant in the reference */
    if (o.d2 && autos) res += q1 * q2 * H1 * H1 * H2 * H2 * f1 * f2 * ren;                     /* :574-597 */
    if (o.g1 || o.g2 || o.g3) {
        /* g_i-g_j autos and crosses: 9 Om^2/(4 a1 a2) x bilinear form in Phi      :598-669, 1289-1368 */
        double bil = e1.phi * e2.phi;
        if (!autos) bil -= e1.phi_auto[0] * e2.phi_auto[0] + e1.phi_auto[1] * e2.phi_auto[1] + e1.phi_auto[2] * e2.phi_auto[2];
        res += h4 * (9 * Om * Om / 4 / a1 / a2 * bil * ren);
        if (o.den)                                                                             /* :840-881 */
            res += h2 * (-(b1 * 3 * Om / 2 / a2 * e2.phi + b2 * 3 * Om / 2 / a1 * e1.phi) * I[I20]);
        if (o.rsd) {                                                                           /* :1040-1095 */
            const double T1 = 2. / 3 * r2 - s2 * chi1 * chi1, T2 = 2. / 3 * r2 - s2 * chi2 * chi2;
            res += h2 * (-(Om / 2. / a2 * f1 * e2.phi + Om / 2. / a1 * f2 * e1.phi) * I[I20]
                         + (3 * Om / 2. / a2 * f1 * e2.phi * T2 + 3 * Om / 2. / a1 * f2 * e1.phi * T1) * I[I22]);
        }
        if (o.d1)                                                                              /* :1110-1149 */
            res += h2 * ((3 * Om / 2. / a2 * H1 * f1 * e2.phi * (chi2 * c - chi1)
                          + 3 * Om / 2. / a1 * H2 * f2 * e1.phi * (chi1 * c - chi2)) * I[I31]);
        if (o.d2)                                                                              /* :1150-1230 */
            res += h2 * (-(3 * q1 * Om / 2. / a2 * H1 * H1 * f1 * e2.phi + 3 * q2 * Om / 2. / a1 * H2 * H2 * f2 * e1.phi) * ren);
    }
    if (o.den && o.d2) res += (q2 * b1 * f2 * H2 * H2 + q1 * b2 * f1 * H1 * H1) * I[I20];      /* :828-839 */
    if (o.rsd && o.d2) {                                                                       /* :1023-1039 */
        const double T1 = 2. / 3 * r2 - s2 * chi1 * chi1, T2 = 2. / 3 * r2 - s2 * chi2 * chi2;
        res += (q2 / 3 * f1 * f2 * H2 * H2 + q1 / 3 * f2 * f1 * H1 * H1) * I[I20]
             - (q2 * f1 * f2 * H2 * H2 * T2 + q1 * f2 * f1 * H1 * H1 * T1) * I[I22];
    }
    if (o.d1 && o.d2)                                                                          /* :1098-1109 */
        res += -(q2 * H1 * H2 * H2 * f1 * f2 * (chi2 * c - chi1) + q1 * H2 * H1 * H1 * f2 * f1 * (chi1 * c - chi2)) * I[I31];

    /* growth: D1(z_mean)^2 in flat-sky, D1(z1) D1(z2) otherwise            src/functions.c:1456-1467 */
    if (o.midpoint) return res_std + (o.flatsky_local ? res * sc.D1m * sc.D1m : res * e1.D1 * e2.D1);
    res += res_std;
    return o.flatsky_local ? res * sc.D1m * sc.D1m : res * e1.D1 * e2.D1;
}

/* ---- local x integrated ------------------------------------------------------------------------
   One half: the local effect sits at the "source" end (distance cs), the integrated effect runs along
   the other line of sight, lambda in [0, co].  Amplitudes follow src/functions.c:1558-3423. */
struct SIHalf {
    double cs, co;
    double a_denlen, a_rsdlen, a_d1len, a_pot;      /* x (co - lambda) W(lambda) x bracket */
    double g4_i20, g4_rsd, g4_d1, g4_ren;           /* x W(lambda) x bracket */
    double g5_i20, g5_rsd, g5_d1, g5_ren;           /* x V(lambda) x {Wa | W} x bracket (co, G_other folded in) */
    int alt_i04;                                    /* |mu| >= 0.999 */
    int any_lens, any_g4, any_g5, need_ren;
    int mid;                                        /* midpoint approximation: den/rsd/d1 x len read the 2-D tables */
    double zs, D1s;                                 /* redshift and growth factor of the local end */
    /* The lensing bracket of the half as polynomials in lambda with the amplitudes folded in (si_half_setup), so that a node
       costs a few Horner steps instead of re-deriving every coefficient from (cs, cos theta):
         br = I00 P00(l) + [I02 P02(l) + I13 P13(l)] / R^2 + I04 P04(l - cs) / R^4 + I11 P11(l) + I20 p20 + I31 p31 + I22 p22 l
       (|mu| >= 0.999: I04 (p04a[0] + p04a[1] l) in place of the quintic, src/functions.c:2199).  The quintic is held in powers of
       l - cs with coefficients in powers of w = 1 - cos(theta) -- the form in which the reference's O(chi^5) terms have already
       cancelled down to O(chi r^4) (see the comment at its derivation in si_half_setup). */
    double p00[2], p02[4], p13[5], p04[6], p04a[2], p11[3], p20, p31, p22;
};

/* src = local end, oth = the end whose line of sight is integrated.  `first` selects the quirks of the
   first (src = 1) or second (src = 2) summand. */
COFFE_HD SIHalf si_half_setup(const Options &o, const Geometry &g, const EndVals &src, const EndVals &oth, const bool first)
{
    SIHalf h;
    const double Om = o.Om, h2 = o.H0h * o.H0h, h4 = h2 * h2;
    h.cs = src.chi; h.co = oth.chi;
    const double m_oth = 2.0 - 5.0 * oth.s;      /* (2 - 5 s) of the lensed galaxy */
    const double q = 3.0 - src.fevo;
    h.a_denlen = (o.den && o.len && !o.flatsky_local_nonlocal) ? h2 * (-3 * Om / 2.) * src.b * m_oth * src.D1 : 0.0;
    h.a_rsdlen = (o.rsd && o.len && !o.flatsky_local_nonlocal) ? h2 * (3 * Om / 2.) * src.f * m_oth * src.D1 : 0.0;
    /* d1-len: both summands use bg.G1 and (2 - 5 s2)      src/functions.c:2490-2493, 2776-2779 */
    {
        const double m_quirk = first ? m_oth : (2.0 - 5.0 * src.s);
        h.a_d1len = (o.d1 && o.len && !o.flatsky_local_nonlocal) ? h2 * (3 * Om / 2.) * src.Hc * src.f * src.G1raw * m_quirk * src.D1 : 0.0;
    }
    h.a_pot = 0.0;
    if (o.len) {
        if (o.d2) h.a_pot += h2 * (-3 * Om / 2.) * q * src.f * src.Hc * src.Hc * m_oth * src.D1;         /* :3000-3030 */
        if (o.g1 || o.g2 || o.g3) h.a_pot += h4 * (9 * Om * Om / 4.) * src.phi * m_oth * src.D1;            /* :3031-3121 */
    }
    h.g4_i20 = h.g4_rsd = h.g4_d1 = h.g4_ren = 0.0;
    if (o.g4) {
        if (o.den) h.g4_i20 = h2 * (-3 * Om) * src.b * m_oth * src.D1;                                      /* :3122-3141 */
        if (o.rsd) h.g4_rsd = h2 * (3 * Om) * src.f * m_oth * src.D1;                                       /* :3164-3191 */
        if (o.d1) h.g4_d1 = h2 * (3 * Om) * src.Hc * src.f * m_oth * src.D1;                               /* :3222-3238 */
        if (o.d2) h.g4_ren += h2 * (-3 * Om) * q * src.f * src.Hc * src.Hc * m_oth * src.D1;                /* :3260-3278 */
        if (o.g1 || o.g2 || o.g3) h.g4_ren += h4 * (9 * Om * Om / 2.) * src.phi * m_oth * src.D1 / src.a;   /* :3300-3320 ... */
    }
    h.g5_i20 = h.g5_rsd = h.g5_d1 = h.g5_ren = 0.0;
    if (o.g5) {
        const double pre = oth.chi * oth.G;      /* chi_other x G_other(z_other) */
        if (o.den) h.g5_i20 = h2 * (-3 * Om) * pre * src.b * src.D1;                                       /* :3142-3163 */
        if (o.rsd) h.g5_rsd = h2 * (3 * Om) * pre * src.f * src.D1;                                        /* :3192-3221 */
        if (o.d1) h.g5_d1 = h2 * (3 * Om) * pre * src.Hc * src.f * src.D1;                                /* :3239-3259 */
        if (o.d2) h.g5_ren += h2 * (-3 * Om) * pre * q * src.f * src.Hc * src.Hc * src.D1;                  /* :3279-3299 */
        if (o.g1 || o.g2 || o.g3) h.g5_ren += h4 * (9 * Om * Om / 2.) * pre * src.phi * src.D1 / src.a;     /* :3321-3423 */
    }
    h.mid = o.midpoint; h.zs = src.z; h.D1s = src.D1;
    h.alt_i04 = fabs(g.mu) >= 0.999;
    h.any_lens = (h.a_denlen != 0.0) || (h.a_rsdlen != 0.0) || (h.a_d1len != 0.0) || (h.a_pot != 0.0);
    h.any_g4 = o.g4 && opt_has_ni(o);
    h.any_g5 = o.g5 && opt_has_ni(o);
    h.need_ren = o.divergent && ((h.g4_ren != 0.0) || (h.g5_ren != 0.0));
    {
        const double cs = h.cs, cs2 = cs * cs, cs3 = cs2 * cs, c = g.c, c2 = c * c, tc = 2 * c2 - 1, s2 = g.s2, w = g.omc;
        const double ad = h.a_denlen, ar = h.a_rsdlen, a1 = h.a_d1len, ap = h.a_pot;
        /* den-len: 2 cs c I11 - cs^2 s2 l I02 / R^2                                                       src/functions.c:1592-1640 */
        /* rsd-len: (l (1 + 3 tc) - 6 cs c) I00 / 15 - (6 cs^3 c - cs^2 (9 c^2 + 11) l + cs c (3 tc + 19) l^2 - 2 (3 tc + 1) l^3) I02 / (21 R^2)
                    - p5 I04 / (35 R^4)                                                                     src/functions.c:2059-2390 */
        /* d1-len:  2 (c (l^2 - 2 cs^2) + cs l (2 tc - 1)) I11 / 15 + 2 c I20 / 3
                    - (4 cs^4 c - cs^3 (c^2 + 9) l + cs^2 c (c^2 + 5) l^2 - 2 cs (tc - 2) l^3 - 2 c l^4) I13 / (15 R^2)    :2490-2560 */
        h.p00[0] = -ar * 6 * cs * c * (1.0 / 15.0); h.p00[1] = ar * (1 + 3 * tc) * (1.0 / 15.0);
        h.p02[0] = -ar * 6 * cs3 * c * (1.0 / 21.0);
        h.p02[1] = ar * cs2 * (9 * c2 + 11) * (1.0 / 21.0) - ad * cs2 * s2;
        h.p02[2] = -ar * cs * c * (3 * tc + 19) * (1.0 / 21.0);
        h.p02[3] = ar * 2 * (3 * tc + 1) * (1.0 / 21.0);
        h.p13[0] = -a1 * 4 * cs2 * cs2 * c * (1.0 / 15.0);
        h.p13[1] = a1 * cs3 * (c2 + 9) * (1.0 / 15.0);
        h.p13[2] = -a1 * cs2 * c * (c2 + 5) * (1.0 / 15.0);
        h.p13[3] = a1 * 2 * cs * (tc - 2) * (1.0 / 15.0);
        h.p13[4] = a1 * 2 * c * (1.0 / 15.0);
        /* the reference's quintic (functions.c:2205-2214) cancels from O(chi^5) to O(chi r^4); expanded in d = l - cs and
           w = 1 - cos(theta) it reads
             -4 d^4 (2 cs + d) + w (4 d^2 (12 cs^3 + 23 cs^2 d + 15 cs d^2 + 3 d^3)
                + w (-6 l^2 (2 cs^3 + 6 cs^2 d + 6 cs d^2 + d^3) + w (2 cs l^2 (2 cs^2 + 4 cs d + 3 d^2) + w cs^2 l^3))),
           which collected in powers of d gives the coefficients below (every bracket a polynomial in the small w) */
        const double k35 = -ar * (1.0 / 35.0);
        h.p04[0] = k35 * cs3 * cs2 * w * w * (w - 2) * (w + 6);
        h.p04[1] = k35 * cs2 * cs2 * w * w * (w * (3 * w + 16) - 60);
        h.p04[2] = k35 * cs3 * w * (w * (w * (3 * w + 26) - 120) + 48);
        h.p04[3] = k35 * cs2 * w * (w * (w * (w + 20) - 114) + 92);
        h.p04[4] = k35 * 2 * cs * (w * (w * (3 * w - 24) + 30) - 4);
        h.p04[5] = k35 * -2 * (w * (3 * w - 6) + 2);
        h.p04a[0] = ar * 4 * cs * (1.0 / 35.0); h.p04a[1] = ar * 4 * (1.0 / 35.0);
        h.p11[0] = ad * 2 * cs * c - a1 * 4 * c * cs2 * (1.0 / 15.0);
        h.p11[1] = a1 * 2 * cs * (2 * tc - 1) * (1.0 / 15.0);
        h.p11[2] = a1 * 2 * c * (1.0 / 15.0);
        h.p20 = a1 * 2 * c * (1.0 / 3.0);
        h.p31 = ap * 2 * cs * c;
        h.p22 = -ap * cs2 * s2;
    }
    return h;
}

/* The part of the integrand behind the table lookup: rr = R^2, iR = 1/R and I = the nine table values at R are passed in, so that the
   kernel can share one lookup between the two nodes at the same distance from the closest approach (item.cuh: si_half_group).
   SPLIT_REN: renormalization(lambda, chi_source) is integrated apart (si_ren_smooth_partial), only I^4_0(R) is added here. */
template <bool MID, bool SPLIT_REN, class TA>
COFFE_HD double si_half_body(const TA &t, const SIHalf &h, const Geometry &g, const double lambda, const double4 los,
                             const double rr, const double iR, const double *I)
{
    const double c = g.c, s2 = g.s2;
    const double cs = h.cs, l = lambda;
    const double dl = l - cs;
    const double R = rr * iR;
    const double W = los.x, V = los.y, Wa = los.w;
    double res = 0.0;
    /* Midpoint approximation (src/functions.c:1626-1700, 2061-2326, 2492-2918): den-len, rsd-len and d1-len read
       I^n_l((z_src + z(lambda))/2, R) from the 2-D tables in place of D1(z_src) D1(z(lambda)) I^n_l(R).  The growth factors stay
       folded into the amplitudes, so the table value is divided by them; z(lambda) and D1(lambda) follow from W = D1/a, Wa = D1 a.
       The potential terms (a_pot, g4, g5) have no such variant and keep the linear tables. */
    double Imid[N_ITAB];
    const double *Is = I;
    if (MID && h.mid && h.any_lens) {
        const double D1l = sqrt(W * Wa), zl = sqrt(W / Wa) - 1.0;
        t.ifam_mid(R, 0.5 * (h.zs + zl), Imid);
        const double inv = 1.0 / (h.D1s * D1l);
#pragma unroll
        for (int k = 0; k < N_ITAB; ++k) Imid[k] *= inv;
        Is = Imid;
    }
    if (h.any_lens) {
        double br = 0.0;
        if (rr != 0.0) {
            /* the brackets as polynomials in l (coefficients and their derivation: si_half_setup) */
            const double irr = iR * iR;
            br = Is[I00] * fma(h.p00[1], l, h.p00[0])
               + Is[I11] * fma(fma(h.p11[2], l, h.p11[1]), l, h.p11[0])
               + Is[I20] * h.p20
               + irr * (Is[I02] * fma(fma(fma(h.p02[3], l, h.p02[2]), l, h.p02[1]), l, h.p02[0])
                        + Is[I13] * fma(fma(fma(fma(h.p13[4], l, h.p13[3]), l, h.p13[2]), l, h.p13[1]), l, h.p13[0]))
               + I[I31] * h.p31 + I[I22] * (h.p22 * l);
            if (!h.alt_i04)
                br += Is[I04] * (irr * irr) * fma(fma(fma(fma(fma(h.p04[5], dl, h.p04[4]), dl, h.p04[3]), dl, h.p04[2]), dl, h.p04[1]), dl, h.p04[0]);
            else br += Is[I04] * fma(h.p04a[1], l, h.p04a[0]);
            res += (h.co - l) * W * br;
        }
        else {
            /* r == 0 limits                                             src/functions.c:1641-1760, 2391-2480, 2631-2990 */
            if (h.a_denlen != 0.0) br += h.a_denlen * 2 * cs * Is[I11];
            if (h.a_d1len != 0.0) br += h.a_d1len * 2 * Is[I20] / 3.;
            if (h.a_pot != 0.0) br += h.a_pot * (2 * cs * c * I[I31] - cs * cs * l * s2 * I[I22]);
            res += (h.co - l) * W * br;
            if (h.a_rsdlen != 0.0) res += h.a_rsdlen * (h.co - cs) * W * (-2 * cs * Is[I00] / 15.);
        }
    }
    if (h.any_g4 || h.any_g5) {
        double ren = 0.0;
        if (h.need_ren)                                                                             /* :1524-1556 */
            ren = SPLIT_REN ? ((rr == 0.0) ? t.ren0(l) + t.ren2d(l, cs) : I[N_ITAB])
                            : ((rr == 0.0) ? t.ren0(l) : I[N_ITAB] - t.ren2d(l, cs));
        /* thirds as multiplications: a division by a constant stays a division (a 15-instruction serial chain) without fast-math */
        const double b_rsd = fma(rr, 2.0 / 3.0, -(s2 * l * l)) * I[I22] - I[I20] * (1.0 / 3.0);
        const double b_d1 = (l * c - cs) * I[I31];
        if (h.any_g4) res += W * (h.g4_i20 * I[I20] + h.g4_rsd * b_rsd + h.g4_d1 * b_d1 + h.g4_ren * ren);
        if (h.any_g5) res += V * (h.g5_i20 * Wa * I[I20] + W * (h.g5_rsd * b_rsd + h.g5_d1 * b_d1 + h.g5_ren * ren));
    }
    return res;
}

/* integrand of one half at lambda (distance along the integrated line of sight), per unit x = lambda/co.
   los = {W, V, U, Wa}(lambda). */
template <bool MID = false, class TA>
COFFE_HD double si_half_eval(const TA &t, const SIHalf &h, const Geometry &g, const double lambda, const double4 los)
{
    /* r^2 = l^2 + cs^2 - 2 cs l cos(theta), written without the cancellation */
    const double dl = lambda - h.cs;
    double rr = dl * dl + 2.0 * h.cs * lambda * g.omc;
    if (rr < 0) rr = 0;
    const double iR = rr > 0.0 ? COFFE_RSQRT(rr) : 0.0;
    double I[N_ITAB9];
    t.ifam9(rr * iR, I, h.need_ren != 0);    /* the eight I^n_l and, for the renormalised potential terms, I^4_0 in the same lookup */
    return si_half_body<MID, false>(t, h, g, lambda, los, rr, iR, I);
}

/* ---- integrated x integrated ------------------------------------------------------------------- */
struct DIConst {
    double k_ll;              /* 9 Om^2 (2-5s1)(2-5s2)/4                        len-len  :3516-3836 */
    double k_4l;              /* 9 Om^2 (2-5s1)(2-5s2)/2                        g4-len   :3876-3918 */
    double k_5la, k_5lb;      /* 9 Om^2/2 (2-5s2) G1 chi1 ; (2-5s1) G2 chi2     g5-len   :3919-3965 */
    double k_44;              /* 9 Om^2 (2-5s1)(2-5s2)                          g4-g4    :3837-3852 */
    double k_55;              /* 9 Om^2 G1 G2 chi1 chi2                         g5-g5    :3853-3875 */
    double k_45a, k_45b;      /* 9 Om^2 G2 (2-5s1) chi2 ; G1 (2-5s2) chi1       g4-g5    :3966-3988 */
    double h4;
    int any_ll, any_pl, any_ren;
    int mid;                  /* midpoint approximation: len-len reads the 2-D tables (src/functions.c:3564-3836) */
};

COFFE_HD DIConst di_setup(const Options &o, const EndVals &e1, const EndVals &e2)
{
    DIConst k;
    const double Om2 = o.Om * o.Om;
    const double m1 = 2.0 - 5.0 * e1.s, m2 = 2.0 - 5.0 * e2.s;
    const bool autos = !o.only_cross;
    /* With flatsky_nonlocal the reference integrand carries no len-len term at all: its flat-sky branch needs
       output_type != MULTIPOLES, which the Python API never sets, and the full-sky branch is `else if (!flatsky_nonlocal)`
       (src/functions.c:3521-3562, SURVEY.md 8a note 7).  The 3-D integral then holds only the g4/g5 terms
       (src/signal.c:737-752), the closed form of src/signal.c:707-735 supplies len-len for the multipoles, and the
       2PCF / redshift-averaged multipoles lose the term -- reproduced as is. */
    const bool skip_ll = o.flatsky_nonlocal != 0;
    k.k_ll = (o.len && autos && !skip_ll) ? 9. * Om2 * m1 * m2 / 4. : 0.0;
    k.k_4l = (o.g4 && o.len) ? 9 * Om2 / 2. * m1 * m2 : 0.0;
    k.k_5la = (o.g5 && o.len) ? 9 * Om2 / 2. * m2 * e1.G * e1.chi : 0.0;
    k.k_5lb = (o.g5 && o.len) ? 9 * Om2 / 2. * m1 * e2.G * e2.chi : 0.0;
    k.k_44 = (o.g4 && autos) ? 9 * Om2 * m1 * m2 : 0.0;
    k.k_55 = (o.g5 && autos) ? 9 * Om2 * e1.G * e2.G * e1.chi * e2.chi : 0.0;
    k.k_45a = (o.g4 && o.g5) ? 9 * Om2 * e2.G * m1 * e2.chi : 0.0;
    k.k_45b = (o.g4 && o.g5) ? 9 * Om2 * e1.G * m2 * e1.chi : 0.0;
    const double h2 = o.H0h * o.H0h;
    k.h4 = h2 * h2;
    k.mid = o.midpoint;
    k.any_ll = k.k_ll != 0.0;
    k.any_pl = (k.k_4l != 0.0) || (k.k_5la != 0.0) || (k.k_5lb != 0.0);
    k.any_ren = o.divergent && ((k.k_44 != 0.0) || (k.k_55 != 0.0) || (k.k_45a != 0.0) || (k.k_45b != 0.0));
    return k;
}

/* integrand per unit (x1, x2) = (l1/chi1, l2/chi2); los_i = {W, V, U, Wa}(l_i) */
template <class TA>
COFFE_HD double di_eval(const TA &t, const DIConst &k, const Geometry &g, const double l1, const double l2,
                        const double4 los1, const double4 los2)
{
    const double c = g.c;
    const double d = l1 - l2;
    double r2 = d * d + 2.0 * l1 * l2 * g.omc;
    if (r2 < 0) r2 = 0;
    const double ir = r2 > 0.0 ? COFFE_RSQRT(r2) : 0.0;     /* 1/R: one reciprocal square root serves R, 1/r^2, 1/r^4 */
    const double R = r2 * ir;
    double I[N_ITAB];
    t.ifam_all(R, I);
    const double ll = l1 * l2;
    const double WW = los1.x * los2.x;
    const double e1 = g.chi1 - l1, e2 = g.chi2 - l2;    /* chi_i (1 - x_i) */
    double res = 0.0;
    if (k.any_ll) {
        /* midpoint approximation: I^n_l((z(l1) + z(l2))/2, R) from the 2-D tables over the growth factors folded into W W */
        double Imid[N_ITAB];
        const double *Il = I;
        if (k.mid) {
            const double z1 = sqrt(los1.x / los1.w) - 1.0, z2 = sqrt(los2.x / los2.w) - 1.0;
            t.ifam_mid(R, 0.5 * (z1 + z2), Imid);
            const double inv = 1.0 / sqrt(los1.x * los1.w * los2.x * los2.w);
#pragma unroll
            for (int j = 0; j < N_ITAB; ++j) Imid[j] *= inv;
            Il = Imid;
        }
        double br;
        if (r2 > 1e-20) {
            const double ir2 = ir * ir;
            const double cm = -g.s2;     /* cos^2(theta) - 1 */
            br = 2 * cm * ll * Il[I00] / 5.
               + 4 * c * Il[I20] / 3.
               + 4 * c * (r2 + 6 * c * ll) * Il[I11] / 15.
               + 2 * cm * ll * (2 * r2 + 3 * c * ll) * Il[I02] / 7. * ir2
               + 2 * c * (2 * r2 * r2 + 12 * c * r2 * ll + 15 * cm * ll * ll) * Il[I13] / 15. * ir2
               + cm * ll * (6 * r2 * r2 + 30 * c * r2 * ll + 35 * cm * ll * ll) * Il[I04] / 35. * ir2 * ir2;
        }
        /* coincident points: the linear branch reads I^1_1(0), the midpoint branch of the reference its (2, 0) table (src/functions.c:3796-3808) */
        else br = 4. + 24. * ll * (k.mid ? Il[I20] : I[I11]) / 15.;
        res += k.k_ll * e1 * e2 * WW * br;
    }
    if (k.any_pl) {
        const double br = (r2 != 0.0) ? 2 * ll * c * I[I31] - ll * ll * g.s2 * I[I22] : 2 * ll * I[I31];
        /* (1 - x2)/x2 = (chi2 - l2)/l2 */
        const double w2 = e2 / l2, w1 = e1 / l1;
        res += WW * br * (k.k_4l * (w2 + w1) + k.k_5la * los1.z * w2 + k.k_5lb * los2.z * w1);
    }
    if (k.any_ren) {
        const double ren = (r2 <= 1e-12) ? t.ren0(l1) : t.i40(R) - t.ren2d(l1, l2);
        res += WW * ren * (k.k_44 + k.k_55 * los1.y * los2.y + k.k_45a * los2.y + k.k_45b * los1.y);
    }
    return res * k.h4;
}

/* line-of-sight functions at lambda through the reference's own two-level lookup (z = z(lambda), then
   the background splines at z): {W = D1/a, V = Hc (f-1), U = Hc (Hc-1), Wa = D1 a} */
COFFE_HD double4 los_point(const Tables &t, const double lambda)
{
    const double z = spline_eval(t.bg.z_of_chi, lambda);
    const BackgroundAtZ b = bg_eval_all(t.bg, z);
    double4 o;
    o.x = b.D1 / b.a;
    o.y = b.Hc * (b.f - 1.0);
    o.z = b.Hc * (b.Hc - 1.0);
    o.w = b.D1 * b.a;
    return o;
}

/* Hermite knot i of the line-of-sight table (tables.cuh: LosView): values at lambda_i = i h and h d/dlambda from
   fourth-order differences of the same two-level lookup (step h/2: wide enough to average over the kinks of the
   underlying Akima interpolants, which are rough at the 1e-7 level; one-sided at the two ends of the table) */
COFFE_HD void los_knot(const Tables &t, const int i, const double h, const double chi_max, double2 out[kLosStride])
{
    double lam = i * h;
    if (lam > chi_max) lam = chi_max;
    const double d = h / 2.0;
    const double4 f0 = los_point(t, lam);
    double dv[4];
    if (lam - 2 * d >= 0.0 && lam + 2 * d <= chi_max) {
        const double4 m2 = los_point(t, lam - 2 * d), m1 = los_point(t, lam - d), p1 = los_point(t, lam + d), p2 = los_point(t, lam + 2 * d);
        dv[0] = (m2.x - 8 * m1.x + 8 * p1.x - p2.x) / 12.; dv[1] = (m2.y - 8 * m1.y + 8 * p1.y - p2.y) / 12.;
        dv[2] = (m2.z - 8 * m1.z + 8 * p1.z - p2.z) / 12.; dv[3] = (m2.w - 8 * m1.w + 8 * p1.w - p2.w) / 12.;
    }
    else {
        const double sg = lam - 2 * d < 0.0 ? 1.0 : -1.0;
        const double4 a1 = los_point(t, lam + sg * d), a2 = los_point(t, lam + sg * 2 * d), a3 = los_point(t, lam + sg * 3 * d),
                      a4 = los_point(t, lam + sg * 4 * d);
        dv[0] = sg * (-25 * f0.x + 48 * a1.x - 36 * a2.x + 16 * a3.x - 3 * a4.x) / 12.;
        dv[1] = sg * (-25 * f0.y + 48 * a1.y - 36 * a2.y + 16 * a3.y - 3 * a4.y) / 12.;
        dv[2] = sg * (-25 * f0.z + 48 * a1.z - 36 * a2.z + 16 * a3.z - 3 * a4.z) / 12.;
        dv[3] = sg * (-25 * f0.w + 48 * a1.w - 36 * a2.w + 16 * a3.w - 3 * a4.w) / 12.;
    }
    /* dv = d (df/dlambda); h f' = 2 dv */
    out[0].x = f0.x; out[0].y = 2 * dv[0]; out[1].x = f0.y; out[1].y = 2 * dv[1];
    out[2].x = f0.z; out[2].y = 2 * dv[2]; out[3].x = f0.w; out[3].y = 2 * dv[3];
    /* fifth slot: z(lambda) itself, {z, h dz/dlambda} -- the kernels' item set-up reads the redshift of the two end points from here
       (one Hermite lookup in shared memory) instead of bisecting the 10^4-knot z(chi) spline through L2 */
    {
        const double z0 = spline_eval(t.bg.z_of_chi, lam);
        double dz;
        if (lam - 2 * d >= 0.0 && lam + 2 * d <= chi_max)
            dz = (spline_eval(t.bg.z_of_chi, lam - 2 * d) - 8 * spline_eval(t.bg.z_of_chi, lam - d) + 8 * spline_eval(t.bg.z_of_chi, lam + d) - spline_eval(t.bg.z_of_chi, lam + 2 * d)) / 12.;
        else {
            const double sg = lam - 2 * d < 0.0 ? 1.0 : -1.0;
            dz = sg * (-25 * z0 + 48 * spline_eval(t.bg.z_of_chi, lam + sg * d) - 36 * spline_eval(t.bg.z_of_chi, lam + sg * 2 * d)
                       + 16 * spline_eval(t.bg.z_of_chi, lam + sg * 3 * d) - 3 * spline_eval(t.bg.z_of_chi, lam + sg * 4 * d)) / 12.;
        }
        out[4].x = z0; out[4].y = 2 * dz;
    }
}

/* ---- flat-sky closed forms for the multipoles ----------------------------------------------------- */
COFFE_HD double factorial_d(const int n)
{
    double r = 1.0;
    for (int k = 2; k <= n; ++k) r *= k;
    return r;
}

COFFE_HD double choose_d(const int n, int m)
{
    if (m > n) return 0.0;
    if (m > n - m) m = n - m;
    double r = 1.0;
    for (int k = 1; k <= m; ++k) r = r * (n - m + k) / k;
    return floor(r + 0.5);
}

/* density-lensing multipole l in the flat-sky limit: finite sum over the half-integer tables
   I^{(l-2k+3)/2}_{(l-2k+1)/2}(r); includes the reference's 2 pi^2 (src/functions.c:4136-4253, src/signal.c:610-631).
   Returns NaN when a needed table was not uploaded. */
COFFE_HD double flatsky_density_lensing_multipole(const Options &o, const Tables &t, const SepConst &sc, const double sep, const int l)
{
    if (l % 2 != 0) return 0.0;
    double sum = 0.0;
    for (int k = 0; k <= l / 2; ++k) {
        const SplineView *s = find_extra(t, l - 2 * k + 3, l - 2 * k + 1, 1);
        if (!s) return nan("");
        const double sign = (k % 2 == 0) ? 1.0 : -1.0;
        sum += sign / ldexp(1.0, k) * choose_d(l, k) * choose_d(2 * l - 2 * k, l) * factorial_d(l / 2 - k) * spline_eval(*s, sep);
    }
    const double pi = 3.14159265358979323846;
    sum *= -3 * o.Om / pi / 8 * (2 * l + 1) * sc.D1m * sc.D1m * (1 + sc.z_mean) * sep
         * ((2 - 5 * sc.sm1) * sc.bm2 + (2 - 5 * sc.sm2) * sc.bm1) * pow(2.0, 1 - l / 2);
    return 2 * pi * pi * sum / sqrt(2 * pi) * o.H0h * o.H0h;
}

/* integrand in x of the flat-sky lensing-lensing multipole l (src/functions.c:4020-4132), without the 2 pi^2 */
COFFE_HD double flatsky_lensing_lensing_integrand(const Options &o, const Tables &t, const SepConst &sc, const SplineView &i1l,
                                                  const double sep, const int l, const double x)
{
    const double pi = 3.14159265358979323846;
    const double lambda = sc.chi_mean * x;
    const double z = spline_eval(t.bg.z_of_chi, lambda);
    const int i = bg_index(t.bg, z);
    const double D1 = cubic_eval(t.bg.D1[i], z - t.bg.z[i]);
    const double fl2 = factorial_d(l / 2);
    double r = (2 * l + 1) * factorial_d(l) / ldexp(1.0, l) / (fl2 * fl2)
             * 9 * o.Om * o.Om * (2 - 5 * sc.sm1) * (2 - 5 * sc.sm2) * sc.chi_mean * sc.chi_mean * sc.chi_mean / 8. / pi
             * spline_eval(i1l, x * sep) * D1 * D1 * (1 + z) * (1 + z) * (x * (1 - x)) * (x * (1 - x));
    const double h2 = o.H0h * o.H0h;
    if (l > 0) r *= x * sep;
    return r * h2 * h2;
}

}  // namespace coffe
