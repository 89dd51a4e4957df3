// Device-side table views and evaluators.
//
// Everything the reference looks up through GSL (`coffe_interp_spline`, src/common.c:704-719) is
// flattened on upload into POD arrays of per-interval cubic coefficients {y, b, c, d} (one double4 per
// interval) so that a lookup is: find interval, one 32-byte load, three FMAs.  The interval search is
// specialised per table family:
//   * background functions of z   -- uniform z grid (src/background.c:446), O(1) index + fix-up;
//   * I^n_l(r) family             -- 43 irregular small-r knots + log-spaced FFTLog knots
//                                    (src/integrals.c:111-126, 524-562), O(1) index from log(r);
//                                    the eight (n,l) tables share knots and are interleaved per knot;
//   * everything else (z(chi), biases, I^4_0, zero-separation renormalisation) -- binary search;
//   * renormalisation(chi1,chi2)  -- uniform 201x201 grid (src/integrals.c:851-907), bicubic Hermite
//                                    patch with node derivatives from natural cubic splines (GSL bicubic).
#pragma once
#include "hd.h"

namespace coffe {

COFFE_HD double cubic_eval(const double4 k, const double dx)
{
    return fma(dx, fma(dx, fma(dx, k.w, k.z), k.y), k.x);
}

/* cubic Hermite basis on the unit interval, value form f0 + h01 (f1 - f0) + h10 m0 + h11 m1 */
struct HermiteBasis { double h01, h10, h11; };

COFFE_HD HermiteBasis hermite_basis(const double u)
{
    HermiteBasis b;
    const double u2 = u * u, um1 = u - 1.0;
    b.h01 = u2 * fma(-2.0, u, 3.0);      /* 3u^2 - 2u^3 */
    b.h10 = u * um1 * um1;
    b.h11 = u2 * um1;
    return b;
}

COFFE_HD double hermite_eval(const HermiteBasis &b, const double f0, const double m0, const double f1, const double m1)
{
    return fma(b.h01, f1 - f0, fma(b.h10, m0, fma(b.h11, m1, f0)));
}

/* index i with x[i] <= v < x[i+1], clamped to [0, n-2] (gsl_interp_bsearch semantics) */
COFFE_HD int knot_search(const double *x, const int n, const double v)
{
    int lo = 0, hi = n - 1;
    while (hi > lo + 1) {
        const int mid = (hi + lo) >> 1;
        if (x[mid] > v) hi = mid; else lo = mid;
    }
    return lo;
}

struct SplineView {
    const double *x;
    const double4 *k;
    int n;
};

COFFE_HD double spline_eval(const SplineView s, const double v)
{
    const int i = knot_search(s.x, s.n, v);
    return cubic_eval(s.k[i], v - s.x[i]);
}

/* spline on irregular positive knots with a bucket table in log(x) that gives the starting interval, so a
   lookup is a couple of forward steps instead of a binary search (I^4_0 has 297 hand-picked knots,
   src/integrals.c:76-96) */
struct HintedSplineView {
    const double *x;
    const double4 *k;
    const int *hint;        /* hint[b] = largest i with x[i] <= exp(log_lo + b/inv_dlog), clamped to [0, n-2] */
    int n, n_hint;
    double log_lo, inv_dlog;
};

COFFE_HD double hspline_eval(const HintedSplineView s, const double v)
{
    int i = 0;
    if (v > s.x[1]) {
#if defined(__CUDA_ARCH__)
        int b = (int)(((double)__logf((float)v) - s.log_lo) * s.inv_dlog) - 1;
#else
        int b = (int)((log(v) - s.log_lo) * s.inv_dlog) - 1;
#endif
        if (b < 0) b = 0;
        if (b > s.n_hint - 1) b = s.n_hint - 1;
        i = s.hint[b];
        while (i < s.n - 2 && v >= s.x[i + 1]) ++i;
    }
    return cubic_eval(s.k[i], v - s.x[i]);
}

/* value and first derivative of the interpolant (the resampler of compact_build.h needs both) */
COFFE_HD double hspline_eval_deriv(const HintedSplineView s, const double v, double *der)
{
    int i = 0;
    if (v > s.x[1]) {
        i = knot_search(s.x, s.n, v);
        if (i > s.n - 2) i = s.n - 2;
    }
    const double4 k = s.k[i];
    const double dx = v - s.x[i];
    *der = fma(dx, fma(dx, 3.0 * k.w, 2.0 * k.z), k.y);
    return cubic_eval(k, dx);
}

/* spline on (nominally) uniform knots: O(1) guess, then fix-up against the real knots so that the
   interval is exactly the one a binary search would return */
struct UniformSplineView {
    const double *x;
    const double4 *k;
    int n;
    double x0, inv_dx;
};

COFFE_HD int uniform_index(const double *x, const int n, const double x0, const double inv_dx, const double v)
{
    int i = (int)((v - x0) * inv_dx);
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    if (v < x[i] && i > 0) --i;
    else if (i < n - 2 && v >= x[i + 1]) ++i;
    return i;
}

COFFE_HD double uspline_eval(const UniformSplineView s, const double v)
{
    const int i = uniform_index(s.x, s.n, s.x0, s.inv_dx, v);
    return cubic_eval(s.k[i], v - s.x[i]);
}

/* background: all functions of z share the z knots */
struct BackgroundView {
    const double *z;          /* n knots, uniform on [0, 15] */
    int n;
    double inv_dz;
    const double4 *a, *Hc, *D1, *f, *G1, *G2, *chi;   /* per-interval cubics, functions of z */
    SplineView z_of_chi;      /* knots = chi(z_i) */
};

struct BackgroundAtZ { double a, Hc, D1, f, G1, G2; };

COFFE_HD int bg_index(const BackgroundView &bg, const double z)
{
    return uniform_index(bg.z, bg.n, 0.0, bg.inv_dz, z);
}

COFFE_HD BackgroundAtZ bg_eval_all(const BackgroundView &bg, const double z)
{
    const int i = bg_index(bg, z);
    const double dz = z - bg.z[i];
    BackgroundAtZ o;
    o.a = cubic_eval(bg.a[i], dz);
    o.Hc = cubic_eval(bg.Hc[i], dz);
    o.D1 = cubic_eval(bg.D1[i], dz);
    o.f = cubic_eval(bg.f[i], dz);
    o.G1 = cubic_eval(bg.G1[i], dz);
    o.G2 = cubic_eval(bg.G2[i], dz);
    return o;
}

COFFE_HD double bg_chi(const BackgroundView &bg, const double z)
{
    const int i = bg_index(bg, z);
    return cubic_eval(bg.chi[i], z - bg.z[i]);
}

/* ---- I^n_l(r) family on shared knots ------------------------------------------------------- */
enum { I00 = 0, I02, I04, I11, I13, I20, I22, I31, N_ITAB };

struct IntegralFamilyView {
    const double *x;        /* n knots: 0, small-r list, then r0*exp(i*dln) */
    const double4 *k;       /* n-1 intervals x N_ITAB cubics, interleaved: k[i*N_ITAB + t] */
    int n;
    int n_small;            /* index of the first log-spaced knot (x[n_small] = r0) */
    double r0, inv_dln;
};

COFFE_HD int ifam_guess(const IntegralFamilyView &t, const double r)
{
    if (r < t.r0) return knot_search(t.x, t.n_small + 1, r);
#if defined(__CUDA_ARCH__)
    int i = t.n_small + (int)((double)__logf((float)(r / t.r0)) * t.inv_dln);
#else
    int i = t.n_small + (int)(log(r / t.r0) * t.inv_dln);
#endif
    if (i > t.n - 2) i = t.n - 2;
    if (i < t.n_small) i = t.n_small;
    return i;
}

/* exact interval (the one a binary search returns): the float log is good to ~1e-6 absolute, far below
   the knot spacing, so the guess is at most one step off and only within ~1e-6 of a knot */
COFFE_HD int ifam_index(const IntegralFamilyView &t, const double r)
{
    int i = ifam_guess(t, r);
    while (i > 0 && r < t.x[i]) --i;
    while (i < t.n - 2 && r >= t.x[i + 1]) ++i;
    return i;
}

COFFE_HD void ifam_eval_all(const IntegralFamilyView &t, const double r, double out[N_ITAB])
{
    /* speculative single round trip: knots and the interleaved record of the guessed interval are loaded
       together; the (rare) wrong guess reloads */
    int i = ifam_guess(t, r);
    double xi = t.x[i];
    const double xn = t.x[i + 1];
    const double4 *k = t.k + (size_t)i * N_ITAB;
    double4 c[N_ITAB];
#pragma unroll
    for (int j = 0; j < N_ITAB; ++j) c[j] = k[j];
    if (!(r >= xi && (r < xn || i == t.n - 2))) {
        i = ifam_index(t, r);
        xi = t.x[i];
        k = t.k + (size_t)i * N_ITAB;
#pragma unroll
        for (int j = 0; j < N_ITAB; ++j) c[j] = k[j];
    }
    const double dx = r - xi;
#pragma unroll
    for (int j = 0; j < N_ITAB; ++j) out[j] = cubic_eval(c[j], dx);
}

/* ---- renormalisation(chi1, chi2), bicubic ---------------------------------------------------- */
struct Renorm2DView {
    const double *x;        /* n knots (same on both axes) */
    const double4 *node;    /* n*n nodes {z, zx, zy, zxy}; index j*n + i with i <-> first argument */
    int n;
    double x0, inv_dx, dx;
};

/* the grid is uniform by construction (chi_j = j/nbins chi_max, src/integrals.c:876-878): the cell index and
   the in-cell coordinates t, u come straight from the argument, no knot loads */
COFFE_HD double renorm2d_eval(const Renorm2DView &t, const double x, const double y)
{
    const double sx = (x - t.x0) * t.inv_dx, sy = (y - t.x0) * t.inv_dx;
    int xi = (int)sx, yi = (int)sy;
    if (xi < 0) xi = 0;
    if (yi < 0) yi = 0;
    if (xi > t.n - 2) xi = t.n - 2;
    if (yi > t.n - 2) yi = t.n - 2;
    const double tt = sx - xi, u = sy - yi;
    const double dx = t.dx, dy = t.dx;
    const double4 n00 = t.node[yi * t.n + xi], n10 = t.node[yi * t.n + xi + 1];
    const double4 n01 = t.node[(yi + 1) * t.n + xi], n11 = t.node[(yi + 1) * t.n + xi + 1];
    const double t2 = tt * tt, t3 = t2 * tt, u2 = u * u, u3 = u2 * u;
    const double h00t = 2 * t3 - 3 * t2 + 1, h10t = (t3 - 2 * t2 + tt) * dx, h01t = -2 * t3 + 3 * t2, h11t = (t3 - t2) * dx;
    const double h00u = 2 * u3 - 3 * u2 + 1, h10u = (u3 - 2 * u2 + u) * dy, h01u = -2 * u3 + 3 * u2, h11u = (u3 - u2) * dy;
    return h00t * (h00u * n00.x + h01u * n01.x + h10u * n00.z + h11u * n01.z)
         + h01t * (h00u * n10.x + h01u * n11.x + h10u * n10.z + h11u * n11.z)
         + h10t * (h00u * n00.y + h01u * n01.y + h10u * n00.w + h11u * n01.w)
         + h11t * (h00u * n10.y + h01u * n11.y + h10u * n10.w + h11u * n11.w);
}

/* ---- line-of-sight functions resampled on a uniform lambda grid --------------------------------
   W = D1/a, V = Hc (f-1), U = Hc (Hc-1), Wa = D1 a evaluated through the reference's own two-level
   lookup z = z(lambda) -> spline(z) at n uniformly spaced lambda and stored as cubic Hermite data
   {f, h f'} per knot: v[5j + 0..3] = {W, hW'}, {V, hV'}, {U, hU'}, {Wa, hWa'}, v[5j + 4] = padding (an 80-byte
   stride puts neighbouring knots on different 16-byte bank groups of shared memory).  These functions are smooth
   on Gpc scales, so with spacing ~18 Mpc the resampling error is ~1e-11 relative (tolerance
   budget; tests/test_physics_cpu.py::test_line_of_sight_resampling_error). */
static const int kLosStride = 5;     /* double2 per knot */
struct LosView {
    const double2 *v;   /* kLosStride double2 per knot */
    int n;
    double h, inv_h;
};

COFFE_HD double4 los_eval(const double2 *v, const int n, const double inv_h, const double lambda)
{
    const double s = lambda * inv_h;
    int j = (int)s;
    if (j < 0) j = 0;
    if (j > n - 2) j = n - 2;
    const HermiteBasis b = hermite_basis(s - j);
    const double2 *k0 = v + (size_t)kLosStride * j, *k1 = k0 + kLosStride;
    double4 o;
    { const double2 p = k0[0], q = k1[0]; o.x = hermite_eval(b, p.x, p.y, q.x, q.y); }
    { const double2 p = k0[1], q = k1[1]; o.y = hermite_eval(b, p.x, p.y, q.x, q.y); }
    { const double2 p = k0[2], q = k1[2]; o.z = hermite_eval(b, p.x, p.y, q.x, q.y); }
    { const double2 p = k0[3], q = k1[3]; o.w = hermite_eval(b, p.x, p.y, q.x, q.y); }
    return o;
}

/* The same table as the integration kernels stage it in shared memory: one 80-byte record per INTERVAL holding the cubic of each of the
   four functions in t = lambda / h - j, constant term as a double, the three higher coefficients as floats (they describe the change
   across one ~18 Mpc interval, a per cent of the value: 6e-8 of that).  A lookup reads one record instead of two knots, and the fields
   are laid out so that the integrated x integrated piece (W, V, U) reads 16-byte words 0..3 and the local x integrated piece
   (W, V, Wa) words 1..4: four loads per lookup instead of six.  Built from the Hermite knots by los_rec_pack. */
struct alignas(16) LosRec {
    double c0z; float c1z, c2z;             /* word 0 */
    float c3z, c3x; double c0x;             /* word 1 */
    double c0y; float c1x, c2x;             /* word 2 */
    float c1y, c2y, c3y, c3w;               /* word 3 */
    double c0w; float c1w, c2w;             /* word 4 */
};
static_assert(sizeof(LosRec) == 16 * kLosStride, "a record takes the room of a knot");

COFFE_HD void los_rec_pack(const double2 *k0, const double2 *k1, LosRec &r)
{
    double c0[4], c1[4], c2[4], c3[4];
    for (int f = 0; f < 4; ++f) {
        const double f0 = k0[f].x, m0 = k0[f].y, f1 = k1[f].x, m1 = k1[f].y, df = f1 - f0;
        c0[f] = f0; c1[f] = m0; c2[f] = 3.0 * df - 2.0 * m0 - m1; c3[f] = m0 + m1 - 2.0 * df;
    }
    r.c0x = c0[0]; r.c1x = (float)c1[0]; r.c2x = (float)c2[0]; r.c3x = (float)c3[0];
    r.c0y = c0[1]; r.c1y = (float)c1[1]; r.c2y = (float)c2[1]; r.c3y = (float)c3[1];
    r.c0z = c0[2]; r.c1z = (float)c1[2]; r.c2z = (float)c2[2]; r.c3z = (float)c3[2];
    r.c0w = c0[3]; r.c1w = (float)c1[3]; r.c2w = (float)c2[3]; r.c3w = (float)c3[3];
}

COFFE_HD double los_cubic(const double c0, const float c1, const float c2, const float c3, const double t)
{
    return fma(fma(fma((double)c3, t, (double)c2), t, (double)c1), t, c0);
}

/* n = number of knots of the table the records were built from (n - 1 records) */
COFFE_HD double4 los_eval(const LosRec *v, const int n, const double inv_h, const double lambda)
{
    const double s = lambda * inv_h;
    int j = (int)s;
    if (j < 0) j = 0;
    if (j > n - 2) j = n - 2;
    const double t = s - j;
    double4 o;
#if defined(__CUDA_ARCH__)
    const int4 *p = reinterpret_cast<const int4 *>(v + j);
    const int4 w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3], w4 = p[4];        /* the unused ones are dropped by the compiler */
    o.x = los_cubic(__hiloint2double(w1.w, w1.z), __int_as_float(w2.z), __int_as_float(w2.w), __int_as_float(w1.y), t);
    o.y = los_cubic(__hiloint2double(w2.y, w2.x), __int_as_float(w3.x), __int_as_float(w3.y), __int_as_float(w3.z), t);
    o.z = los_cubic(__hiloint2double(w0.y, w0.x), __int_as_float(w0.z), __int_as_float(w0.w), __int_as_float(w1.x), t);
    o.w = los_cubic(__hiloint2double(w4.y, w4.x), __int_as_float(w4.z), __int_as_float(w4.w), __int_as_float(w3.w), t);
#else
    const LosRec &r = v[j];
    o.x = los_cubic(r.c0x, r.c1x, r.c2x, r.c3x, t);
    o.y = los_cubic(r.c0y, r.c1y, r.c2y, r.c3y, t);
    o.z = los_cubic(r.c0z, r.c1z, r.c2z, r.c3z, t);
    o.w = los_cubic(r.c0w, r.c1w, r.c2w, r.c3w, t);
#endif
    return o;
}

/* ---- I^n_l(z, r) tables of the midpoint approximation (coffe_integral_t.result2d, src/integrals.c:663-767) ------------------
   GSL's bicubic interpolant on the (z, r) grid: per node {f, f_z, f_r, f_zr} (table_build.h: bicubic_nodes restates how GSL
   estimates the derivatives), evaluated with the Hermite patch of the cell; z knots uniform, r knots = the family's own.
   node[t][j*nz + i] <-> table t at (z_i, r_j).  Read through L1/L2: this is the path of the non-linear P(k), not the benchmark's. */
struct Mid2DView {
    const double *z;
    const double4 *node[N_ITAB];
    int nz, nr;
    double z0, inv_dz;
};

COFFE_HD void mid2d_eval_all(const Mid2DView &m, const IntegralFamilyView &fam, double zq, double r, double out[N_ITAB])
{
    int iz = (int)((zq - m.z0) * m.inv_dz);
    if (iz < 0) iz = 0;
    if (iz > m.nz - 2) iz = m.nz - 2;
    while (iz > 0 && zq < m.z[iz]) --iz;
    while (iz < m.nz - 2 && zq >= m.z[iz + 1]) ++iz;
    const int ir = ifam_index(fam, r);
    const double dz = m.z[iz + 1] - m.z[iz], dr = fam.x[ir + 1] - fam.x[ir];
    const double tt = (zq - m.z[iz]) / dz, u = (r - fam.x[ir]) / dr;
    const double t2 = tt * tt, t3 = t2 * tt, u2 = u * u, u3 = u2 * u;
    const double h00t = 2 * t3 - 3 * t2 + 1, h10t = (t3 - 2 * t2 + tt) * dz, h01t = -2 * t3 + 3 * t2, h11t = (t3 - t2) * dz;
    const double h00u = 2 * u3 - 3 * u2 + 1, h10u = (u3 - 2 * u2 + u) * dr, h01u = -2 * u3 + 3 * u2, h11u = (u3 - u2) * dr;
    const size_t a = (size_t)ir * m.nz + iz, b = a + m.nz;
#pragma unroll
    for (int k = 0; k < N_ITAB; ++k) {
        const double4 n00 = m.node[k][a], n10 = m.node[k][a + 1], n01 = m.node[k][b], n11 = m.node[k][b + 1];
        out[k] = h00t * (h00u * n00.x + h01u * n01.x + h10u * n00.z + h11u * n01.z)
               + h01t * (h00u * n10.x + h01u * n11.x + h10u * n10.z + h11u * n11.z)
               + h10t * (h00u * n00.y + h01u * n01.y + h10u * n00.w + h11u * n01.w)
               + h11t * (h00u * n10.y + h01u * n11.y + h10u * n10.w + h11u * n11.w);
    }
}

/* z(lambda) from the fifth slot of the line-of-sight table */
COFFE_HD double los_z(const double2 *v, const int n, const double inv_h, const double lambda)
{
    const double s = lambda * inv_h;
    int j = (int)s;
    if (j < 0) j = 0;
    if (j > n - 2) j = n - 2;
    const HermiteBasis b = hermite_basis(s - j);
    const double2 p = v[(size_t)kLosStride * j + 4], q = v[(size_t)kLosStride * (j + 1) + 4];
    return hermite_eval(b, p.x, p.y, q.x, q.y);
}

/* an I^n_l table outside the interleaved family: I^1_l (flat-sky lensing-lensing multipoles, src/integrals.c:
   1085-1247) and the half-integer tables (n/2, l/2) of the flat-sky density-lensing multipoles (:1249-1413) */
struct ExtraSpline { int n, l, half; SplineView s; };
static const int kMaxExtra = 24;

/* everything a kernel needs, passed by value */
struct Tables {
    BackgroundView bg;
    SplineView bias[8];        /* galaxy_bias1..4, magnification_bias1,2, evolution_bias1,2 */
    IntegralFamilyView ifam;
    HintedSplineView i40;      /* I^4_0(r) (divergent, renormalised) */
    SplineView ren0;           /* renormalization_zero_separation */
    Renorm2DView ren2d;
    LosView los;
    ExtraSpline extra[kMaxExtra];   /* every other uploaded I^n_l (flat-sky closed forms), looked up by key */
    int n_extra;
    int has_divergent;
    Mid2DView mid;             /* nz = 0 unless the 2-D tables were uploaded */
};

COFFE_HD const SplineView *find_extra(const Tables &t, const int n, const int l, const int half)
{
    for (int i = 0; i < t.n_extra; ++i)
        if (t.extra[i].n == n && t.extra[i].l == l && t.extra[i].half == half) return &t.extra[i].s;
    return nullptr;
}

enum { B_GAL1 = 0, B_GAL2, B_GAL3, B_GAL4, B_MAG1, B_MAG2, B_EVO1, B_EVO2 };

}  // namespace coffe
