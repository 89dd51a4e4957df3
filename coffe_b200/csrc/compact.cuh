// Shared-memory resident tables of the integrated-pieces kernel (K1 v2).
//
// The line-of-sight integrands evaluate the eight I^n_l(R) of the family, I^4_0(R) and the line-of-sight
// functions at every node.  Read through L1/L2 from the full tables (10^4 knots x 256 B per interval,
// scattered 32-byte records) those lookups bound the kernel (profiles/r1_items_kernel_v1_summary.txt:
// l1tex at 78 % of peak, FP64 pipe at 22 %).  Here everything a node needs lives in ONE blob that each
// CTA copies into shared memory once:
//   * R < r_lo (= 1 Mpc for the reference's tables): the reference's own small-r knots
//     (src/integrals.c:111-126), exact per-interval cubics, found through a bucket table indexed by the
//     exponent and top two mantissa bits of R;
//   * R >= r_lo: the family resampled as cubic Hermite data {f, h f'} on a binary "floating-point" grid:
//     every quarter octave [2^e (1 + q/4), 2^e (1 + (q+1)/4)) is a region split into 2^m equal intervals,
//     m chosen per region by the builder (compact_build.h) so that the resampling error stays below a
//     tolerance within a knot budget.  The interval index, the left knot and 1/h come straight from the
//     bits of R -- no logarithm, no knot array, no fix-up loop.  All functions of a knot are
//     interleaved in one 112-byte record (kKnotBytes below);
//   * I^4_0: resampled onto the same binary grid as the ninth function of a knot (what the integrated x
//     integrated kernel reads); the reference's 297 hand-picked knots (src/integrals.c:76-96) are kept as
//     exact cubics behind a bucket table on exponent + 5 mantissa bits for R < r_lo and for the
//     local x integrated piece.
// The resampling error is part of the tolerance budget; the exact-table kernel stays
// available (coffe_b200_set_kernel_variant) and the tests compare the two.
#pragma once
#include "tables.cuh"

namespace coffe {

static const int kKnotStride = 18;      /* doubles per knot in the builder's working copy: 9 x {f, h f'}: the eight I^n_l of the family, then I^4_0
                                           (zero without the divergent tables) */
/* A knot as it is staged in shared memory: the nine values as doubles, the nine scaled slopes h f' as floats
   (24-bit mantissa, rounded to nearest by the builder; one conversion instruction to unpack), 4 bytes of padding: 112 B instead of 144.  The integrated kernels run at 81 % of the shared-memory /
   L1 data pipe (profiles/r1_di_kernel_cfg5_v40_ncu.txt), so bytes per lookup are what they pay for.  A slope enters the Hermite
   form with weight |h10|, |h11| <= 4/27 and |h f'| is a fraction of the local envelope, so its 6e-8 relative rounding moves
   the interpolant by < 4e-8 of the envelope -- the builder's error check (region_fill) runs on the rounded slopes and
   reports it inside max_rel_error.  A 112-byte stride is 7 bank groups of 16 bytes: neighbouring knots fall on different ones. */
static const int kKnotBytes = 112;
static const int N_ITAB9 = N_ITAB + 1;
static const int kSmallStride = 34;     /* doubles per small-r interval: 8 x double4 + 2 padding */

/* byte offsets are relative to the start of the blob and multiples of 16 */
struct CompactHeader {
    int bytes;                 /* size of the blob */
    /* regions of the binary grid: id = (hi >> 18) - region0, hi = upper 32 bits of R */
    int off_region, n_region, region0;
    int off_knots, n_knots;
    /* small-r exact records */
    int off_shint, n_shint, shint0;
    int off_sx, off_sk, sn;
    /* I^4_0 */
    int off_i40hint, n_i40hint, i40hint0;
    int off_i40x, off_i40k, i40n;
    double r_lo;               /* 2^e_min: first knot of the binary grid */
    double max_rel_error;      /* achieved resampling error of the builder (relative to the local envelope), I^n_l family */
    double max_rel_error_i40;  /* the same for I^4_0 */
    int valid;
};

#if defined(__CUDA_ARCH__)
COFFE_D int hi_word(const double v) { return __double2hiint(v); }
COFFE_D double from_hi_word(const int hi) { return __hiloint2double(hi, 0); }
#else
inline int hi_word(const double v) { long long b; memcpy(&b, &v, 8); return (int)(b >> 32); }
inline double from_hi_word(const int hi) { const long long b = (long long)(unsigned int)hi << 32; double v; memcpy(&v, &b, 8); return v; }
#endif

/* h f' as stored: a float (24-bit mantissa: 6e-8 relative, it enters the Hermite form with weight <= 4/27), one conversion
   instruction to unpack.  (Until the end of round 2: the upper 32 bits of the double, 20 mantissa bits, unpacked by two moves.) */
COFFE_HD int slope_word(const double v)
{
    const float f = (float)v;
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int w; memcpy(&w, &f, 4); return w;
#endif
}
COFFE_HD double slope_unpack(const int w)
{
#if defined(__CUDA_ARCH__)
    return (double)__int_as_float(w);
#else
    float f; memcpy(&f, &w, 4); return (double)f;
#endif
}
COFFE_HD double slope_rounded(const double v) { return slope_unpack(slope_word(v)); }

/* the 16-byte words of a packed knot */
struct KnotWord { int x, y, z, w; };
COFFE_HD double word_pair(const int lo, const int hi)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    const long long b = ((long long)(unsigned int)hi << 32) | (unsigned int)lo; double v; memcpy(&v, &b, 8); return v;
#endif
}

/* Layout of a packed knot (byte offsets): two groups of four functions, {f0 f1 | f2 f3 | d0 d1 d2 d3} at 0 and
   {f4 f5 | f6 f7 | d4 d5 d6 d7} at 48 (values as doubles, slopes as 32-bit words), then {f8 | d8, padding} at 96: a group is
   evaluated as soon as its three 16-byte words of the two knots have arrived, so few registers are live at a time. */
static const int kKnotGroupBytes = 48;
COFFE_HD int knot_value_offset(const int f) { return f < 8 ? (f >> 2) * kKnotGroupBytes + (f & 3) * 8 : 96; }
COFFE_HD int knot_slope_offset(const int f) { return f < 8 ? (f >> 2) * kKnotGroupBytes + 32 + (f & 3) * 4 : 104; }

COFFE_HD void knot_words(const unsigned char *p, KnotWord &w)
{
#if defined(__CUDA_ARCH__)
    const int4 v = *reinterpret_cast<const int4 *>(p);
    w.x = v.x; w.y = v.y; w.z = v.z; w.w = v.w;
#else
    memcpy(&w, p, 16);
#endif
}

/* Hermite interpolant of functions 4g..4g+3 between the knots at k0 and k1 */
template <class HB>
COFFE_HD void knot_group_eval(const unsigned char *k0, const unsigned char *k1, const int g, const HB &hb, double *out)
{
    const double2 *q0 = reinterpret_cast<const double2 *>(k0 + g * kKnotGroupBytes), *q1 = reinterpret_cast<const double2 *>(k1 + g * kKnotGroupBytes);
    const double2 a0 = q0[0], a1 = q0[1], b0 = q1[0], b1 = q1[1];
    KnotWord s0, s1;
    knot_words(k0 + g * kKnotGroupBytes + 32, s0);
    knot_words(k1 + g * kKnotGroupBytes + 32, s1);
    out[0] = hermite_eval(hb, a0.x, slope_unpack(s0.x), b0.x, slope_unpack(s1.x));
    out[1] = hermite_eval(hb, a0.y, slope_unpack(s0.y), b0.y, slope_unpack(s1.y));
    out[2] = hermite_eval(hb, a1.x, slope_unpack(s0.z), b1.x, slope_unpack(s1.z));
    out[3] = hermite_eval(hb, a1.y, slope_unpack(s0.w), b1.y, slope_unpack(s1.w));
}

/* the ninth function (I^4_0) */
template <class HB>
COFFE_HD double knot_tail_eval(const unsigned char *k0, const unsigned char *k1, const HB &hb)
{
    KnotWord t0, t1;
    knot_words(k0 + 96, t0);
    knot_words(k1 + 96, t1);
    return hermite_eval(hb, word_pair(t0.x, t0.y), slope_unpack(t0.z), word_pair(t1.x, t1.y), slope_unpack(t1.z));
}

/* view of a staged blob (shared memory on the device, host memory in the CPU harness) */
struct CompactView {
    const unsigned char *base;
    CompactHeader h;
};

COFFE_HD void compact_ifam_all(const CompactView &c, const double r, double out[N_ITAB])
{
    const int hi = hi_word(r);
    int reg = (hi >> 18) - c.h.region0;
    if (reg < 0) {
        /* small-r exact records */
        int b = (hi >> 18) - c.h.shint0;
        if (b < 0) b = 0;
        if (b > c.h.n_shint - 1) b = c.h.n_shint - 1;
        const int *hint = reinterpret_cast<const int *>(c.base + c.h.off_shint);
        const double *sx = reinterpret_cast<const double *>(c.base + c.h.off_sx);
        int i = hint[b];
        while (i < c.h.sn - 2 && r >= sx[i + 1]) ++i;
        const double dx = r - sx[i];
        const double4 *k = reinterpret_cast<const double4 *>(c.base + c.h.off_sk + (size_t)i * kSmallStride * 8);
#pragma unroll
        for (int t = 0; t < N_ITAB; ++t) out[t] = cubic_eval(k[t], dx);
        return;
    }
    if (reg > c.h.n_region - 1) reg = c.h.n_region - 1;
    const int2 rg = reinterpret_cast<const int2 *>(c.base + c.h.off_region)[reg];     /* {first knot, 18 - m} */
    int idx = rg.x + ((hi & 0x3FFFF) >> rg.y);
    if (idx > c.h.n_knots - 2) idx = c.h.n_knots - 2;
    const double xj = from_hi_word(hi & (int)(~0u << rg.y));
    /* 1/h = 2^(m + 2 - E), m = 18 - rg.y, E = (hi >> 20) - 1023 */
    const double inv_h = from_hi_word((2046 + 20 - rg.y - (hi >> 20)) << 20);
    const double u = (r - xj) * inv_h;
    const double u2 = u * u, um1 = u - 1.0;
    const double h01 = u2 * fma(-2.0, u, 3.0);      /* 3u^2 - 2u^3 */
    const double h10 = u * um1 * um1;
    const double h11 = u2 * um1;
    const unsigned char *k0 = c.base + c.h.off_knots + (size_t)idx * kKnotBytes;
    HermiteBasis hb; hb.h01 = h01; hb.h10 = h10; hb.h11 = h11;
    knot_group_eval(k0, k0 + kKnotBytes, 0, hb, out);
    knot_group_eval(k0, k0 + kKnotBytes, 1, hb, out + 4);
}

COFFE_HD double compact_i40(const CompactView &c, const double r);

/* the eight I^n_l and (when want_i40) I^4_0 at r: one index computation, one Hermite basis, 2 x 7 16-byte loads */
COFFE_HD void compact_ifam9(const CompactView &c, const double r, double out[N_ITAB9], const bool want_i40)
{
    const int hi = hi_word(r);
    int reg = (hi >> 18) - c.h.region0;
    if (COFFE_UNLIKELY(reg < 0)) {
        compact_ifam_all(c, r, out);
        out[N_ITAB] = want_i40 ? compact_i40(c, r) : 0.0;
        return;
    }
    if (reg > c.h.n_region - 1) reg = c.h.n_region - 1;
    const int2 rg = reinterpret_cast<const int2 *>(c.base + c.h.off_region)[reg];
    int idx = rg.x + ((hi & 0x3FFFF) >> rg.y);
    if (idx > c.h.n_knots - 2) idx = c.h.n_knots - 2;
    const double xj = from_hi_word(hi & (int)(~0u << rg.y));
    const double inv_h = from_hi_word((2046 + 20 - rg.y - (hi >> 20)) << 20);
    const HermiteBasis hb = hermite_basis((r - xj) * inv_h);
    const unsigned char *k0 = c.base + c.h.off_knots + (size_t)idx * kKnotBytes;
    knot_group_eval(k0, k0 + kKnotBytes, 0, hb, out);
    knot_group_eval(k0, k0 + kKnotBytes, 1, hb, out + 4);
    out[N_ITAB] = want_i40 ? knot_tail_eval(k0, k0 + kKnotBytes, hb) : 0.0;
}

COFFE_HD double compact_i40(const CompactView &c, const double r)
{
    const int hi = hi_word(r);
    int b = (hi >> 15) - c.h.i40hint0;
    if (b < 0) b = 0;
    if (b > c.h.n_i40hint - 1) b = c.h.n_i40hint - 1;
    const int *hint = reinterpret_cast<const int *>(c.base + c.h.off_i40hint);
    const double *x = reinterpret_cast<const double *>(c.base + c.h.off_i40x);
    int i = hint[b];
    while (i < c.h.i40n - 2 && r >= x[i + 1]) ++i;
    const double4 k = reinterpret_cast<const double4 *>(c.base + c.h.off_i40k)[i];
    return cubic_eval(k, r - x[i]);
}

/* ---- table accessors: the integrands are written once against this interface ------------------- */
struct ExactAccess {
    const Tables *t;
    COFFE_HD void ifam_all(const double r, double out[N_ITAB]) const { ifam_eval_all(t->ifam, r, out); }
    COFFE_HD void ifam9(const double r, double out[N_ITAB9], const bool want_i40) const
    {
        ifam_eval_all(t->ifam, r, out);
        out[N_ITAB] = want_i40 ? hspline_eval(t->i40, r) : 0.0;
    }
    COFFE_HD double i40(const double r) const { return hspline_eval(t->i40, r); }
    COFFE_HD double ren2d(const double a, const double b) const { return renorm2d_eval(t->ren2d, a, b); }
    COFFE_HD double ren0(const double l) const { return spline_eval(t->ren0, l); }
    COFFE_HD void ifam_mid(const double r, const double z, double out[N_ITAB]) const { mid2d_eval_all(t->mid, t->ifam, z, r, out); }
};

struct CompactAccess {
    const Tables *t;           /* renormalisation grid and zero-separation spline stay in global memory (L2) */
    CompactView c;
    COFFE_HD void ifam_all(const double r, double out[N_ITAB]) const { compact_ifam_all(c, r, out); }
    COFFE_HD void ifam9(const double r, double out[N_ITAB9], const bool want_i40) const { compact_ifam9(c, r, out, want_i40); }
    COFFE_HD double i40(const double r) const { return compact_i40(c, r); }
    COFFE_HD double ren2d(const double a, const double b) const { return renorm2d_eval(t->ren2d, a, b); }
    COFFE_HD double ren0(const double l) const { return spline_eval(t->ren0, l); }
    /* the 2-D tables are never resampled into shared memory: the host routes midpoint runs to the exact-table kernel */
    COFFE_HD void ifam_mid(const double r, const double z, double out[N_ITAB]) const { mid2d_eval_all(t->mid, t->ifam, z, r, out); }
};

}  // namespace coffe
