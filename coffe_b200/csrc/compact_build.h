// Host-side construction of the shared-memory blob described in compact.cuh (one-off per table upload).
//
// The binary grid is adaptive: every quarter-octave region starts with two intervals and the region with
// the largest resampling error is refined (intervals doubled) until every region is below `tol` or the knot
// budget is spent.  The error of a region is measured at three interior points of every interval against
// the exact table (the reference's Akima interpolant), relative to the local envelope of the function
// (largest |f| within +-19 % in r), floored at 1e-2 max(r |f|) / r: a line-of-sight integral weighs f(R) with
// the measure dR, so where r |f| is below 1 % of its maximum (the tiny wiggly tails of I^0_0, I^0_2, I^0_4
// beyond a Gpc) the tolerated absolute error stops shrinking with |f| and the budget goes where it matters.  Knot derivatives come from the
// 6-point Lagrange polynomial through the neighbouring reference knots: the Akima slopes themselves are only
// second-order accurate, which would limit the resampled table to ~1e-5.
#pragma once
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "assemble.h"
#include "compact.cuh"

namespace coffe {

/* default knot budget and tolerance of the binary grid */
static const int kCompactKnotBudget = 800;
static const double kCompactTolerance = 3e-9;
/* I^4_0 only enters the renormalised potential terms (relative weight ~1e-4 of the signal): looser target */
static const double kCompactToleranceI40 = 1e-6;

struct CompactBuild {
    std::vector<double> blob;       /* 16-byte aligned payload (vector<double> storage is 16-byte aligned by the allocator) */
    CompactHeader h;
    std::vector<int> region_m;      /* subdivisions 2^m per region (diagnostics) */
};

namespace build {

/* where r |f(r)| falls below this fraction of its maximum the tolerated error stops shrinking with |f| */
static const double kImportanceFloor = 1e-2;

struct FamilyRef {
    const Tables *t;
    int m, n_small;
    std::vector<double> y[N_ITAB];      /* reference values at the knots */
    std::vector<double> env[N_ITAB];    /* envelope used to normalise errors */
    bool div = false;                   /* I^4_0 present: ninth function of every knot */
    double G40 = 0.0;                   /* max r |I^4_0(r)| over its knots */
};

/* I^4_0 is the reference's Akima interpolant through 297 knots; it is resampled with its own exact derivative
   (cubic Hermite reproduces a cubic, so only binary-grid intervals that straddle one of the 297 knots carry an
   error).  Envelope as for the family: largest |f| within +-19 % in r, floored where r |f| is unimportant. */
inline double i40_envelope(const FamilyRef &F, const double r)
{
    double a = 0.0;
    for (int q = -4; q <= 4; ++q) a = std::max(a, std::fabs(hspline_eval(F.t->i40, r * (1.0 + 0.0475 * q))));
    return std::max(a, kImportanceFloor * F.G40 / std::max(r, 1e-3));
}

inline void family_ref_init(const Tables &t, FamilyRef &F)
{
    F.t = &t; F.m = t.ifam.n; F.n_small = t.ifam.n_small;
    const int m = F.m;
    for (int f = 0; f < N_ITAB; ++f) { F.y[f].resize(m); F.env[f].resize(m); }
    for (int i = 0; i < m; ++i) {
        double v[N_ITAB];
        ifam_eval_all(t.ifam, t.ifam.x[i], v);
        for (int f = 0; f < N_ITAB; ++f) F.y[f][i] = v[f];
    }
    F.div = t.has_divergent != 0;
    if (F.div)
        for (int i = 1; i < t.i40.n; ++i) F.G40 = std::max(F.G40, t.i40.x[i] * std::fabs(hspline_eval(t.i40, t.i40.x[i])));
    /* windows in knots of the log-spaced part: +-19 % and a factor 4 below */
    const double dln = t.ifam.inv_dln > 0 ? 1.0 / t.ifam.inv_dln : 0.01;
    const int w1 = std::max(2, (int)(std::log(1.19) / dln)), w4 = std::max(4, (int)(std::log(4.0) / dln));
    (void)w4;
    for (int f = 0; f < N_ITAB; ++f) {
        /* G = max r |f(r)|: the weight of the function in a line-of-sight integral (measure dR) */
        double G = 0.0;
        for (int i = F.n_small; i < m; ++i) G = std::max(G, t.ifam.x[i] * std::fabs(F.y[f][i]));
        for (int i = 0; i < m; ++i) {
            double a = 0;
            for (int j = std::max(0, i - w1); j <= std::min(m - 1, i + w1); ++j) a = std::max(a, std::fabs(F.y[f][j]));
            const double r = std::max(t.ifam.x[i], 1e-3);
            F.env[f][i] = std::max(a, kImportanceFloor * G / r);
        }
    }
}

inline double clamp_r(const FamilyRef &F, const double r) { return std::min(r, F.t->ifam.x[F.m - 1]); }

/* values and derivatives of the eight functions at r */
inline void family_value_deriv(const FamilyRef &F, double r, double val[N_ITAB], double der[N_ITAB])
{
    const double *x = F.t->ifam.x;
    const bool beyond = r > x[F.m - 1];
    r = clamp_r(F, r);
    ifam_eval_all(F.t->ifam, r, val);
    if (beyond) { for (int f = 0; f < N_ITAB; ++f) der[f] = 0.0; return; }
    const int np = 6;
    int i = ifam_index(F.t->ifam, r);
    int lo = i - np / 2 + 1;
    lo = std::max(lo, F.n_small);
    lo = std::min(lo, F.m - np);
    double w[np];
    for (int a = 0; a < np; ++a) {
        double s = 0.0;
        for (int b = 0; b < np; ++b) {
            if (b == a) continue;
            double p = 1.0 / (x[lo + a] - x[lo + b]);
            for (int c = 0; c < np; ++c)
                if (c != a && c != b) p *= (r - x[lo + c]) / (x[lo + a] - x[lo + c]);
            s += p;
        }
        w[a] = s;
    }
    for (int f = 0; f < N_ITAB; ++f) {
        double d = 0.0;
        for (int a = 0; a < np; ++a) d += w[a] * F.y[f][lo + a];
        der[f] = d;
    }
}

struct Region {
    double lo, hi;      /* [lo, hi) */
    int m;
    double err;         /* worst error of the region in units of its tolerance (family: tol, I^4_0: kCompactToleranceI40) */
    double err_fam, err_i40;
    std::vector<double> knots;      /* (2^m + 1) x kKnotStride */
};

inline void region_fill(const FamilyRef &F, Region &R, const double tol)
{
    const int n = 1 << R.m;
    const double h = (R.hi - R.lo) / n;
    R.knots.assign((size_t)(n + 1) * kKnotStride, 0.0);
    for (int j = 0; j <= n; ++j) {
        double v[N_ITAB], d[N_ITAB];
        family_value_deriv(F, R.lo + j * h, v, d);
        /* slopes as they are stored (compact.cuh: kKnotBytes): the error check below sees the rounded values */
        for (int f = 0; f < N_ITAB; ++f) { R.knots[(size_t)j * kKnotStride + 2 * f] = v[f]; R.knots[(size_t)j * kKnotStride + 2 * f + 1] = slope_rounded(h * d[f]); }
        if (F.div) {
            double der = 0.0;
            R.knots[(size_t)j * kKnotStride + 2 * N_ITAB] = hspline_eval_deriv(F.t->i40, R.lo + j * h, &der);
            R.knots[(size_t)j * kKnotStride + 2 * N_ITAB + 1] = slope_rounded(h * der);
        }
    }
    /* error at three interior points per interval */
    double worst = 0.0, worst40 = 0.0;
    const double *x = F.t->ifam.x;
    for (int j = 0; j < n; ++j)
        for (int q = 1; q <= 3; ++q) {
            const double u = 0.25 * q, r = R.lo + (j + u) * h;
            if (r >= x[F.m - 1]) continue;
            double ex[N_ITAB];
            ifam_eval_all(F.t->ifam, r, ex);
            const int i = ifam_index(F.t->ifam, r);
            const double u2 = u * u, um1 = u - 1.0, h01 = u2 * (3.0 - 2.0 * u), h10 = u * um1 * um1, h11 = u2 * um1;
            const double *k0 = &R.knots[(size_t)j * kKnotStride], *k1 = k0 + kKnotStride;
            for (int f = 0; f < N_ITAB; ++f) {
                const double got = k0[2 * f] + h01 * (k1[2 * f] - k0[2 * f]) + h10 * k0[2 * f + 1] + h11 * k1[2 * f + 1];
                const double e = std::fabs(got - ex[f]) / std::max(F.env[f][i], 1e-300);
                worst = std::max(worst, e);
            }
            if (F.div) {
                const double got = k0[2 * N_ITAB] + h01 * (k1[2 * N_ITAB] - k0[2 * N_ITAB]) + h10 * k0[2 * N_ITAB + 1] + h11 * k1[2 * N_ITAB + 1];
                worst40 = std::max(worst40, std::fabs(got - hspline_eval(F.t->i40, r)) / i40_envelope(F, r));
            }
        }
    R.err_fam = worst; R.err_i40 = worst40;
    R.err = std::max(worst / tol, worst40 / kCompactToleranceI40);
}

}  // namespace build

/* builds the blob from an assembled arena; `knot_budget` bounds the knots of the binary grid */
inline bool build_compact(const Tables &t, const Layout &L, const int knot_budget, const double tol, CompactBuild &out, std::string &err)
{
    using namespace build;
    CompactHeader &H = out.h;
    std::memset(&H, 0, sizeof(H));
    if (!L.has_ifam || L.ifam_inv_dln <= 0.0 || L.ifam_n_small < 1) { err = "compact tables need the log-spaced I^n_l family"; return false; }
    FamilyRef F;
    family_ref_init(t, F);
    const double *x = t.ifam.x;
    const double r0 = x[L.ifam_n_small], r_max = x[L.ifam_n - 1];
    int e_min = (int)std::ceil(std::log2(r0) - 1e-12);
    const double r_lo = std::ldexp(1.0, e_min);
    if (!(r_max > 4.0 * r_lo)) { err = "I^n_l tables too short for the compact grid"; return false; }
    /* ---- small-r exact records: every reference interval below r_lo ---- */
    int sn = L.ifam_n_small + 1;
    while (sn < L.ifam_n && x[sn - 1] < r_lo) ++sn;         /* knots x[0..sn-1], x[sn-1] >= r_lo */
    /* ---- regions ---- */
    std::vector<Region> regs;
    for (int e = e_min;; ++e) {
        bool done = false;
        for (int q = 0; q < 4; ++q) {
            Region R;
            R.lo = std::ldexp(1.0 + 0.25 * q, e); R.hi = std::ldexp(1.0 + 0.25 * (q + 1), e);
            R.m = 1;
            region_fill(F, R, tol);
            regs.push_back(std::move(R));
            if (regs.back().hi >= r_max) { done = true; break; }
        }
        if (done) break;
        if (regs.size() > 400) { err = "compact grid: too many regions"; return false; }
    }
    auto total_knots = [&]() { size_t n = 0; for (auto &R : regs) n += (size_t)(1 << R.m) + 1; return n; };
    const int m_cap = 7;
    for (;;) {
        /* the worst region whose refinement still fits the budget */
        const size_t left = (size_t)knot_budget - std::min((size_t)knot_budget, total_knots());
        int worst = -1;
        for (int i = 0; i < (int)regs.size(); ++i)
            if (regs[i].m < m_cap && regs[i].hi <= r_max && regs[i].err > 1.0 && (size_t)(1 << regs[i].m) <= left
                && (worst < 0 || regs[i].err > regs[worst].err)) worst = i;
        if (worst < 0) break;
        regs[worst].m += 1;
        region_fill(F, regs[worst], tol);
    }
    double max_err = 0.0, max_err40 = 0.0;
    for (auto &R : regs)
        if (R.hi <= r_max) { max_err = std::max(max_err, R.err_fam); max_err40 = std::max(max_err40, R.err_i40); }      /* the last region straddles the end of the table: never reached by a valid point */
    /* ---- lay out the blob ---- */
    std::vector<double> &B = out.blob;
    B.clear();
    auto alloc_bytes = [&](size_t bytes) { const size_t off = B.size() * 8; B.resize(B.size() + ((bytes + 15) / 16) * 2, 0.0); return (int)off; };
    const int n_region = (int)regs.size();
    H.off_region = alloc_bytes((size_t)n_region * 8);
    H.n_region = n_region;
    H.region0 = (1023 + e_min) << 2;
    /* small hint: buckets of (hi >> 18) from the bucket of x[1] up to the bucket below r_lo */
    {
        const int b_lo = hi_word(x[1]) >> 18, b_hi = (hi_word(r_lo) >> 18) - 1;
        H.shint0 = b_lo; H.n_shint = std::max(1, b_hi - b_lo + 1);
        H.off_shint = alloc_bytes((size_t)H.n_shint * 4);
        int *hint = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(B.data()) + H.off_shint);
        int i = 0;
        for (int b = 0; b < H.n_shint; ++b) {
            const double edge = from_hi_word((b + b_lo) << 18);
            while (i < sn - 2 && x[i + 1] <= edge) ++i;
            hint[b] = b == 0 ? 0 : i;
        }
    }
    H.sn = sn;
    H.off_sx = alloc_bytes((size_t)sn * 8);
    H.off_sk = alloc_bytes((size_t)(sn - 1) * kSmallStride * 8);
    /* I^4_0 */
    const bool div = L.has_divergent != 0;
    if (div) {
        const int n = t.i40.n;
        const int b_lo = hi_word(t.i40.x[1]) >> 15, b_hi = hi_word(t.i40.x[n - 1]) >> 15;
        H.i40hint0 = b_lo; H.n_i40hint = b_hi - b_lo + 1; H.i40n = n;
        H.off_i40hint = alloc_bytes((size_t)H.n_i40hint * 4);
        H.off_i40x = alloc_bytes((size_t)n * 8);
        H.off_i40k = alloc_bytes((size_t)(n - 1) * 32);
    }
    H.off_knots = alloc_bytes(0);
    size_t nk = total_knots();
    H.n_knots = (int)nk;
    alloc_bytes(nk * kKnotBytes);
    H.bytes = (int)(B.size() * 8);
    H.r_lo = r_lo;
    H.max_rel_error = max_err;
    H.max_rel_error_i40 = max_err40;
    unsigned char *base = reinterpret_cast<unsigned char *>(B.data());
    {
        int2 *rg = reinterpret_cast<int2 *>(base + H.off_region);
        unsigned char *kn = base + H.off_knots;
        int first = 0;
        out.region_m.clear();
        for (int i = 0; i < n_region; ++i) {
            rg[i].x = first; rg[i].y = 18 - regs[i].m;
            /* pack (compact.cuh: knot_value_offset / knot_slope_offset) */
            const int nk_reg = (1 << regs[i].m) + 1;
            for (int j = 0; j < nk_reg; ++j) {
                const double *src = &regs[i].knots[(size_t)j * kKnotStride];
                unsigned char *dst = kn + (size_t)(first + j) * kKnotBytes;
                for (int f = 0; f < N_ITAB9; ++f) {
                    std::memcpy(dst + knot_value_offset(f), &src[2 * f], 8);
                    const int w = slope_word(src[2 * f + 1]);
                    std::memcpy(dst + knot_slope_offset(f), &w, 4);
                }
            }
            first += nk_reg;
            out.region_m.push_back(regs[i].m);
        }
    }
    {
        double *sx = reinterpret_cast<double *>(base + H.off_sx);
        double *sk = reinterpret_cast<double *>(base + H.off_sk);
        for (int i = 0; i < sn; ++i) sx[i] = x[i];
        for (int i = 0; i < sn - 1; ++i)
            for (int f = 0; f < N_ITAB; ++f) {
                const double4 c = t.ifam.k[(size_t)i * N_ITAB + f];
                double *d = sk + (size_t)i * kSmallStride + 4 * f;
                d[0] = c.x; d[1] = c.y; d[2] = c.z; d[3] = c.w;
            }
    }
    if (div) {
        const int n = t.i40.n;
        int *hint = reinterpret_cast<int *>(base + H.off_i40hint);
        double *ix = reinterpret_cast<double *>(base + H.off_i40x);
        double *ik = reinterpret_cast<double *>(base + H.off_i40k);
        int i = 0;
        for (int b = 0; b < H.n_i40hint; ++b) {
            const double edge = from_hi_word((b + H.i40hint0) << 15);
            while (i < n - 2 && t.i40.x[i + 1] <= edge) ++i;
            hint[b] = b == 0 ? 0 : i;
        }
        for (int j = 0; j < n; ++j) ix[j] = t.i40.x[j];
        for (int j = 0; j < n - 1; ++j) { const double4 c = t.i40.k[j]; ik[4 * j] = c.x; ik[4 * j + 1] = c.y; ik[4 * j + 2] = c.z; ik[4 * j + 3] = c.w; }
    }
    H.valid = 1;
    return true;
}

/* bytes of the blob a launch must stage so that every R <= r_need is covered */
inline int compact_stage_bytes(const CompactBuild &cb, const double r_need)
{
    const CompactHeader &H = cb.h;
    const unsigned char *base = reinterpret_cast<const unsigned char *>(cb.blob.data());
    const int2 *rg = reinterpret_cast<const int2 *>(base + H.off_region);
    int reg = (hi_word(r_need) >> 18) - H.region0 + 1;        /* one region of slack */
    if (reg < 0) reg = 0;
    int knots = H.n_knots;
    if (reg + 1 < H.n_region) knots = rg[reg + 1].x;
    return (H.off_knots + knots * kKnotBytes + 15) / 16 * 16;
}

}  // namespace coffe
