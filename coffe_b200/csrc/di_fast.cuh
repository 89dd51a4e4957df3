// Integrated x integrated piece (functions_double_integrated, src/functions.c:3460-4016, under the Cuhre
// call of src/signal.c:737-752) organised for the inner loop of the sm_100a kernel.
//
// Same integrand as di_eval (physics.cuh), evaluated on the ridge panels of quadrature.cuh, but
//   * everything that depends only on the work item is folded into DIFast (the polynomial coefficients of
//     the len-len bracket in Horner form, H0^4 included), everything that depends only on the outer node
//     lambda_1 into DIOuter; the potential x lensing weights (1-x)/x are multiplied out against the bracket's
//     own lambda_1 lambda_2 factor, so the inner loop has no division;
//   * the separation of the two points is R^2 = d^2 + lambda_1^2 sin^2(theta), d = lambda_2 - lambda_1 cos(theta):
//     the panels on the two sides of the ridge are mirror images, so as long as neither side is clipped by the
//     ends of the line of sight the nodes at -d and +d share R and therefore ONE lookup of the nine I^n_l
//     (the dominant shared-memory traffic); only the remainder of the longer side is evaluated singly;
//   * renormalisation(lambda_1, lambda_2) is read from the row collapsed at lambda_1 (tables.cuh: ren_row_fill),
//     filled once per outer node by the lanes of the warp.
// The functions are lane-strided like item.cuh: `nlanes` cooperating lanes (a warp on the device; the CPU
// harness calls them with nlanes = n_panel lanes in turn), lane = panel slot * n_panel + node.
#pragma once
#include "item.cuh"

namespace coffe {

struct DIFast {
    /* len-len bracket, premultiplied by k_ll H0^4:
       br = ll [I00 k00 + I02 (k02a + k02b q) + I04 (k04a + q (k04b + k04c q)) + I13 k13c q]
          + (k11a R^2 + k11b ll)(I11 + I13) + I20 k20,        ll = l1 l2, q = ll / R^2 */
    double k00, k20, k11a, k11b, k02a, k02b, k13c, k04a, k04b, k04c;
    double k_ll0, k_ll1;        /* R^2 <= 1e-20 fallback: k_ll0 + k_ll1 ll I11       src/functions.c:3814-3833 */
    /* potential x lensing: G = 2c I31 - ll s2 I22 */
    double two_c, s2, k_4l, k_5la, k_5lb;
    /* renormalised potential terms */
    double k_44, k_55, k_45a, k_45b;
    double c, chi1, chi2;
    int any_ll, any_pl, any_ren;
};

COFFE_HD DIFast di_fast_setup(const DIConst &k, const Geometry &g)
{
    DIFast f;
    const double c = g.c, cm = -g.s2, a = k.k_ll * k.h4;
    f.k00 = a * 2 * cm / 5.;
    f.k20 = a * 4 * c / 3.;
    f.k11a = a * 4 * c / 15.;
    f.k11b = a * 24 * c * c / 15.;
    f.k02a = a * 4 * cm / 7.;
    f.k02b = a * 6 * cm * c / 7.;
    f.k13c = a * 2 * c * cm;
    f.k04a = a * 6 * cm / 35.;
    f.k04b = a * 30 * c * cm / 35.;
    f.k04c = a * cm * cm;
    f.k_ll0 = a * 4.; f.k_ll1 = a * 24. / 15.;
    f.two_c = 2 * c; f.s2 = g.s2;
    f.k_4l = k.k_4l * k.h4; f.k_5la = k.k_5la * k.h4; f.k_5lb = k.k_5lb * k.h4;
    f.k_44 = k.k_44 * k.h4; f.k_55 = k.k_55 * k.h4; f.k_45a = k.k_45a * k.h4; f.k_45b = k.k_45b * k.h4;
    f.c = c; f.chi1 = g.chi1; f.chi2 = g.chi2;
    f.any_ll = k.any_ll; f.any_pl = k.any_pl; f.any_ren = k.any_ren;
    return f;
}

/* constants of one outer node lambda_1; W1 = los1.x is folded into every amplitude */
struct DIOuter {
    double l1, c0, eps2;        /* ridge position l1 cos(theta), squared impact parameter l1^2 sin^2(theta) */
    double e1W1;                /* (chi1 - l1) W1 */
    double A1, E4, E5;          /* potential x lensing: W1 l1 (k_4l + k_5la U1), W1 e1 k_4l, W1 e1 k_5lb */
    double B0, B1;              /* renormalised terms: W1 (k_44 + k_45b V1), W1 (k_45a + k_55 V1) */
};

COFFE_HD DIOuter di_outer_setup(const DIFast &f, const double l1, const double4 los1)
{
    DIOuter o;
    o.l1 = l1; o.c0 = l1 * f.c; o.eps2 = l1 * l1 * f.s2;
    const double e1 = f.chi1 - l1, W1 = los1.x;
    o.e1W1 = e1 * W1;
    o.A1 = W1 * l1 * (f.k_4l + f.k_5la * los1.z);
    o.E4 = W1 * e1 * f.k_4l;
    o.E5 = W1 * e1 * f.k_5lb;
    o.B0 = W1 * (f.k_44 + f.k_45b * los1.y);
    o.B1 = W1 * (f.k_45a + f.k_55 * los1.y);
    return o;
}

/* everything of a node that depends on |d| only: shared by the mirror nodes at -d and +d */
struct DIShared {
    double r2, ir2;
    double I00k, I20k, I13kq;   /* I00 k00, I20 k20, I13 k13c / R^2 (so that x ll gives I13 k13c q) */
    double S, P0;               /* I11 + I13, k11a R^2 */
    double I02, I04, I11, I22, twocI31, I40;
};

template <class TA>
COFFE_HD DIShared di_shared(const TA &ta, const DIFast &f, const DIOuter &o, const double d)
{
    DIShared s;
    const double r2 = fma(d, d, o.eps2);
    const double ir = r2 > 0.0 ? COFFE_RSQRT(r2) : 0.0;
    double I[N_ITAB9];
    ta.ifam9(r2 * ir, I, f.any_ren != 0);
    s.r2 = r2; s.ir2 = ir * ir;
    s.I00k = I[I00] * f.k00; s.I20k = I[I20] * f.k20; s.I13kq = I[I13] * f.k13c * s.ir2;
    s.S = I[I11] + I[I13]; s.P0 = f.k11a * r2;
    s.I02 = I[I02]; s.I04 = I[I04]; s.I11 = I[I11]; s.I22 = I[I22]; s.twocI31 = f.two_c * I[I31]; s.I40 = I[N_ITAB];
    return s;
}

/* what the inner loop needs to read the tables of the second line of sight */
template <class TA>
struct DISideCtx {
    const TA *ta;
    const double2 *los; int los_n; double los_inv_h;
    const Renorm2DView *r2d; const double2 *row; int n_row;
};

/* integrand at lambda_2 = l2 (per unit lambda_2, H0^4 and all amplitudes included).  ALL: every group of terms is
   on and R^2 > 1e-12 is guaranteed by the caller (squared impact parameter above it), so the body is branch-free. */
template <bool ALL, class TA>
COFFE_HD double di_side(const DISideCtx<TA> &cx, const DIFast &f, const DIOuter &o, const DIShared &s, const double l2)
{
    const double4 los2 = los_eval(cx.los, cx.los_n, cx.los_inv_h, l2);
    const double ll = o.l1 * l2, e2 = f.chi2 - l2;
    double acc = 0.0;
    if (ALL || f.any_ll) {
        double br;
        if (ALL || s.r2 > 1e-20) {
            const double q = ll * s.ir2;
            double t = fma(s.I02, fma(f.k02b, q, f.k02a), s.I00k);
            t = fma(s.I04, fma(fma(f.k04c, q, f.k04b), q, f.k04a), t);
            t = fma(s.I13kq, ll, t);
            br = fma(fma(f.k11b, ll, s.P0), s.S, s.I20k);
            br = fma(ll, t, br);
        }
        else br = fma(f.k_ll1 * ll, s.I11, f.k_ll0);
        acc = o.e1W1 * e2 * br;
    }
    if (ALL || f.any_pl) {
        const double G = s.twocI31 - (f.s2 * ll) * s.I22;
        const double F = fma(o.A1, e2, l2 * fma(o.E5, los2.z, o.E4));
        acc = fma(G, F, acc);
    }
    if (ALL || f.any_ren) {
        const double ren = (!ALL && s.r2 <= 1e-12) ? cx.ta->ren0(o.l1) : s.I40 - ren_row_eval(*cx.r2d, cx.row, cx.n_row, l2);
        acc = fma(ren, fma(o.B1, los2.y, o.B0), acc);
    }
    return los2.x * acc;
}

/* the panels of one inner integral: a prefix of K_pair mirror pairs, then the remainders of the two sides */
struct DIPanels {
    PanelPlan plan;
    double a_neg, b_neg, b_pos;
    int k_neg, k_pos, k_pair;
};

COFFE_HD DIPanels di_panels(const Rules &q, const DIFast &f, const DIOuter &o)
{
    DIPanels p;
    p.plan = panel_plan(sqrt(o.eps2), q.eps_min, q.panel_w_di, q.panel_du_di, q.tail_ratio_di, q.rpow_di, q.geo_shift_di);
    p.b_neg = o.c0;
    p.a_neg = o.c0 > f.chi2 ? o.c0 - f.chi2 : 0.0;
    p.b_pos = f.chi2 - o.c0;
    p.k_neg = panel_count_by_break(p.plan, p.b_neg);
    p.k_pos = p.b_pos > 0.0 ? panel_count_by_break(p.plan, p.b_pos) : 0;
    p.k_pair = 0;
    if (p.a_neg == 0.0 && p.k_pos > 0 && p.k_neg > 0) {
        const double bmin = p.b_neg < p.b_pos ? p.b_neg : p.b_pos;
        const int kc = p.k_neg < p.k_pos ? p.k_neg : p.k_pos;           /* panels of the shorter side */
        p.k_pair = panel_break(p.plan, kc) > bmin ? kc - 1 : kc;        /* its last panel is clipped unless a break falls on the end */
    }
    return p;
}

/* clipped panel k of one side: [lo, lo + w) in |d|; w = 0 if empty */
COFFE_HD void di_single_panel(const DIPanels &p, const bool neg, const int k, double &lo, double &w)
{
    lo = panel_break(p.plan, k);
    double up = panel_break(p.plan, k + 1);
    const double a = neg ? p.a_neg : 0.0, b = neg ? p.b_neg : p.b_pos;
    if (lo < a) lo = a;
    if (up > b) up = b;
    w = up > lo ? up - lo : 0.0;
}

/* inner integral over lambda_2 in [0, chi2] at one outer node: this lane's partial sum */
template <bool ALL, class TA>
COFFE_HD double di_inner_partial(const DISideCtx<TA> &cx, const Rules &q, const DIFast &f, const DIOuter &o,
                                 const int lane, const int nlanes)
{
    const int np = q.n_panel_di;
    const int ppi = nlanes >= np ? nlanes / np : 1;
    const int lp = lane / np, j = lane - lp * np;
    if (nlanes >= np && lp >= ppi) return 0.0;
    const double *px = q.panel_x ? q.panel_x : q.gl_x + rule_offset(np), *pw = q.panel_w ? q.panel_w : q.gl_w + rule_offset(np);
    const DIPanels p = di_panels(q, f, o);
    double acc = 0.0;
    /* nodes this lane owns: j, j + nlanes, ... (only when a panel has more nodes than there are lanes) */
    for (int jj = nlanes >= np ? j : lane; jj < np; jj += nlanes >= np ? np : nlanes) {
        const double xj = px[jj], wj = pw[jj];
        /* mirror pairs */
        for (int k = nlanes >= np ? lp : 0; k < p.k_pair; k += ppi) {
            const double lo = panel_break(p.plan, k), w = panel_break(p.plan, k + 1) - lo;
            const double d = fma(w, xj, lo);
            const DIShared s = di_shared(*cx.ta, f, o, d);
            const double v = di_side<ALL>(cx, f, o, s, o.c0 - d) + di_side<ALL>(cx, f, o, s, o.c0 + d);
            acc = fma(w * wj, v, acc);
        }
        /* remainders: negative side k_pair..k_neg-1, then positive side k_pair..k_pos-1 */
        const int n_negl = p.k_neg - p.k_pair, n_left = n_negl + (p.k_pos - p.k_pair);
        for (int e = nlanes >= np ? lp : 0; e < n_left; e += ppi) {
            const bool neg = e < n_negl;
            double lo, w;
            di_single_panel(p, neg, p.k_pair + (neg ? e : e - n_negl), lo, w);
            if (w == 0.0) continue;
            const double d = fma(w, xj, lo);
            const DIShared s = di_shared(*cx.ta, f, o, d);
            acc = fma(w * wj, di_side<ALL>(cx, f, o, s, neg ? o.c0 - d : o.c0 + d), acc);
        }
    }
    return acc;
}

#if defined(__CUDA_ARCH__)
/* The same sum on a full warp, with the panel bookkeeping done by the warp instead of by every lane: lane k holds the break
   B_k, the panel counts are ballots, and a panel's ends are two shuffles (profiles/README.md: the scalar bookkeeping -- doubling
   loops, per-iteration break formulas -- was 15 % of the kernel's instructions with ~3 inner iterations per outer node).  Same
   breaks, same counts, same nodes, same summation order as di_inner_partial: bit-identical.  Only valid when a side never has more
   than 31 panels and the panel rule is staged (q.panel_x): the host checks the configured plan (rules_build.h: di_side_panels) and
   launches the WARP_PANELS instantiation of the kernel. */
template <bool ALL, class TA>
__device__ __forceinline__ double di_inner_partial_warp(const DISideCtx<TA> &cx, const Rules &q, const DIFast &f, const DIOuter &o, const int lane)
{
    const unsigned int full = 0xffffffffu;
    const int np = q.n_panel_di;            /* <= 32 (the caller checks) */
    const int ppi = 32 / np;
    const int lp = lane / np, j = lane - lp * np;
    const bool on = lp < ppi;
    const PanelPlan plan = panel_plan(sqrt(o.eps2), q.eps_min, q.panel_w_di, q.panel_du_di, q.tail_ratio_di, q.rpow_di, q.geo_shift_di);
    const double brk = panel_break(plan, lane);
    const double b_neg = o.c0, a_neg = o.c0 > f.chi2 ? o.c0 - f.chi2 : 0.0, b_pos = f.chi2 - o.c0;
    const int k_neg = b_neg > 0.0 ? __popc(__ballot_sync(full, brk < b_neg)) : 0;
    const int k_pos = b_pos > 0.0 ? __popc(__ballot_sync(full, brk < b_pos)) : 0;
    int k_pair = 0;
    if (a_neg == 0.0 && k_pos > 0 && k_neg > 0) {
        const double bmin = b_neg < b_pos ? b_neg : b_pos;
        const int kc = k_neg < k_pos ? k_neg : k_pos;
        k_pair = __shfl_sync(full, brk, kc) > bmin ? kc - 1 : kc;
    }
    const double xj = q.panel_x[j], wj = q.panel_w[j];
    double acc = 0.0;
    for (int k0 = 0; k0 < k_pair; k0 += ppi) {
        const int k = k0 + lp;
        const double lo = __shfl_sync(full, brk, k & 31), up = __shfl_sync(full, brk, (k + 1) & 31);
        if (on && k < k_pair) {
            const double w = up - lo;
            const double d = fma(w, xj, lo);
            const DIShared s = di_shared(*cx.ta, f, o, d);
            const double v = di_side<ALL>(cx, f, o, s, o.c0 - d) + di_side<ALL>(cx, f, o, s, o.c0 + d);
            acc = fma(w * wj, v, acc);
        }
    }
    const int n_negl = k_neg - k_pair, n_left = n_negl + (k_pos - k_pair);
    for (int e0 = 0; e0 < n_left; e0 += ppi) {
        const int e = e0 + lp;
        const bool neg = e < n_negl;
        const int k = k_pair + (neg ? e : e - n_negl);
        double lo = __shfl_sync(full, brk, k & 31), up = __shfl_sync(full, brk, (k + 1) & 31);
        const double a = neg ? a_neg : 0.0, b = neg ? b_neg : b_pos;
        if (lo < a) lo = a;
        if (up > b) up = b;
        if (on && e < n_left && up > lo) {
            const double w = up - lo;
            const double d = fma(w, xj, lo);
            const DIShared s = di_shared(*cx.ta, f, o, d);
            acc = fma(w * wj, di_side<ALL>(cx, f, o, s, neg ? o.c0 - d : o.c0 + d), acc);
        }
    }
    return acc;
}
#endif

COFFE_HD unsigned int di_inner_eval_count(const Rules &q, const DIFast &f, const DIOuter &o)
{
    const DIPanels p = di_panels(q, f, o);
    unsigned int n = 2u * p.k_pair;
    for (int k = p.k_pair; k < p.k_neg; ++k) { double lo, w; di_single_panel(p, true, k, lo, w); if (w > 0.0) ++n; }
    for (int k = p.k_pair; k < p.k_pos; ++k) { double lo, w; di_single_panel(p, false, k, lo, w); if (w > 0.0) ++n; }
    return n * q.n_panel_di;
}

/* integrated x integrated piece per unit (x1, x2).  Outer nodes strided over `ngroups` groups of `nlanes` lanes;
   `row` is the group's own buffer of t.ren2d.n double2 (renormalisation row), unused without the divergent terms. */
template <class TA, bool WARP_PANELS = false>
COFFE_HD double di_fast_partial(const TA &ta, const Tables &t, const double2 *los, const Rules &q, const Geometry &g, const DIConst &k,
                                double2 *row, const int group, const int ngroups, const int lane, const int nlanes,
                                const bool serial_fill = false /* CPU harness: lanes run one after the other, each fills the whole row */)
{
    const DIFast f = di_fast_setup(k, g);
    const int n_outer = q.n_outer > 0 ? q.n_outer : auto_n_outer(g.sep);
    const double *outer_x = q.gl_x + rule_offset(n_outer), *outer_w = q.gl_w + rule_offset(n_outer);
    DISideCtx<TA> cx;
    cx.ta = &ta; cx.los = los; cx.los_n = t.los.n; cx.los_inv_h = t.los.inv_h;
    cx.r2d = &t.ren2d; cx.row = row;
    cx.n_row = f.any_ren ? ren_row_count(t.ren2d, g.chi2) : 0;
    double acc = 0.0;
    for (int io = group; io < n_outer; io += ngroups) {
        const double l1 = g.chi1 * outer_x[io];
        const DIOuter o = di_outer_setup(f, l1, los_eval(los, t.los.n, t.los.inv_h, l1));
        if (f.any_ren) {
            COFFE_SYNC_LANES();             /* the previous node's readers are done */
            ren_row_fill(t.ren2d, l1, cx.n_row, row, serial_fill ? 0 : lane, serial_fill ? 1 : nlanes);
            COFFE_SYNC_LANES();
        }
        const bool all = f.any_ll && f.any_pl && f.any_ren && o.eps2 > 1e-12;
#if defined(__CUDA_ARCH__) && !defined(COFFE_NO_WARP_PANELS)
        if (WARP_PANELS) {
            acc = fma(outer_w[io], all ? di_inner_partial_warp<true>(cx, q, f, o, lane) : di_inner_partial_warp<false>(cx, q, f, o, lane), acc);
            continue;
        }
#endif
        acc = fma(outer_w[io], all ? di_inner_partial<true>(cx, q, f, o, lane, nlanes) : di_inner_partial<false>(cx, q, f, o, lane, nlanes), acc);
    }
    return acc / g.chi2;
}

COFFE_HD unsigned int di_fast_eval_count(const Rules &q, const Geometry &g, const DIConst &k)
{
    const DIFast f = di_fast_setup(k, g);
    const int n_outer = q.n_outer > 0 ? q.n_outer : auto_n_outer(g.sep);
    const double *outer_x = q.gl_x + rule_offset(n_outer);
    unsigned int n = 0;
    double4 unit; unit.x = unit.y = unit.z = unit.w = 1.0;
    for (int io = 0; io < n_outer; ++io) n += di_inner_eval_count(q, f, di_outer_setup(f, g.chi1 * outer_x[io], unit));
    return n;
}

}  // namespace coffe
