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
//   * of the renormalised potential terms only I^4_0(R) rides on the ridge panels; renormalisation(lambda_1, lambda_2), which has no
//     structure on the scale of the impact parameter, has its own inner rule (di_ren_smooth_partial).
// Lane layout: groups of n_panel lanes, one group per outer node, lane j of a group = node j of every panel (di_inner_group);
// the CPU harness calls the same functions for the n_panel lanes of one group in turn.
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
    f.k00 = a * 2 * cm * (1.0 / 5.0);
    f.k20 = a * 4 * c * (1.0 / 3.0);
    f.k11a = a * 4 * c * (1.0 / 15.0);
    f.k11b = a * 24 * c * c * (1.0 / 15.0);
    f.k02a = a * 4 * cm * (1.0 / 7.0);
    f.k02b = a * 6 * cm * c * (1.0 / 7.0);
    f.k13c = a * 2 * c * cm;
    f.k04a = a * 6 * cm * (1.0 / 35.0);
    f.k04b = a * 30 * c * cm * (1.0 / 35.0);
    f.k04c = a * cm * cm;
    f.k_ll0 = a * 4.; f.k_ll1 = a * 24. * (1.0 / 15.0);
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
template <class TA, class LP = const double2 *>
struct DISideCtx {
    const TA *ta;
    LP los; int los_n; double los_inv_h;        /* Hermite knots (const double2 *) or the staged records (const LosRec *) */
    const Renorm2DView *r2d;
};

/* integrand at lambda_2 = l2 (per unit lambda_2, H0^4 and all amplitudes included).  ALL: every group of terms is
   on and R^2 > 1e-12 is guaranteed by the caller (squared impact parameter above it), so the body is branch-free. */
template <bool ALL, class TA, class LP>
COFFE_HD double di_side(const DISideCtx<TA, LP> &cx, const DIFast &f, const DIOuter &o, const DIShared &s, const double l2)
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
        /* renormalization(l1, l2) itself is integrated apart (di_ren_smooth_partial): only I^4_0(R) carries the cusp */
        const double ren = (!ALL && s.r2 <= 1e-12) ? cx.ta->ren0(o.l1) + renorm2d_eval(*cx.r2d, o.l1, l2) : s.I40;
        acc = fma(ren, fma(o.B1, los2.y, o.B0), acc);
    }
    return los2.x * acc;
}

/* The renormalised potential terms are (I^4_0(R) - renormalization(l1, l2)) x smooth line-of-sight factors (src/functions.c:3838-4016).
   The two halves are of the same size (both carry the infrared part of the spectrum) and cancel to a few per cent, but only I^4_0(R)
   has structure on the scale of the impact parameter: renormalization(l1, l2), a bicubic interpolant on a ~30 Mpc grid, varies on
   scales of hundreds of Mpc, with its curvature concentrated around the diagonal l2 = l1.  Its inner integral therefore gets its own
   rule -- two Gauss-Legendre panels split at the diagonal, n_ren nodes together (one per lane on the device) -- instead of riding on
   the ridge panels: against a composite rule on the cells of the grid 2 x 16 nodes leave 4e-8 of this half (CPU experiment, three
   distances), i.e. < 1e-6 of the difference.  Returns this lane's share of - integral dl2 W2 (B0 + B1 V2) renormalization(l1, l2). */
template <class LP>
COFFE_HD double di_ren_smooth_partial(const Renorm2DView &r2d, const LP los, const int los_n, const double los_inv_h, const Rules &q,
                                      const DIFast &f, const DIOuter &o, const int lane, const int nlanes)
{
    const int half = q.n_ren / 2;
    const double *x = q.gl_x + rule_offset(half), *w = q.gl_w + rule_offset(half);
    const double split = o.l1 < f.chi2 ? o.l1 : f.chi2;
    double acc = 0.0;
    for (int i = lane; i < 2 * half; i += nlanes) {
        const bool upper = i >= half;
        const int jj = upper ? i - half : i;
        const double lo = upper ? split : 0.0, wd = upper ? f.chi2 - split : split;
        const double l2 = fma(wd, x[jj], lo);
        const double4 los2 = los_eval(los, los_n, los_inv_h, l2);
        acc = fma(wd * w[jj], los2.x * renorm2d_eval(r2d, o.l1, l2) * fma(o.B1, los2.y, o.B0), acc);
    }
    return -acc;
}

/* The inner integral over lambda_2 in [0, chi2] at one outer node, on ONE GROUP of n_panel lanes: lane j owns node j of every panel
   and walks through the units of the integral in sequence (quadrature.cuh: RidgeUnits -- mirror pairs share one lookup of the nine
   I^n_l).  A warp runs 32 / n_panel groups side by side, each on its own outer node (engine.cu: di_items_smem_kernel), so the
   lanes stay busy whatever the panel counts are and the set-up of an outer node is paid once per group, not once per warp.
   ONE loop over the units and ONE copy of the integrand (di_shared, and di_side in a loop over the two sides that is not unrolled):
   the kernel is bound by its instruction footprint -- with separate loops for the pairs (two inlined sides) and the single panels
   it measured 12 % slower, with a third copy for the last panels 25 %.  Returns this lane's partial sum. */
COFFE_HD PanelPlan di_plan(const Rules &q, const DIOuter &o)
{
    /* impact parameter and its inverse from one reciprocal square root */
    const double e2 = fmax(o.eps2, q.eps_min * q.eps_min), ie = COFFE_RSQRT(e2);
    return panel_plan(e2 * ie, q.eps_min, q.panel_w_di, q.panel_du_di, q.tail_ratio_di, q.rpow_di, q.geo_shift_di, q.geo_L_di, q.tail_lnr_di, ie, q.inv_w_di, q.inv_geo_L_di);
}

template <bool ALL, class TA, class LP>
COFFE_HD double di_inner_group(const DISideCtx<TA, LP> &cx, const Rules &q, const DIFast &f, const DIOuter &o, const int j)
{
    const int np = q.n_panel_di;
    const double xj = q.panel_x ? q.panel_x[j] : q.gl_x[rule_offset(np) + j], wj = q.panel_w ? q.panel_w[j] : q.gl_w[rule_offset(np) + j];
    const PanelPlan plan = di_plan(q, o);
    RidgeUnits u;
    ridge_units_init(u, plan, o.c0, f.chi2, xj, wj, plan.lnr > 0.0 ? (q.panel_p ? q.panel_p[j] : exp(plan.lnr * xj)) : 0.0);
    double acc = 0.0;
    for (int iu = 0; iu < u.n_units; ++iu) {
        double d, wgt;
        const int sides = ridge_unit_node(u, iu, d, wgt);
        if (sides) {
            const DIShared s = di_shared(*cx.ta, f, o, d);
            double v = 0.0;
#pragma unroll 1
            for (int sd = 0; sd < 2; ++sd)
                if (sides & (1 << sd)) v += di_side<ALL>(cx, f, o, s, sd == 0 ? o.c0 - d : o.c0 + d);
            acc = fma(wgt, v, acc);
        }
    }
    return acc;
}

COFFE_HD unsigned int di_inner_eval_count(const Rules &q, const DIFast &f, const DIOuter &o)
{
    return ridge_units_count(di_plan(q, o), o.c0, f.chi2) * q.n_panel_di;
}

/* integrated x integrated piece per unit (x1, x2): outer nodes strided over `ngroups` groups of n_panel lanes, `j` = this lane's node.
   The loop is uniform over the groups (a group without an outer node in the last round idles): no lane leaves early, so the device
   kernels may use warp-wide votes inside. */
template <class TA, class LP>
COFFE_HD double di_fast_partial(const TA &ta, const Tables &t, const LP los, const Rules &q, const Geometry &g, const DIConst &k,
                                const int group, const int ngroups, const int j)
{
    const DIFast f = di_fast_setup(k, g);
    const int n_outer = q.n_outer > 0 ? q.n_outer : auto_n_outer(g.sep, q.n_outer_min, q.n_outer_step, q.n_panel_di);
    const double *outer_x = q.gl_x + rule_offset(n_outer), *outer_w = q.gl_w + rule_offset(n_outer);
    DISideCtx<TA, LP> cx;
    cx.ta = &ta; cx.los = los; cx.los_n = t.los.n; cx.los_inv_h = t.los.inv_h;
    cx.r2d = &t.ren2d;
    double acc = 0.0;
    for (int io0 = 0; io0 < n_outer; io0 += ngroups) {
        const int io = io0 + group;
        const bool valid = group < ngroups && io < n_outer;
        const double l1 = g.chi1 * outer_x[valid ? io : io0];
        const DIOuter o = di_outer_setup(f, l1, los_eval(los, t.los.n, t.los.inv_h, l1));
        bool all = f.any_ll && f.any_pl && f.any_ren && o.eps2 > 1e-12;
#if defined(__CUDA_ARCH__)
        all = __all_sync(0xffffffffu, all);       /* one instantiation per round for the whole warp */
#endif
        if (valid) {
            double v = all ? di_inner_group<true>(cx, q, f, o, j) : di_inner_group<false>(cx, q, f, o, j);
            if (f.any_ren) v += di_ren_smooth_partial(t.ren2d, los, t.los.n, t.los.inv_h, q, f, o, j, q.n_panel_di);
            acc = fma(outer_w[io], v, acc);
        }
    }
    return acc / g.chi2;
}

COFFE_HD unsigned int di_fast_eval_count(const Rules &q, const Geometry &g, const DIConst &k)
{
    const DIFast f = di_fast_setup(k, g);
    const int n_outer = q.n_outer > 0 ? q.n_outer : auto_n_outer(g.sep, q.n_outer_min, q.n_outer_step, q.n_panel_di);
    const double *outer_x = q.gl_x + rule_offset(n_outer);
    unsigned int n = 0;
    double4 unit; unit.x = unit.y = unit.z = unit.w = 1.0;
    for (int io = 0; io < n_outer; ++io) n += di_inner_eval_count(q, f, di_outer_setup(f, g.chi1 * outer_x[io], unit));
    return n;
}

}  // namespace coffe
