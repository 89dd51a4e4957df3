// Per-work-item integration: one (z_mean, separation, mu) point -> the three pieces the reference's
// dispatcher returns (coffe_integrate, src/signal.c:472-771): local x local (closed form),
// local x integrated (1-D line-of-sight integral per half) and integrated x integrated (2-D).
//
// Written as per-lane partial sums: lane j of a group of n_panel lanes owns node j of every panel and the caller reduces the
// partials (warp shuffle on the device; the CPU debugging harness calls the lanes of a group in turn), so the device kernels and
// the harness execute the same arithmetic on the same nodes.  (ridge_integrate / si_half_partial / di_partial below are the plain
// panel-list form of the same integrals: the midpoint-approximation mode, whose 2-D tables rule out the shared lookup of mirror
// nodes, still runs on them.)
#pragma once
#include "physics.cuh"
#include "quadrature.cuh"

namespace coffe {

struct Rules {
    const double *gl_x, *gl_w;      /* bank of Gauss-Legendre rules on [0,1] (quadrature.cuh: rule_offset) */
    const double *rpow;             /* tail_ratio^n, n < kMaxTail */
    int n_panel_si, n_panel_di;
    int n_outer;                    /* 0 = auto_n_outer(sep) */
    int n_outer_min; double n_outer_step;    /* by separation: max(n_outer_min, 10 + sep / n_outer_step) */
    double panel_w_mpc, panel_du, eps_min, tail_ratio;
    double panel_w_di, panel_du_di;  /* panel plan of the inner integral of the integrated x integrated piece */
    const double *panel_x, *panel_w; /* nodes / weights of the n_panel_di rule when staged separately (else null: read from the bank) */
    const double *panel_p;           /* tail_ratio_di^x_j for the nodes of that rule, when staged (log-mapped tail panels) */
    const double *panel_x_si, *panel_w_si;   /* the same for the n_panel_si rule */
    const double *panel_p_si;                /* tail_ratio^x_j for the nodes of that rule, when staged */
    const double *rpow_di;           /* tail_ratio_di^n, n < kMaxTail */
    double tail_ratio_di;
    int geo_shift_di;
    int n_ren;                       /* nodes of the inner rule of renormalization(l1, l2) (di_fast.cuh: di_ren_smooth_partial), even */
    double tail_lnr, tail_lnr_di;    /* log of the tail ratio when the tail panels carry their nodes in log(d) (quadrature.cuh); 0 = in d */
    double inv_w, inv_w_di, inv_geo_L, inv_geo_L_di;   /* reciprocals of the panel widths and of geo_L (0 where that is 0): rules_reciprocals */
    double geo_L, geo_L_di;          /* log of the growth factor of the sinh-map geometric panels (quadrature.cuh); 0 = doubling panels */
};

inline void rules_reciprocals(Rules &q)
{
    q.inv_w = 1.0 / q.panel_w_mpc; q.inv_w_di = 1.0 / q.panel_w_di;
    q.inv_geo_L = q.geo_L > 0.0 ? 1.0 / q.geo_L : 0.0; q.inv_geo_L_di = q.geo_L_di > 0.0 ? 1.0 / q.geo_L_di : 0.0;
}

/* panels one ridge_integrate call can hold in its scratch area (engine.cu checks the configured plan) */
static const int kMaxPanels = 256;

/* integral over lambda in [0, hi] of f(lambda), with a ridge (closest approach) at `center` of width
   eps.  Cooperative over `nlanes` lanes (one warp on the device; 1 in the CPU harness): first the lanes
   fill `scratch` (2*kMaxPanels doubles shared by the lanes) with {start, signed width} of every clipped
   panel, then each lane owns one Gauss-Legendre node index and strides over the panels.  Returns this
   lane's partial sum. */
template <class F>
COFFE_HD double ridge_integrate(const F &f, const double hi, const double center, const double eps,
                                const Rules &q, const int n_panel, const int lane, const int nlanes, double *scratch)
{
    const bool staged = q.panel_x_si && n_panel == q.n_panel_si;
    const double *panel_x = staged ? q.panel_x_si : q.gl_x + rule_offset(n_panel), *panel_w = staged ? q.panel_w_si : q.gl_w + rule_offset(n_panel);
    const PanelPlan plan = panel_plan(eps, q.eps_min, q.panel_w_mpc, q.panel_du, q.tail_ratio, q.rpow, 1, q.geo_L, q.tail_lnr);
    /* negative side: |d| in [a_neg, b_neg]; positive side: |d| in [0, b_pos] (center >= 0 always) */
    const double b_neg = center;
    const double a_neg = center > hi ? center - hi : 0.0;
    const double b_pos = hi - center;
    const int k_neg = panel_count(plan, b_neg);
    const int k_pos = b_pos > 0.0 ? panel_count(plan, b_pos) : 0;
    int total = k_neg + k_pos;
    if (total > kMaxPanels) total = kMaxPanels;
    /* scratch: the ends of every clipped panel in the variable of its rule (quadrature.cuh: panel_map) */
    for (int p = lane; p < total; p += nlanes) {
        const bool neg = p < k_neg;
        const int k = neg ? p : p - k_neg;
        double lo = panel_break(plan, k), up = panel_break(plan, k + 1);
        const double a = neg ? a_neg : 0.0, b = neg ? b_neg : b_pos;
        const bool lo_cl = lo < a, up_cl = up > b;
        if (lo_cl) lo = a;
        if (up_cl) up = b;
        if (!(up > lo)) up = lo;
        else panel_map_ends(plan, panel_map(plan, k), k, lo, up, lo_cl, up_cl, lo, up);
        scratch[2 * p] = lo;
        scratch[2 * p + 1] = up;
    }
    COFFE_SYNC_LANES();
    double acc = 0.0;
    if (nlanes >= n_panel) {
        const int ppi = nlanes / n_panel;           /* panels in flight per sweep */
        const int lp = lane / n_panel, j = lane - lp * n_panel;
        if (lp < ppi) {
            const double xj = panel_x[j], wj = panel_w[j];
            for (int p = lp; p < total; p += ppi) {
                const double s0 = scratch[2 * p], s1 = scratch[2 * p + 1];
                if (!(s1 > s0)) continue;
                const bool neg = p < k_neg;
                double d, wgt;
                panel_node_ends(plan, panel_map(plan, neg ? p : p - k_neg), s0, s1, xj, wj, d, wgt);
                acc += wgt * f(neg ? center - d : center + d);
            }
        }
    }
    else {
        for (int p = 0; p < total; ++p) {
            const double s0 = scratch[2 * p], s1 = scratch[2 * p + 1];
            if (!(s1 > s0)) continue;
            const bool neg = p < k_neg;
            const int mapped = panel_map(plan, neg ? p : p - k_neg);
            for (int jj = lane; jj < n_panel; jj += nlanes) {
                double d, wgt;
                panel_node_ends(plan, mapped, s0, s1, panel_x[jj], panel_w[jj], d, wgt);
                acc += wgt * f(neg ? center - d : center + d);
            }
        }
    }
    COFFE_SYNC_LANES();
    return acc;
}

/* number of integrand evaluations ridge_integrate performs (empty clipped panels included: they cost
   the index arithmetic but no evaluation, so they are not counted) */
COFFE_HD unsigned int ridge_eval_count(const double hi, const double center, const double eps, const Rules &q, const int n_panel)
{
    const PanelPlan plan = panel_plan(eps, q.eps_min, q.panel_w_mpc, q.panel_du, q.tail_ratio, q.rpow, 1, q.geo_L, q.tail_lnr);
    const double b_neg = center, a_neg = center > hi ? center - hi : 0.0, b_pos = hi - center;
    const int k_neg = panel_count(plan, b_neg);
    const int k_pos = b_pos > 0.0 ? panel_count(plan, b_pos) : 0;
    unsigned int n = 0;
    for (int p = 0; p < k_neg + k_pos; ++p) {
        const bool neg = p < k_neg;
        const int k = neg ? p : p - k_neg;
        double lo = panel_break(plan, k), up = panel_break(plan, k + 1);
        const double a = neg ? a_neg : 0.0, b = neg ? b_neg : b_pos;
        if (lo < a) lo = a;
        if (up > b) up = b;
        if (up - lo > 0.0) n += n_panel;
    }
    return n;
}

template <class TA, bool MID = false>
struct SIFunctor {
    const TA *t; const SIHalf *h; const Geometry *g; const double2 *los; int los_n; double los_inv_h;
    COFFE_HD double operator()(const double lambda) const
    {
        return si_half_eval<MID>(*t, *h, *g, lambda, los_eval(los, los_n, los_inv_h, lambda));
    }
};

/* one half of the local x integrated piece, per unit x (the reference integrates x in [0,1], src/signal.c:584-606) */
template <class TA, bool MID = false>
COFFE_HD double si_half_partial(const TA &t, const LosView &lv, const Rules &q, const Geometry &g, const SIHalf &h,
                                const double2 *los, const int lane, const int nlanes, double *scratch)
{
    SIFunctor<TA, MID> f{&t, &h, &g, los, lv.n, lv.inv_h};
    const double center = h.cs * g.c;
    const double eps = h.cs * sqrt(g.s2);
    return ridge_integrate(f, h.co, center, eps, q, q.n_panel_si, lane, nlanes, scratch) / h.co;
}

COFFE_HD unsigned int si_eval_count(const Rules &q, const Geometry &g, const SIHalf &h)
{
    return ridge_eval_count(h.co, h.cs * g.c, h.cs * sqrt(g.s2), q, q.n_panel_si);
}

/* ---- one half on ONE GROUP of n_panel_si lanes (the layout of di_fast.cuh: di_inner_group) ---------------------------------------
   Lane j owns node j of every panel and walks through the units of the half's line-of-sight integral in sequence (quadrature.cuh:
   RidgeUnits).  R^2 = d^2 + cs^2 sin^2(theta) depends on |d| = |lambda - cs cos(theta)| only, so the two nodes of a mirror pair
   share one lookup of the nine I^n_l here too; the brackets, polynomials in lambda (physics.cuh: si_half_body), are evaluated per
   side.  A warp runs 32 / n_panel_si halves side by side (engine.cu: si_items_smem_kernel).
   renormalization(lambda, chi_source), the smooth half of the renormalised potential terms, is integrated apart on two
   Gauss-Legendre panels split at lambda = chi_source (the reasoning of di_fast.cuh: di_ren_smooth_partial): no column of the
   renormalisation grid per half in shared memory, which is what lets several halves share a warp. */
template <class LP>
COFFE_HD double si_ren_smooth_partial(const Renorm2DView &r2d, const LP los, const int los_n, const double los_inv_h, const Rules &q,
                                      const SIHalf &h, const int j, const int nlanes)
{
    const int half = q.n_ren / 2;
    const double *x = q.gl_x + rule_offset(half), *w = q.gl_w + rule_offset(half);
    const double split = h.cs < h.co ? h.cs : h.co;
    double acc = 0.0;
    for (int i = j; i < 2 * half; i += nlanes) {
        const bool upper = i >= half;
        const int jj = upper ? i - half : i;
        const double lo = upper ? split : 0.0, wd = upper ? h.co - split : split;
        const double l = fma(wd, x[jj], lo);
        const double4 ls = los_eval(los, los_n, los_inv_h, l);
        acc = fma(wd * w[jj], renorm2d_eval(r2d, l, h.cs) * ls.x * fma(h.g5_ren, ls.y, h.g4_ren), acc);
    }
    return -acc;
}

COFFE_HD PanelPlan si_plan(const Rules &q, const Geometry &g, const SIHalf &h)
{
    /* impact parameter cs sin(theta) and its inverse from one reciprocal square root (as di_plan) */
    const double e2 = fmax(h.cs * h.cs * g.s2, q.eps_min * q.eps_min), ie = COFFE_RSQRT(e2);
    return panel_plan(e2 * ie, q.eps_min, q.panel_w_mpc, q.panel_du, q.tail_ratio, q.rpow, 1, q.geo_L, q.tail_lnr, ie, q.inv_w, q.inv_geo_L);
}

template <class TA, class LP>
COFFE_HD double si_half_group(const TA &t, const Renorm2DView &r2d, const LosView &lv, const Rules &q, const Geometry &g, const SIHalf &h,
                              const LP los, const int j)
{
    const int np = q.n_panel_si;
    const double xj = q.panel_x_si ? q.panel_x_si[j] : q.gl_x[rule_offset(np) + j], wj = q.panel_w_si ? q.panel_w_si[j] : q.gl_w[rule_offset(np) + j];
    const double center = h.cs * g.c, eps2 = h.cs * h.cs * g.s2;
    const PanelPlan plan = si_plan(q, g, h);
    RidgeUnits u;
    ridge_units_init(u, plan, center, h.co, xj, wj, plan.lnr > 0.0 ? (q.panel_p_si ? q.panel_p_si[j] : exp(plan.lnr * xj)) : 0.0);
    double acc = 0.0;
    for (int iu = 0; iu < u.n_units; ++iu) {
        double d, wgt;
        const int sides = ridge_unit_node(u, iu, d, wgt);
        if (sides) {
            const double rr = fma(d, d, eps2), iR = rr > 0.0 ? COFFE_RSQRT(rr) : 0.0;
            double I[N_ITAB9];
            t.ifam9(rr * iR, I, h.need_ren != 0);
            double v = 0.0;
#pragma unroll 1
            for (int sd = 0; sd < 2; ++sd)
                if (sides & (1 << sd)) {
                    const double l = sd == 0 ? center - d : center + d;
                    v += si_half_body<false, true>(t, h, g, l, los_eval(los, lv.n, lv.inv_h, l), rr, iR, I);
                }
            acc = fma(wgt, v, acc);
        }
    }
    if (h.need_ren) acc += si_ren_smooth_partial(r2d, los, lv.n, lv.inv_h, q, h, j, np);
    return acc / h.co;
}

COFFE_HD unsigned int si_group_eval_count(const Rules &q, const Geometry &g, const SIHalf &h)
{
    return ridge_units_count(si_plan(q, g, h), h.cs * g.c, h.co) * q.n_panel_si;
}

template <class TA>
struct DIFunctor {
    const TA *t; const DIConst *k; const Geometry *g; const double2 *los; int los_n; double los_inv_h;
    double l1; double4 los1;
    COFFE_HD double operator()(const double l2) const
    {
        return di_eval(*t, *k, *g, l1, l2, los1, los_eval(los, los_n, los_inv_h, l2));
    }
};

/* integrated x integrated piece per unit (x1, x2): outer nodes strided over `ngroups` groups of
   `nlanes` lanes (warps of a CTA on the device) */
template <class TA>
COFFE_HD double di_partial(const TA &t, const LosView &lv, const Rules &q, const Geometry &g, const DIConst &k, const double2 *los,
                           const int group, const int ngroups, const int lane, const int nlanes, double *scratch)
{
    double acc = 0.0;
    const double sth = sqrt(g.s2);
    const int n_outer = q.n_outer > 0 ? q.n_outer : auto_n_outer(g.sep, q.n_outer_min, q.n_outer_step);
    const double *outer_x = q.gl_x + rule_offset(n_outer), *outer_w = q.gl_w + rule_offset(n_outer);
    for (int io = group; io < n_outer; io += ngroups) {
        const double l1 = g.chi1 * outer_x[io];
        DIFunctor<TA> f{&t, &k, &g, los, lv.n, lv.inv_h, l1, los_eval(los, lv.n, lv.inv_h, l1)};
        acc += outer_w[io] * ridge_integrate(f, g.chi2, l1 * g.c, l1 * sth, q, q.n_panel_di, lane, nlanes, scratch);
    }
    return acc / g.chi2;
}

COFFE_HD unsigned int di_eval_count(const Rules &q, const Geometry &g)
{
    const double sth = sqrt(g.s2);
    const int n_outer = q.n_outer > 0 ? q.n_outer : auto_n_outer(g.sep, q.n_outer_min, q.n_outer_step);
    const double *outer_x = q.gl_x + rule_offset(n_outer);
    unsigned int n = 0;
    for (int io = 0; io < n_outer; ++io) {
        const double l1 = g.chi1 * outer_x[io];
        n += ridge_eval_count(g.chi2, l1 * g.c, l1 * sth, q, q.n_panel_di);
    }
    return n;
}

}  // namespace coffe
