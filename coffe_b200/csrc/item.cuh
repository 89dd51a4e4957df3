// Per-work-item integration: one (z_mean, separation, mu) point -> the three pieces the reference's
// dispatcher returns (coffe_integrate, src/signal.c:472-771): local x local (closed form),
// local x integrated (1-D line-of-sight integral per half) and integrated x integrated (2-D).
//
// Written as lane-strided partial sums: a caller with `nlanes` cooperating lanes passes its lane id and
// reduces the returned partials (warp shuffle on the device; nlanes = 1 in the CPU debugging harness),
// so the device kernels and the harness execute the same arithmetic on the same nodes.
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
    const double *rpow_di;           /* tail_ratio_di^n, n < kMaxTail */
    double tail_ratio_di;
    int geo_shift_di;
    int n_ren;                       /* nodes of the inner rule of renormalization(l1, l2) (di_fast.cuh: di_ren_smooth_partial), even */
    double tail_lnr, tail_lnr_di;    /* log of the tail ratio when the tail panels carry their nodes in log(d) (quadrature.cuh); 0 = in d */
    double geo_L, geo_L_di;          /* log of the growth factor of the sinh-map geometric panels (quadrature.cuh); 0 = doubling panels */
};

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

#if defined(__CUDA_ARCH__)
/* The same integral on a full warp with the panel bookkeeping held by the warp (the scheme of di_fast.cuh: di_inner_partial_warp):
   lane k computes the break B_k once, the panel counts of the two sides are two ballots, a panel's ends two shuffles.  With ~3.5
   sweeps per lane and half, the scalar bookkeeping of ridge_integrate (panel plan, two counting loops, panel list through shared
   memory, two barriers) cost as many instructions as the evaluations themselves (ncu r2: 713 thread instructions per evaluation
   against 342 in the integrated x integrated kernel).  Same breaks, nodes and weights; the panel count is the by-break definition
   (quadrature.cuh: panel_count_by_break), so a break within rounding of the end of a side may add a panel a few ulp wide.
   Valid when a side never has more than 63 panels and the rule is staged (q.panel_x_si): the host checks the configured plan. */
template <class TA>
__device__ __forceinline__ double si_half_partial_warp(const TA &t, const LosView &lv, const Rules &q, const Geometry &g, const SIHalf &h,
                                                       const double2 *los, const int lane)
{
    const unsigned int full = 0xffffffffu;
    const int np = q.n_panel_si;            /* <= 32 (the caller checks) */
    const int ppi = 32 / np;
    const int lp = lane / np, j = lane - lp * np;
    const bool on = lp < ppi;
    const double center = h.cs * g.c, eps = h.cs * sqrt(g.s2), hi = h.co;
    const PanelPlan plan = panel_plan(eps, q.eps_min, q.panel_w_mpc, q.panel_du, q.tail_ratio, q.rpow, 1, q.geo_L, q.tail_lnr);
    /* breaks B_0 .. B_63 in two registers per lane (the default plan needs up to 36 per side: 18 geometric from the eps floor,
       3 uniform, 15 tail panels at ratio 1.5) */
    const double brk0 = panel_break(plan, lane), brk1 = panel_break(plan, lane + 32);
    const double b_neg = center, a_neg = center > hi ? center - hi : 0.0, b_pos = hi - center;
    const int k_neg = b_neg > 0.0 ? __popc(__ballot_sync(full, brk0 < b_neg)) + __popc(__ballot_sync(full, brk1 < b_neg)) : 0;
    const int k_pos = b_pos > 0.0 ? __popc(__ballot_sync(full, brk0 < b_pos)) + __popc(__ballot_sync(full, brk1 < b_pos)) : 0;
    const double xj = q.panel_x_si[j], wj = q.panel_w_si[j];
    const int n_all = k_neg + k_pos;
    double acc = 0.0;
    for (int e0 = 0; e0 < n_all; e0 += ppi) {
        const int e = e0 + lp;
        const bool neg = e < k_neg;
        const int k = neg ? e : e - k_neg;
        const double lo0 = __shfl_sync(full, brk0, k & 31), lo1 = __shfl_sync(full, brk1, k & 31);
        const double up0 = __shfl_sync(full, brk0, (k + 1) & 31), up1 = __shfl_sync(full, brk1, (k + 1) & 31);
        double lo = k < 32 ? lo0 : lo1, up = k + 1 < 32 ? up0 : up1;
        const double a = neg ? a_neg : 0.0, b = neg ? b_neg : b_pos;
        const bool lo_cl = lo < a, up_cl = up > b;
        if (lo_cl) lo = a;
        if (up_cl) up = b;
        if (on && e < n_all && up > lo) {
            double d, wgt;
            panel_node(plan, k, lo, up, lo_cl, up_cl, xj, wj, d, wgt);
            const double lambda = neg ? center - d : center + d;
            acc = fma(wgt, si_half_eval(t, h, g, lambda, los_eval(los, lv.n, lv.inv_h, lambda)), acc);
        }
    }
    return acc / h.co;
}
#endif

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
