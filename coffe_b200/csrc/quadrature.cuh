// Fixed-node quadrature that replaces the reference's adaptive integrators on this path
// (Cuhre, src/common.c:1074-1118; QAG-61, src/common.c:1004-1040).
//
// The line-of-sight integrands are smooth functions of the integration variable multiplied by
// I^n_l(sqrt(d^2 + eps^2)) brackets, where d is the distance along the line of sight from the point
// of closest approach ("ridge") and eps the impact parameter: a cusp of width eps at d = 0, structure
// on the BAO scale out to a few hundred Mpc and a slowly varying power-law tail beyond.  A tensor
// Gauss-Legendre grid in (x1, x2) cannot resolve the cusp (this is why Cuhre exhausts its 750 000
// evaluation budget); instead each inner integral is split at |d| = B_k,
//     B_0 = 0, B_k = eps 2^(k-1)            while the panel width is below W   (geometric, resolves the cusp)
//     then steps of W                       out to Du                           (uniform, resolves BAO-scale structure)
//     then doubling again                   out to the end of the line of sight (power-law tail)
// with an n-point Gauss-Legendre rule per panel.  On a geometric panel the nearest singularity of
// f(sqrt(d^2+eps^2)) (at d = +-i eps) is at least one panel length away, so the rule converges like
// rho^(-2n) with rho >= 3.7 independent of eps.
#pragma once
#include "hd.h"

namespace coffe {

/* powers of the tail ratio, rpow[n] = ratio^n for n < kMaxTail (built on the host with the rule bank) */
static const int kMaxTail = 48;

struct PanelPlan {
    const double *rpow;
    double eps, inv_eps, W, inv_W, ratio;
    int shift;      /* log2 of the growth factor of the geometric panels (1: doubling, 2: x4) */
    int J;          /* B_k = eps 2^(shift (k-1)) for 1 <= k <= J */
    int M;          /* B_k = B_J + (k-J) W for J < k <= J+M */
    double BJ, BJM;
    double L;       /* > 0: the sinh map.  B_k = eps sinh(k L) for k <= J with J L = asinh(W / eps), i.e. B_J = W exactly, and the
                       nodes of a geometric panel are Gauss-Legendre in t = asinh(d / eps) (panel_node) */
    double lnr;     /* > 0: log(ratio), and the nodes of a tail panel are Gauss-Legendre in log(d): the tails are power laws times
                       line-of-sight factors that are polynomials in d, entire in log(d), whereas in d a panel [b, ratio b] sees
                       the singularity of d^-p at the origin (rho = 3.7 for ratio 3, 2 for ratio 9) */
};

/* asinh(x), x >= 0, and (up / lo)^(1/4): written on the reciprocal square root, which costs the device a third of a square root
   plus a division (the plan and the clipped last panels ask for them once or twice per line of sight) */
COFFE_HD double asinh_pos(const double x)
{
    const double s = fma(x, x, 1.0);
    return log(fma(s, COFFE_RSQRT(s), x));
}
COFFE_HD double fourth_root_ratio(const double up, const double lo)
{
    const double il = COFFE_RSQRT(lo), x = up * il * il;        /* up / lo */
    const double s = x * COFFE_RSQRT(x);                        /* its square root */
    return s * COFFE_RSQRT(s);
}

COFFE_HD PanelPlan panel_plan(double eps, const double eps_min, const double W, const double Du, const double ratio, const double *rpow,
                              const int shift = 1, const double geo_L = 0.0, const double tail_lnr = 0.0, double inv_eps = 0.0,
                              double inv_W = 0.0, double inv_geo_L = 0.0 /* reciprocals the caller holds already (0: computed here) */)
{
    PanelPlan p;
    p.lnr = tail_lnr;
    p.rpow = rpow;
    p.shift = shift;
    if (eps < eps_min) { eps = eps_min; inv_eps = 0.0; }
    if (!(inv_eps > 0.0)) inv_eps = 1.0 / eps;
    if (!(inv_W > 0.0)) inv_W = 1.0 / W;
    p.eps = eps; p.inv_eps = inv_eps; p.W = W; p.inv_W = inv_W; p.ratio = ratio;
    p.L = 0.0;
    int J = 0;
    double b = 0.0;
    if (eps < W && geo_L > 0.0) {
        /* sinh map: f(sqrt(d^2 + eps^2)) = f(eps cosh t) under d = eps sinh t has its singularities at t = +- i pi/2 whatever the
           panel, so a Gauss-Legendre rule in t converges like rho^(-2n) with (rho - 1/rho) = 2 pi / L on a panel of length L in t:
           rho = 4.7 for L = ln 4.  The region [0, W] is cut into J equal panels in t, the fewest with L <= geo_L. */
        const double T = asinh_pos(W * inv_eps);
        J = (int)ceil(inv_geo_L > 0.0 ? T * inv_geo_L : T / geo_L);
        if (J < 1) J = 1;
        p.L = T / J;
        b = W;
    }
    else if (COFFE_UNLIKELY(eps < W)) {          /* doubling panels (legacy plan); otherwise the cusp is wider than a uniform panel: uniform panels from d = 0 */
        /* J - 1 = smallest m >= 1 with eps 2^(shift m) >= W, from the exponents (exact: the same J and B_J as repeated doubling) */
        const int e0 = ilogb(W) - ilogb(eps);
        const int n = ldexp(eps, e0) >= W ? e0 : e0 + 1;         /* smallest n with eps 2^n >= W; n >= 1 here */
        const int m = (n + shift - 1) / shift;
        J = m + 1; b = ldexp(eps, shift * m);
    }
    p.J = J; p.BJ = b;
    int M = 0;
    if (b < Du) M = (int)ceil((Du - b) * inv_W);
    p.M = M; p.BJM = b + M * W;
    return p;
}

/* break k < J of the geometric region (out of line: the kernels' hot loops ask for it a few times per line of sight) */
COFFE_COLD double panel_break_geometric(const double eps, const double L, const int shift, const int k)
{
    return L > 0.0 ? eps * sinh(k * L) : ldexp(eps, shift * (k - 1));
}

COFFE_HD double panel_break(const PanelPlan &p, const int k)
{
    if (k <= 0) return 0.0;
    if (k == p.J) return p.BJ;
    if (COFFE_UNLIKELY(k < p.J)) return panel_break_geometric(p.eps, p.L, p.shift, k);
    if (k <= p.J + p.M) return p.BJ + (k - p.J) * p.W;
    const int n = k - p.J - p.M;
    return p.BJM * p.rpow[n < kMaxTail ? n : kMaxTail - 1];
}

/* number of panels needed to cover [0, b] */
COFFE_HD int panel_count(const PanelPlan &p, const double b)
{
    if (b <= 0.0) return 0;
    int k = 1;
    if (COFFE_LIKELY(b > p.BJM)) {
        k = p.J + p.M;
        int n = 1;
        while (n < kMaxTail - 1 && p.BJM * p.rpow[n] < b) ++n;      /* smallest n with B_(k+n) >= b */
        return k + n;
    }
    if (b > p.BJ || p.J == 0) return p.J + (int)ceil((b - p.BJ) * p.inv_W);
    if (COFFE_LIKELY(p.L > 0.0)) {
        k = (int)ceil(asinh_pos(b * p.inv_eps) / p.L);
        return k < 1 ? 1 : k > p.J ? p.J : k;
    }
    double e = p.eps;
    while (e < b) { e = ldexp(e, p.shift); ++k; }
    return k;
}

/* the same count from the breaks themselves: smallest k >= 1 with B_k >= b (CPU tests of the closed forms); it differs from
   panel_count only when b sits within rounding of a break, where the extra panel is a few ulp wide. */
COFFE_HD int panel_count_by_break(const PanelPlan &p, const double b)
{
    if (b <= 0.0) return 0;
    const int k_max = p.J + p.M + kMaxTail - 1;
    int k = 1;
    while (k < k_max && panel_break(p, k) < b) ++k;
    return k;
}

/* Node and weight (in |d|) of abscissa x (weight w, rule on [0,1]) of panel k clipped to [lo, up]: Gauss-Legendre in d, or in
   t = asinh(d / eps) on the geometric panels of a sinh-map plan.  lo_clipped / up_clipped: that end is not the panel's own break. */
/* variable of the rule of panel k: 0 = d, 1 = t = asinh(d / eps) (geometric panels of a sinh-map plan), 2 = log(d) (tail panels
   of a log-tail plan) */
COFFE_HD int panel_map(const PanelPlan &p, const int k) { return (p.L > 0.0 && k < p.J) ? 1 : (p.lnr > 0.0 && k >= p.J + p.M) ? 2 : 0; }

/* the same from the panel's ends s0 < s1 in the variable of its rule */
COFFE_HD void panel_node_ends(const PanelPlan &p, const int map, const double s0, const double s1, const double x, const double w,
                              double &d, double &wgt)
{
    const double h = s1 - s0;
    if (map == 1) {
        const double e = exp(fma(h, x, s0)), ie = 1.0 / e;
        d = (0.5 * p.eps) * (e - ie);
        wgt = (0.5 * p.eps) * (e + ie) * (h * w);
        return;
    }
    if (map == 2) {
        d = exp(fma(h, x, s0));
        wgt = d * (h * w);
        return;
    }
    d = fma(h, x, s0);
    wgt = h * w;
}

/* the ends of panel k clipped to [lo, up] in the variable of its rule (lo > 0 on a tail panel) */
COFFE_HD void panel_map_ends(const PanelPlan &p, const int map, const int k, const double lo, const double up, const bool lo_clipped,
                             const bool up_clipped, double &s0, double &s1)
{
    if (map == 1) {
        s0 = lo_clipped ? asinh(lo / p.eps) : k * p.L;
        s1 = up_clipped ? asinh(up / p.eps) : (k + 1) * p.L;
    }
    else if (map == 2) { s0 = log(lo); s1 = log(up); }
    else { s0 = lo; s1 = up; }
}

COFFE_HD void panel_node(const PanelPlan &p, const int k, const double lo, const double up, const bool lo_clipped, const bool up_clipped,
                         const double x, const double w, double &d, double &wgt)
{
    const int map = panel_map(p, k);
    double s0, s1;
    panel_map_ends(p, map, k, lo, up, lo_clipped, up_clipped, s0, s1);
    panel_node_ends(p, map, s0, s1, x, w, d, wgt);
}

/* node of a sinh-map panel clipped to [lo, up], out of line (scalars only): returns d, *wgt = weight */
COFFE_COLD double panel_node_sinh(const double eps, const double L, const int k, const double lo, const double up, const bool lo_clipped,
                                  const bool up_clipped, const double x, const double w, double *wgt)
{
    const double s0 = lo_clipped ? asinh(lo / eps) : k * L, s1 = up_clipped ? asinh(up / eps) : (k + 1) * L;
    const double h = s1 - s0, e = exp(fma(h, x, s0)), ie = 1.0 / e;
    *wgt = (0.5 * eps) * (e + ie) * (h * w);
    return (0.5 * eps) * (e - ie);
}

/* ---- the panels of one ridge integral as a sequence of UNITS, for one lane that owns node j of every panel -------------------
   Integral over lambda in [0, hi] with the closest approach at `center` (>= 0): |d| = |lambda - center| runs over
   [a_neg, b_neg] on the near side and [0, b_pos] on the far side.  Where neither side is clipped the panels of the two sides are
   mirror images: such a MIRROR PAIR is one unit with two nodes, center - d and center + d, which share R = sqrt(d^2 + eps^2) and
   therefore one table lookup.  Order of the units (ridge_unit_node):
     1. the two last panels (near side k_neg-1, far side k_pos-1), each clipped by the end of its line of sight -- first, while
        the groups of lanes that walk different integrals side by side in a warp are still in step, so that a clipped tail panel
        with its two square roots holds nobody up;
     2. the unclipped panels k_pair .. of the longer side that have no partner;
     3. the mirror pairs k = 0 .. k_pair-1; on the sinh-map panels (S, C) = (sinh, cosh)(k L) follow from a two-term recurrence
        with (S1, C1) = (sinh, cosh)(L) and the node from the addition theorem with (sj, cj) = (sinh, cosh)(L x_j): one exp per
        lane and integral; on the log-mapped tail panels the node is the panel's start times ratio^x_j (pj).
   When the closest approach lies beyond the end of the line of sight (center > hi) only [a_neg, b_neg] of the near side exists:
   its panels in order, clipped at both ends. */
struct RidgeUnits {
    PanelPlan plan;
    double a_neg, b_neg, b_pos;
    int k_neg, k_pos, k_pair;
    bool ordinary, neg_long;
    int n_single, n_units;
    double xj, wj, pj;
    double S, C, sj, cj, S1, C1;
};

COFFE_HD void ridge_units_init(RidgeUnits &u, const PanelPlan &plan, const double center, const double hi, const double xj, const double wj,
                               const double pj /* ratio^x_j of a log-tail plan, 0 otherwise */)
{
    u.plan = plan;
    u.xj = xj; u.wj = wj; u.pj = pj;
    u.b_neg = center;
    u.a_neg = center > hi ? center - hi : 0.0;
    u.b_pos = hi - center;
    u.k_neg = panel_count(plan, u.b_neg);
    u.k_pos = u.b_pos > 0.0 ? panel_count(plan, u.b_pos) : 0;
    /* the last panel of the shorter side is clipped (or ends within rounding of the end of the line of sight): not paired */
    const int kc = u.k_neg < u.k_pos ? u.k_neg : u.k_pos;
    u.k_pair = (u.a_neg == 0.0 && kc > 0) ? kc - 1 : 0;
    u.ordinary = u.a_neg == 0.0;
    u.neg_long = u.k_neg > u.k_pos;
    const int k_long = u.neg_long ? u.k_neg : u.k_pos;
    u.n_single = u.ordinary ? 2 + (k_long - 1 > u.k_pair ? k_long - 1 - u.k_pair : 0) : u.k_neg;
    u.n_units = u.n_single + u.k_pair;
    u.S = 0.0; u.C = 1.0; u.sj = 0.0; u.cj = 1.0; u.S1 = 0.0; u.C1 = 1.0;
    if (plan.L > 0.0 && u.k_pair > 0) {
        const double e = exp(plan.L * xj), ie = 1.0 / e, e1 = exp(plan.L), ie1 = 1.0 / e1;
        u.sj = 0.5 * (e - ie); u.cj = 0.5 * (e + ie); u.S1 = 0.5 * (e1 - ie1); u.C1 = 0.5 * (e1 + ie1);
    }
}

/* unit iu = 0 .. n_units-1, to be asked for in order: |d| of this lane's node, its weight, and which sides carry it
   (bit 0: center - d, bit 1: center + d; 0: empty) */
COFFE_HD int ridge_unit_node(RidgeUnits &u, const int iu, double &d, double &wgt)
{
    const PanelPlan &plan = u.plan;
    const double xj = u.xj, wj = u.wj;
    d = 0.0; wgt = 0.0;
    if (iu >= u.n_single) {
        const int k = iu - u.n_single;
        if (k < plan.J) {
            if (COFFE_LIKELY(plan.L > 0.0)) {
                d = plan.eps * fma(u.S, u.cj, u.C * u.sj);
                wgt = (plan.eps * plan.L * wj) * fma(u.C, u.cj, u.S * u.sj);
                const double Sn = fma(u.S, u.C1, u.C * u.S1);
                u.C = fma(u.C, u.C1, u.S * u.S1);
                u.S = Sn;
            }
            else {
                const double lo = panel_break(plan, k), wd = panel_break(plan, k + 1) - lo;
                d = fma(wd, xj, lo); wgt = wd * wj;
            }
        }
        else if (k < plan.J + plan.M) { d = fma(plan.W, xj, plan.BJ + (k - plan.J) * plan.W); wgt = plan.W * wj; }
        else {
            const int n = k - plan.J - plan.M;
            const double lo = plan.BJM * plan.rpow[n < kMaxTail ? n : kMaxTail - 1];
            if (COFFE_LIKELY(plan.lnr > 0.0)) { d = lo * u.pj; wgt = d * (plan.lnr * wj); }
            else {
                const double wd = plan.BJM * plan.rpow[n + 1 < kMaxTail ? n + 1 : kMaxTail - 1] - lo;
                d = fma(wd, xj, lo); wgt = wd * wj;
            }
        }
        return 3;
    }
    const int e = iu;
    const bool neg = u.ordinary ? (e == 0 || (e >= 2 && u.neg_long)) : true;
    const int k = !u.ordinary ? e : e == 0 ? u.k_neg - 1 : e == 1 ? u.k_pos - 1 : u.k_pair + e - 2;
    if (k < 0) return 0;
    const int map = panel_map(plan, k);
    if (u.ordinary && map == 1) {
        /* sinh-map panel [k L, (k + 1) L] in t = asinh(d / eps); a last panel ends at the end of its line of sight */
        const double b = neg ? u.b_neg : u.b_pos;
        const double s0 = k * plan.L, s1 = e < 2 ? fmin((k + 1) * plan.L, asinh_pos(b * plan.inv_eps)) : (k + 1) * plan.L;
        if (!(s1 > s0)) return 0;
        const double h = s1 - s0, ex = exp(fma(h, xj, s0)), ie = 1.0 / ex;
        d = (0.5 * plan.eps) * (ex - ie); wgt = (0.5 * plan.eps) * (ex + ie) * (h * wj);
        return neg ? 1 : 2;
    }
    double lo = panel_break(plan, k), up = panel_break(plan, k + 1);
    const double a = neg ? u.a_neg : 0.0, b = neg ? u.b_neg : u.b_pos;
    const bool lo_cl = lo < a, up_cl = up > b;
    if (lo_cl) lo = a;
    if (up_cl) up = b;
    if (!(up > lo)) return 0;
    if (map == 2) {
        if (lo_cl || up_cl) {
            /* clipped tail panel [lo, up]: d = lo (1 + a x)^4 reaches up at x = 1 with a = (up/lo)^(1/4) - 1.  Like log(d) it keeps
               d^-p far from its singularity (rho = 9 at up/lo = 6) at the price of a fourth root instead of log and exp */
            const double a4 = fourth_root_ratio(up, lo) - 1.0, t = fma(a4, xj, 1.0), t2 = t * t;
            d = lo * (t2 * t2); wgt = (4.0 * a4 * lo * wj) * (t2 * t);
        }
        else { d = lo * u.pj; wgt = d * (plan.lnr * wj); }
    }
    else if (COFFE_UNLIKELY(map == 1)) d = panel_node_sinh(plan.eps, plan.L, k, lo, up, lo_cl, up_cl, xj, wj, &wgt);
    else { const double wd = up - lo; d = fma(wd, xj, lo); wgt = wd * wj; }
    return neg ? 1 : 2;
}

/* integrand evaluations the units of one ridge integral ask for, per node of the panel rule */
COFFE_HD unsigned int ridge_units_count(const PanelPlan &plan, const double center, const double hi)
{
    RidgeUnits u;
    ridge_units_init(u, plan, center, hi, 0.5, 1.0, 1.0);
    unsigned int n = 0;
    for (int iu = 0; iu < u.n_units; ++iu) {
        double d, w;
        const int sides = ridge_unit_node(u, iu, d, w);
        n += (sides & 1) + ((sides >> 1) & 1);
    }
    return n;
}

/* runtime quadrature configuration (mirrors coffe_b200_quadrature, include/coffe_b200.h) */
struct QuadConfig {
    int n_mu_ni;            /* Gauss-Legendre nodes in mu for the local x local piece */
    int n_mu_panel;         /* nodes per transverse-separation panel of the mu rule of the integrated pieces */
    double mu_rho0, mu_panel_w, mu_du;     /* panel plan in r_perp = r sqrt(1-mu^2) [Mpc] */
    int n_outer;            /* outer line-of-sight nodes of the integrated x integrated piece; 0 = by separation */
    int n_outer_min; double n_outer_step;  /* by separation: max(n_outer_min, 10 + sep / n_outer_step) */
    int n_panel_si, n_panel_di;            /* nodes per inner panel */
    int n_z;                /* nodes of the redshift average */
    double panel_w;         /* W  [Mpc] */
    double panel_du;        /* Du [Mpc] */
    double eps_min;         /* floor on eps [Mpc] */
    double tail_ratio;      /* growth factor of the tail panels beyond Du */
    /* the integrated x integrated piece is smoother (two line-of-sight integrals) and dominates the cost: it has
       its own, coarser mu rule and inner panel plan (DESIGN.md section 3.1 holds the convergence evidence) */
    int n_mu_panel_di;
    double panel_w_di, panel_du_di;
    double mu_panel_w_di;   /* width of the uniform r_perp panels of its mu rule */
    double tail_ratio_di;   /* growth factor of its tail panels */
    int geo_shift_di;       /* log2 of the growth factor of its geometric panels (1 or 2) */
    /* growth factor per geometric panel under the sinh map (panel_plan); <= 1: the doubling panels with Gauss-Legendre nodes in d */
    double geo_growth, geo_growth_di;
    int tail_log, tail_log_di;     /* tail panels with Gauss-Legendre nodes in log(d) instead of d */
    double mu_tail;         /* growth factor of the r_perp panels of the mu rules beyond mu_du */
    double mu_growth, mu_growth_di; /* growth factor of the log-mapped geometric r_perp panels of the mu rules; <= 1: doubling panels in phi */
    int mu_switch_di;       /* the integrated x integrated mu rule keeps the break at |mu| = 0.999 (only the local x integrated integrand switches there) */
    int n_ren;              /* nodes of the inner rule of renormalization(l1, l2) in the integrated x integrated piece (even) */
};

/* bank of Gauss-Legendre rules on [0,1] for n = 1..kMaxRule, rule n stored at offset n(n-1)/2 */
static const int kMaxRule = 128;
COFFE_HD int rule_offset(const int n) { return n * (n - 1) / 2; }

/* outer node count when QuadConfig::n_outer == 0: the inner integral as a function of lambda_1 carries
   the structure of I(R) stretched by 1/sin(theta) ~ chi/r_perp, so wider separations need more nodes.
   Measured on 16 (z, r) points against a rule with every count doubled (errors relative to the piece itself,
   worst multipole): r <= 224 Mpc is converged to the 2e-6 floor of the other rules with 24 nodes; r = 299 Mpc
   needs 28 (1e-6; 24 gives 1e-5), r = 448 Mpc needs 44 (2e-6; 36 gives 1e-5, 24 gives 2e-4): independent of z.
   The floor of 30 keeps the l = 4 multipole at r = 60-150 Mpc below 3e-6 (26 nodes: 6e-6). */
COFFE_HD int auto_n_outer(const double sep, const int n_min = 30, const double step = 12.0, const int n_panel = 0)
{
    int n = 10 + (int)(sep / step);
    if (n < n_min) n = n_min;
    /* the kernel runs 32 / n_panel outer nodes side by side (di_fast.cuh): a count that does not fill its last round costs the same
       as the next multiple, so take the multiple */
    if (n_panel > 0 && n_panel <= 32) { const int g = 32 / n_panel; n = (n + g - 1) / g * g; }
    if (n > 96) n = 96;
    return n;
}

}  // namespace coffe
