// TEST INFRASTRUCTURE ONLY -- CPU debugging harness.
//
// Compiles the hot path's host/device headers (coffe_b200/csrc/*.cuh) with plain g++ so that the
// physics transcription and the fixed-node quadrature can be checked against the oracle in the
// GPU-less build container.  It is never linked into libcoffe_b200.so and nothing in the product
// imports it; the `-m gpu` parity tests exercise the real sm_100a kernels through the C ABI.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../coffe_b200/csrc/assemble.h"
#include "../../coffe_b200/csrc/compact_build.h"
#include "../../coffe_b200/csrc/di_fast.cuh"
#include "../../coffe_b200/csrc/rules_build.h"

using namespace coffe;

struct HostSim {
    Arena arena;
    Layout layout;
    Tables t;
    Options o;
    QuadConfig qc;
    std::vector<double> bx, bw, rpow, rpow_di;
    std::vector<LosRec> los_rec;    /* the line-of-sight table as the shared-memory kernels stage it (tables.cuh) */
    Rules q;
    std::string err;
    CompactBuild cb;
    int variant = 0;        /* 0 = exact tables, 1 = compact (shared-memory) tables */
    int legacy_di = 0;      /* 1 = di_eval on generic ridge panels instead of di_fast.cuh */
    int di_lanes = 1;       /* lanes emulated by the di_fast path */
    bool di_half = false;   /* tables_di_symmetric: the integrated x integrated mu rule covers mu > 0 only */
    bool pops_identical = false, mag_constant = false;
    bool si_half() const { return pops_identical && (mag_constant || !(o.d1 && o.len)); }
    CompactAccess compact() const { return CompactAccess{&t, CompactView{reinterpret_cast<const unsigned char *>(cb.blob.data()), cb.h}}; }
};

static void set_rules(HostSim *s)
{
    if (s->bx.empty()) build_rule_bank(s->bx, s->bw);
    Rules &q = s->q;
    s->rpow = build_rpow(s->qc.tail_ratio);
    q.gl_x = s->bx.data(); q.gl_w = s->bw.data(); q.rpow = s->rpow.data();
    q.n_panel_si = s->qc.n_panel_si; q.n_panel_di = s->qc.n_panel_di; q.n_outer = s->qc.n_outer; q.n_outer_min = s->qc.n_outer_min; q.n_outer_step = s->qc.n_outer_step;
    q.panel_w_mpc = s->qc.panel_w; q.panel_du = s->qc.panel_du; q.eps_min = s->qc.eps_min; q.tail_ratio = s->qc.tail_ratio;
    q.panel_w_di = s->qc.panel_w_di; q.panel_du_di = s->qc.panel_du_di;
    q.panel_x = q.panel_w = q.panel_p = nullptr;
    q.panel_x_si = q.panel_w_si = q.panel_p_si = nullptr;
    s->rpow_di = build_rpow(s->qc.tail_ratio_di);
    q.rpow_di = s->rpow_di.data(); q.tail_ratio_di = s->qc.tail_ratio_di; q.geo_shift_di = s->qc.geo_shift_di;
    q.n_ren = s->qc.n_ren;
    q.tail_lnr = s->qc.tail_log ? std::log(s->qc.tail_ratio) : 0.0;
    q.tail_lnr_di = s->qc.tail_log_di ? std::log(s->qc.tail_ratio_di) : 0.0;
    q.geo_L = s->qc.geo_growth > 1.0 ? std::log(s->qc.geo_growth) : 0.0;
    q.geo_L_di = s->qc.geo_growth_di > 1.0 ? std::log(s->qc.geo_growth_di) : 0.0;
    rules_reciprocals(q);
}

extern "C" {

void *hs_create(const coffe_b200_tables *T, const coffe_b200_options *opt)
{
    HostSim *s = new HostSim();
    if (!assemble(*T, s->arena, s->layout, s->err)) { fprintf(stderr, "hostsim: %s\n", s->err.c_str()); delete s; return nullptr; }
    s->t = make_view(s->layout, s->arena.buf.data());
    s->di_half = tables_di_symmetric(*T);
    s->pops_identical = tables_populations_identical(*T);
    s->mag_constant = tables_magnification_constant(*T);
    /* line-of-sight table */
    double2 *los = reinterpret_cast<double2 *>(s->arena.buf.data() + s->layout.los);
    for (int i = 0; i < s->layout.los_n; ++i) los_knot(s->t, i, s->layout.los_h, s->layout.chi_max_table, los + (size_t)kLosStride * i);
    s->los_rec.resize(s->layout.los_n);
    for (int i = 0; i + 1 < s->layout.los_n; ++i) los_rec_pack(los + (size_t)kLosStride * i, los + (size_t)kLosStride * (i + 1), s->los_rec[i]);
    Options &o = s->o;
    o.den = opt->den; o.rsd = opt->rsd; o.len = opt->len; o.d1 = opt->d1; o.d2 = opt->d2;
    o.g1 = opt->g1; o.g2 = opt->g2; o.g3 = opt->g3; o.g4 = opt->g4; o.g5 = opt->g5;
    o.only_cross = opt->only_cross_correlations;
    o.flatsky_local = opt->flatsky_local; o.flatsky_local_nonlocal = opt->flatsky_local_nonlocal; o.flatsky_nonlocal = opt->flatsky_nonlocal;
    o.midpoint = opt->midpoint;
    o.divergent = s->layout.has_divergent;
    o.out_multipoles = 1;
    o.Om = T->omega_matter; o.H0h = T->h / 2997.92458;
    s->qc = default_quadrature();
    set_rules(s);
    return s;
}

void hs_destroy(void *h) { delete (HostSim *)h; }

/* the ridge rule of quadrature.cuh (RidgeUnits: the unit walk both integration kernels use) on a closed-form test integrand, without a
   context: integral over lambda in [0, hi] of (1 + lambda / hi) / (d^2 + eps^2)^p, d = lambda - center, with n_panel nodes per panel
   and the given plan; p = 0 checks that the units tile [0, hi] exactly once */
double hs_ridge_rule(double center, double hi, double eps, double eps_min, double W, double Du, double ratio, double growth, int tail_log,
                     int n_panel, double p, double *n_evals)
{
    const std::vector<double> rp = build_rpow(ratio);
    std::vector<double> x, w;
    build::gauss_legendre_01(n_panel, x, w);
    const PanelPlan plan = panel_plan(eps, eps_min, W, Du, ratio, rp.data(), 1, growth > 1.0 ? std::log(growth) : 0.0, tail_log ? std::log(ratio) : 0.0);
    double acc = 0.0, n = 0.0;
    for (int j = 0; j < n_panel; ++j) {
        RidgeUnits u;
        ridge_units_init(u, plan, center, hi, x[j], w[j], tail_log ? std::exp(std::log(ratio) * x[j]) : 0.0);
        for (int iu = 0; iu < u.n_units; ++iu) {
            double d, wgt;
            const int sides = ridge_unit_node(u, iu, d, wgt);
            for (int sd = 0; sd < 2; ++sd)
                if (sides & (1 << sd)) {
                    const double lam = sd == 0 ? center - d : center + d;
                    acc += wgt * (1.0 + lam / hi) * std::pow(d * d + eps * eps, -p);
                    n += 1.0;
                }
        }
    }
    if (n_evals) *n_evals = n;
    return acc;
}

/* panel plan of quadrature.cuh without a context: out = {J, M, B_J, B_(J+M), panel_count(b), panel_count_by_break(b),
   J and B_J by repeated doubling (the definition the closed form must reproduce)} */
void hs_panel_plan(double eps, double eps_min, double W, double Du, double ratio, int shift, double b, double *out)
{
    const std::vector<double> rp = build_rpow(ratio);
    const PanelPlan p = panel_plan(eps, eps_min, W, Du, ratio, rp.data(), shift);
    out[0] = p.J; out[1] = p.M; out[2] = p.BJ; out[3] = p.BJM;
    out[4] = panel_count(p, b); out[5] = panel_count_by_break(p, b);
    double e = eps < eps_min ? eps_min : eps;
    int J = 0; double bj = 0.0;
    if (e < W) { J = 1; bj = e; while (bj < W) { bj = ldexp(bj, shift); ++J; } }
    out[6] = J; out[7] = bj;
}

/* the two closed forms the panel plan evaluates on the reciprocal square root (quadrature.cuh) */
double hs_asinh_pos(double x) { return asinh_pos(x); }
double hs_fourth_root_ratio(double up, double lo) { return fourth_root_ratio(up, lo); }

/* the line-of-sight table at lambda through the Hermite knots (out[0..3]) and through the staged per-interval records (out[4..7]) */
void hs_los_both(void *h, double lambda, double *out)
{
    HostSim *s = (HostSim *)h;
    const double4 a = los_eval(s->t.los.v, s->t.los.n, s->t.los.inv_h, lambda);
    const double4 b = los_eval((const LosRec *)s->los_rec.data(), s->t.los.n, s->t.los.inv_h, lambda);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}

/* a slope as the packed knots store it (compact.cuh) */
double hs_slope_rounded(double v) { return slope_rounded(v); }

/* builds the compact tables (knot budget, tolerance) and selects them; returns the number of knots (< 0: failed).
   info[0] = achieved max relative error, info[1] = blob bytes, info[2] = the error of I^4_0, info[3..] = subdivisions m per region (up to n_info-2) */
int hs_use_compact(void *h, int knot_budget, double tol, double *info, int n_info)
{
    HostSim *s = (HostSim *)h;
    if (knot_budget == 0) { s->variant = 0; return 0; }
    if (knot_budget < 0) { knot_budget = kCompactKnotBudget; tol = kCompactTolerance; }      /* the library's defaults */
    if (!build_compact(s->t, s->layout, knot_budget, tol, s->cb, s->err)) { fprintf(stderr, "hostsim: %s\n", s->err.c_str()); return -1; }
    s->variant = 1;
    if (info && n_info >= 2) {
        info[0] = s->cb.h.max_rel_error; info[1] = s->cb.h.bytes;
        if (n_info >= 3) info[2] = s->cb.h.max_rel_error_i40;
        for (int i = 0; i + 3 < n_info && i < (int)s->cb.region_m.size(); ++i) info[i + 3] = s->cb.region_m[i];
    }
    return s->cb.h.n_knots;
}

/* the eight I^n_l and I^4_0 at r through the selected tables: out[0..7], out[8] */
void hs_tables_at(void *h, double r, double *out)
{
    HostSim *s = (HostSim *)h;
    if (s->variant) { const CompactAccess ca = s->compact(); ca.ifam_all(r, out); out[8] = s->layout.has_divergent ? ca.i40(r) : 0.0; }
    else { const ExactAccess ea{&s->t}; ea.ifam_all(r, out); out[8] = s->layout.has_divergent ? ea.i40(r) : 0.0; }
}

void hs_set_di_mode(void *h, int legacy, int lanes)
{
    HostSim *s = (HostSim *)h;
    s->legacy_di = legacy; s->di_lanes = lanes > 0 ? lanes : 1;
}

void hs_set_quadrature(void *h, const coffe_b200_quadrature *u)
{
    HostSim *s = (HostSim *)h;
    s->qc = merge_quadrature(default_quadrature(), *u);
    set_rules(s);
}

double hs_ni(void *h, double z, double mu, double sep)
{
    HostSim *s = (HostSim *)h;
    const SepConst sc = sep_const(s->o, s->t, z, sep);
    const Geometry g = geometry(sc.chi_mean, sep, mu);
    const EndVals e1 = end_eval(s->o, s->t, g.chi1, 1), e2 = end_eval(s->o, s->t, g.chi2, 2);
    return ni_eval(s->o, s->t, sc, g, e1, e2);
}

/* both halves at the same x, as functions_single_integrated evaluates them; exact LOS (no resampling)
   when exact != 0 */
double hs_si(void *h, double z, double mu, double sep, double x, int exact)
{
    HostSim *s = (HostSim *)h;
    const double chi_mean = bg_chi(s->t.bg, z);
    const Geometry g = geometry(chi_mean, sep, mu);
    const EndVals e1 = end_eval(s->o, s->t, g.chi1, 1), e2 = end_eval(s->o, s->t, g.chi2, 2);
    const SIHalf ha = si_half_setup(s->o, g, e1, e2, true), hb = si_half_setup(s->o, g, e2, e1, false);
    const double la = g.chi2 * x, lb = g.chi1 * x;
    const double4 losa = exact ? los_point(s->t, la) : los_eval(s->t.los.v, s->t.los.n, s->t.los.inv_h, la);
    const double4 losb = exact ? los_point(s->t, lb) : los_eval(s->t.los.v, s->t.los.n, s->t.los.inv_h, lb);
    if (s->variant) { const CompactAccess ca = s->compact(); return si_half_eval(ca, ha, g, la, losa) + si_half_eval(ca, hb, g, lb, losb); }
    const ExactAccess ea{&s->t};
    if (s->o.midpoint) return si_half_eval<true>(ea, ha, g, la, losa) + si_half_eval<true>(ea, hb, g, lb, losb);
    return si_half_eval(ea, ha, g, la, losa) + si_half_eval(ea, hb, g, lb, losb);
}

double hs_di(void *h, double z, double mu, double sep, double x1, double x2, int exact)
{
    HostSim *s = (HostSim *)h;
    const double chi_mean = bg_chi(s->t.bg, z);
    const Geometry g = geometry(chi_mean, sep, mu);
    const EndVals e1 = end_eval(s->o, s->t, g.chi1, 1), e2 = end_eval(s->o, s->t, g.chi2, 2);
    const DIConst k = di_setup(s->o, e1, e2);
    const double l1 = g.chi1 * x1, l2 = g.chi2 * x2;
    const double4 los1 = exact ? los_point(s->t, l1) : los_eval(s->t.los.v, s->t.los.n, s->t.los.inv_h, l1);
    const double4 los2 = exact ? los_point(s->t, l2) : los_eval(s->t.los.v, s->t.los.n, s->t.los.inv_h, l2);
    if (s->variant) return di_eval(s->compact(), k, g, l1, l2, los1, los2);
    return di_eval(ExactAccess{&s->t}, k, g, l1, l2, los1, los2);
}

/* the three pieces at one (z, sep, mu) with the fixed-node rules: out[0..2] = NI, SI, DI (pieces not in `mask`
   = 1 | 2 | 4 are left at zero) */
static void item_pieces(void *h, double z, double sep, double mu, int mask, double *out)
{
    HostSim *s = (HostSim *)h;
    const Options &o = s->o;
    const SepConst sc = sep_const(o, s->t, z, sep);
    const Geometry g = geometry(sc.chi_mean, sep, mu);
    const EndVals e1 = end_eval(o, s->t, g.chi1, 1), e2 = end_eval(o, s->t, g.chi2, 2);
    out[0] = out[1] = out[2] = 0.0;
    if (opt_has_ni(o) && (mask & 1)) out[0] = ni_eval(o, s->t, sc, g, e1, e2);
    if (opt_has_si(o) && (mask & 2)) {
        const SIHalf ha = si_half_setup(o, g, e1, e2, true), hb = si_half_setup(o, g, e2, e1, false);
        double scratch[2 * kMaxPanels];
        /* the kernel's lane decomposition: the n_panel_si lanes of one group in turn; the midpoint mode keeps the plain ridge rule */
        double acc = 0.0;
        if (o.midpoint) {
            const ExactAccess ea{&s->t};
            acc = si_half_partial<ExactAccess, true>(ea, s->t.los, s->q, g, ha, s->t.los.v, 0, 1, scratch) + si_half_partial<ExactAccess, true>(ea, s->t.los, s->q, g, hb, s->t.los.v, 0, 1, scratch);
        }
        else
            for (int j = 0; j < s->q.n_panel_si; ++j) {
                if (s->variant) { const CompactAccess ca = s->compact(); const LosRec *lr = s->los_rec.data(); acc += si_half_group(ca, s->t.ren2d, s->t.los, s->q, g, ha, lr, j) + si_half_group(ca, s->t.ren2d, s->t.los, s->q, g, hb, lr, j); }
                else { const ExactAccess ea{&s->t}; acc += si_half_group(ea, s->t.ren2d, s->t.los, s->q, g, ha, s->t.los.v, j) + si_half_group(ea, s->t.ren2d, s->t.los, s->q, g, hb, s->t.los.v, j); }
            }
        out[1] = acc;
    }
    if (opt_has_integrated(o) && (mask & 4)) {
        const DIConst k = di_setup(o, e1, e2);
        if (s->legacy_di) {
            /* the straightforward evaluator (di_eval on the same panels, no pairing): cross-check of di_fast.cuh */
            double scratch[2 * kMaxPanels];
            Rules ql = s->q;
            ql.panel_w_mpc = s->q.panel_w_di; ql.panel_du = s->q.panel_du_di; ql.tail_ratio = s->q.tail_ratio_di; ql.rpow = s->q.rpow_di;
            if (s->variant) out[2] = di_partial(s->compact(), s->t.los, ql, g, k, s->t.los.v, 0, 1, 0, 1, scratch);
            else out[2] = di_partial(ExactAccess{&s->t}, s->t.los, ql, g, k, s->t.los.v, 0, 1, 0, 1, scratch);
        }
        else {
            /* the kernel's lane decomposition: the n_panel lanes of one group in turn */
            double acc = 0.0;
            for (int j = 0; j < s->q.n_panel_di; ++j) {
                if (s->variant) acc += di_fast_partial(s->compact(), s->t, (const LosRec *)s->los_rec.data(), s->q, g, k, 0, 1, j);
                else acc += di_fast_partial(ExactAccess{&s->t}, s->t, s->t.los.v, s->q, g, k, 0, 1, j);
            }
            out[2] = acc;
        }
    }
}

void hs_item(void *h, double z, double sep, double mu, double *out) { item_pieces(h, z, sep, mu, 7, out); }

/* multipoles of the three pieces: out[piece*n_l + il]; same rules as the device path (the local x integrated and
   integrated x integrated pieces each on their own mu rule) */
void hs_multipoles(void *h, double z, double sep, const int *ls, int n_l, double *out)
{
    HostSim *s = (HostSim *)h;
    for (int p = 1; p < 3; ++p) {
        std::vector<double> mu, w;
        build_mu_rule(s->qc, sep, mu, w, p, p == 2 ? s->di_half : s->si_half());
        const int nm = (int)mu.size();
        std::vector<double> v((size_t)3 * nm);
#pragma omp parallel for schedule(dynamic, 1)
        for (int im = 0; im < nm; ++im) item_pieces(h, z, sep, mu[im], p == 1 ? 2 : 4, &v[(size_t)3 * im]);
        for (int il = 0; il < n_l; ++il) {
            double acc = 0.0;
            if (!((p == 2 ? s->di_half : s->si_half()) && (ls[il] & 1)))
                for (int im = 0; im < nm; ++im) acc += w[im] * legendre_p(ls[il], mu[im]) * v[(size_t)3 * im + p];
            out[p * n_l + il] = (2 * ls[il] + 1) * acc;
        }
    }
    const int nn = s->qc.n_mu_ni;
    const double *x = s->bx.data() + rule_offset(nn), *wn = s->bw.data() + rule_offset(nn);
    for (int il = 0; il < n_l; ++il) {
        double acc = 0.0;
        if (opt_has_ni(s->o))
            for (int j = 0; j < nn; ++j) acc += wn[j] * legendre_p(ls[il], 2 * x[j] - 1) * hs_ni(h, z, 2 * x[j] - 1, sep);
        out[il] = (2 * ls[il] + 1) * acc;
    }
}

/* integrand evaluations of the fixed-node rules at one (z, sep): out[0] = SI, out[1] = DI, summed over the mu items */
void hs_eval_counts(void *h, double z, double sep, double *out)
{
    HostSim *s = (HostSim *)h;
    std::vector<double> mu, w;
    const double chi_mean = bg_chi(s->t.bg, z);
    out[0] = out[1] = 0.0;
    build_mu_rule(s->qc, sep, mu, w, 1, s->si_half());
    for (size_t im = 0; im < mu.size(); ++im) {
        const Geometry g = geometry(chi_mean, sep, mu[im]);
        const EndVals e1 = end_eval(s->o, s->t, g.chi1, 1), e2 = end_eval(s->o, s->t, g.chi2, 2);
        const SIHalf ha = si_half_setup(s->o, g, e1, e2, true), hb = si_half_setup(s->o, g, e2, e1, false);
        out[0] += si_group_eval_count(s->q, g, ha) + si_group_eval_count(s->q, g, hb);
    }
    build_mu_rule(s->qc, sep, mu, w, 2, s->di_half);
    for (size_t im = 0; im < mu.size(); ++im) {
        const Geometry g = geometry(chi_mean, sep, mu[im]);
        const EndVals e1 = end_eval(s->o, s->t, g.chi1, 1), e2 = end_eval(s->o, s->t, g.chi2, 2);
        out[1] += di_fast_eval_count(s->q, g, di_setup(s->o, e1, e2));
    }
}

/* mu rule of one piece (1 = local x integrated, 2 = integrated x integrated) for one separation under the quadrature configuration in
   force: returns the node count, fills mu / w (capacity cap) */
int hs_mu_rule(void *h, double sep, int piece, int half_only, int cap, double *mu_out, double *w_out)
{
    HostSim *s = (HostSim *)h;
    std::vector<double> mu, w;
    build_mu_rule(s->qc, sep, mu, w, piece, half_only != 0);
    const int n = (int)mu.size();
    for (int i = 0; i < n && i < cap; ++i) { mu_out[i] = mu[i]; w_out[i] = w[i]; }
    return n;
}

int hs_n_mu(void *h, double sep)
{
    HostSim *s = (HostSim *)h;
    std::vector<double> mu, w;
    build_mu_rule(s->qc, sep, mu, w);
    return (int)mu.size();
}

/* flat-sky closed forms of one (z, sep, l): out[0] = density-lensing multipole, out[1] = lensing-lensing multipole
   (same arithmetic as flatsky_multipoles_kernel, n_x Gauss-Legendre nodes in x) */
void hs_flatsky(void *h, double z, double sep, int l, int n_x, double *out)
{
    HostSim *s = (HostSim *)h;
    const SepConst sc = sep_const(s->o, s->t, z, sep);
    out[0] = flatsky_density_lensing_multipole(s->o, s->t, sc, sep, l);
    out[1] = 0.0;
    const SplineView *sv = find_extra(s->t, 1, l, 0);
    if (!sv) { out[1] = nan(""); return; }
    const double *x = s->bx.data() + rule_offset(n_x), *w = s->bw.data() + rule_offset(n_x);
    double acc = 0.0;
    for (int j = 0; j < n_x; ++j) acc += w[j] * flatsky_lensing_lensing_integrand(s->o, s->t, sc, *sv, sep, l, x[j]);
    out[1] = 2 * 3.14159265358979323846 * 3.14159265358979323846 * acc;
}

}  // extern "C"
