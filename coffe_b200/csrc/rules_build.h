// Host-side construction of the quadrature rule arrays described by QuadConfig (quadrature.cuh).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/coffe_b200.h"
#include "quadrature.cuh"
#include "table_build.h"

namespace coffe {

/* |mu| at which the reference switches the rsd-len bracket (src/functions.c:2199) */
static const double kMuSwitch = 0.999;

inline QuadConfig default_quadrature()
{
    QuadConfig q;
    q.n_mu_ni = 32;
    q.n_mu_panel = 8; q.mu_rho0 = 4.0; q.mu_panel_w = 32.0; q.mu_du = 256.0;
    q.n_outer = 0; q.n_outer_min = 30; q.n_outer_step = 12.0;
    q.n_panel_si = 8; q.n_panel_di = 8;
    q.n_z = 8;
    q.panel_w = 64.0; q.panel_du = 128.0; q.eps_min = 1e-3; q.tail_ratio = 1.5;
    q.n_mu_panel_di = 5; q.panel_w_di = 64.0; q.panel_du_di = 192.0; q.mu_panel_w_di = 64.0; q.tail_ratio_di = 6.0; q.geo_shift_di = 1;
    q.geo_growth = 4.0; q.geo_growth_di = 8.0;
    q.n_ren = 24;
    q.mu_growth = 3.0; q.mu_growth_di = 1.0; q.mu_switch_di = 1; q.mu_tail = 1.5;
    q.tail_log = 0; q.tail_log_di = 1;
    return q;
}

/* Defaults when the 1e-5 gate is measured against the TOTAL a call returns and the local terms (density, RSD) are on: the two
   line-of-sight pieces are then at most a few per cent of every multipole of the benchmarked grid, and rules sized for 1e-5 of
   each piece spend 30 % of the step on digits nobody can see.  Chosen on GPU sweeps over 2 000 (z, r) points of cfg5 against the
   doubled rules (profiles/r2_quad_sweep_{a,b,c}.log): 6 instead of 8 nodes per inner panel of the integrated x integrated piece
   (5 panels x 6 nodes fill 30 lanes), tail growth 3 instead of 2.5, 4 instead of 5 mu nodes per r_perp panel; 64 instead of 32 Mpc
   mu panels, 7 instead of 8 mu nodes and 6 instead of 8 inner nodes for the local x integrated piece, first r_perp break of the mu rules
   at 8 instead of 4 Mpc: worst total error 2.9e-6 of the l = 0..4 scale (piece rules: 6.5e-7; profiles/r2_quad_sweep_e.log), 0.49 /
   0.47 of the evaluations.  Guarded by the dense tests of tests/test_gpu_parity.py. */
inline QuadConfig total_target_quadrature()
{
    QuadConfig q = default_quadrature();
    q.n_panel_di = 6; q.n_mu_panel_di = 4; q.n_ren = 6;
    q.n_outer_min = 15; q.n_outer_step = 15.0;
    q.panel_w_di = 96.0;
    q.mu_growth = 8.0; q.mu_switch_di = 0;
    q.tail_log = 1; q.tail_ratio = 3.0;
    q.mu_panel_w = 64.0; q.n_mu_panel = 7; q.n_panel_si = 6;
    q.mu_rho0 = 8.0;
    return q;
}

inline QuadConfig merge_quadrature(const QuadConfig &base, const coffe_b200_quadrature &u)
{
    QuadConfig q = base;
    if (u.n_mu_ni > 0) q.n_mu_ni = u.n_mu_ni;
    if (u.n_mu_panel > 0) q.n_mu_panel = u.n_mu_panel;
    if (u.mu_rho0 > 0) q.mu_rho0 = u.mu_rho0;
    if (u.mu_panel_width > 0) q.mu_panel_w = u.mu_panel_width;
    if (u.mu_uniform_extent > 0) q.mu_du = u.mu_uniform_extent;
    if (u.n_outer > 0) q.n_outer = u.n_outer;
    if (u.n_outer < 0) q.n_outer = 0;
    if (u.mu_tail_ratio > 1.0) q.mu_tail = u.mu_tail_ratio;
    if (u.n_outer_min > 0) q.n_outer_min = u.n_outer_min;
    if (u.n_outer_step > 0) q.n_outer_step = u.n_outer_step;
    if (u.n_panel_si > 0) q.n_panel_si = u.n_panel_si;
    if (u.n_panel_di > 0) q.n_panel_di = u.n_panel_di;
    if (u.n_z > 0) q.n_z = u.n_z;
    if (u.panel_width > 0) q.panel_w = u.panel_width;
    if (u.panel_uniform_extent > 0) q.panel_du = u.panel_uniform_extent;
    if (u.eps_min > 0) q.eps_min = u.eps_min;
    if (u.tail_ratio > 1.0) q.tail_ratio = u.tail_ratio;
    if (u.n_mu_panel_di > 0) q.n_mu_panel_di = u.n_mu_panel_di;
    if (u.panel_width_di > 0) q.panel_w_di = u.panel_width_di;
    if (u.panel_uniform_extent_di > 0) q.panel_du_di = u.panel_uniform_extent_di;
    if (u.mu_panel_width_di > 0) q.mu_panel_w_di = u.mu_panel_width_di;
    if (u.tail_ratio_di > 1.0) q.tail_ratio_di = u.tail_ratio_di;
    if (u.geometric_shift_di > 0) q.geo_shift_di = u.geometric_shift_di;
    if (u.geometric_growth > 0) q.geo_growth = u.geometric_growth;
    if (u.geometric_growth_di > 0) q.geo_growth_di = u.geometric_growth_di;
    if (u.n_renorm > 0) q.n_ren = u.n_renorm & ~1;
    if (u.tail_map > 0) q.tail_log = u.tail_map == 2;
    if (u.tail_map_di > 0) q.tail_log_di = u.tail_map_di == 2;
    if (u.mu_geometric_growth > 0) q.mu_growth = u.mu_geometric_growth;
    if (u.mu_geometric_growth_di > 0) q.mu_growth_di = u.mu_geometric_growth_di;
    if (u.mu_switch_break_di > 0) q.mu_switch_di = u.mu_switch_break_di == 1;
    return q;
}

/* upper bound on the geometric panels of one side: doubling panels (x 2^shift), or the sinh map with `growth` per panel */
inline int geometric_panels(const double eps_min, const double W, const int shift, const double growth)
{
    if (!(eps_min < W)) return 0;
    if (growth > 1.0) return (int)std::ceil(std::asinh(W / eps_min) / std::log(growth)) + 1;
    return (int)std::ceil(std::log2(W / eps_min) / shift) + 2;
}

/* upper bound on the panels of one line-of-sight integral (both sides of the ridge) */
inline int max_panels(const QuadConfig &q, const double extent = 2.0e4)
{
    const int geo = geometric_panels(q.eps_min, q.panel_w, 1, q.geo_growth);
    const int uni = (int)std::ceil(q.panel_du / q.panel_w) + 1;
    const double start = std::max(q.panel_du, q.panel_w);
    const int tail = extent > start ? (int)std::ceil(std::log(extent / start) / std::log(q.tail_ratio)) + 2 : 0;
    if (tail >= kMaxTail) return 1 << 30;
    return 2 * (geo + uni + tail);
}

/* upper bound on the panels of ONE side of the ridge of the integrated x integrated plan (its own knobs) */
inline int di_side_panels(const QuadConfig &q, const double extent = 2.0e4)
{
    const int geo = geometric_panels(q.eps_min, q.panel_w_di, q.geo_shift_di, q.geo_growth_di);
    const int uni = (int)std::ceil(q.panel_du_di / q.panel_w_di) + 1;
    const double start = std::max(q.panel_du_di, q.panel_w_di);
    const int tail = extent > start ? (int)std::ceil(std::log(extent / start) / std::log(q.tail_ratio_di)) + 2 : 0;
    if (tail >= kMaxTail) return 1 << 30;
    return geo + uni + tail;
}

/* the same bound for the local x integrated plan */
inline int si_side_panels(const QuadConfig &q, const double extent = 2.0e4)
{
    const int geo = geometric_panels(q.eps_min, q.panel_w, 1, q.geo_growth);
    const int uni = (int)std::ceil(q.panel_du / q.panel_w) + 1;
    const double start = std::max(q.panel_du, q.panel_w);
    const int tail = extent > start ? (int)std::ceil(std::log(extent / start) / std::log(q.tail_ratio)) + 2 : 0;
    if (tail >= kMaxTail) return 1 << 30;
    return geo + uni + tail;
}

inline bool quadrature_valid(const QuadConfig &q, const int panel_capacity)
{
    if (!(q.panel_w > 0 && q.eps_min > 0 && q.tail_ratio > 1.0 && q.panel_du >= 0)) return false;
    if (max_panels(q) > panel_capacity) return false;
    return q.n_mu_ni >= 2 && q.n_mu_ni <= kMaxRule && q.n_mu_panel >= 2 && q.n_mu_panel <= kMaxRule &&
           q.n_outer >= 0 && q.n_outer <= kMaxRule && q.n_panel_si >= 2 && q.n_panel_si <= 32 &&
           q.n_panel_di >= 2 && q.n_panel_di <= 32 &&        /* a panel's nodes are the lanes of one group of a warp */ q.n_z >= 1 && q.n_z <= kMaxRule &&
           q.panel_w > 0 && q.panel_du >= 0 && q.eps_min > 0 && q.tail_ratio > 1.0 &&
           q.mu_rho0 > 0 && q.mu_panel_w > 0 && q.mu_du >= 0 &&
           q.n_mu_panel_di >= 2 && q.n_mu_panel_di <= kMaxRule && q.panel_w_di > 0 && q.panel_du_di >= 0 && q.mu_panel_w_di > 0 && q.tail_ratio_di > 1.0 &&
           q.geo_shift_di >= 1 && q.geo_shift_di <= 3 && q.geo_growth >= 0 && q.geo_growth_di >= 0 && q.mu_tail > 1.0 && q.n_ren >= 2 && q.n_ren <= 2 * kMaxRule;
}

/* bank of Gauss-Legendre rules on [0,1], rule n at offset n(n-1)/2 */
inline void build_rule_bank(std::vector<double> &bx, std::vector<double> &bw)
{
    bx.assign(rule_offset(kMaxRule + 1), 0.0);
    bw.assign(rule_offset(kMaxRule + 1), 0.0);
    std::vector<double> x, w;
    for (int n = 1; n <= kMaxRule; ++n) {
        build::gauss_legendre_01(n, x, w);
        std::copy(x.begin(), x.end(), bx.begin() + rule_offset(n));
        std::copy(w.begin(), w.end(), bw.begin() + rule_offset(n));
    }
}

inline std::vector<double> build_rpow(const double ratio)
{
    std::vector<double> v(kMaxTail);
    for (int n = 0; n < kMaxTail; ++n) v[n] = std::pow(ratio, (double)n);
    return v;
}

/* mu rule of the integrated pieces for one separation: the integrands are functions of the transverse
   separation r_perp = r sin(phi), mu = cos(phi), that peak sharply at r_perp -> 0 (|mu| -> 1) and carry
   BAO-scale structure in r_perp.  Panels in phi with break points at r_perp = B_k (the same geometric /
   uniform / geometric plan as the line-of-sight integrals) plus, for the local x integrated piece, one at |mu| = 0.999
   where the reference switches formulas; n-point Gauss-Legendre per panel; mirrored to mu < 0.  Weights are for the measure
   dmu/2, i.e. xi_l = (2l+1) sum_j w_j P_l(mu_j) xi(mu_j)      (src/signal.c:97-110, 541).
   piece: 1 = local x integrated, 2 = integrated x integrated (its own node count, panel width, growth and switch break).
   With a growth factor > 1 the geometric region [rho0, panel width] is cut into equal panels in log(r_perp) with the nodes
   Gauss-Legendre in log(r_perp): the small-r_perp behaviour of the integrated pieces is logarithmic. */
inline void build_mu_rule(const QuadConfig &q, const double sep, std::vector<double> &mu, std::vector<double> &w, const int piece = 1,
                          const bool half_only = false)
{
    const bool di = piece == 2;
    const int n_mu_panel = di ? q.n_mu_panel_di : q.n_mu_panel;
    const double mu_panel_w = di ? q.mu_panel_w_di : q.mu_panel_w;
    const double growth = di ? q.mu_growth_di : q.mu_growth;
    const bool switch_break = di ? q.mu_switch_di != 0 : true;
    mu.clear(); w.clear();
    std::vector<double> bp;
    bp.push_back(0.0);
    double b = q.mu_rho0;
    double log_lo = 0.0, log_hi = 0.0;      /* region of the log-mapped panels */
    const double w_log = std::min(mu_panel_w, 0.5 * sep);      /* in log(r_perp) only where phi is still far from pi/2 */
    if (growth > 1.0 && b < w_log) {
        const int J = std::max(1, (int)std::ceil(std::log(w_log / b) / std::log(growth) - 1e-12));
        const double ratio = std::pow(w_log / b, 1.0 / J);
        log_lo = b;
        for (int k = 0; k < J; ++k) { bp.push_back(b); b = k + 1 == J ? w_log : b * ratio; }
        log_hi = w_log;
    }
    while (b < mu_panel_w && b < sep) { bp.push_back(b); b *= 2.0; }
    while (b < q.mu_du && b < sep) { bp.push_back(b); b += mu_panel_w; }
    while (b < sep) { bp.push_back(b); b *= q.mu_tail; }
    bp.push_back(sep);
    if (switch_break) bp.push_back(sep * std::sqrt(1.0 - kMuSwitch * kMuSwitch));
    std::sort(bp.begin(), bp.end());
    std::vector<double> x, wt;
    build::gauss_legendre_01(n_mu_panel, x, wt);
    for (size_t k = 0; k + 1 < bp.size(); ++k) {
        const double a = std::asin(std::min(1.0, bp[k] / sep)), c = std::asin(std::min(1.0, bp[k + 1] / sep));
        if (!(c - a > 1e-14)) continue;
        if (log_hi > log_lo && bp[k] >= log_lo && bp[k + 1] <= log_hi) {
            /* nodes in u = log(r_perp): phi = asin(e^u / r), dphi/du = tan(phi) */
            const double u0 = std::log(bp[k]), u1 = std::log(bp[k + 1]);
            for (int j = 0; j < n_mu_panel; ++j) {
                const double sphi = std::exp(u0 + (u1 - u0) * x[j]) / sep, cphi = std::sqrt(1.0 - sphi * sphi);
                mu.push_back(cphi);
                w.push_back((u1 - u0) * wt[j] * (sphi / cphi) * sphi / 2.0);
            }
            continue;
        }
        for (int j = 0; j < n_mu_panel; ++j) {
            const double phi = a + (c - a) * x[j];
            mu.push_back(std::cos(phi));
            w.push_back((c - a) * wt[j] * std::sin(phi) / 2.0);
        }
    }
    const size_t n = mu.size();
    if (half_only) {
        /* xi(mu) = xi(-mu): the mu < 0 half is the mirror image, its weight moves onto the mu > 0 nodes (even multipoles) */
        for (size_t i = 0; i < n; ++i) w[i] *= 2.0;
        return;
    }
    for (size_t i = 0; i < n; ++i) { mu.push_back(-mu[i]); w.push_back(w[i]); }
}

}  // namespace coffe
