// Host-side assembly of the uploaded knots (coffe_b200_tables, include/coffe_b200.h) into one flat
// arena of doubles plus a `Tables` view onto it.  The same arena is either copied to HBM (engine.cu)
// or viewed in place by the CPU debugging harness (tests/hostsim).
#pragma once
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/coffe_b200.h"
#include "table_build.h"
#include "tables.cuh"

namespace coffe {

struct Arena {
    std::vector<double> buf;
    /* offsets are multiples of 4 doubles so every double4 stays 32-byte aligned */
    size_t alloc(const size_t n)
    {
        const size_t off = buf.size();
        buf.resize(off + ((n + 3) / 4) * 4, 0.0);
        return off;
    }
    size_t put(const double *p, const size_t n)
    {
        const size_t off = alloc(n);
        if (n) std::memcpy(buf.data() + off, p, n * sizeof(double));
        return off;
    }
    size_t put(const std::vector<build::Coef> &k) { return put(reinterpret_cast<const double *>(k.data()), 4 * k.size()); }
};

struct SplineOff { size_t x = 0, k = 0; int n = 0; };
struct HintOff { size_t hint = 0; int n_hint = 0; double log_lo = 0, inv_dlog = 0; };

struct Layout {
    size_t bg_z = 0; int bg_n = 0; double bg_inv_dz = 0;
    size_t bg_a = 0, bg_Hc = 0, bg_D1 = 0, bg_f = 0, bg_G1 = 0, bg_G2 = 0, bg_chi = 0;
    SplineOff z_of_chi, bias[8], i40, ren0, extra[kMaxExtra];
    int extra_n[kMaxExtra] = {0}, extra_l[kMaxExtra] = {0}, extra_half[kMaxExtra] = {0}, n_extra = 0;
    HintOff i40_hint;
    size_t ifam_x = 0, ifam_k = 0; int ifam_n = 0, ifam_n_small = 0; double ifam_r0 = 1, ifam_inv_dln = 0;
    size_t ren2d_x = 0, ren2d_node = 0; int ren2d_n = 0; double ren2d_x0 = 0, ren2d_inv_dx = 0, ren2d_dx = 0;
    size_t los = 0; int los_n = 0; double los_h = 0;
    int has_divergent = 0, has_ifam = 0, interp_kind = 5;
    size_t mid_z = 0, mid_node[N_ITAB] = {0}; int mid_nz = 0, mid_nr = 0; double mid_z0 = 0, mid_inv_dz = 0;     /* 2-D I^n_l(z, r) tables */
    double chi_max_table = 0;     /* largest chi of the background table */
    double r_max_table = 0;       /* largest r of the I^n_l family */
};

inline SplineView view_of(const SplineOff &s, const double *base)
{
    SplineView v;
    v.x = base + s.x;
    v.k = reinterpret_cast<const double4 *>(base + s.k);
    v.n = s.n;
    return v;
}

inline Tables make_view(const Layout &L, const double *base)
{
    Tables t;
    std::memset(&t, 0, sizeof(t));
    t.bg.z = base + L.bg_z; t.bg.n = L.bg_n; t.bg.inv_dz = L.bg_inv_dz;
    t.bg.a = reinterpret_cast<const double4 *>(base + L.bg_a);
    t.bg.Hc = reinterpret_cast<const double4 *>(base + L.bg_Hc);
    t.bg.D1 = reinterpret_cast<const double4 *>(base + L.bg_D1);
    t.bg.f = reinterpret_cast<const double4 *>(base + L.bg_f);
    t.bg.G1 = reinterpret_cast<const double4 *>(base + L.bg_G1);
    t.bg.G2 = reinterpret_cast<const double4 *>(base + L.bg_G2);
    t.bg.chi = reinterpret_cast<const double4 *>(base + L.bg_chi);
    t.bg.z_of_chi = view_of(L.z_of_chi, base);
    for (int i = 0; i < 8; ++i) t.bias[i] = view_of(L.bias[i], base);
    t.ifam.x = base + L.ifam_x;
    t.ifam.k = reinterpret_cast<const double4 *>(base + L.ifam_k);
    t.ifam.n = L.ifam_n; t.ifam.n_small = L.ifam_n_small; t.ifam.r0 = L.ifam_r0; t.ifam.inv_dln = L.ifam_inv_dln;
    {
        const SplineView v = view_of(L.i40, base);
        t.i40.x = v.x; t.i40.k = v.k; t.i40.n = v.n;
        t.i40.hint = reinterpret_cast<const int *>(base + L.i40_hint.hint);
        t.i40.n_hint = L.i40_hint.n_hint; t.i40.log_lo = L.i40_hint.log_lo; t.i40.inv_dlog = L.i40_hint.inv_dlog;
    }
    t.ren0 = view_of(L.ren0, base);
    t.ren2d.x = base + L.ren2d_x;
    t.ren2d.node = reinterpret_cast<const double4 *>(base + L.ren2d_node);
    t.ren2d.n = L.ren2d_n; t.ren2d.x0 = L.ren2d_x0; t.ren2d.inv_dx = L.ren2d_inv_dx; t.ren2d.dx = L.ren2d_dx;
    t.los.v = reinterpret_cast<const double2 *>(base + L.los);
    t.los.n = L.los_n; t.los.h = L.los_h; t.los.inv_h = L.los_h > 0 ? 1.0 / L.los_h : 0.0;
    t.n_extra = L.n_extra;
    for (int i = 0; i < L.n_extra; ++i) {
        t.extra[i].n = L.extra_n[i]; t.extra[i].l = L.extra_l[i]; t.extra[i].half = L.extra_half[i];
        t.extra[i].s = view_of(L.extra[i], base);
    }
    t.has_divergent = L.has_divergent;
    t.mid.nz = L.mid_nz; t.mid.nr = L.mid_nr; t.mid.z0 = L.mid_z0; t.mid.inv_dz = L.mid_inv_dz;
    t.mid.z = base + L.mid_z;
    for (int k = 0; k < N_ITAB; ++k) t.mid.node[k] = reinterpret_cast<const double4 *>(base + L.mid_node[k]);
    return t;
}

/* The per-interval cubic coefficients of every uploaded spline (~25 tables of 10^4 knots: 45 ms when built one after the other, most
   of a table upload) are independent of each other: assemble() builds them all up front on a few host threads and the code below
   picks them up by the address of their ordinates. */
struct CoefCache {
    std::map<const double *, std::vector<build::Coef>> done;
    std::vector<std::pair<const double *, std::pair<const double *, int>>> jobs;      /* y -> (x, n) */
    void want(const coffe_b200_spline &s) { if (s.x && s.y && s.n >= 2 && !done.count(s.y)) { done[s.y]; jobs.push_back({s.y, {s.x, s.n}}); } }
    void run(const int kind)
    {
        const unsigned hw = std::thread::hardware_concurrency();
        const int n_threads = (int)std::max(1u, std::min<unsigned>(hw ? hw : 4u, std::min<size_t>(jobs.size(), 16)));
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([this, t, n_threads, kind]() {
                for (size_t j = t; j < jobs.size(); j += n_threads) {
                    std::vector<build::Coef> k;
                    const double *x = jobs[j].second.first;
                    bool ok = true;
                    for (int i = 0; i + 1 < jobs[j].second.second && ok; ++i) ok = x[i + 1] > x[i];
                    if (ok && build::spline_coefs(kind, x, jobs[j].first, jobs[j].second.second, k)) done.find(jobs[j].first)->second = std::move(k);
                }
            });
        for (auto &t : th) t.join();
    }
    /* false: not precomputed (or the table was rejected): the caller builds it itself and reports the error */
    bool take(const double *y, std::vector<build::Coef> &k)
    {
        auto it = done.find(y);
        if (it == done.end() || it->second.empty()) return false;
        k = std::move(it->second);
        done.erase(it);
        return true;
    }
};

inline CoefCache *&active_coef_cache()
{
    static thread_local CoefCache *c = nullptr;
    return c;
}

inline bool cached_spline_coefs(const int kind, const double *x, const double *y, const int n, std::vector<build::Coef> &k)
{
    if (active_coef_cache() && active_coef_cache()->take(y, k)) return true;
    return build::spline_coefs(kind, x, y, n, k);
}

inline bool put_spline(Arena &A, const coffe_b200_spline &s, const int kind, SplineOff &out, std::string &err, const char *what)
{
    if (!s.x || !s.y || s.n < 2) { err = std::string("missing table: ") + what; return false; }
    for (int i = 0; i + 1 < s.n; ++i)
        if (!(s.x[i + 1] > s.x[i])) { err = std::string("knots not strictly increasing in table: ") + what; return false; }
    std::vector<build::Coef> k;
    if (!cached_spline_coefs(kind, s.x, s.y, s.n, k)) { err = std::string("unsupported interpolation type or too few knots for table: ") + what; return false; }
    out.x = A.put(s.x, s.n);
    out.k = A.put(k);
    out.n = s.n;
    return true;
}

static const int kFamilyN[N_ITAB] = {0, 0, 0, 1, 1, 2, 2, 3};
static const int kFamilyL[N_ITAB] = {0, 2, 4, 1, 3, 0, 2, 1};

inline const coffe_b200_integral *find_integral(const coffe_b200_tables &T, const int n, const int l, const int half)
{
    for (int i = 0; i < T.n_integrals; ++i)
        if (T.integrals[i].n == n && T.integrals[i].l == l && (T.integrals[i].half_integer != 0) == (half != 0)) return &T.integrals[i];
    return nullptr;
}

/* The integrated x integrated integrand is invariant under the exchange of the two galaxies, f(mu; x1, x2) = f(-mu; x2, x1)
   (checked against the reference's functions_double_integrated), whenever the two populations carry the same magnification
   and evolution bias -- the only population-dependent inputs of that piece are s_i(z) and G_i(z).  Then xi_DI(mu) = xi_DI(-mu):
   the multipole sums need the mu > 0 nodes only and the odd multipoles of the piece vanish. */
inline bool same_knots(const coffe_b200_spline &a, const coffe_b200_spline &b)
{
    if (!a.x || !b.x || a.n != b.n) return false;
    for (int i = 0; i < a.n; ++i) {
        const bool nan_a = a.y[i] != a.y[i], nan_b = b.y[i] != b.y[i];
        if (a.x[i] != b.x[i] || (nan_a != nan_b) || (!nan_a && a.y[i] != b.y[i])) return false;
    }
    return true;
}

inline bool tables_di_symmetric(const coffe_b200_tables &T)
{
    return same_knots(T.bias[COFFE_B200_BIAS_MAGNIFICATION1], T.bias[COFFE_B200_BIAS_MAGNIFICATION2]) &&
           same_knots(T.background[COFFE_B200_BG_G1], T.background[COFFE_B200_BG_G2]);
}

/* The local x integrated piece is the sum of two halves with the roles of the galaxies exchanged, so it has the same mirror
   symmetry when the two populations are identical in every bias -- except for one reference quirk: the second half of the
   d1 x lensing term takes (2 - 5 s) at its own end instead of the other one (src/functions.c:2776-2779, SURVEY.md 8a note 1),
   which only cancels when the magnification bias does not depend on redshift.  (Checked numerically: identical populations with
   s(z) = 0.1 + 0.05 z break the symmetry of the piece at the 6e-4 level, constant s keeps it to the last bit.) */
inline bool tables_populations_identical(const coffe_b200_tables &T)
{
    return tables_di_symmetric(T) &&
           same_knots(T.bias[COFFE_B200_BIAS_GALAXY1], T.bias[COFFE_B200_BIAS_GALAXY2]) &&
           same_knots(T.bias[COFFE_B200_BIAS_EVOLUTION1], T.bias[COFFE_B200_BIAS_EVOLUTION2]);
}

inline bool tables_magnification_constant(const coffe_b200_tables &T)
{
    const coffe_b200_spline &s = T.bias[COFFE_B200_BIAS_MAGNIFICATION1];
    if (!s.y || s.n < 1) return false;
    for (int i = 1; i < s.n; ++i)
        if (s.y[i] != s.y[0]) return false;
    return true;
}

/* number of uniformly spaced lambda samples of the line-of-sight table */
static const int kLosSamples = 512;

inline bool assemble_serial(const coffe_b200_tables &T, Arena &A, Layout &L, std::string &err);

inline bool assemble(const coffe_b200_tables &T, Arena &A, Layout &L, std::string &err)
{
    /* every spline's coefficients up front, in parallel; assemble_serial then only validates and lays out */
    CoefCache cache;
    for (int i = 0; i < COFFE_B200_BG_COUNT; ++i) cache.want(T.background[i]);
    for (int i = 0; i < COFFE_B200_BIAS_COUNT; ++i) cache.want(T.bias[i]);
    for (int i = 0; i < T.n_integrals; ++i) cache.want(T.integrals[i].result);
    if (T.n_renorm > 0) cache.want(T.renorm_zero_separation);
    cache.run(T.interp_method);
    active_coef_cache() = &cache;
    const bool ok = assemble_serial(T, A, L, err);
    active_coef_cache() = nullptr;
    return ok;
}

inline bool assemble_serial(const coffe_b200_tables &T, Arena &A, Layout &L, std::string &err)
{
    const int kind = T.interp_method;
    L.interp_kind = kind;
    /* ---- background ---- */
    const coffe_b200_spline &chi = T.background[COFFE_B200_BG_COMOVING_DISTANCE];
    if (!chi.x || chi.n < 5) { err = "missing background tables"; return false; }
    const int n = chi.n;
    L.bg_n = n;
    L.bg_z = A.put(chi.x, n);
    L.bg_inv_dz = (n - 1) / (chi.x[n - 1] - chi.x[0]);
    if (chi.x[0] != 0.0) { err = "background z grid must start at 0"; return false; }
    struct { int idx; size_t *off; const char *name; } fz[] = {
        {COFFE_B200_BG_A, &L.bg_a, "a"}, {COFFE_B200_BG_CONFORMAL_HZ, &L.bg_Hc, "conformal_Hz"},
        {COFFE_B200_BG_D1, &L.bg_D1, "D1"}, {COFFE_B200_BG_F, &L.bg_f, "f"}, {COFFE_B200_BG_G1, &L.bg_G1, "G1"},
        {COFFE_B200_BG_G2, &L.bg_G2, "G2"}, {COFFE_B200_BG_COMOVING_DISTANCE, &L.bg_chi, "comoving_distance"}};
    for (auto &e : fz) {
        const coffe_b200_spline &s = T.background[e.idx];
        if (!s.x || !s.y || s.n != n) { err = std::string("background table missing or of different size: ") + e.name; return false; }
        for (int i = 0; i < n; ++i)
            if (s.x[i] != chi.x[i]) { err = std::string("background tables must share their z knots: ") + e.name; return false; }
        std::vector<build::Coef> k;
        if (!cached_spline_coefs(kind, s.x, s.y, n, k)) { err = "unsupported interpolation type for the background"; return false; }
        *e.off = A.put(k);
    }
    if (!put_spline(A, T.background[COFFE_B200_BG_Z_AS_CHI], kind, L.z_of_chi, err, "z_as_chi")) return false;
    L.chi_max_table = T.background[COFFE_B200_BG_Z_AS_CHI].x[T.background[COFFE_B200_BG_Z_AS_CHI].n - 1];
    static const char *bias_names[8] = {"galaxy_bias1", "galaxy_bias2", "galaxy_bias3", "galaxy_bias4",
                                        "magnification_bias1", "magnification_bias2", "evolution_bias1", "evolution_bias2"};
    for (int i = 0; i < 8; ++i) {
        /* galaxy_bias3/4 only matter for the covariance; tolerate their absence */
        if ((i == 2 || i == 3) && !T.bias[i].x) { L.bias[i] = SplineOff(); continue; }
        if (!put_spline(A, T.bias[i], kind, L.bias[i], err, bias_names[i])) return false;
    }
    /* ---- I^n_l family on shared knots (absent when only the covariance is wanted) ---- */
    L.has_ifam = T.n_integrals > 0;
    if (L.has_ifam) {
    const coffe_b200_integral *fam[N_ITAB];
    for (int t = 0; t < N_ITAB; ++t) {
        fam[t] = find_integral(T, kFamilyN[t], kFamilyL[t], 0);
        if (!fam[t]) { err = "missing I^n_l table (need (0,0),(0,2),(0,4),(1,1),(1,3),(2,0),(2,2),(3,1))"; return false; }
        if (fam[t]->result.n != fam[0]->result.n) { err = "I^n_l tables must share their knots"; return false; }
    }
    const int m = fam[0]->result.n;
    const double *fx = fam[0]->result.x;
    for (int t = 1; t < N_ITAB; ++t)
        for (int i = 0; i < m; ++i)
            if (fam[t]->result.x[i] != fx[i]) { err = "I^n_l tables must share their knots"; return false; }
    {
        std::vector<build::Coef> inter((size_t)(m - 1) * N_ITAB);
        for (int t = 0; t < N_ITAB; ++t) {
            std::vector<build::Coef> k;
            if (!cached_spline_coefs(kind, fx, fam[t]->result.y, m, k)) { err = "unsupported interpolation type for I^n_l"; return false; }
            for (int i = 0; i < m - 1; ++i) inter[(size_t)i * N_ITAB + t] = k[i];
        }
        L.ifam_x = A.put(fx, m);
        L.ifam_k = A.put(inter);
        L.ifam_n = m;
        L.r_max_table = fx[m - 1];
        /* longest geometric suffix: x[i] = r0 * q^(i - n_small) */
        int ns = m - 2;
        const double lq = std::log(fx[m - 1] / fx[m - 2]);
        while (ns > 0 && fx[ns - 1] > 0.0 && std::fabs(std::log(fx[ns] / fx[ns - 1]) - lq) < 1e-7 * lq) --ns;
        if (m - ns < 16) { ns = m - 1; L.ifam_r0 = fx[m - 1]; L.ifam_inv_dln = 0.0; }   /* not log-spaced: always search */
        else { L.ifam_r0 = fx[ns]; L.ifam_inv_dln = (m - 1 - ns) / std::log(fx[m - 1] / fx[ns]); }
        L.ifam_n_small = ns;
    }
    }
    /* ---- 2-D tables of the midpoint approximation: all eight family members or none ---- */
    L.mid_nz = 0;
    if (L.has_ifam) {
        const coffe_b200_integral *fam2[N_ITAB];
        int have = 0;
        for (int t = 0; t < N_ITAB; ++t) { fam2[t] = find_integral(T, kFamilyN[t], kFamilyL[t], 0); have += fam2[t]->n_z2d > 0 && fam2[t]->z2d && fam2[t]->values2d; }
        if (have != 0 && have != N_ITAB) { err = "the 2-D tables I^n_l(z, r) must be given for all eight family members or for none"; return false; }
        if (have == N_ITAB) {
            const int nz = fam2[0]->n_z2d, nr = fam2[0]->result.n;
            if (nz < 4) { err = "2-D tables need at least four redshift knots"; return false; }
            for (int t = 0; t < N_ITAB; ++t) {
                if (fam2[t]->n_z2d != nz) { err = "2-D tables must share their redshift knots"; return false; }
                for (int i = 0; i < nz; ++i)
                    if (fam2[t]->z2d[i] != fam2[0]->z2d[i] || (i > 0 && !(fam2[t]->z2d[i] > fam2[t]->z2d[i - 1]))) { err = "2-D tables must share strictly increasing redshift knots"; return false; }
            }
            L.mid_z = A.put(fam2[0]->z2d, nz);
            L.mid_nz = nz; L.mid_nr = nr;
            L.mid_z0 = fam2[0]->z2d[0];
            L.mid_inv_dz = (nz - 1) / (fam2[0]->z2d[nz - 1] - fam2[0]->z2d[0]);
            /* bicubic node data per table, built in parallel (1e6 nodes each) */
            std::vector<std::vector<build::Coef>> nodes(N_ITAB);
            std::vector<std::thread> th;
            const double *fx = fam2[0]->result.x;
            for (int t = 0; t < N_ITAB; ++t)
                th.emplace_back([&, t]() { nodes[t] = build::bicubic_nodes(fam2[t]->z2d, fx, fam2[t]->values2d, nz, nr); });
            for (auto &x : th) x.join();
            for (int t = 0; t < N_ITAB; ++t) L.mid_node[t] = A.put(nodes[t]);
        }
    }
    /* ---- divergent extras ---- */
    L.has_divergent = 0;
    if (L.has_ifam && T.n_renorm > 0) {
        const coffe_b200_integral *i40 = find_integral(T, 4, 0, 0);
        if (!i40 || !T.renorm_chi || !T.renorm_z) { err = "renormalisation grid given without I^4_0"; return false; }
        if (!put_spline(A, i40->result, kind, L.i40, err, "I^4_0")) return false;
        {
            /* bucket table in log(r) over [x_1, x_(n-1)] (x_0 may be tiny but is positive, src/integrals.c:76) */
            const coffe_b200_spline &sp = i40->result;
            if (!(sp.x[1] > 0.0)) { err = "I^4_0 knots must be positive"; return false; }
            const int nb = 1024;
            const double lo = std::log(sp.x[1]), hi = std::log(sp.x[sp.n - 1]);
            std::vector<double> packed((nb + 1) / 2 + 1, 0.0);
            int *hint = reinterpret_cast<int *>(packed.data());
            int i = 0;
            for (int b = 0; b < nb; ++b) {
                const double edge = std::exp(lo + (hi - lo) * b / nb);
                while (i < sp.n - 2 && sp.x[i + 1] <= edge) ++i;
                hint[b] = i;
            }
            L.i40_hint.hint = A.put(packed.data(), packed.size());
            L.i40_hint.n_hint = nb; L.i40_hint.log_lo = lo; L.i40_hint.inv_dlog = nb / (hi - lo);
        }
        if (!put_spline(A, T.renorm_zero_separation, kind, L.ren0, err, "renormalization_zero_separation")) return false;
        const int nr = T.n_renorm;
        if (nr < 4) { err = "renormalisation grid too small"; return false; }
        const std::vector<build::Coef> node = build::bicubic_nodes(T.renorm_chi, T.renorm_chi, T.renorm_z, nr, nr);
        L.ren2d_x = A.put(T.renorm_chi, nr);
        L.ren2d_node = A.put(node);
        L.ren2d_n = nr;
        L.ren2d_x0 = T.renorm_chi[0];
        L.ren2d_inv_dx = (nr - 1) / (T.renorm_chi[nr - 1] - T.renorm_chi[0]);
        L.ren2d_dx = (T.renorm_chi[nr - 1] - T.renorm_chi[0]) / (nr - 1);
        for (int i = 0; i < nr; ++i)
            if (std::fabs(T.renorm_chi[i] - (L.ren2d_x0 + i * L.ren2d_dx)) > 1e-9 * L.ren2d_dx * nr) { err = "renormalisation grid must be uniform"; return false; }
        L.has_divergent = 1;
    }
    /* ---- every other uploaded table (flat-sky closed forms), kept by key ---- */
    L.n_extra = 0;
    for (int i = 0; i < T.n_integrals; ++i) {
        const coffe_b200_integral &it = T.integrals[i];
        const bool half = it.half_integer != 0;
        bool in_family = false;
        for (int t = 0; t < N_ITAB; ++t) in_family = in_family || (!half && it.n == kFamilyN[t] && it.l == kFamilyL[t]);
        if (in_family || (!half && it.n == 4 && it.l == 0)) continue;
        if (L.n_extra == kMaxExtra) { err = "too many extra I^n_l tables"; return false; }
        if (!put_spline(A, it.result, kind, L.extra[L.n_extra], err, "extra I^n_l")) return false;
        L.extra_n[L.n_extra] = it.n; L.extra_l[L.n_extra] = it.l; L.extra_half[L.n_extra] = half;
        ++L.n_extra;
    }
    /* ---- line-of-sight table (filled later by evaluating the splines above) ---- */
    L.los_n = kLosSamples;
    L.los_h = L.chi_max_table / (kLosSamples - 1);
    L.los = A.alloc((size_t)2 * kLosStride * kLosSamples);
    return true;
}

}  // namespace coffe
