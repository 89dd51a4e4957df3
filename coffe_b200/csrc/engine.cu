// libcoffe_b200: sm_100a kernels and C ABI of the signal-evaluation path.
//
// Work decomposition:
//   point  = one (z, separation) pair of the output grid (for redshift-averaged multipoles: one
//            (bin, separation, z-node) triple);
//   item   = one (point, mu-node) pair.  The integrated pieces (local x integrated, integrated x
//            integrated) of the 2-point function at that (z, r, mu) are computed by ONE CTA:
//            thread 0 folds the end-point amplitudes into shared memory, the line-of-sight table is
//            staged in shared memory, warps stride over the outer line-of-sight nodes and lanes over
//            the inner panel nodes, partial sums are combined with warp shuffles;
//   the local x local piece is a 1-D Gauss-Legendre sum in mu, one warp per point;
//   a last kernel forms (2l+1) sum_mu w P_l(mu) xi(mu) per point in a fixed order (deterministic).
// This replaces the three OpenMP sweeps of coffe_multipoles_init / coffe_corrfunc_init /
// coffe_average_multipoles_init (src/multipoles.c:88-125, src/corrfunc.c:95-132,
// src/average_multipoles.c:112-151) and the Cuhre / QAG calls under coffe_integrate (src/signal.c:472-771).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "ctx.h"

namespace {

constexpr int kItemThreads = 128;
constexpr int kItemWarps = kItemThreads / 32;
constexpr int kMaxL = 8;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

/* ---- set-up: line-of-sight table -------------------------------------------------------------- */
__global__ void los_build_kernel(const __grid_constant__ Tables t, double4 *out, const int n, const double h, const double chi_max)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lam = i * h;
    if (lam > chi_max) lam = chi_max;
    out[i] = los_point(t, lam);
}

/* ---- local x local piece ----------------------------------------------------------------------- */
struct LList { int l[kMaxL]; int n; };

/* one warp per point: out[point*n_l + il] = (2l+1) int_0^1 du P_l(mu) xi_NI(mu), mu = 2u-1   (src/signal.c:83-111, 521-542) */
__global__ void __launch_bounds__(128) ni_multipoles_kernel(const __grid_constant__ Tables t, const __grid_constant__ Options o,
                                                            const double *__restrict__ pz, const double *__restrict__ psep, const int n_points,
                                                            const double *__restrict__ gl_x, const double *__restrict__ gl_w, const int n_mu,
                                                            const LList ls, double *__restrict__ out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_points) return;
    const double z = pz[warp], sep = psep[warp];
    const SepConst sc = sep_const(o, t, z, sep);
    double acc[kMaxL];
#pragma unroll
    for (int i = 0; i < kMaxL; ++i) acc[i] = 0.0;
    const double *x = gl_x + rule_offset(n_mu), *w = gl_w + rule_offset(n_mu);
    for (int j = lane; j < n_mu; j += 32) {
        const double mu = 2.0 * x[j] - 1.0;
        const Geometry g = geometry(sc.chi_mean, sep, mu);
        const EndVals e1 = end_eval(o, t, g.chi1, 1), e2 = end_eval(o, t, g.chi2, 2);
        const double v = w[j] * ni_eval(o, t, sc, g, e1, e2);
#pragma unroll
        for (int i = 0; i < kMaxL; ++i)
            if (i < ls.n) acc[i] += v * legendre_p(ls.l[i], mu);
    }
#pragma unroll
    for (int i = 0; i < kMaxL; ++i) {
        if (i < ls.n) {
            const double s = warp_sum(acc[i]);
            if (lane == 0) out[(size_t)warp * ls.n + i] = (2 * ls.l[i] + 1) * s;
        }
    }
}

/* one thread per (point, mu): plain evaluation (src/signal.c:515-520) */
__global__ void ni_points_kernel(const __grid_constant__ Tables t, const __grid_constant__ Options o,
                                 const double *__restrict__ pz, const double *__restrict__ psep,
                                 const int *__restrict__ item_point, const double *__restrict__ item_mu, const int n_items,
                                 double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const int p = item_point[i];
    const double z = pz[p], sep = psep[p], mu = item_mu[i];
    const SepConst sc = sep_const(o, t, z, sep);
    const Geometry g = geometry(sc.chi_mean, sep, mu);
    const EndVals e1 = end_eval(o, t, g.chi1, 1), e2 = end_eval(o, t, g.chi2, 2);
    out[i] = ni_eval(o, t, sc, g, e1, e2);
}

/* ---- integrated pieces: one CTA per item ------------------------------------------------------- */
struct ItemShared {
    Geometry g;
    SIHalf ha, hb;
    DIConst k;
    double red[kItemWarps][2];
    double panels[kItemWarps][2 * kMaxPanels];      /* per-warp panel list of ridge_integrate */
};

#ifndef COFFE_ITEM_MIN_BLOCKS
#define COFFE_ITEM_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(kItemThreads, COFFE_ITEM_MIN_BLOCKS) integrated_items_kernel(
    const __grid_constant__ Tables t, const __grid_constant__ Options o, const __grid_constant__ Rules q,
    const double *__restrict__ pz, const double *__restrict__ psep,
    const int *__restrict__ item_point, const double *__restrict__ item_mu, const int n_items,
    const int los_stage_n, double *__restrict__ out_si, double *__restrict__ out_di, unsigned int *__restrict__ out_evals)
{
    extern __shared__ double4 los_s[];
    __shared__ ItemShared sh;
    const int item = blockIdx.x;
    if (item >= n_items) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool has_si = opt_has_si(o), has_di = opt_has_integrated(o);

    /* stage the line-of-sight table {W, V, U, Wa}(lambda) */
    const double4 *los = t.los.v;
    if (los_stage_n > 0) {
        for (int i = tid; i < los_stage_n; i += kItemThreads) los_s[i] = t.los.v[i];
        los = los_s;
    }
    if (tid == 0) {
        const int p = item_point[item];
        const double z = pz[p], sep = psep[p], mu = item_mu[item];
        const double chi_mean = bg_chi(t.bg, z);
        sh.g = geometry(chi_mean, sep, mu);
        const EndVals e1 = end_eval(o, t, sh.g.chi1, 1), e2 = end_eval(o, t, sh.g.chi2, 2);
        if (has_si) {
            sh.ha = si_half_setup(o, sh.g, e1, e2, true);
            sh.hb = si_half_setup(o, sh.g, e2, e1, false);
        }
        if (has_di) sh.k = di_setup(o, e1, e2);
    }
    __syncthreads();

    double di = 0.0, si = 0.0;
    if (out_evals && tid < 2) {
        /* integrand evaluations of this item (roofline accounting): [2*item] = SI, [2*item+1] = DI */
        out_evals[2 * item + tid] = tid == 0 ? (has_si ? si_eval_count(q, sh.g, sh.ha) + si_eval_count(q, sh.g, sh.hb) : 0u)
                                             : (has_di ? di_eval_count(q, sh.g) : 0u);
    }
    /* local x integrated: one warp per half (1-D, ~3% of the work); the other warps go straight to the 2-D piece */
    const ExactAccess ta{&t};
    if (has_si && warp < 2) si = si_half_partial(ta, t.los, q, sh.g, warp == 0 ? sh.ha : sh.hb, los, lane, 32, sh.panels[warp]);
    if (has_di) di = di_partial(ta, t.los, q, sh.g, sh.k, los, warp, kItemWarps, lane, 32, sh.panels[warp]);
    di = warp_sum(di);
    si = warp_sum(si);
    if (lane == 0) { sh.red[warp][0] = si; sh.red[warp][1] = di; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < kItemWarps; ++w) { a += sh.red[w][0]; b += sh.red[w][1]; }
        out_si[item] = a;
        out_di[item] = b;
    }
}


/* ---- integrated pieces, shared-memory tables: persistent CTAs, one warp per item ------------------
   Every CTA copies the compact tables (compact.cuh) and the line-of-sight table into shared memory once,
   then its warps draw items from a global counter until none are left.  A warp owns its item: lanes 0-2
   fold the end-point amplitudes, all lanes stride over the (panel, node) pairs of the inner ridge rule
   while the outer line-of-sight nodes are visited in order, so the result of an item does not depend on
   which warp computed it (bit-reproducible for any grid size or GPU count). */
struct WarpState {
    EndVals e[2];
    SIHalf h[2];
    DIConst k;
};

constexpr int kSmemThreads = 384;

__global__ void __launch_bounds__(kSmemThreads, 1) items_smem_kernel(
    const __grid_constant__ Tables t, const __grid_constant__ Options o, const __grid_constant__ Rules q,
    const __grid_constant__ CompactHeader ch, const uint4 *__restrict__ blob, const int stage_bytes,
    const double *__restrict__ pz, const double *__restrict__ psep,
    const int *__restrict__ item_point, const double *__restrict__ item_mu, const int n_items,
    const int los_stage_n, const int panel_cap, double *__restrict__ out_si, double *__restrict__ out_di,
    unsigned int *__restrict__ out_evals, unsigned int *__restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int n_warps = kSmemThreads / 32;
    /* layout: [compact blob prefix][line-of-sight table][per-warp state][per-warp panel lists] */
    double4 *los_s = reinterpret_cast<double4 *>(smem + stage_bytes);
    WarpState *ws_all = reinterpret_cast<WarpState *>(smem + stage_bytes + (size_t)los_stage_n * sizeof(double4));
    double *scratch = reinterpret_cast<double *>(ws_all + n_warps) + (size_t)warp * 2 * panel_cap;
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < stage_bytes / 16; i += kSmemThreads) dst[i] = blob[i];
        for (int i = tid; i < los_stage_n; i += kSmemThreads) los_s[i] = t.los.v[i];
    }
    __syncthreads();
    WarpState &ws = ws_all[warp];
    CompactAccess ta;
    ta.t = &t; ta.c.base = smem; ta.c.h = ch;
    LosView lv = t.los;
    lv.v = los_s;
    const bool has_si = opt_has_si(o), has_di = opt_has_integrated(o);

    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= (unsigned int)n_items) break;
        const int p = item_point[item];
        const double z = pz[p], sep = psep[p], mu = item_mu[item];
        const Geometry g = geometry(bg_chi(t.bg, z), sep, mu);
        if (lane < 2) ws.e[lane] = end_eval(o, t, lane == 0 ? g.chi1 : g.chi2, lane + 1);
        __syncwarp();
        if (lane < 2 && has_si) ws.h[lane] = si_half_setup(o, g, ws.e[lane], ws.e[1 - lane], lane == 0);
        if (lane == 2 && has_di) ws.k = di_setup(o, ws.e[0], ws.e[1]);
        __syncwarp();
        if (out_evals && lane < 2)
            out_evals[2 * item + lane] = lane == 0 ? (has_si ? si_eval_count(q, g, ws.h[0]) + si_eval_count(q, g, ws.h[1]) : 0u)
                                                   : (has_di ? di_eval_count(q, g) : 0u);
        double si = 0.0, di = 0.0;
        if (has_si) {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const SIHalf h = ws.h[half];
                si += si_half_partial(ta, lv, q, g, h, los_s, lane, 32, scratch);
            }
        }
        if (has_di) {
            const DIConst k = ws.k;
            di = di_partial(ta, lv, q, g, k, los_s, 0, 1, lane, 32, scratch);
        }
        si = warp_sum(si);
        di = warp_sum(di);
        if (lane == 0) { out_si[item] = si; out_di[item] = di; }
        __syncwarp();
    }
}

/* ---- mu reduction: one warp per point, fixed summation order ---------------------------------- */
/* pm[(point*3 + piece)*n_l + il]; piece 0 (NI) is copied from ni[] */
__global__ void reduce_multipoles_kernel(const int *__restrict__ point_begin, const int n_points,
                                         const double *__restrict__ item_mu, const double *__restrict__ item_w,
                                         const double *__restrict__ si, const double *__restrict__ di,
                                         const double *__restrict__ ni, const LList ls, double *__restrict__ pm)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_points) return;
    const int b = point_begin[warp], e = point_begin[warp + 1];
    for (int il = 0; il < ls.n; ++il) {
        double a = 0.0, c = 0.0;
        for (int i = b + lane; i < e; i += 32) {
            const double f = item_w[i] * legendre_p(ls.l[il], item_mu[i]);
            a += f * si[i];
            c += f * di[i];
        }
        a = warp_sum(a); c = warp_sum(c);
        if (lane == 0) {
            const double pre = 2 * ls.l[il] + 1;
            pm[((size_t)warp * 3 + 0) * ls.n + il] = ni ? ni[(size_t)warp * ls.n + il] : 0.0;
            pm[((size_t)warp * 3 + 1) * ls.n + il] = pre * a;
            pm[((size_t)warp * 3 + 2) * ls.n + il] = pre * c;
        }
    }
}

/* ---- flat-sky closed forms (src/signal.c:610-631, 707-735) ------------------------------------------- */
/* adds the density-lensing closed form to piece 1 and the 1-D lensing-lensing integral to piece 2 of
   pm[(point*3 + piece)*n_l + il]; one warp per (point, l) */
__global__ void flatsky_multipoles_kernel(const __grid_constant__ Tables t, const __grid_constant__ Options o,
                                          const double *__restrict__ pz, const double *__restrict__ psep, const int n_points,
                                          const double *__restrict__ gl_x, const double *__restrict__ gl_w, const int n_x,
                                          const LList ls, const int do_denlen, const int do_lenlen, double *__restrict__ pm)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_points * ls.n) return;
    const int point = warp / ls.n, il = warp % ls.n, l = ls.l[il];
    const double z = pz[point], sep = psep[point];
    const SepConst sc = sep_const(o, t, z, sep);
    if (do_denlen && lane == 0) pm[((size_t)point * 3 + 1) * ls.n + il] += flatsky_density_lensing_multipole(o, t, sc, sep, l);
    if (do_lenlen && l % 2 == 0) {
        const SplineView *s = find_extra(t, 1, l, 0);
        double acc = 0.0;
        if (s) {
            const double *x = gl_x + rule_offset(n_x), *w = gl_w + rule_offset(n_x);
            for (int j = lane; j < n_x; j += 32) acc += w[j] * flatsky_lensing_lensing_integrand(o, t, sc, *s, sep, l, x[j]);
        }
        else acc = nan("");
        acc = warp_sum(acc);
        if (lane == 0) pm[((size_t)point * 3 + 2) * ls.n + il] += 2 * 3.14159265358979323846 * 3.14159265358979323846 * acc;
    }
}

/* ---- redshift average ---------------------------------------------------------------------------
   points (bin, r, iz): z = z1 + (z2 - z1) u, z1 = z(chi(z_min) + r/2), z2 = z(chi(z_max) - r/2), weight
   w_u H0 h / (Hc(z) (1+z))                                            src/signal.c:262-283 */
__global__ void average_points_kernel(const __grid_constant__ Tables t, const double H0h,
                                      const double *__restrict__ zmin, const double *__restrict__ zmax, const int n_bins,
                                      const double *__restrict__ sep, const int n_sep,
                                      const double *__restrict__ gl_x, const double *__restrict__ gl_w, const int n_z,
                                      double *__restrict__ pz, double *__restrict__ psep, double *__restrict__ pw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bins * n_sep * n_z) return;
    const int iz = i % n_z, ir = (i / n_z) % n_sep, ib = i / (n_z * n_sep);
    const double r = sep[ir];
    const double z1 = spline_eval(t.bg.z_of_chi, bg_chi(t.bg, zmin[ib]) + r / 2.);
    const double z2 = spline_eval(t.bg.z_of_chi, bg_chi(t.bg, zmax[ib]) - r / 2.);
    const double z = (z2 - z1) * gl_x[rule_offset(n_z) + iz] + z1;
    const int k = bg_index(t.bg, z);
    const double Hc = cubic_eval(t.bg.Hc[k], z - t.bg.z[k]);
    pz[i] = z; psep[i] = r;
    pw[i] = gl_w[rule_offset(n_z) + iz] / Hc / (1 + z) * H0h;
}

/* out[group*3*n_l + ...] = sum over the n_z points of a group of pw * pm */
__global__ void average_reduce_kernel(const double *__restrict__ pm, const double *__restrict__ pw, const int n_groups,
                                      const int n_z, const int n_l, double *__restrict__ gm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_groups * 3 * n_l) return;
    const int g = i / (3 * n_l), rest = i % (3 * n_l);
    double a = 0.0;
    for (int k = 0; k < n_z; ++k) a += pw[(size_t)g * n_z + k] * pm[((size_t)g * n_z + k) * 3 * n_l + rest];
    gm[i] = a;
}

/* FP64 FMA throughput probe: 8 independent chains per thread, `iters` x 8 DFMA each */
__global__ void __launch_bounds__(256) dfma_probe_kernel(double *out, const int iters, const double a, const double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace

/* ================================================================================================= */
static std::string g_create_error;

namespace {

Rules make_rules(const coffe_b200_ctx *ctx)
{
    Rules q;
    q.gl_x = ctx->d_gl_x; q.gl_w = ctx->d_gl_w; q.rpow = ctx->d_rpow;
    q.n_panel_si = ctx->qc.n_panel_si; q.n_panel_di = ctx->qc.n_panel_di;
    q.n_outer = ctx->qc.n_outer;
    q.panel_w_mpc = ctx->qc.panel_w; q.panel_du = ctx->qc.panel_du; q.eps_min = ctx->qc.eps_min; q.tail_ratio = ctx->qc.tail_ratio;
    return q;
}

/* host-side range check of one point: the reference aborts (non-finite integrand) outside the tables */
int check_point(coffe_b200_ctx *ctx, const double z, const double sep, const bool need_renorm)
{
    const Tables &t = ctx->t_host;
    const double zmax = t.bg.z[t.bg.n - 1];
    if (!(z > 0.0) || !(z < zmax) || !(sep > 0.0)) { ctx->err = "redshift or separation outside the background table"; return COFFE_B200_ERANGE; }
    const double chi = bg_chi(t.bg, z);
    if (!(chi - sep / 2. > 0.0)) { ctx->err = "separation larger than twice the comoving distance to z_mean"; return COFFE_B200_ERANGE; }
    if (!(chi + sep / 2. < ctx->layout.chi_max_table)) { ctx->err = "chi_mean + sep/2 outside the background table"; return COFFE_B200_ERANGE; }
    if (!(sep <= ctx->layout.r_max_table)) { ctx->err = "separation outside the I^n_l tables"; return COFFE_B200_ERANGE; }
    if (need_renorm) {
        const double rmax = t.ren2d.x[t.ren2d.n - 1];
        if (!(chi + sep / 2. <= rmax)) {
            ctx->err = "chi_mean + sep/2 outside the renormalisation grid (it extends to chi(max(z_mean + deltaz)), src/integrals.c:858-871)";
            return COFFE_B200_ERANGE;
        }
    }
    return COFFE_B200_OK;
}

bool needs_renorm(const Options &o) { return o.d2 || o.g1 || o.g2 || o.g3 || o.g4 || o.g5; }

/* a batch of points with their mu items, on the device */
struct Batch {
    int n_points = 0, n_items = 0;
    DevBuf<double> pz, psep, item_mu, item_w, si, di, ni, pm;
    DevBuf<int> item_point, point_begin;
    DevBuf<unsigned int> evals;     /* optional: per-item integrand evaluation counts */
    double chi_max = 0;     /* largest chi2 over the batch: how much of the line-of-sight table to stage */
};

int launch_integrated(coffe_b200_ctx *ctx, Batch &b)
{
    if (b.n_items == 0) return COFFE_B200_OK;
    if (!opt_has_integrated(ctx->o)) {
        CK(cudaMemsetAsync(b.si.p, 0, sizeof(double) * b.n_items, ctx->stream));
        CK(cudaMemsetAsync(b.di.p, 0, sizeof(double) * b.n_items, ctx->stream));
        return COFFE_B200_OK;
    }
    int stage_n = (int)std::ceil(b.chi_max / ctx->layout.los_h) + 4;
    if (stage_n > ctx->layout.los_n) stage_n = ctx->layout.los_n;
    if (ctx->variant == 1 && ctx->compact.h.valid) {
        /* shared-memory tables: blob prefix covering R <= 2 chi_max, line-of-sight table, per-warp state and panel lists */
        const int panel_cap = max_panels(ctx->qc);
        const int stage_bytes = compact_stage_bytes(ctx->compact, 2.0 * b.chi_max);
        const size_t need = (size_t)stage_bytes + (size_t)stage_n * sizeof(double4)
                            + (size_t)(kSmemThreads / 32) * (sizeof(WarpState) + (size_t)2 * panel_cap * sizeof(double));
        if (need <= (size_t)ctx->smem_optin) {
            CompactHeader ch = ctx->compact.h;
            ch.n_knots = (stage_bytes - ch.off_knots) / (kKnotStride * 8);
            CK(cudaFuncSetAttribute(items_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            CK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int), ctx->stream));
            const int warps = kSmemThreads / 32;
            int grid = std::min(ctx->n_sm, (b.n_items + warps - 1) / warps);
            if (grid < 1) grid = 1;
            items_smem_kernel<<<grid, kSmemThreads, need, ctx->stream>>>(
                ctx->t, ctx->o, make_rules(ctx), ch, reinterpret_cast<const uint4 *>(ctx->d_blob), stage_bytes, b.pz.p, b.psep.p,
                b.item_point.p, b.item_mu.p, b.n_items, stage_n, panel_cap, b.si.p, b.di.p, b.evals.p, ctx->d_counter);
            CK(cudaGetLastError());
            ctx->last_launches += 1;
            ctx->last_variant = 1;
            return COFFE_B200_OK;
        }
    }
    ctx->last_variant = 0;
    size_t smem = (size_t)stage_n * sizeof(double4);
    if (smem > 160 * 1024) { stage_n = 0; smem = 0; }     /* read the table through L1/L2 instead */
    CK(cudaFuncSetAttribute(integrated_items_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    integrated_items_kernel<<<b.n_items, kItemThreads, smem, ctx->stream>>>(
        ctx->t, ctx->o, make_rules(ctx), b.pz.p, b.psep.p, b.item_point.p, b.item_mu.p, b.n_items, stage_n, b.si.p, b.di.p, b.evals.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    return COFFE_B200_OK;
}

LList make_llist(const int *l, const int n_l)
{
    LList ls;
    ls.n = n_l;
    for (int i = 0; i < kMaxL; ++i) ls.l[i] = i < n_l ? l[i] : 0;
    return ls;
}

/* per-point multipoles of the three pieces into b.pm[(point*3 + piece)*n_l + il] */
int run_multipole_batch(coffe_b200_ctx *ctx, Batch &b, const LList &ls, const bool closed_forms)
{
    ctx->o.out_multipoles = closed_forms ? 1 : 0;
    const bool has_ni = opt_has_ni(ctx->o);
    if (has_ni) {
        const int blocks = (b.n_points * 32 + 127) / 128;
        ni_multipoles_kernel<<<blocks, 128, 0, ctx->stream>>>(ctx->t, ctx->o, b.pz.p, b.psep.p, b.n_points, ctx->d_gl_x, ctx->d_gl_w,
                                                              ctx->qc.n_mu_ni, ls, b.ni.p);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    int rc = launch_integrated(ctx, b);
    if (rc) return rc;
    const int blocks = (b.n_points * 32 + 127) / 128;
    reduce_multipoles_kernel<<<blocks, 128, 0, ctx->stream>>>(b.point_begin.p, b.n_points, b.item_mu.p, b.item_w.p, b.si.p, b.di.p,
                                                              has_ni ? b.ni.p : nullptr, ls, b.pm.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    /* flat-sky closed forms replace the corresponding full-sky integrals (src/signal.c:610-631, 707-735) */
    const Options &o = ctx->o;
    const int do_denlen = closed_forms && o.flatsky_local_nonlocal && o.len && o.den;
    const int do_lenlen = closed_forms && o.flatsky_nonlocal && o.len && !o.only_cross;
    if (do_denlen || do_lenlen) {
        const int nb = (b.n_points * ls.n * 32 + 127) / 128;
        flatsky_multipoles_kernel<<<nb, 128, 0, ctx->stream>>>(ctx->t, ctx->o, b.pz.p, b.psep.p, b.n_points, ctx->d_gl_x, ctx->d_gl_w, 96, ls,
                                                               do_denlen, do_lenlen, b.pm.p);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    return COFFE_B200_OK;
}

/* build the mu items of a list of points (host), upload everything */
int build_multipole_batch(coffe_b200_ctx *ctx, const std::vector<double> &pz, const std::vector<double> &psep, Batch &b,
                          const bool points_on_device)
{
    const int n = (int)psep.size();
    b.n_points = n;
    std::vector<int> item_point, point_begin(n + 1, 0);
    std::vector<double> item_mu, item_w, mu, w;
    const bool integrated = opt_has_integrated(ctx->o);
    double last_sep = -1.0;
    for (int p = 0; p < n; ++p) {
        point_begin[p] = (int)item_mu.size();
        if (!integrated) continue;
        if (psep[p] != last_sep) { build_mu_rule(ctx->qc, psep[p], mu, w); last_sep = psep[p]; }
        for (size_t j = 0; j < mu.size(); ++j) { item_point.push_back(p); item_mu.push_back(mu[j]); item_w.push_back(w[j]); }
    }
    point_begin[n] = (int)item_mu.size();
    b.n_items = (int)item_mu.size();
    if (!points_on_device) {
        CK(b.pz.upload(pz, ctx->stream));
        CK(b.psep.upload(psep, ctx->stream));
    }
    CK(b.item_point.upload(item_point, ctx->stream));
    CK(b.point_begin.upload(point_begin, ctx->stream));
    CK(b.item_mu.upload(item_mu, ctx->stream));
    CK(b.item_w.upload(item_w, ctx->stream));
    return COFFE_B200_OK;
}

int check_finite(coffe_b200_ctx *ctx, const double *v, const size_t n)
{
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(v[i])) { ctx->err = "non-finite result (an argument left the range of an uploaded table)"; return COFFE_B200_ENONFINITE; }
    return COFFE_B200_OK;
}

int check_l(coffe_b200_ctx *ctx, const int *l, const int n_l)
{
    if (!l || n_l < 1 || n_l > kMaxL) { ctx->err = "need between 1 and 8 multipoles"; return COFFE_B200_EINVAL; }
    for (int i = 0; i < n_l; ++i)
        if (l[i] < 0 || l[i] > 64) { ctx->err = "multipole out of range"; return COFFE_B200_EINVAL; }
    return COFFE_B200_OK;
}

}  // namespace

/* a prepared multipole computation: the (z, r) items [begin, end) with their mu rules resident in HBM */
struct coffe_b200_plan {
    coffe_b200_ctx *ctx = nullptr;
    Batch b;
    LList ls;
    long n = 0;
};

namespace {

int plan_build(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep, const int *l, int n_l,
               long begin, long end, coffe_b200_plan &pl)
{
    if (!ctx->have_tables || !ctx->layout.has_ifam) { ctx->err = "no tables uploaded (the signal needs the I^n_l family)"; return COFFE_B200_ENOTABLES; }
    if (n_z < 0 || n_sep < 0 || (n_z > 0 && !z_mean) || (n_sep > 0 && !sep)) { ctx->err = "bad grid arguments"; return COFFE_B200_EINVAL; }
    int rc = check_l(ctx, l, n_l);
    if (rc) return rc;
    const long total = (long)n_z * n_sep;
    if (begin < 0 || end > total || begin > end) { ctx->err = "item range outside the grid"; return COFFE_B200_EINVAL; }
    const Options &o = ctx->o;
    /* tables of the flat-sky closed forms must have been uploaded for every requested even l */
    for (int i = 0; i < n_l; ++i) {
        if (l[i] % 2) continue;
        if (o.flatsky_nonlocal && o.len && !o.only_cross && !find_extra(ctx->t_host, 1, l[i], 0)) {
            ctx->err = "flat-sky lensing-lensing multipoles need the I^1_l table of every even l (src/integrals.c:1085-1247)";
            return COFFE_B200_ENOTABLES;
        }
        if (o.flatsky_local_nonlocal && o.len && o.den)
            for (int k = 0; k <= l[i] / 2; ++k)
                if (!find_extra(ctx->t_host, l[i] - 2 * k + 3, l[i] - 2 * k + 1, 1)) {
                    ctx->err = "flat-sky density-lensing multipoles need the half-integer I^n_l tables (src/integrals.c:1249-1413)";
                    return COFFE_B200_ENOTABLES;
                }
    }
    pl.ctx = ctx; pl.n = end - begin; pl.ls = make_llist(l, n_l);
    if (end == begin) return COFFE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    std::vector<double> pz, psep;
    const bool ren = needs_renorm(ctx->o);
    if (ren && !ctx->layout.has_divergent) { ctx->err = "d2/g1..g5 need the divergent I^4_0 tables (par->divergent)"; return COFFE_B200_EINVAL; }
    pz.reserve(pl.n); psep.reserve(pl.n);
    for (long it = begin; it < end; ++it) {
        const double z = z_mean[it / n_sep], s = sep[it % n_sep];
        rc = check_point(ctx, z, s, ren && opt_has_ni(ctx->o));
        if (rc) return rc;
        pz.push_back(z); psep.push_back(s);
        const double chi2 = bg_chi(ctx->t_host.bg, z) + s / 2.;
        if (chi2 > pl.b.chi_max) pl.b.chi_max = chi2;
    }
    rc = build_multipole_batch(ctx, pz, psep, pl.b, false);
    if (rc) return rc;
    CK(pl.b.ni.alloc((size_t)pl.b.n_points * n_l));
    CK(pl.b.pm.alloc((size_t)pl.b.n_points * 3 * n_l));
    CK(pl.b.si.alloc(pl.b.n_items));
    CK(pl.b.di.alloc(pl.b.n_items));
    return COFFE_B200_OK;
}

int plan_run(coffe_b200_plan &pl)
{
    coffe_b200_ctx *ctx = pl.ctx;
    ctx->last_ms = 0; ctx->last_launches = 0;
    if (pl.n == 0) return COFFE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const int rc = run_multipole_batch(ctx, pl.b, pl.ls, true);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return COFFE_B200_OK;
}

int plan_sync(coffe_b200_plan &pl)
{
    coffe_b200_ctx *ctx = pl.ctx;
    if (pl.n == 0) return COFFE_B200_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    return COFFE_B200_OK;
}

/* host result hm[(k*3 + piece)*n_l + il], k = item - begin */
int plan_fetch(coffe_b200_plan &pl, std::vector<double> &hm)
{
    coffe_b200_ctx *ctx = pl.ctx;
    hm.assign((size_t)pl.n * 3 * pl.ls.n, 0.0);
    if (pl.n == 0) return COFFE_B200_OK;
    CK(cudaMemcpyAsync(hm.data(), pl.b.pm.p, hm.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return check_finite(ctx, hm.data(), hm.size());
}

/* shared body of multipoles / multipoles_range / multipoles_pieces */
int multipoles_core(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep, const int *l, int n_l,
                    long begin, long end, std::vector<double> &hm)
{
    coffe_b200_plan pl;
    int rc = plan_build(ctx, z_mean, n_z, sep, n_sep, l, n_l, begin, end, pl);
    if (rc) return rc;
    rc = plan_run(pl);
    if (rc) return rc;
    rc = plan_fetch(pl, hm);          /* the copy is ordered after the kernels on the same stream */
    if (rc) return rc;
    return plan_sync(pl);
}

}  // namespace

/* ================================================================================================= */
extern "C" {

int coffe_b200_create(int device, coffe_b200_ctx **out)
{
    if (!out) return COFFE_B200_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libcoffe_b200 has no CPU fallback)";
        return COFFE_B200_ECUDA;
    }
    if (device < 0 || device >= count) { g_create_error = "device index out of range"; return COFFE_B200_EINVAL; }
    coffe_b200_ctx *ctx = new coffe_b200_ctx();
    ctx->device = device;
    auto fail = [&](const char *what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        delete ctx;
        return COFFE_B200_ECUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return fail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return fail("cudaEventCreate", e);
    std::vector<double> bx, bw;
    build_rule_bank(bx, bw);
    if ((e = cudaMalloc(&ctx->d_gl_x, bx.size() * sizeof(double))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc(&ctx->d_gl_w, bw.size() * sizeof(double))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemcpy(ctx->d_gl_x, bx.data(), bx.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    if ((e = cudaMemcpy(ctx->d_gl_w, bw.data(), bw.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    ctx->qc = default_quadrature();
    {
        const std::vector<double> rp = build_rpow(ctx->qc.tail_ratio);
        if ((e = cudaMalloc(&ctx->d_rpow, kMaxTail * sizeof(double))) != cudaSuccess) return fail("cudaMalloc", e);
        if ((e = cudaMemcpy(ctx->d_rpow, rp.data(), kMaxTail * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
        if ((e = cudaMalloc(&ctx->d_counter, sizeof(unsigned int))) != cudaSuccess) return fail("cudaMalloc", e);
        int v = 0;
        if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return fail("cudaDeviceGetAttribute", e);
        ctx->n_sm = v;
        if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) return fail("cudaDeviceGetAttribute", e);
        ctx->smem_optin = v;
    }
    std::memset(&ctx->o, 0, sizeof(ctx->o));
    ctx->o.den = 1;
    *out = ctx;
    return COFFE_B200_OK;
}

void coffe_b200_destroy(coffe_b200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->d_arena) cudaFree(ctx->d_arena);
    if (ctx->d_gl_x) cudaFree(ctx->d_gl_x);
    if (ctx->d_gl_w) cudaFree(ctx->d_gl_w);
    if (ctx->d_rpow) cudaFree(ctx->d_rpow);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->d_blob) cudaFree(ctx->d_blob);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *coffe_b200_last_error(const coffe_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int coffe_b200_upload_tables(coffe_b200_ctx *ctx, const coffe_b200_tables *T)
{
    if (!ctx || !T) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    Arena A;
    Layout L;
    if (!assemble(*T, A, L, ctx->err)) return COFFE_B200_EINVAL;
    if (!(T->omega_matter > 0) || !(T->h > 0)) { ctx->err = "omega_matter and h must be positive"; return COFFE_B200_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    if (ctx->d_arena) { cudaFree(ctx->d_arena); ctx->d_arena = nullptr; }
    ctx->have_tables = false;
    CK(cudaMalloc(&ctx->d_arena, A.buf.size() * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->d_arena, A.buf.data(), A.buf.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->arena = std::move(A);
    ctx->layout = L;
    ctx->t = make_view(L, ctx->d_arena);
    ctx->t_host = make_view(L, ctx->arena.buf.data());
    los_build_kernel<<<(L.los_n + 127) / 128, 128, 0, ctx->stream>>>(ctx->t, reinterpret_cast<double4 *>(ctx->d_arena + L.los), L.los_n, L.los_h,
                                                                     L.chi_max_table);
    CK(cudaGetLastError());
    /* shared-memory tables of the integrated-pieces kernel (absent when the family is not log-spaced: exact-table kernel) */
    if (ctx->d_blob) { cudaFree(ctx->d_blob); ctx->d_blob = nullptr; }
    ctx->compact = CompactBuild();
    if (L.has_ifam) {
        std::string why;
        if (build_compact(ctx->t_host, L, kCompactKnotBudget, kCompactTolerance, ctx->compact, why)) {
            CK(cudaMalloc(&ctx->d_blob, ctx->compact.blob.size() * sizeof(double)));
            CK(cudaMemcpyAsync(ctx->d_blob, ctx->compact.blob.data(), ctx->compact.blob.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        }
        else ctx->compact.h.valid = 0;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->omega_matter = T->omega_matter; ctx->h = T->h;
    ctx->o.Om = T->omega_matter;
    ctx->o.H0h = T->h / 2997.92458;          /* COFFE_H0 = 1/2997.92458 h/Mpc, src/common.h:75-77 */
    ctx->o.divergent = L.has_divergent;
    ctx->have_tables = true;
    return COFFE_B200_OK;
}

int coffe_b200_set_options(coffe_b200_ctx *ctx, const coffe_b200_options *opt)
{
    if (!ctx || !opt) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    Options &o = ctx->o;
    o.den = opt->den != 0; o.rsd = opt->rsd != 0; o.len = opt->len != 0; o.d1 = opt->d1 != 0; o.d2 = opt->d2 != 0;
    o.g1 = opt->g1 != 0; o.g2 = opt->g2 != 0; o.g3 = opt->g3 != 0; o.g4 = opt->g4 != 0; o.g5 = opt->g5 != 0;
    o.only_cross = opt->only_cross_correlations != 0;
    o.flatsky_local = opt->flatsky_local != 0;
    o.flatsky_local_nonlocal = opt->flatsky_local_nonlocal != 0;
    o.flatsky_nonlocal = opt->flatsky_nonlocal != 0;
    return COFFE_B200_OK;
}

int coffe_b200_set_quadrature(coffe_b200_ctx *ctx, const coffe_b200_quadrature *q)
{
    if (!ctx || !q) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    const QuadConfig merged = merge_quadrature(default_quadrature(), *q);
    if (!quadrature_valid(merged, kMaxPanels)) {
        ctx->err = "quadrature sizes out of range (node counts 2..128; the line-of-sight panel plan must fit " + std::to_string(kMaxPanels) + " panels)";
        return COFFE_B200_EINVAL;
    }
    if (merged.tail_ratio != ctx->qc.tail_ratio) {
        const std::vector<double> rp = build_rpow(merged.tail_ratio);
        CK(cudaSetDevice(ctx->device));
        CK(cudaMemcpy(ctx->d_rpow, rp.data(), kMaxTail * sizeof(double), cudaMemcpyHostToDevice));
    }
    ctx->qc = merged;
    return COFFE_B200_OK;
}

int coffe_b200_set_kernel_variant(coffe_b200_ctx *ctx, int variant)
{
    if (!ctx || (variant != 0 && variant != 1)) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->variant = variant;
    return COFFE_B200_OK;
}

int coffe_b200_table_info(const coffe_b200_ctx *ctx, int *variant_used, int *compact_knots, int *compact_bytes, double *compact_max_rel_error)
{
    if (!ctx) return COFFE_B200_EINVAL;
    if (variant_used) *variant_used = ctx->last_variant;
    if (compact_knots) *compact_knots = ctx->compact.h.valid ? ctx->compact.h.n_knots : 0;
    if (compact_bytes) *compact_bytes = ctx->compact.h.valid ? ctx->compact.h.bytes : 0;
    if (compact_max_rel_error) *compact_max_rel_error = ctx->compact.h.valid ? ctx->compact.h.max_rel_error : 0.0;
    return COFFE_B200_OK;
}

int coffe_b200_get_quadrature(const coffe_b200_ctx *ctx, coffe_b200_quadrature *q)
{
    if (!ctx || !q) return COFFE_B200_EINVAL;
    const QuadConfig &c = ctx->qc;
    q->n_mu_ni = c.n_mu_ni; q->n_mu_panel = c.n_mu_panel; q->mu_rho0 = c.mu_rho0; q->mu_panel_width = c.mu_panel_w;
    q->mu_uniform_extent = c.mu_du; q->n_outer = c.n_outer; q->n_panel_si = c.n_panel_si; q->n_panel_di = c.n_panel_di;
    q->n_z = c.n_z; q->panel_width = c.panel_w; q->panel_uniform_extent = c.panel_du; q->eps_min = c.eps_min; q->tail_ratio = c.tail_ratio;
    return COFFE_B200_OK;
}

int coffe_b200_multipoles_pieces(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                                 const int *l, int n_l, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    std::vector<double> hm;
    const int rc = multipoles_core(ctx, z_mean, n_z, sep, n_sep, l, n_l, 0, (long)n_z * n_sep, hm);
    if (rc) return rc;
    for (int piece = 0; piece < 3; ++piece)
        for (int iz = 0; iz < n_z; ++iz)
            for (int il = 0; il < n_l; ++il)
                for (int ir = 0; ir < n_sep; ++ir)
                    out[(((size_t)piece * n_z + iz) * n_l + il) * n_sep + ir] = hm[(((size_t)iz * n_sep + ir) * 3 + piece) * n_l + il];
    return COFFE_B200_OK;
}

int coffe_b200_multipoles_range(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                                const int *l, int n_l, long item_begin, long item_end, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    std::vector<double> hm;
    const int rc = multipoles_core(ctx, z_mean, n_z, sep, n_sep, l, n_l, item_begin, item_end, hm);
    if (rc) return rc;
    const size_t n = (size_t)(item_end - item_begin);
    for (size_t k = 0; k < n; ++k)
        for (int il = 0; il < n_l; ++il)
            out[(size_t)il * n + k] = hm[(k * 3 + 0) * n_l + il] + hm[(k * 3 + 1) * n_l + il] + hm[(k * 3 + 2) * n_l + il];
    return COFFE_B200_OK;
}

int coffe_b200_multipoles(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                          const int *l, int n_l, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    std::vector<double> hm;
    const int rc = multipoles_core(ctx, z_mean, n_z, sep, n_sep, l, n_l, 0, (long)n_z * n_sep, hm);
    if (rc) return rc;
    for (int iz = 0; iz < n_z; ++iz)
        for (int il = 0; il < n_l; ++il)
            for (int ir = 0; ir < n_sep; ++ir) {
                const size_t k = (size_t)iz * n_sep + ir;
                out[((size_t)iz * n_l + il) * n_sep + ir] = hm[(k * 3 + 0) * n_l + il] + hm[(k * 3 + 1) * n_l + il] + hm[(k * 3 + 2) * n_l + il];
            }
    return COFFE_B200_OK;
}

int coffe_b200_corrfunc(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *mu, int n_mu,
                        const double *sep, int n_sep, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->have_tables || !ctx->layout.has_ifam) { ctx->err = "no tables uploaded (the signal needs the I^n_l family)"; return COFFE_B200_ENOTABLES; }
    if (!z_mean || !sep || !mu || n_z < 0 || n_sep < 0 || n_mu < 0) { ctx->err = "bad grid arguments"; return COFFE_B200_EINVAL; }
    ctx->last_ms = 0; ctx->last_launches = 0;
    const size_t n_out = (size_t)n_z * n_mu * n_sep;
    if (n_out == 0) return COFFE_B200_OK;
    for (int i = 0; i < n_mu; ++i)
        if (!(mu[i] > -1.0 && mu[i] < 1.0)) { ctx->err = "mu must lie strictly inside (-1, 1)"; return COFFE_B200_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    const bool ren = needs_renorm(ctx->o);
    if (ren && !ctx->layout.has_divergent) { ctx->err = "d2/g1..g5 need the divergent I^4_0 tables (par->divergent)"; return COFFE_B200_EINVAL; }
    Batch b;
    std::vector<double> pz, psep, item_mu;
    std::vector<int> item_point;
    for (int iz = 0; iz < n_z; ++iz)
        for (int ir = 0; ir < n_sep; ++ir) {
            const int rc = check_point(ctx, z_mean[iz], sep[ir], ren && opt_has_ni(ctx->o));
            if (rc) return rc;
            pz.push_back(z_mean[iz]); psep.push_back(sep[ir]);
            const double chi2 = bg_chi(ctx->t_host.bg, z_mean[iz]) + sep[ir] / 2.;
            if (chi2 > b.chi_max) b.chi_max = chi2;
        }
    /* items in output order (iz, imu, ir) */
    for (int iz = 0; iz < n_z; ++iz)
        for (int im = 0; im < n_mu; ++im)
            for (int ir = 0; ir < n_sep; ++ir) { item_point.push_back(iz * n_sep + ir); item_mu.push_back(mu[im]); }
    b.n_points = (int)pz.size(); b.n_items = (int)item_mu.size();
    CK(b.pz.upload(pz, ctx->stream));
    CK(b.psep.upload(psep, ctx->stream));
    CK(b.item_point.upload(item_point, ctx->stream));
    CK(b.item_mu.upload(item_mu, ctx->stream));
    CK(b.si.alloc(b.n_items));
    CK(b.di.alloc(b.n_items));
    CK(b.ni.alloc(b.n_items));
    ctx->o.out_multipoles = 0;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (opt_has_ni(ctx->o)) {
        ni_points_kernel<<<(b.n_items + 127) / 128, 128, 0, ctx->stream>>>(ctx->t, ctx->o, b.pz.p, b.psep.p, b.item_point.p, b.item_mu.p, b.n_items, b.ni.p);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    else CK(cudaMemsetAsync(b.ni.p, 0, sizeof(double) * b.n_items, ctx->stream));
    int rc = launch_integrated(ctx, b);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    std::vector<double> ni(n_out), si(n_out), di(n_out);
    CK(cudaMemcpyAsync(ni.data(), b.ni.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(si.data(), b.si.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(di.data(), b.di.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    for (size_t i = 0; i < n_out; ++i) out[i] = ni[i] + si[i] + di[i];
    return check_finite(ctx, out, n_out);
}

int coffe_b200_average_multipoles(coffe_b200_ctx *ctx, const double *z_min, const double *z_max, int n_bins,
                                  const double *sep, int n_sep, const int *l, int n_l, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->have_tables || !ctx->layout.has_ifam) { ctx->err = "no tables uploaded (the signal needs the I^n_l family)"; return COFFE_B200_ENOTABLES; }
    if (!z_min || !z_max || !sep || n_bins < 0 || n_sep < 0) { ctx->err = "bad grid arguments"; return COFFE_B200_EINVAL; }
    int rc = check_l(ctx, l, n_l);
    if (rc) return rc;
    ctx->last_ms = 0; ctx->last_launches = 0;
    const int n_groups = n_bins * n_sep, nz = ctx->qc.n_z;
    if (n_groups == 0) return COFFE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    const bool ren = needs_renorm(ctx->o);
    if (ren && !ctx->layout.has_divergent) { ctx->err = "d2/g1..g5 need the divergent I^4_0 tables (par->divergent)"; return COFFE_B200_EINVAL; }
    Batch b;
    std::vector<double> psep((size_t)n_groups * nz), pz_dummy;
    const Tables &th = ctx->t_host;
    for (int ib = 0; ib < n_bins; ++ib) {
        if (!(z_min[ib] > 0.0) || !(z_max[ib] > z_min[ib]) || !(z_max[ib] < th.bg.z[th.bg.n - 1])) { ctx->err = "bad redshift bin"; return COFFE_B200_ERANGE; }
        const double c1 = bg_chi(th.bg, z_min[ib]), c2 = bg_chi(th.bg, z_max[ib]);
        for (int ir = 0; ir < n_sep; ++ir) {
            if (!(sep[ir] > 0.0) || !(c1 + sep[ir] / 2. < c2 - sep[ir] / 2.)) {
                ctx->err = "separation does not fit inside the redshift bin (need chi(z_min) + r/2 < chi(z_max) - r/2)";
                return COFFE_B200_ERANGE;
            }
            if (ren && opt_has_ni(ctx->o) && !(c2 <= th.ren2d.x[th.ren2d.n - 1])) { ctx->err = "redshift bin outside the renormalisation grid"; return COFFE_B200_ERANGE; }
            for (int k = 0; k < nz; ++k) psep[((size_t)ib * n_sep + ir) * nz + k] = sep[ir];
        }
        if (c2 > b.chi_max) b.chi_max = c2;
    }
    /* points on the device */
    DevBuf<double> d_zmin, d_zmax, d_sep, pw, gm;
    CK(d_zmin.upload(std::vector<double>(z_min, z_min + n_bins), ctx->stream));
    CK(d_zmax.upload(std::vector<double>(z_max, z_max + n_bins), ctx->stream));
    CK(d_sep.upload(std::vector<double>(sep, sep + n_sep), ctx->stream));
    const int n_points = n_groups * nz;
    CK(b.pz.alloc(n_points));
    CK(b.psep.alloc(n_points));
    CK(pw.alloc(n_points));
    CK(gm.alloc((size_t)n_groups * 3 * n_l));
    rc = build_multipole_batch(ctx, pz_dummy, psep, b, true);
    if (rc) return rc;
    CK(b.ni.alloc((size_t)b.n_points * n_l));
    CK(b.pm.alloc((size_t)b.n_points * 3 * n_l));
    CK(b.si.alloc(b.n_items));
    CK(b.di.alloc(b.n_items));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    average_points_kernel<<<(n_points + 127) / 128, 128, 0, ctx->stream>>>(ctx->t, ctx->o.H0h, d_zmin.p, d_zmax.p, n_bins, d_sep.p, n_sep,
                                                                          ctx->d_gl_x, ctx->d_gl_w, nz, b.pz.p, b.psep.p, pw.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    rc = run_multipole_batch(ctx, b, make_llist(l, n_l), false);
    if (rc) return rc;
    average_reduce_kernel<<<(n_groups * 3 * n_l + 127) / 128, 128, 0, ctx->stream>>>(b.pm.p, pw.p, n_groups, nz, n_l, gm.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    std::vector<double> hm((size_t)n_groups * 3 * n_l);
    CK(cudaMemcpyAsync(hm.data(), gm.p, hm.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    for (int ib = 0; ib < n_bins; ++ib)
        for (int il = 0; il < n_l; ++il)
            for (int ir = 0; ir < n_sep; ++ir) {
                const size_t g = (size_t)ib * n_sep + ir;
                out[((size_t)ib * n_l + il) * n_sep + ir] = hm[(g * 3 + 0) * n_l + il] + hm[(g * 3 + 1) * n_l + il] + hm[(g * 3 + 2) * n_l + il];
            }
    return check_finite(ctx, out, (size_t)n_groups * n_l);
}

int coffe_b200_last_timing(const coffe_b200_ctx *ctx, double *kernel_ms, int *n_launches)
{
    if (!ctx) return COFFE_B200_EINVAL;
    if (kernel_ms) *kernel_ms = ctx->last_ms;
    if (n_launches) *n_launches = ctx->last_launches;
    return COFFE_B200_OK;
}

/* ---- prepared computations (bench: inputs resident in HBM before the timed region) ---------------- */
int coffe_b200_plan_multipoles(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                               const int *l, int n_l, long item_begin, long item_end, coffe_b200_plan **out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    *out = nullptr;
    std::lock_guard<std::mutex> lock(ctx->mu);
    coffe_b200_plan *pl = new coffe_b200_plan();
    const int rc = plan_build(ctx, z_mean, n_z, sep, n_sep, l, n_l, item_begin, item_end, *pl);
    if (rc) { delete pl; return rc; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { delete pl; ctx->err = "upload failed"; return COFFE_B200_ECUDA; }
    *out = pl;
    return COFFE_B200_OK;
}

int coffe_b200_plan_execute(coffe_b200_plan *plan, double *kernel_ms)
{
    if (!plan) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(plan->ctx->mu);
    int rc = plan_run(*plan);
    if (rc) return rc;
    rc = plan_sync(*plan);
    if (kernel_ms) *kernel_ms = plan->ctx->last_ms;
    return rc;
}

int coffe_b200_plan_fetch(coffe_b200_plan *plan, double *out)
{
    if (!plan || !out) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(plan->ctx->mu);
    std::vector<double> hm;
    const int rc = plan_fetch(*plan, hm);
    if (rc) return rc;
    const size_t n = (size_t)plan->n;
    const int n_l = plan->ls.n;
    for (size_t k = 0; k < n; ++k)
        for (int il = 0; il < n_l; ++il)
            out[(size_t)il * n + k] = hm[(k * 3 + 0) * n_l + il] + hm[(k * 3 + 1) * n_l + il] + hm[(k * 3 + 2) * n_l + il];
    return COFFE_B200_OK;
}

int coffe_b200_plan_count(coffe_b200_plan *plan, double *n_items, double *n_ni_evals, double *n_si_evals, double *n_di_evals)
{
    if (!plan) return COFFE_B200_EINVAL;
    coffe_b200_ctx *ctx = plan->ctx;
    std::lock_guard<std::mutex> lock(ctx->mu);
    Batch &b = plan->b;
    double si = 0, di = 0;
    if (b.n_items > 0 && opt_has_integrated(ctx->o)) {
        CK(cudaSetDevice(ctx->device));
        CK(b.evals.alloc((size_t)2 * b.n_items));
        ctx->o.out_multipoles = 1;
        int rc = launch_integrated(ctx, b);
        if (rc) return rc;
        std::vector<unsigned int> h((size_t)2 * b.n_items);
        CK(cudaMemcpyAsync(h.data(), b.evals.p, h.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(b.evals.p); b.evals.p = nullptr; b.evals.n = 0;
        for (int i = 0; i < b.n_items; ++i) { si += h[2 * i]; di += h[2 * i + 1]; }
    }
    if (n_items) *n_items = b.n_items;
    if (n_ni_evals) *n_ni_evals = opt_has_ni(ctx->o) ? (double)b.n_points * ctx->qc.n_mu_ni : 0.0;
    if (n_si_evals) *n_si_evals = si;
    if (n_di_evals) *n_di_evals = di;
    return COFFE_B200_OK;
}

void coffe_b200_plan_destroy(coffe_b200_plan *plan)
{
    if (!plan) return;
    cudaSetDevice(plan->ctx->device);
    delete plan;
}

/* achievable FP64 FMA rate of this device [TFLOP/s]: the roofline denominator */
int coffe_b200_fp64_peak(coffe_b200_ctx *ctx, double *tflops)
{
    if (!ctx || !tflops) return COFFE_B200_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    DevBuf<double> out;
    CK(out.alloc((size_t)blocks * threads));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        dfma_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(out.p, iters, 0.999999, 1e-9);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *tflops = best;
    return COFFE_B200_OK;
}

}  // extern "C"
