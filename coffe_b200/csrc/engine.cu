// libcoffe_b200: sm_100a kernels and C ABI of the signal-evaluation path.
//
// Work decomposition:
//   point  = one (z, separation) pair of the output grid (for redshift-averaged multipoles: one
//            (bin, separation, z-node) triple);
//   item   = one (point, mu-node) pair.  The integrated pieces (local x integrated, integrated x
//            integrated) of the 2-point function at that (z, r, mu) come from three kernels per piece:
//            item_setup_kernel (one thread per item: end-point values folded into amplitudes / polynomial
//            coefficients, one record per item in HBM), then a persistent kernel with the I^n_l and
//            line-of-sight tables in shared memory, in which every GROUP of n_panel lanes walks one
//            line-of-sight integral (di_items_smem_kernel: one warp per item, one outer node per group;
//            si_items_smem_kernel: one half per group), partial sums combined with warp shuffles;
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
#include <chrono>
#include <cstdlib>

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
__global__ void los_build_kernel(const __grid_constant__ Tables t, double2 *out, const int n, const double h, const double chi_max)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 k[kLosStride];
    los_knot(t, i, h, chi_max, k);
    for (int c = 0; c < kLosStride; ++c) out[(size_t)kLosStride * i + c] = k[c];
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

/* ---- integrated pieces, exact tables through L1/L2: one CTA per item (variant 0, cross-check) -----
   mode 1: local x integrated (one warp per half); mode 2: integrated x integrated (outer nodes strided
   over the warps).  Dynamic shared memory: [line-of-sight table]. */
struct ItemShared {
    Geometry g;
    SIHalf ha, hb;
    DIConst k;
    double red[kItemWarps];
    double panels[kItemWarps][2 * kMaxPanels];      /* per-warp panel list of ridge_integrate */
};

#ifndef COFFE_ITEM_MIN_BLOCKS
#define COFFE_ITEM_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(kItemThreads, COFFE_ITEM_MIN_BLOCKS) integrated_items_kernel(
    const __grid_constant__ Tables t, const __grid_constant__ Options o, const __grid_constant__ Rules q, const int mode,
    const double *__restrict__ pz, const double *__restrict__ psep,
    const int *__restrict__ item_point, const double *__restrict__ item_mu, const int n_items,
    const int los_stage_n, double *__restrict__ out, unsigned int *__restrict__ out_evals)
{
    extern __shared__ double2 los_s[];
    __shared__ ItemShared sh;
    const int item = blockIdx.x;
    if (item >= n_items) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    /* stage the line-of-sight table {W, V, U, Wa}(lambda) */
    const double2 *los = t.los.v;
    if (los_stage_n > 0) {
        for (int i = tid; i < kLosStride * los_stage_n; i += kItemThreads) los_s[i] = t.los.v[i];
        los = los_s;
    }
    if (tid == 0) {
        const int p = item_point[item];
        const double z = pz[p], sep = psep[p], mu = item_mu[item];
        const double chi_mean = bg_chi(t.bg, z);
        sh.g = geometry(chi_mean, sep, mu);
        const EndVals e1 = end_eval(o, t, sh.g.chi1, 1), e2 = end_eval(o, t, sh.g.chi2, 2);
        if (mode == 1) {
            sh.ha = si_half_setup(o, sh.g, e1, e2, true);
            sh.hb = si_half_setup(o, sh.g, e2, e1, false);
        }
        else sh.k = di_setup(o, e1, e2);
    }
    __syncthreads();

    if (out_evals && tid == 0)
        out_evals[item] = mode == 1 ? (o.midpoint ? si_eval_count(q, sh.g, sh.ha) + si_eval_count(q, sh.g, sh.hb)
                                                  : si_group_eval_count(q, sh.g, sh.ha) + si_group_eval_count(q, sh.g, sh.hb))
                                    : (o.midpoint ? di_eval_count(q, sh.g) : di_fast_eval_count(q, sh.g, sh.k));
    const ExactAccess ta{&t};
    double v = 0.0;
    if (mode == 1) {
        if (warp < 2)
            /* midpoint: the 2-D tables depend on z(lambda), no shared lookup for mirror nodes: the plain ridge rule.  Otherwise the
               group walk of the shared-memory kernel, one group per half here (lanes beyond n_panel_si idle: this is the cross-check path) */
            v = o.midpoint ? si_half_partial<ExactAccess, true>(ta, t.los, q, sh.g, warp == 0 ? sh.ha : sh.hb, los, lane, 32, sh.panels[warp])
                           : (lane < q.n_panel_si ? si_half_group(ta, t.ren2d, t.los, q, sh.g, warp == 0 ? sh.ha : sh.hb, los, lane) : 0.0);
    }
    /* midpoint approximation: the mirror nodes of di_fast share one table lookup, which the 2-D tables (z differs between the
       mirror nodes) do not allow: the plain ridge rule over di_eval, same panels and nodes */
    else if (o.midpoint) v = di_partial(ta, t.los, q, sh.g, sh.k, los, warp, kItemWarps, lane, 32, sh.panels[warp]);
    else {
        /* groups of n_panel lanes, one outer node each (di_fast.cuh); a rule with more than 32 nodes per panel has no lane layout */
        const int np = q.n_panel_di, ppi = 32 / np, grp = lane / np;
        v = di_fast_partial(ta, t, los, q, sh.g, sh.k, grp < ppi ? warp * ppi + grp : kItemWarps * ppi, kItemWarps * ppi, lane - grp * np);
    }
    v = warp_sum(v);
    if (lane == 0) sh.red[warp] = v;
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kItemWarps; ++w) a += sh.red[w];
        out[item] = a;
    }
}


/* ---- integrated pieces, shared-memory tables: persistent CTAs ---------------------------------------
   Every CTA copies the compact tables (compact.cuh) and the line-of-sight table into shared memory once,
   then its warps draw work from a global counter until none is left.  A group of n_panel lanes owns one
   line-of-sight integral and sums it in a fixed order, so a result does not depend on which warp computed it
   (bit-reproducible for any grid size or GPU count).  Two kernels so that each gets its own register
   allocation: si_items_smem_kernel (local x integrated: two 1-D integrals per item, one per group) and
   di_items_smem_kernel (integrated x integrated: the 2-D integral, one warp per item and one outer node per
   group; 86 % of the device time of a cfg5 step). */
/* doubles of the small read-only block staged next to the tables: powers of the tail ratio, then the nodes and
   weights of the inner panel rule */
constexpr int kDISmallDoubles = kMaxTail + 3 * 32;

/* What an item needs before its line-of-sight integrals start: the values at its two end points (background, biases; a dozen
   dependent spline lookups through L2) folded into the amplitudes of the two halves / of the integrated x integrated terms.  Computed
   by item_setup_kernel, one THREAD per item, ahead of the integration kernels: inside them this chain ran on one or two lanes of a
   warp while the other thirty waited (a third of the local x integrated kernel's time). */
struct SIItemSetup { Geometry g; SIHalf h[2]; };
struct DIItemSetup { DIConst k; };

constexpr int kSetupThreads = 32;
/* words of a staged record: one more than the record so that the threads of a warp write different banks */
constexpr int kSetupStride = (int)(sizeof(SIItemSetup) / 8) + 1;

template <int mode>
__global__ void item_setup_kernel(const __grid_constant__ Tables t, const __grid_constant__ Options o,
                                  const double *__restrict__ pz, const double *__restrict__ psep,
                                  const int *__restrict__ item_point, const double *__restrict__ item_mu, const int n_items,
                                  SIItemSetup *__restrict__ si_out, DIItemSetup *__restrict__ di_out)
{
    /* the 768-byte records of the local x integrated items go through shared memory so that the block writes them as one
       contiguous run (a record per thread straight to global memory: 96 stores of a 768-byte stride per warp, 0.9 ms per cfg5 step) */
    __shared__ double stage[mode == 1 ? kSetupThreads * kSetupStride : 1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i < n_items;
    if (have) {
        const int p = item_point[i];
        const Geometry g = geometry(bg_chi(t.bg, pz[p]), psep[p], item_mu[i]);
        /* redshift of the end points from the line-of-sight table (fifth slot), as the integration kernels read it along the way */
        const EndVals e1 = end_eval_at(o, t, g.chi1, los_z(t.los.v, t.los.n, t.los.inv_h, g.chi1), 1);
        const EndVals e2 = end_eval_at(o, t, g.chi2, los_z(t.los.v, t.los.n, t.los.inv_h, g.chi2), 2);
        if (mode == 1) {
            SIItemSetup *rec = reinterpret_cast<SIItemSetup *>(stage + (size_t)threadIdx.x * kSetupStride);
            rec->g = g;
            rec->h[0] = si_half_setup(o, g, e1, e2, true);
            rec->h[1] = si_half_setup(o, g, e2, e1, false);
        }
        else di_out[i].k = di_setup(o, e1, e2);
    }
    if (mode != 1) return;
    __syncthreads();
    constexpr int kWords = (int)(sizeof(SIItemSetup) / 8);
    const int first = blockIdx.x * blockDim.x, n_here = min((int)blockDim.x, n_items - first);
    double *dst = reinterpret_cast<double *>(si_out + first);
    for (int e = threadIdx.x; e < n_here * kWords; e += blockDim.x) dst[e] = stage[(e / kWords) * kSetupStride + e % kWords];
}

constexpr int kSIThreads = 384;

/* Local x integrated piece, one HALF (item, source end) per group of n_panel_si lanes, 32 / n_panel_si halves side by side in a warp
   (item.cuh: si_half_group).  A warp draws a batch of 32 / n_panel_si consecutive halves from a global counter; each group copies
   the record of its half's item (geometry + both halves, written by item_setup_kernel) into its slot of shared memory -- read from
   there with group-uniform addresses -- integrates, and adds its half to out[item] (two addends on a zeroed value: the order does
   not matter).  The batches are drawn two ahead and the next record is in registers before the current half is done. */
struct SIGroupSlot { Geometry g; SIHalf h; };

__global__ void __launch_bounds__(kSIThreads, 1) si_items_smem_kernel(
    const __grid_constant__ Tables t, const __grid_constant__ Options o, const __grid_constant__ Rules q_in,
    const __grid_constant__ CompactHeader ch, const uint4 *__restrict__ blob, const int stage_bytes,
    const SIItemSetup *__restrict__ item_setup, const int n_items,
    const int los_stage_n, double *__restrict__ out,
    unsigned int *__restrict__ out_evals, unsigned int *__restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const unsigned int full = 0xffffffffu;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int np = q_in.n_panel_si, ppi = 32 / np;          /* n_panel_si <= 32: the host checks */
    const int grp = lane / np, j = lane - grp * np;
    const bool lane_on = grp < ppi;
    /* layout: [compact blob prefix][line-of-sight table][tail powers, panel rule][per-warp: ppi slots] */
    LosRec *los_s = reinterpret_cast<LosRec *>(smem + stage_bytes);          /* one record per interval (tables.cuh) in the room of the knots */
    double *small = reinterpret_cast<double *>(los_s + los_stage_n);
    SIGroupSlot *slots = reinterpret_cast<SIGroupSlot *>(small + kDISmallDoubles) + (size_t)warp * ppi;
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < stage_bytes / 16; i += kSIThreads) dst[i] = blob[i];
        for (int i = tid; i + 1 < los_stage_n; i += kSIThreads) los_rec_pack(t.los.v + (size_t)kLosStride * i, t.los.v + (size_t)kLosStride * (i + 1), los_s[i]);
        if (tid < kMaxTail) small[tid] = q_in.rpow[tid];
        if (tid < np) {
            small[kMaxTail + tid] = q_in.gl_x[rule_offset(np) + tid];
            small[kMaxTail + 32 + tid] = q_in.gl_w[rule_offset(np) + tid];
            small[kMaxTail + 64 + tid] = exp(q_in.tail_lnr * q_in.gl_x[rule_offset(np) + tid]);
        }
    }
    __syncthreads();
    CompactAccess ta;
    ta.t = &t; ta.c.base = smem; ta.c.h = ch;
    const LosView lv = t.los;           /* n and 1/h; the values are read from the staged records */
    Rules q = q_in;
    q.rpow = small;
    q.panel_x_si = small + kMaxTail; q.panel_w_si = small + kMaxTail + 32; q.panel_p_si = small + kMaxTail + 64;

    /* words of a slot: the geometry of the item, then the half */
    constexpr int kGeoWords = (int)(sizeof(Geometry) / 8), kHalfWords = (int)(sizeof(SIHalf) / 8), kWords = kGeoWords + kHalfWords;
    constexpr int kRecWords = (int)(sizeof(SIItemSetup) / 8);
    static_assert(sizeof(Geometry) % 8 == 0 && sizeof(SIHalf) % 8 == 0 && offsetof(SIItemSetup, h) == sizeof(Geometry), "records are copied in 8-byte words");
    constexpr int kPre = 9;                                 /* words per lane held in registers: enough for n_panel_si >= 6 */
    const double *records = reinterpret_cast<const double *>(item_setup);
    const unsigned int n_halves = 2u * (unsigned int)n_items;
    auto draw = [&]() {
        unsigned int i = 0;
        if (lane == 0) i = atomicAdd(counter, (unsigned int)ppi);
        return __shfl_sync(full, i, 0);
    };
    /* word i of the slot of half hh: the geometry of its item, then the half */
    auto word_of = [&](const unsigned int hh, const int i) {
        const int half_off = kGeoWords + (int)(hh & 1u) * kHalfWords;
        return records + (size_t)(hh >> 1) * kRecWords + (i < kGeoWords ? i : half_off + (i - kGeoWords));
    };
    double pre[kPre];
    /* batch q (drawn as first = q ppi): the halves of the SAME source end of ppi consecutive items -- neighbours in mu of one point,
       whose integrals have nearly the same number of panels, so that the groups of a warp finish together (batches of consecutive
       halves, which mix the two ends of an item, left 22 of 32 lanes busy) */
    auto half_of = [&](const unsigned int batch) {
        const unsigned int q = batch / (unsigned int)ppi;
        return 2u * ((q >> 1) * (unsigned int)ppi + (unsigned int)grp) + (q & 1u);
    };
    auto prefetch = [&](const unsigned int batch) {
        const unsigned int hh = half_of(batch);
        if (lane_on && hh < n_halves)
#pragma unroll
            for (int w = 0; w < kPre; ++w)
                if (j + np * w < kWords) pre[w] = *word_of(hh, j + np * w);
    };
    unsigned int first = draw(), next = draw();
    prefetch(first);
    /* batches cover the items in pairs of (ppi items, end 0), (ppi items, end 1): the last pair may reach beyond the list */
    const unsigned int n_batched = 2u * (unsigned int)ppi * (((unsigned int)n_items + ppi - 1) / ppi);
    while (first < n_batched) {
        const unsigned int hh = half_of(first);
        const bool valid = lane_on && hh < n_halves;
        SIGroupSlot &slot = slots[lane_on ? grp : 0];
        __syncwarp();                   /* the previous batch's readers are done */
        if (valid) {
            double *dst = reinterpret_cast<double *>(&slot);
#pragma unroll
            for (int w = 0; w < kPre; ++w)
                if (j + np * w < kWords) dst[j + np * w] = pre[w];
            for (int i = j + np * kPre; i < kWords; i += np) dst[i] = *word_of(hh, i);      /* fewer than 6 lanes per group: the rest directly */
        }
        __syncwarp();
        prefetch(next);
        const unsigned int after = draw();
        double si = 0.0;
        if (valid) {
            si = si_half_group(ta, t.ren2d, lv, q, slot.g, slot.h, (const LosRec *)los_s, j);
            if (out_evals && j == 0) atomicAdd(&out_evals[hh >> 1], si_group_eval_count(q, slot.g, slot.h));
        }
        /* sum over the lanes of the group: np <= 32 consecutive lanes */
        for (int off = 1; off < np; off <<= 1) {
            const double other = __shfl_down_sync(full, si, off);
            if (j + off < np) si += other;
        }
        if (valid && j == 0) atomicAdd(&out[hh >> 1], si);
        first = next; next = after;
    }
}

#ifndef COFFE_DI_THREADS
#define COFFE_DI_THREADS 384
#endif
constexpr int kDIThreads = COFFE_DI_THREADS;        /* upper bound (register allocation); the launch may use fewer warps */

/* Lane layout (di_fast.cuh): the lanes of a warp are 32 / n_panel groups of n_panel lanes; a group owns ONE outer line-of-sight node
   at a time and walks through the panels of its inner integral in sequence (lane = node of the panel), so the outer nodes in flight
   fill the warp whatever their panel counts are.  (Until round 2 a warp held one outer node and spread its panels over the lanes, five
   at a time: with ~8 mirror pairs and ~3 single panels per outer node 30 % of the sweeps ran empty, and every lane paid the set-up of
   every outer node.) */
__global__ void __launch_bounds__(kDIThreads, 1) di_items_smem_kernel(
    const __grid_constant__ Tables t, const __grid_constant__ Options o, const __grid_constant__ Rules q,
    const __grid_constant__ CompactHeader ch, const uint4 *__restrict__ blob, const int stage_bytes,
    const double *__restrict__ pz, const double *__restrict__ psep,
    const int *__restrict__ item_point, const double *__restrict__ item_mu, const DIItemSetup *__restrict__ item_setup, const int n_items,
    const int los_stage_n, double *__restrict__ out,
    unsigned int *__restrict__ out_evals, unsigned int *__restrict__ counter)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_threads = blockDim.x;
    const int np = q.n_panel_di, ppi = 32 / np;             /* n_panel_di <= 32: the host checks */
    const int grp = lane / np, j = lane - grp * np;
    /* layout: [compact blob prefix][line-of-sight table][tail powers, panel rule] */
    LosRec *los_s = reinterpret_cast<LosRec *>(smem + stage_bytes);
    double *small = reinterpret_cast<double *>(los_s + los_stage_n);
    {
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < stage_bytes / 16; i += n_threads) dst[i] = blob[i];
        for (int i = tid; i + 1 < los_stage_n; i += n_threads) los_rec_pack(t.los.v + (size_t)kLosStride * i, t.los.v + (size_t)kLosStride * (i + 1), los_s[i]);
        if (tid < kMaxTail) small[tid] = q.rpow_di[tid];
        if (tid < np) {
            small[kMaxTail + tid] = q.gl_x[rule_offset(np) + tid];
            small[kMaxTail + 32 + tid] = q.gl_w[rule_offset(np) + tid];
            small[kMaxTail + 64 + tid] = exp(q.tail_lnr_di * q.gl_x[rule_offset(np) + tid]);
        }
    }
    __syncthreads();
    CompactAccess ta;
    ta.t = &t; ta.c.base = smem; ta.c.h = ch;
    Rules ql = q;
    ql.rpow_di = small;
    ql.panel_x = small + kMaxTail; ql.panel_w = small + kMaxTail + 32; ql.panel_p = small + kMaxTail + 64;

    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= (unsigned int)n_items) break;
        const int p = item_point[item];
        const double z = pz[p], sep = psep[p], mu = item_mu[item];
        const Geometry g = geometry(bg_chi(t.bg, z), sep, mu);
        const DIConst k = item_setup[item].k;
        if (out_evals && lane == 0) out_evals[item] = di_fast_eval_count(ql, g, k);
        double di = di_fast_partial(ta, t, (const LosRec *)los_s, ql, g, k, grp < ppi ? grp : ppi, ppi, j);
        di = warp_sum(di);
        if (lane == 0) out[item] = di;
        __syncwarp();
    }
}

/* items of a batch: item (point p, node j of the mu rule of p's separation); one thread per point */
__global__ void expand_items_kernel(const int *__restrict__ point_rule, const int *__restrict__ point_begin, const int n_points,
                                    const int *__restrict__ rule_begin, const double *__restrict__ rule_mu, const double *__restrict__ rule_w,
                                    int *__restrict__ item_point, double *__restrict__ item_mu, double *__restrict__ item_w)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_points) return;
    const int r = point_rule[p], b = rule_begin[r], n = rule_begin[r + 1] - b, dst = point_begin[p];
    for (int j = 0; j < n; ++j) { item_point[dst + j] = p; item_mu[dst + j] = rule_mu[b + j]; item_w[dst + j] = rule_w[b + j]; }
}

/* integrand evaluations of every item of one list, accumulated per point (integers: the sum does not depend on the order);
   mode 1 = local x integrated (two halves), mode 2 = integrated x integrated.  Used to cut cost-balanced blocks for the GPUs. */
__global__ void item_cost_kernel(const __grid_constant__ Tables t, const __grid_constant__ Options o, const __grid_constant__ Rules q,
                                 const int mode, const double *__restrict__ pz, const double *__restrict__ psep,
                                 const int *__restrict__ item_point, const double *__restrict__ item_mu, const int n_items,
                                 const unsigned int weight, unsigned long long *__restrict__ cost)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const int p = item_point[i];
    const Geometry g = geometry(bg_chi(t.bg, pz[p]), psep[p], item_mu[i]);
    const EndVals e1 = end_eval(o, t, g.chi1, 1), e2 = end_eval(o, t, g.chi2, 2);
    unsigned int c;
    if (mode == 1) c = o.midpoint ? si_eval_count(q, g, si_half_setup(o, g, e1, e2, true)) + si_eval_count(q, g, si_half_setup(o, g, e2, e1, false))
                                  : si_group_eval_count(q, g, si_half_setup(o, g, e1, e2, true)) + si_group_eval_count(q, g, si_half_setup(o, g, e2, e1, false));
    else c = di_fast_eval_count(q, g, di_setup(o, e1, e2));
    atomicAdd(&cost[p], (unsigned long long)c * weight);
}

/* ---- mu reduction: one warp per point, fixed summation order ---------------------------------- */
/* pm[(point*3 + piece)*n_l + il]; piece 0 (NI) is copied from ni[]; pieces 1 and 2 each have their own mu rule */
__global__ void reduce_multipoles_kernel(const int n_points,
                                         const int *__restrict__ si_begin, const double *__restrict__ si_mu, const double *__restrict__ si_w,
                                         const double *__restrict__ si, const int si_half,
                                         const int *__restrict__ di_begin, const double *__restrict__ di_mu, const double *__restrict__ di_w,
                                         const double *__restrict__ di, const int di_half,
                                         const double *__restrict__ ni, const LList ls, double *__restrict__ pm)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_points) return;
    for (int il = 0; il < ls.n; ++il) {
        double a = 0.0, c = 0.0;
        if (!(si_half && (ls.l[il] & 1)))
            for (int i = si_begin[warp] + lane; i < si_begin[warp + 1]; i += 32) a += si_w[i] * legendre_p(ls.l[il], si_mu[i]) * si[i];
        if (!(di_half && (ls.l[il] & 1)))
            for (int i = di_begin[warp] + lane; i < di_begin[warp + 1]; i += 32) c += di_w[i] * legendre_p(ls.l[il], di_mu[i]) * di[i];
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
                                      const double *__restrict__ zmin, const double *__restrict__ zmax, const double *__restrict__ gsep,
                                      const int n_groups, const double *__restrict__ gl_x, const double *__restrict__ gl_w, const int n_z,
                                      double *__restrict__ pz, double *__restrict__ psep, double *__restrict__ pw)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_groups * n_z) return;
    const int iz = i % n_z, ig = i / n_z;
    const double r = gsep[ig];
    const double z1 = spline_eval(t.bg.z_of_chi, bg_chi(t.bg, zmin[ig]) + r / 2.);
    const double z2 = spline_eval(t.bg.z_of_chi, bg_chi(t.bg, zmax[ig]) - r / 2.);
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
    q.n_outer = ctx->qc.n_outer; q.n_outer_min = ctx->qc.n_outer_min; q.n_outer_step = ctx->qc.n_outer_step;
    q.panel_w_mpc = ctx->qc.panel_w; q.panel_du = ctx->qc.panel_du; q.eps_min = ctx->qc.eps_min; q.tail_ratio = ctx->qc.tail_ratio;
    q.panel_w_di = ctx->qc.panel_w_di; q.panel_du_di = ctx->qc.panel_du_di;
    q.panel_x = q.panel_w = q.panel_p = nullptr;
    q.panel_x_si = q.panel_w_si = q.panel_p_si = nullptr;
    q.rpow_di = ctx->d_rpow_di; q.tail_ratio_di = ctx->qc.tail_ratio_di; q.geo_shift_di = ctx->qc.geo_shift_di;
    q.n_ren = ctx->qc.n_ren;
    q.tail_lnr = ctx->qc.tail_log ? std::log(ctx->qc.tail_ratio) : 0.0;
    q.tail_lnr_di = ctx->qc.tail_log_di ? std::log(ctx->qc.tail_ratio_di) : 0.0;
    q.geo_L = ctx->qc.geo_growth > 1.0 ? std::log(ctx->qc.geo_growth) : 0.0;
    q.geo_L_di = ctx->qc.geo_growth_di > 1.0 ? std::log(ctx->qc.geo_growth_di) : 0.0;
    rules_reciprocals(q);
    return q;
}

/* (re)derive the rules in force: defaults of the accuracy target, then the caller's non-zero knobs.  want_pieces: the call returns
   the pieces themselves (coffe_integrate per flag_integral), which an automatic target always serves at piece accuracy. */
/* How large the integrated terms are against the density term, up to geometry: max over the knots of the uploaded bias tables of
   |2 - 5 s(z)| / b(z), worst of the two populations.  2 for the default biases (b = 1, s = 0), under which the two line-of-sight
   pieces are at most a few per cent of every multipole of the benchmarked grid -- the premise of the total target. */
double integrated_share_factor(const coffe_b200_ctx *ctx)
{
    if (!ctx->have_tables) return 2.0;
    const Tables &t = ctx->t_host;
    double worst = 0.0;
    for (int pop = 0; pop < 2; ++pop) {
        const SplineView b = t.bias[pop == 0 ? B_GAL1 : B_GAL2], s = t.bias[pop == 0 ? B_MAG1 : B_MAG2];
        if (!b.x || !s.x || b.n < 2 || s.n < 2) continue;
        for (int i = 0; i < b.n; ++i) {
            const double z = b.x[i], bz = std::fabs(spline_eval(b, z)), sz = spline_eval(s, std::min(std::max(z, s.x[0]), s.x[s.n - 1]));
            worst = std::max(worst, std::fabs(2.0 - 5.0 * sz) / std::max(bz, 1e-3));
        }
    }
    return worst;
}

int apply_quadrature(coffe_b200_ctx *ctx, const bool want_pieces)
{
    int target = ctx->user_q.accuracy_target;
    /* automatic: the total target needs the density term to set the scale of the multipoles (RSD alone is an order of magnitude
       smaller) and biases under which the integrated terms stay small against it: twice the default ratio at most */
    if (target != COFFE_B200_TARGET_PIECE && target != COFFE_B200_TARGET_TOTAL)
        target = (!want_pieces && ctx->o.den && !ctx->o.midpoint && integrated_share_factor(ctx) <= 4.0) ? COFFE_B200_TARGET_TOTAL
                                                                                                         : COFFE_B200_TARGET_PIECE;
    QuadConfig base = target == COFFE_B200_TARGET_TOTAL ? total_target_quadrature() : default_quadrature();
    /* potential terms without lensing: the integrated x integrated piece is then smooth in mu but its l = 4 moment sets the scale of
       what 5 nodes per r_perp panel leave (1.0e-5 of the piece at r = 30 Mpc; 6 nodes: 1.9e-6, profiles/r2_piece_sweep_b.log) */
    if (target != COFFE_B200_TARGET_TOTAL && !ctx->o.len && (ctx->o.g4 || ctx->o.g5)) base.n_mu_panel_di = 6;
    const QuadConfig merged = merge_quadrature(base, ctx->user_q);
    if (!quadrature_valid(merged, kMaxPanels)) {
        ctx->err = "quadrature sizes out of range (node counts 2..128, nodes per line-of-sight panel 2..32; the line-of-sight panel plan must fit " + std::to_string(kMaxPanels) + " panels)";
        return COFFE_B200_EINVAL;
    }
    if (std::memcmp(&merged, &ctx->qc, sizeof(QuadConfig)) == 0 && target == ctx->target_in_force) return COFFE_B200_OK;
    if (merged.tail_ratio != ctx->qc.tail_ratio) {
        const std::vector<double> rp = build_rpow(merged.tail_ratio);
        CK(cudaSetDevice(ctx->device));
        CK(cudaMemcpyAsync(ctx->d_rpow, rp.data(), kMaxTail * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (merged.tail_ratio_di != ctx->qc.tail_ratio_di) {
        const std::vector<double> rp = build_rpow(merged.tail_ratio_di);
        CK(cudaSetDevice(ctx->device));
        CK(cudaMemcpyAsync(ctx->d_rpow_di, rp.data(), kMaxTail * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->qc = merged;
    ctx->target_in_force = target;
    ctx->generation++;
    return COFFE_B200_OK;
}

/* host-side range check of one point: the reference aborts (non-finite integrand) outside the tables */
/* chi = comoving distance to z when the caller has it (a grid call asks for the same redshift n_sep times), else < 0 */
int check_point(coffe_b200_ctx *ctx, const double z, const double sep, const bool need_renorm, double chi = -1.0)
{
    const Tables &t = ctx->t_host;
    const double zmax = t.bg.z[t.bg.n - 1];
    if (!(z > 0.0) || !(z < zmax) || !(sep > 0.0)) { ctx->err = "redshift or separation outside the background table"; return COFFE_B200_ERANGE; }
    if (chi < 0.0) chi = bg_chi(t.bg, z);
    if (!(chi - sep / 2. > 0.0)) { ctx->err = "separation larger than twice the comoving distance to z_mean"; return COFFE_B200_ERANGE; }
    if (!(chi + sep / 2. < ctx->layout.chi_max_table)) { ctx->err = "chi_mean + sep/2 outside the background table"; return COFFE_B200_ERANGE; }
    if (!(sep <= ctx->layout.r_max_table)) { ctx->err = "separation outside the I^n_l tables"; return COFFE_B200_ERANGE; }
    if (ctx->o.midpoint && ctx->layout.mid_nz > 0) {
        /* the 2-D tables are read at redshifts up to z(chi_mean + sep/2) (the reference's grid ends at z = 2.97, src/integrals.c:679-689) */
        const double z_hi = spline_eval(t.bg.z_of_chi, chi + sep / 2.);
        if (!(z_hi <= t.mid.z[t.mid.nz - 1])) { ctx->err = "z(chi_mean + sep/2) beyond the redshift grid of the 2-D tables"; return COFFE_B200_ERANGE; }
    }
    if (need_renorm) {
        const double rmax = t.ren2d.x[t.ren2d.n - 1];
        if (!(chi + sep / 2. <= rmax)) {
            ctx->err = "chi_mean + sep/2 outside the renormalisation grid (it extends to chi(max(z_mean + deltaz)), src/integrals.c:858-871)";
            return COFFE_B200_ERANGE;
        }
    }
    return COFFE_B200_OK;
}

/* midpoint approximation: needs the 2-D tables; the reference's flat-sky variants of it (src/functions.c:136-155 ff., 4046-4248) and
   its covariance branch (src/covariance.c, P(k, z) per bin) are not built */
int check_midpoint(coffe_b200_ctx *ctx)
{
    const Options &o = ctx->o;
    if (!o.midpoint) return COFFE_B200_OK;
    if (ctx->layout.mid_nz == 0) { ctx->err = "the midpoint approximation needs the 2-D tables I^n_l(z, r) (coffe_b200_integral.values2d)"; return COFFE_B200_ENOTABLES; }
    if (o.flatsky_local || o.flatsky_local_nonlocal || o.flatsky_nonlocal) { ctx->err = "the midpoint approximation is not available together with the flat-sky switches"; return COFFE_B200_EINVAL; }
    return COFFE_B200_OK;
}

bool needs_renorm(const Options &o) { return o.d2 || o.g1 || o.g2 || o.g3 || o.g4 || o.g5; }

/* does any term read renormalisation(chi_a, chi_b) up to chi_mean + sep/2?  The local potential terms (d2, g1..g3) at the end
   points and against g4/g5 along the line of sight, and the g4-g4 / g5-g5 / g4-g5 terms of the integrated x integrated piece
   (DIConst.any_ren) -- the latter also when no local effect is on.  The kernels clamp the cell index and would extrapolate the
   patch beyond the grid; the reference aborts there, so such a point is refused (ERANGE). */
bool renorm_range_needed(const Options &o)
{
    return o.d2 || o.g1 || o.g2 || o.g3 || ((o.g4 || o.g5) && (!o.only_cross || (o.g4 && o.g5)));
}

/* the (point, mu) work items of one integrated piece, on the device */
struct ItemList {
    int n = 0;
    DevBuf<int> point, begin;       /* item -> point; point -> first item (n_points + 1) */
    DevBuf<double> mu, w, val;
    DevBuf<unsigned char> setup;    /* SIItemSetup / DIItemSetup per item, (re)written by item_setup_kernel in front of every launch */
};

/* a batch of points with their mu items, on the device */
struct Batch {
    int n_points = 0;
    DevBuf<double> pz, psep, ni, pm;
    ItemList si, di;                /* local x integrated and integrated x integrated pieces: separate mu rules */
    DevBuf<unsigned int> evals;     /* optional: per-item integrand evaluation counts of the list being launched */
    double chi_max = 0;     /* largest chi2 over the batch: how much of the line-of-sight table to stage */
    bool di_half = false;   /* the integrated x integrated items cover mu > 0 only (mirror symmetry): odd multipoles of the piece are zero */
    bool si_half = false;   /* the same for the local x integrated items */
    void release()
    {
        pz.release(); psep.release(); ni.release(); pm.release(); evals.release();
        for (ItemList *L : {&si, &di}) { L->point.release(); L->begin.release(); L->mu.release(); L->w.release(); L->val.release(); L->setup.release(); L->n = 0; }
        n_points = 0;
    }
};

/* mode 1: local x integrated over b.si; mode 2: integrated x integrated over b.di */
int launch_integrated(coffe_b200_ctx *ctx, Batch &b, const int mode, unsigned int *evals = nullptr)
{
    ItemList &L = mode == 1 ? b.si : b.di;
    if (L.n == 0) return COFFE_B200_OK;
    const bool active = mode == 1 ? opt_has_si(ctx->o) : opt_has_integrated(ctx->o);
    if (!active) {
        CK(cudaMemsetAsync(L.val.p, 0, sizeof(double) * L.n, ctx->stream));
        return COFFE_B200_OK;
    }
    int stage_n = (int)std::ceil(b.chi_max / ctx->layout.los_h) + 3;
    if (stage_n > ctx->layout.los_n) stage_n = ctx->layout.los_n;
    const size_t los_bytes = (size_t)kLosStride * stage_n * sizeof(double2);
    if (ctx->variant == 1 && ctx->compact.h.valid && !ctx->o.midpoint && (mode == 1 ? ctx->qc.n_panel_si <= 32 : ctx->qc.n_panel_di <= 32)) {
        /* shared-memory tables: blob prefix covering R <= 2 chi_max, line-of-sight table, per-warp state */
        const int stage_bytes = compact_stage_bytes(ctx->compact, 2.0 * b.chi_max);
        CompactHeader ch = ctx->compact.h;
        ch.n_knots = (stage_bytes - ch.off_knots) / kKnotBytes;
        const size_t fixed = (size_t)stage_bytes + los_bytes + kDISmallDoubles * sizeof(double);
        /* per warp: one slot (geometry + half) per group of the local x integrated kernel */
        const size_t per_warp = mode == 1 ? (size_t)(32 / ctx->qc.n_panel_si) * sizeof(SIGroupSlot) : 0;
        int warps = mode == 1 ? kSIThreads / 32 : kDIThreads / 32;
        const int threads = warps * 32;
        const size_t need = fixed + (size_t)warps * per_warp;
        if (warps >= 4 && need <= (size_t)ctx->smem_optin) {
            CK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int), ctx->stream));
            /* end-point values and amplitudes of every item, one thread each */
            const size_t setup_bytes = (size_t)L.n * (mode == 1 ? sizeof(SIItemSetup) : sizeof(DIItemSetup));
            if (L.setup.n != setup_bytes) CK(L.setup.alloc(setup_bytes));
            if (mode == 1)
                item_setup_kernel<1><<<(L.n + kSetupThreads - 1) / kSetupThreads, kSetupThreads, 0, ctx->stream>>>(
                    ctx->t, ctx->o, b.pz.p, b.psep.p, L.point.p, L.mu.p, L.n, reinterpret_cast<SIItemSetup *>(L.setup.p), nullptr);
            else
                item_setup_kernel<2><<<(L.n + 127) / 128, 128, 0, ctx->stream>>>(
                    ctx->t, ctx->o, b.pz.p, b.psep.p, L.point.p, L.mu.p, L.n, nullptr, reinterpret_cast<DIItemSetup *>(L.setup.p));
            CK(cudaGetLastError());
            ctx->last_launches += 1;
            int grid = std::min(ctx->n_sm, (L.n + warps - 1) / warps);
            if (grid < 1) grid = 1;
            if (mode == 1) {
                /* every half adds itself to its item's value (and count) */
                CK(cudaMemsetAsync(L.val.p, 0, sizeof(double) * L.n, ctx->stream));
                if (evals) CK(cudaMemsetAsync(evals, 0, sizeof(unsigned int) * L.n, ctx->stream));
                auto kernel = si_items_smem_kernel;
                CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
                kernel<<<grid, threads, need, ctx->stream>>>(
                    ctx->t, ctx->o, make_rules(ctx), ch, reinterpret_cast<const uint4 *>(ctx->d_blob), stage_bytes,
                    reinterpret_cast<const SIItemSetup *>(L.setup.p), L.n, stage_n, L.val.p, evals, ctx->d_counter);
            }
            else {
                auto kernel = di_items_smem_kernel;
                CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
                kernel<<<grid, threads, need, ctx->stream>>>(
                    ctx->t, ctx->o, make_rules(ctx), ch, reinterpret_cast<const uint4 *>(ctx->d_blob), stage_bytes, b.pz.p, b.psep.p,
                    L.point.p, L.mu.p, reinterpret_cast<const DIItemSetup *>(L.setup.p), L.n, stage_n, L.val.p, evals, ctx->d_counter);
            }
            CK(cudaGetLastError());
            ctx->last_launches += 1;
            ctx->last_variant = 1;
            return COFFE_B200_OK;
        }
    }
    ctx->last_variant = 0;
    size_t smem = los_bytes;
    if (smem > 160 * 1024) { ctx->err = "line-of-sight table does not fit in shared memory"; return COFFE_B200_EINVAL; }
    CK(cudaFuncSetAttribute(integrated_items_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    integrated_items_kernel<<<L.n, kItemThreads, smem, ctx->stream>>>(
        ctx->t, ctx->o, make_rules(ctx), mode, b.pz.p, b.psep.p, L.point.p, L.mu.p, L.n, stage_n, L.val.p, evals);
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
    CK(cudaEventRecord(ctx->ev_k[0], ctx->stream));
    int rc = launch_integrated(ctx, b, 1);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_k[1], ctx->stream));
    CK(cudaEventRecord(ctx->ev_k[2], ctx->stream));
    rc = launch_integrated(ctx, b, 2);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_k[3], ctx->stream));
    ctx->have_kernel_times = true;
    const int blocks = (b.n_points * 32 + 127) / 128;
    reduce_multipoles_kernel<<<blocks, 128, 0, ctx->stream>>>(b.n_points, b.si.begin.p, b.si.mu.p, b.si.w.p, b.si.val.p, b.si_half ? 1 : 0,
                                                              b.di.begin.p, b.di.mu.p, b.di.w.p, b.di.val.p, b.di_half ? 1 : 0,
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

/* mu items of one piece for a list of points.  The mu rule depends on the separation only: the host builds one
   rule per distinct separation (`sep`, n_sep entries; point p uses rule point_rule[p]) and uploads those; the
   (point, mu, weight) items are expanded on the device. */
int build_item_list(coffe_b200_ctx *ctx, const double *sep, const int n_sep, const std::vector<int> &point_rule, const DevBuf<int> &d_point_rule,
                    const int piece, const bool half_only, const bool active, const int slot, ItemList &L)
{
    const int n = (int)point_rule.size();
    std::vector<int> point_begin(n + 1, 0);
    coffe_b200_ctx::MuRuleCache &mc = ctx->mu_cache[slot];
    const bool hit = mc.valid && mc.piece == piece && mc.half_only == half_only &&
                     mc.active == active && (int)mc.sep.size() == n_sep && std::memcmp(&mc.qc, &ctx->qc, sizeof(QuadConfig)) == 0 &&
                     (n_sep == 0 || std::memcmp(mc.sep.data(), sep, sizeof(double) * n_sep) == 0);
    if (!hit) {
        mc.valid = false;
        mc.rule_begin.assign(n_sep + 1, 0);
        mc.rule_mu.clear(); mc.rule_w.clear();
        std::vector<double> mu, w;
        for (int r = 0; r < n_sep; ++r) {
            mc.rule_begin[r] = (int)mc.rule_mu.size();
            if (!active) continue;
            build_mu_rule(ctx->qc, sep[r], mu, w, piece, half_only);
            mc.rule_mu.insert(mc.rule_mu.end(), mu.begin(), mu.end());
            mc.rule_w.insert(mc.rule_w.end(), w.begin(), w.end());
        }
        mc.rule_begin[n_sep] = (int)mc.rule_mu.size();
        mc.sep.assign(sep, sep + n_sep);
        std::memcpy(&mc.qc, &ctx->qc, sizeof(QuadConfig));
        mc.piece = piece; mc.half_only = half_only; mc.active = active;
        mc.valid = true;
    }
    const std::vector<int> &rule_begin = mc.rule_begin;
    const std::vector<double> &rule_mu = mc.rule_mu, &rule_w = mc.rule_w;
    static const bool trace = std::getenv("COFFE_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    long total = 0;
    for (int p = 0; p < n; ++p) { point_begin[p] = (int)total; total += rule_begin[point_rule[p] + 1] - rule_begin[point_rule[p]]; }
    if (total > 0x7fffffffL) { ctx->err = "too many (point, mu) work items for one call"; return COFFE_B200_EINVAL; }
    point_begin[n] = (int)total;
    L.n = (int)total;
    DevBuf<int> d_rule_begin;
    DevBuf<double> d_rule_mu, d_rule_w;
    CK(d_rule_begin.upload(rule_begin, ctx->stream));
    CK(d_rule_mu.upload(rule_mu, ctx->stream));
    CK(d_rule_w.upload(rule_w, ctx->stream));
    CK(L.begin.upload(point_begin, ctx->stream));
    CK(L.point.alloc(L.n));
    CK(L.mu.alloc(L.n));
    CK(L.w.alloc(L.n));
    CK(L.val.alloc(L.n));
    if (L.n > 0) {
        expand_items_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_point_rule.p, L.begin.p, n, d_rule_begin.p, d_rule_mu.p, d_rule_w.p,
                                                                   L.point.p, L.mu.p, L.w.p);
        CK(cudaGetLastError());
    }
    if (trace) std::fprintf(stderr, "coffe_b200 trace: item list of %d items: uploads, allocations and expansion launch %.3f ms\n", L.n,
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    /* the staging buffers above are freed in stream order (cudaFreeAsync) when they go out of scope */
    return COFFE_B200_OK;
}

/* both item lists of a batch of points; point p has separation sep[point_rule[p]] */
int build_multipole_batch(coffe_b200_ctx *ctx, const double *sep, const int n_sep, const std::vector<int> &point_rule, Batch &b)
{
    b.n_points = (int)point_rule.size();
    DevBuf<int> d_point_rule;
    CK(d_point_rule.upload(point_rule, ctx->stream));
    b.si_half = ctx->pops_identical && (ctx->mag_constant || !(ctx->o.d1 && ctx->o.len));
    int rc = build_item_list(ctx, sep, n_sep, point_rule, d_point_rule, 1, b.si_half, opt_has_si(ctx->o), 0, b.si);
    if (rc) return rc;
    b.di_half = ctx->di_mu_symmetric;
    return build_item_list(ctx, sep, n_sep, point_rule, d_point_rule, 2, b.di_half,
                           opt_has_integrated(ctx->o), 1, b.di);
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

/* the *_pieces entry points: piece-accuracy rules for the duration of the call (see coffe_b200_quadrature.accuracy_target) */
struct PieceRules {
    coffe_b200_ctx *ctx;
    int rc;
    explicit PieceRules(coffe_b200_ctx *c) : ctx(c), rc(apply_quadrature(c, true)) {}
    ~PieceRules() { apply_quadrature(ctx, false); }
};

/* a prepared multipole computation: the (z, r) items [begin, end) with their mu rules resident in HBM */
struct coffe_b200_plan {
    coffe_b200_ctx *ctx = nullptr;  /* cleared when the context is destroyed first: the plan is then inert */
    Batch b;
    LList ls;
    long n = 0;
    unsigned long generation = 0;   /* ctx->generation the item lists, stage sizes and mirror-symmetry decisions were built under */
    DevBuf<double> full;            /* gathered result of all ranks (coffe_b200_plan_execute_gather) */
    long n_total = 0;
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
    rc = check_midpoint(ctx);
    if (rc) return rc;
    pl.ctx = ctx; pl.n = end - begin; pl.ls = make_llist(l, n_l); pl.generation = ctx->generation;
    if (end == begin) return COFFE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    std::vector<double> pz, psep;
    std::vector<int> point_rule;
    const bool ren = needs_renorm(ctx->o);
    if (ren && !ctx->layout.has_divergent) { ctx->err = "d2/g1..g5 need the divergent I^4_0 tables (par->divergent)"; return COFFE_B200_EINVAL; }
    pz.reserve(pl.n); psep.reserve(pl.n); point_rule.reserve(pl.n);
    static const bool trace = std::getenv("COFFE_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    const bool ren_range = ren && renorm_range_needed(ctx->o);
    const double z_table_max = ctx->t_host.bg.z[ctx->t_host.bg.n - 1];
    long iz_at = -1;
    double chi_at = -1.0;           /* comoving distance to the redshift of row iz_at: one lookup per row, not per point */
    for (long it = begin; it < end; ++it) {
        const long iz = it / n_sep;
        const double z = z_mean[iz], s = sep[it - iz * n_sep];
        if (iz != iz_at) { iz_at = iz; chi_at = (z > 0.0 && z < z_table_max) ? bg_chi(ctx->t_host.bg, z) : -1.0; }
        rc = check_point(ctx, z, s, ren_range, chi_at);
        if (rc) return rc;
        pz.push_back(z); psep.push_back(s); point_rule.push_back((int)(it - iz * n_sep));
        const double chi2 = chi_at + s / 2.;
        if (chi2 > pl.b.chi_max) pl.b.chi_max = chi2;
    }
    const double t_check = ms_since();
    CK(pl.b.pz.upload(pz, ctx->stream));
    CK(pl.b.psep.upload(psep, ctx->stream));
    const double t_up = ms_since();
    rc = build_multipole_batch(ctx, sep, n_sep, point_rule, pl.b);
    if (rc) return rc;
    const double t_batch = ms_since();
    CK(pl.b.ni.alloc((size_t)pl.b.n_points * n_l));
    CK(pl.b.pm.alloc((size_t)pl.b.n_points * 3 * n_l));
    if (trace) std::fprintf(stderr, "coffe_b200 trace: plan_build: range checks %.3f ms, point upload %.3f ms, item lists %.3f ms, output buffers %.3f ms\n",
                            t_check, t_up - t_check, t_batch - t_up, ms_since() - t_batch);
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
    const size_t n = (size_t)pl.n * 3 * pl.ls.n;
    if (pl.n == 0) { hm.clear(); return COFFE_B200_OK; }
    if (ctx->h_stage_cap < n) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr; ctx->h_stage_cap = 0;
        const size_t cap = n + n / 4;
        if (cudaMallocHost(reinterpret_cast<void **>(&ctx->h_stage), cap * sizeof(double)) == cudaSuccess) ctx->h_stage_cap = cap;
        else { ctx->h_stage = nullptr; (void)cudaGetLastError(); }          /* no pinned memory to be had: copy into the vector itself */
    }
    if (ctx->h_stage) {
        CK(cudaMemcpyAsync(ctx->h_stage, pl.b.pm.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        hm.assign(ctx->h_stage, ctx->h_stage + n);
    }
    else {
        hm.assign(n, 0.0);
        CK(cudaMemcpyAsync(hm.data(), pl.b.pm.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return check_finite(ctx, hm.data(), hm.size());
}

/* shared body of multipoles / multipoles_range / multipoles_pieces */
int multipoles_core(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep, const int *l, int n_l,
                    long begin, long end, std::vector<double> &hm)
{
    /* COFFE_B200_TRACE=1: host-side phase times of the call on stderr (set-up and launch, result copy queued, device done) */
    static const bool trace = std::getenv("COFFE_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&](const std::chrono::steady_clock::time_point &a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    coffe_b200_plan pl;
    int rc = plan_build(ctx, z_mean, n_z, sep, n_sep, l, n_l, begin, end, pl);
    if (rc) return rc;
    const double t_build = ms_since(t0);
    rc = plan_run(pl);
    if (rc) return rc;
    const double t_run = ms_since(t0);
    rc = plan_fetch(pl, hm);          /* the copy is ordered after the kernels on the same stream */
    if (rc) return rc;
    const double t_fetch = ms_since(t0);
    rc = plan_sync(pl);
    if (trace) std::fprintf(stderr, "coffe_b200 trace: plan_build %.3f ms, launches queued at %.3f ms, fetch returned at %.3f ms, synchronised at %.3f ms\n",
                            t_build, t_run, t_fetch, ms_since(t0));
    return rc;
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
    for (int i = 0; i < 4; ++i)
        if ((e = cudaEventCreate(&ctx->ev_k[i])) != cudaSuccess) return fail("cudaEventCreate", e);
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
        const std::vector<double> rpd = build_rpow(ctx->qc.tail_ratio_di);
        if ((e = cudaMalloc(&ctx->d_rpow_di, kMaxTail * sizeof(double))) != cudaSuccess) return fail("cudaMalloc", e);
        if ((e = cudaMemcpy(ctx->d_rpow_di, rpd.data(), kMaxTail * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
        if ((e = cudaMalloc(&ctx->d_counter, sizeof(unsigned int))) != cudaSuccess) return fail("cudaMalloc", e);
        int v = 0;
        if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return fail("cudaDeviceGetAttribute", e);
        ctx->n_sm = v;
        if ((e = cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) return fail("cudaDeviceGetAttribute", e);
        ctx->smem_optin = v;
    }
    {
        /* keep the stream-ordered pool warm between calls */
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            /* lazy module loading would otherwise put the load of the two large kernels inside the first call */
            {
                cudaFuncAttributes fa;
                cudaFuncGetAttributes(&fa, di_items_smem_kernel);
                cudaFuncGetAttributes(&fa, si_items_smem_kernel);
                cudaFuncGetAttributes(&fa, item_setup_kernel<1>);
                cudaFuncGetAttributes(&fa, item_setup_kernel<2>);
                cudaFuncGetAttributes(&fa, ni_multipoles_kernel);
                cudaFuncGetAttributes(&fa, reduce_multipoles_kernel);
                cudaFuncGetAttributes(&fa, expand_items_kernel);
                cudaGetLastError();
            }
            /* first use of the pool costs ~40 ms (measured inside the first call): pay it here, once per context */
            /* ... and reserve what a full-grid call takes from it (~1 GB of item set-up records for 32 x 1000 points; growing the pool
               by that much measured 100-250 ms inside the first call).  COFFE_B200_POOL_MB overrides the reservation (0 = none). */
            size_t reserve_mb = 1024;
            if (const char *e = std::getenv("COFFE_B200_POOL_MB")) reserve_mb = (size_t)std::max(0L, std::atol(e));
            void *warm = nullptr;
            if (cudaMallocAsync(&warm, std::max<size_t>(reserve_mb, 1) << 20, ctx->stream) == cudaSuccess) { cudaFreeAsync(warm, ctx->stream); cudaStreamSynchronize(ctx->stream); }
            else cudaGetLastError();
        }
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
    {
        /* plans that outlive their context become inert: their device buffers go back to the pool while the stream still exists */
        std::lock_guard<std::mutex> lock(ctx->mu);
        AllocScope scope(ctx->stream);
        for (coffe_b200_plan *pl : ctx->live_plans) { pl->b.release(); pl->full.release(); pl->ctx = nullptr; }
        ctx->live_plans.clear();
        cudaStreamSynchronize(ctx->stream);
    }
    if (ctx->comm) coffe_b200_comm_destroy(ctx);
    for (int i = 0; i < 2; ++i)
        if (ctx->ev_g[i]) cudaEventDestroy(ctx->ev_g[i]);
    if (ctx->d_arena) cudaFree(ctx->d_arena);
    if (ctx->d_gl_x) cudaFree(ctx->d_gl_x);
    if (ctx->d_gl_w) cudaFree(ctx->d_gl_w);
    if (ctx->d_rpow) cudaFree(ctx->d_rpow);
    if (ctx->d_rpow_di) cudaFree(ctx->d_rpow_di);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->d_blob) cudaFree(ctx->d_blob);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 4; ++i)
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *coffe_b200_last_error(const coffe_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

/* host half of a table upload: validation, per-interval coefficients, arena layout, shared-memory table build -- independent of the
   device, so a group does it once for all its members */
struct PreparedTables {
    Arena A;
    Layout L;
    CompactBuild compact;
    bool has_compact = false;
    double omega_matter = 0, h = 0;
    bool di_mu_symmetric = false, pops_identical = false, mag_constant = false;
};

int coffe_internal_prepare_tables(const coffe_b200_tables *T, PreparedTables &P, std::string &err)
{
    if (!assemble(*T, P.A, P.L, err)) return COFFE_B200_EINVAL;
    if (!(T->omega_matter > 0) || !(T->h > 0)) { err = "omega_matter and h must be positive"; return COFFE_B200_EINVAL; }
    P.omega_matter = T->omega_matter; P.h = T->h;
    P.di_mu_symmetric = tables_di_symmetric(*T);
    P.pops_identical = tables_populations_identical(*T);
    P.mag_constant = tables_magnification_constant(*T);
    P.has_compact = false;
    if (P.L.has_ifam) {
        /* the shared-memory tables of the integrated-pieces kernels are built from the exact interval cubics of the host arena; the
           line-of-sight table is not needed for that */
        const Tables th = make_view(P.L, P.A.buf.data());
        std::string why;
        P.has_compact = build_compact(th, P.L, kCompactKnotBudget, kCompactTolerance, P.compact, why);
    }
    return COFFE_B200_OK;
}

/* device half: copy the arena and the blob, build the line-of-sight table there.  `P` may be shared between contexts (read only);
   with `own` the context takes the buffers instead of copying them */
int coffe_internal_commit_tables(coffe_b200_ctx *ctx, PreparedTables &P, const bool own)
{
    CK(cudaSetDevice(ctx->device));
    /* the device buffers are kept across uploads and only grow: a new cosmology per call is the MCMC inner loop, and cudaFree /
       cudaMalloc beside a warm stream-ordered pool measured up to 100 ms there */
    ctx->have_tables = false;
    if (ctx->d_arena_cap < P.A.buf.size()) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_arena) { cudaFree(ctx->d_arena); ctx->d_arena = nullptr; ctx->d_arena_cap = 0; }
        CK(cudaMalloc(&ctx->d_arena, P.A.buf.size() * sizeof(double)));
        ctx->d_arena_cap = P.A.buf.size();
    }
    CK(cudaMemcpyAsync(ctx->d_arena, P.A.buf.data(), P.A.buf.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const Layout &L = P.L;
    ctx->layout = L;
    ctx->t = make_view(L, ctx->d_arena);
    los_build_kernel<<<(L.los_n + 127) / 128, 128, 0, ctx->stream>>>(ctx->t, reinterpret_cast<double2 *>(ctx->d_arena + L.los), L.los_n, L.los_h,
                                                                     L.chi_max_table);
    CK(cudaGetLastError());
    if (P.has_compact) {
        if (ctx->d_blob_cap < P.compact.blob.size()) {
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->d_blob) { cudaFree(ctx->d_blob); ctx->d_blob = nullptr; ctx->d_blob_cap = 0; }
            CK(cudaMalloc(&ctx->d_blob, P.compact.blob.size() * sizeof(double)));
            ctx->d_blob_cap = P.compact.blob.size();
        }
        CK(cudaMemcpyAsync(ctx->d_blob, P.compact.blob.data(), P.compact.blob.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (own) { ctx->arena = std::move(P.A); ctx->compact = std::move(P.compact); }
    else { ctx->arena = P.A; ctx->compact = P.compact; }
    if (!P.has_compact) { ctx->compact = CompactBuild(); ctx->compact.h.valid = 0; }
    ctx->t_host = make_view(L, ctx->arena.buf.data());
    ctx->omega_matter = P.omega_matter; ctx->h = P.h;
    ctx->o.Om = P.omega_matter;
    ctx->o.H0h = P.h / 2997.92458;          /* COFFE_H0 = 1/2997.92458 h/Mpc, src/common.h:75-77 */
    ctx->o.divergent = L.has_divergent;
    ctx->di_mu_symmetric = P.di_mu_symmetric;
    ctx->pops_identical = P.pops_identical;
    ctx->mag_constant = P.mag_constant;
    ctx->have_tables = true;
    if (const int rq = apply_quadrature(ctx, false)) return rq;       /* the automatic accuracy target looks at the bias tables */
    ctx->generation++;
    return COFFE_B200_OK;
}

int coffe_b200_upload_tables(coffe_b200_ctx *ctx, const coffe_b200_tables *T)
{
    if (!ctx || !T) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    static const bool trace = std::getenv("COFFE_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    PreparedTables P;
    int rc = coffe_internal_prepare_tables(T, P, ctx->err);
    if (rc) return rc;
    const double t_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    rc = coffe_internal_commit_tables(ctx, P, true);
    if (trace) std::fprintf(stderr, "coffe_b200 trace: upload_tables: host assembly + shared-memory tables %.3f ms, device copies + line-of-sight table %.3f ms\n",
                            t_host, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() - t_host);
    return rc;
}

int coffe_b200_set_options(coffe_b200_ctx *ctx, const coffe_b200_options *opt)
{
    if (!ctx || !opt) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    Options &o = ctx->o;
    o.den = opt->den != 0; o.rsd = opt->rsd != 0; o.len = opt->len != 0; o.d1 = opt->d1 != 0; o.d2 = opt->d2 != 0;
    o.g1 = opt->g1 != 0; o.g2 = opt->g2 != 0; o.g3 = opt->g3 != 0; o.g4 = opt->g4 != 0; o.g5 = opt->g5 != 0;
    o.only_cross = opt->only_cross_correlations != 0;
    o.flatsky_local = opt->flatsky_local != 0;
    o.flatsky_local_nonlocal = opt->flatsky_local_nonlocal != 0;
    o.flatsky_nonlocal = opt->flatsky_nonlocal != 0;
    o.midpoint = opt->midpoint != 0;
    ctx->generation++;
    return apply_quadrature(ctx, false);
}

int coffe_b200_set_quadrature(coffe_b200_ctx *ctx, const coffe_b200_quadrature *q)
{
    if (!ctx || !q) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    if (q->accuracy_target < 0 || q->accuracy_target > 2) { ctx->err = "accuracy_target must be 0 (automatic), 1 (piece) or 2 (total)"; return COFFE_B200_EINVAL; }
    const coffe_b200_quadrature keep = ctx->user_q;
    ctx->user_q = *q;
    const int rc = apply_quadrature(ctx, false);
    if (rc) ctx->user_q = keep;
    return rc;
}

int coffe_b200_set_kernel_variant(coffe_b200_ctx *ctx, int variant)
{
    if (!ctx || (variant != 0 && variant != 1)) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    ctx->variant = variant;
    ctx->generation++;
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
    q->n_mu_panel_di = c.n_mu_panel_di; q->panel_width_di = c.panel_w_di; q->panel_uniform_extent_di = c.panel_du_di;
    q->mu_panel_width_di = c.mu_panel_w_di; q->tail_ratio_di = c.tail_ratio_di; q->geometric_shift_di = c.geo_shift_di;
    q->geometric_growth = c.geo_growth; q->geometric_growth_di = c.geo_growth_di; q->n_renorm = c.n_ren;
    q->tail_map = c.tail_log ? 2 : 1; q->tail_map_di = c.tail_log_di ? 2 : 1;
    q->mu_geometric_growth = c.mu_growth; q->mu_geometric_growth_di = c.mu_growth_di; q->mu_switch_break_di = c.mu_switch_di ? 1 : 2;
    q->n_outer_min = c.n_outer_min; q->n_outer_step = c.n_outer_step; q->mu_tail_ratio = c.mu_tail;
    q->accuracy_target = ctx->target_in_force;
    return COFFE_B200_OK;
}

int coffe_b200_multipoles_pieces(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                                 const int *l, int n_l, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    PieceRules pr(ctx);
    if (pr.rc) return pr.rc;
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
    COFFE_ENTER(ctx);
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
    COFFE_ENTER(ctx);
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

/* 2PCF pieces of the flattened output entries (iz*n_mu + imu)*n_sep + ir in [begin, end): pieces[piece*(end-begin) + k] */
static int corrfunc_core(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *mu, int n_mu,
                         const double *sep, int n_sep, long begin, long end, std::vector<double> &pieces)
{
    if (!ctx->have_tables || !ctx->layout.has_ifam) { ctx->err = "no tables uploaded (the signal needs the I^n_l family)"; return COFFE_B200_ENOTABLES; }
    if (n_z < 0 || n_sep < 0 || n_mu < 0 || (n_z > 0 && !z_mean) || (n_sep > 0 && !sep) || (n_mu > 0 && !mu)) { ctx->err = "bad grid arguments"; return COFFE_B200_EINVAL; }
    ctx->last_ms = 0; ctx->last_launches = 0;
    const long total = (long)n_z * n_mu * n_sep;
    if (begin < 0 || end > total || begin > end) { ctx->err = "entry range outside the grid"; return COFFE_B200_EINVAL; }
    const size_t n_out = (size_t)(end - begin);
    pieces.assign(3 * n_out, 0.0);
    if (n_out == 0) return COFFE_B200_OK;
    { const int rcm = check_midpoint(ctx); if (rcm) return rcm; }
    for (int i = 0; i < n_mu; ++i)
        if (!(mu[i] > -1.0 && mu[i] < 1.0)) { ctx->err = "mu must lie strictly inside (-1, 1)"; return COFFE_B200_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    const bool ren = needs_renorm(ctx->o);
    if (ren && !ctx->layout.has_divergent) { ctx->err = "d2/g1..g5 need the divergent I^4_0 tables (par->divergent)"; return COFFE_B200_EINVAL; }
    Batch b;
    std::vector<double> pz, psep, item_mu;
    std::vector<int> item_point;
    /* points: the (iz, ir) pairs the range touches, in (iz, ir) order; items in output order (iz, imu, ir) */
    const int iz_lo = (int)(begin / ((long)n_mu * n_sep)), iz_hi = (int)((end - 1) / ((long)n_mu * n_sep));
    for (int iz = iz_lo; iz <= iz_hi; ++iz)
        for (int ir = 0; ir < n_sep; ++ir) {
            pz.push_back(z_mean[iz]); psep.push_back(sep[ir]);
        }
    std::vector<char> used(pz.size(), 0);
    item_point.reserve(n_out); item_mu.reserve(n_out);
    for (long e = begin; e < end; ++e) {
        const int ir = (int)(e % n_sep), im = (int)((e / n_sep) % n_mu), iz = (int)(e / ((long)n_sep * n_mu));
        const int p = (iz - iz_lo) * n_sep + ir;
        item_point.push_back(p); item_mu.push_back(mu[im]);
        used[p] = 1;
    }
    for (size_t p = 0; p < pz.size(); ++p) {
        if (!used[p]) continue;
        const int rc = check_point(ctx, pz[p], psep[p], ren && renorm_range_needed(ctx->o));
        if (rc) return rc;
        const double chi2 = bg_chi(ctx->t_host.bg, pz[p]) + psep[p] / 2.;
        if (chi2 > b.chi_max) b.chi_max = chi2;
    }
    b.n_points = (int)pz.size();
    const int n_items = (int)item_mu.size();
    CK(b.pz.upload(pz, ctx->stream));
    CK(b.psep.upload(psep, ctx->stream));
    for (ItemList *L : {&b.si, &b.di}) {
        L->n = n_items;
        CK(L->point.upload(item_point, ctx->stream));
        CK(L->mu.upload(item_mu, ctx->stream));
        CK(L->val.alloc(n_items));
    }
    CK(b.ni.alloc(n_items));
    ctx->o.out_multipoles = 0;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (opt_has_ni(ctx->o)) {
        ni_points_kernel<<<(n_items + 127) / 128, 128, 0, ctx->stream>>>(ctx->t, ctx->o, b.pz.p, b.psep.p, b.si.point.p, b.si.mu.p, n_items, b.ni.p);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    else CK(cudaMemsetAsync(b.ni.p, 0, sizeof(double) * n_items, ctx->stream));
    int rc = launch_integrated(ctx, b, 1);
    if (rc) return rc;
    rc = launch_integrated(ctx, b, 2);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaMemcpyAsync(pieces.data(), b.ni.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(pieces.data() + n_out, b.si.val.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(pieces.data() + 2 * n_out, b.di.val.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    return check_finite(ctx, pieces.data(), pieces.size());
}

int coffe_internal_corrfunc_range(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *mu, int n_mu,
                                  const double *sep, int n_sep, long begin, long end, double *out_sum)
{
    /* The 2-point function at a fixed mu always runs under the piece rules: the total target sizes the rules of the integrated
       pieces against the l = 0..4 scale of a multipole, i.e. after the mu integral; at a single (mu, r) the local terms can be
       close to a zero crossing and the caller compares the value with itself. */
    PieceRules pr(ctx);
    if (pr.rc) return pr.rc;
    std::vector<double> pc;
    const int rc = corrfunc_core(ctx, z_mean, n_z, mu, n_mu, sep, n_sep, begin, end, pc);
    if (rc) return rc;
    const size_t n = (size_t)(end - begin);
    for (size_t i = 0; i < n; ++i) out_sum[i] = pc[i] + pc[n + i] + pc[2 * n + i];
    return COFFE_B200_OK;
}

int coffe_b200_corrfunc(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *mu, int n_mu,
                        const double *sep, int n_sep, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    return coffe_internal_corrfunc_range(ctx, z_mean, n_z, mu, n_mu, sep, n_sep, 0, (long)n_z * n_mu * n_sep, out);
}

int coffe_b200_corrfunc_range_pieces(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *mu, int n_mu,
                                     const double *sep, int n_sep, long begin, long end, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    PieceRules pr(ctx);
    if (pr.rc) return pr.rc;
    std::vector<double> pc;
    const int rc = corrfunc_core(ctx, z_mean, n_z, mu, n_mu, sep, n_sep, begin, end, pc);
    if (rc) return rc;
    std::memcpy(out, pc.data(), pc.size() * sizeof(double));
    return COFFE_B200_OK;
}

/* redshift-averaged multipoles of the flattened (ibin*n_sep + ir) groups [begin, end): hm[(k*3 + piece)*n_l + il] */
int coffe_internal_average_core(coffe_b200_ctx *ctx, const double *z_min, const double *z_max, int n_bins,
                                const double *sep, int n_sep, const int *l, int n_l, long begin, long end, std::vector<double> &hm)
{
    if (!ctx->have_tables || !ctx->layout.has_ifam) { ctx->err = "no tables uploaded (the signal needs the I^n_l family)"; return COFFE_B200_ENOTABLES; }
    if (n_bins < 0 || n_sep < 0 || (n_bins > 0 && (!z_min || !z_max)) || (n_sep > 0 && !sep)) { ctx->err = "bad grid arguments"; return COFFE_B200_EINVAL; }
    int rc = check_l(ctx, l, n_l);
    if (rc) return rc;
    ctx->last_ms = 0; ctx->last_launches = 0;
    const long total = (long)n_bins * n_sep;
    if (begin < 0 || end > total || begin > end) { ctx->err = "group range outside the grid"; return COFFE_B200_EINVAL; }
    const int n_groups = (int)(end - begin), nz = ctx->qc.n_z;
    hm.assign((size_t)n_groups * 3 * n_l, 0.0);
    if (n_groups == 0) return COFFE_B200_OK;
    { const int rcm = check_midpoint(ctx); if (rcm) return rcm; }
    CK(cudaSetDevice(ctx->device));
    const bool ren = needs_renorm(ctx->o);
    if (ren && !ctx->layout.has_divergent) { ctx->err = "d2/g1..g5 need the divergent I^4_0 tables (par->divergent)"; return COFFE_B200_EINVAL; }
    Batch b;
    std::vector<int> point_rule((size_t)n_groups * nz);
    std::vector<double> g_zmin(n_groups), g_zmax(n_groups), g_sep(n_groups);
    const Tables &th = ctx->t_host;
    for (long gi = begin; gi < end; ++gi) {
        const int ib = (int)(gi / n_sep), ir = (int)(gi % n_sep), k = (int)(gi - begin);
        if (!(z_min[ib] > 0.0) || !(z_max[ib] > z_min[ib]) || !(z_max[ib] < th.bg.z[th.bg.n - 1])) { ctx->err = "bad redshift bin"; return COFFE_B200_ERANGE; }
        const double c1 = bg_chi(th.bg, z_min[ib]), c2 = bg_chi(th.bg, z_max[ib]);
        if (!(sep[ir] > 0.0) || !(c1 + sep[ir] / 2. < c2 - sep[ir] / 2.)) {
            ctx->err = "separation does not fit inside the redshift bin (need chi(z_min) + r/2 < chi(z_max) - r/2)";
            return COFFE_B200_ERANGE;
        }
        if (ren && renorm_range_needed(ctx->o) && !(c2 <= th.ren2d.x[th.ren2d.n - 1])) { ctx->err = "redshift bin outside the renormalisation grid"; return COFFE_B200_ERANGE; }
        for (int j = 0; j < nz; ++j) point_rule[(size_t)k * nz + j] = ir;
        g_zmin[k] = z_min[ib]; g_zmax[k] = z_max[ib]; g_sep[k] = sep[ir];
        if (c2 > b.chi_max) b.chi_max = c2;
    }
    /* points on the device: one (zmin, zmax, sep) triple per group */
    DevBuf<double> d_zmin, d_zmax, d_sep, pw, gm;
    CK(d_zmin.upload(g_zmin, ctx->stream));
    CK(d_zmax.upload(g_zmax, ctx->stream));
    CK(d_sep.upload(g_sep, ctx->stream));
    const int n_points = n_groups * nz;
    CK(b.pz.alloc(n_points));
    CK(b.psep.alloc(n_points));
    CK(pw.alloc(n_points));
    CK(gm.alloc((size_t)n_groups * 3 * n_l));
    rc = build_multipole_batch(ctx, sep, n_sep, point_rule, b);
    if (rc) return rc;
    CK(b.ni.alloc((size_t)b.n_points * n_l));
    CK(b.pm.alloc((size_t)b.n_points * 3 * n_l));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    average_points_kernel<<<(n_points + 127) / 128, 128, 0, ctx->stream>>>(ctx->t, ctx->o.H0h, d_zmin.p, d_zmax.p, d_sep.p, n_groups,
                                                                          ctx->d_gl_x, ctx->d_gl_w, nz, b.pz.p, b.psep.p, pw.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    rc = run_multipole_batch(ctx, b, make_llist(l, n_l), false);
    if (rc) return rc;
    average_reduce_kernel<<<(n_groups * 3 * n_l + 127) / 128, 128, 0, ctx->stream>>>(b.pm.p, pw.p, n_groups, nz, n_l, gm.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaMemcpyAsync(hm.data(), gm.p, hm.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    return check_finite(ctx, hm.data(), hm.size());
}

int coffe_b200_average_multipoles(coffe_b200_ctx *ctx, const double *z_min, const double *z_max, int n_bins,
                                  const double *sep, int n_sep, const int *l, int n_l, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    std::vector<double> hm;
    const int rc = coffe_internal_average_core(ctx, z_min, z_max, n_bins, sep, n_sep, l, n_l, 0, (long)n_bins * n_sep, hm);
    if (rc) return rc;
    for (int ib = 0; ib < n_bins; ++ib)
        for (int il = 0; il < n_l; ++il)
            for (int ir = 0; ir < n_sep; ++ir) {
                const size_t g = (size_t)ib * n_sep + ir;
                out[((size_t)ib * n_l + il) * n_sep + ir] = hm[(g * 3 + 0) * n_l + il] + hm[(g * 3 + 1) * n_l + il] + hm[(g * 3 + 2) * n_l + il];
            }
    return COFFE_B200_OK;
}

int coffe_b200_average_multipoles_range_pieces(coffe_b200_ctx *ctx, const double *z_min, const double *z_max, int n_bins,
                                               const double *sep, int n_sep, const int *l, int n_l, long begin, long end, double *out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    PieceRules pr(ctx);
    if (pr.rc) return pr.rc;
    std::vector<double> hm;
    const int rc = coffe_internal_average_core(ctx, z_min, z_max, n_bins, sep, n_sep, l, n_l, begin, end, hm);
    if (rc) return rc;
    const size_t n = (size_t)(end - begin);
    for (size_t k = 0; k < n; ++k)
        for (int piece = 0; piece < 3; ++piece)
            for (int il = 0; il < n_l; ++il) out[((size_t)piece * n_l + il) * n + k] = hm[(k * 3 + piece) * n_l + il];
    return COFFE_B200_OK;
}

int coffe_b200_last_timing(const coffe_b200_ctx *ctx, double *kernel_ms, int *n_launches)
{
    if (!ctx) return COFFE_B200_EINVAL;
    if (kernel_ms) *kernel_ms = ctx->last_ms;
    if (n_launches) *n_launches = ctx->last_launches;
    return COFFE_B200_OK;
}

int coffe_b200_last_kernel_times(coffe_b200_ctx *ctx, double *si_ms, double *di_ms)
{
    if (!ctx) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    if (si_ms) *si_ms = 0.0;
    if (di_ms) *di_ms = 0.0;
    if (!ctx->have_kernel_times) return COFFE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, ctx->ev_k[0], ctx->ev_k[1]));
    CK(cudaEventElapsedTime(&b, ctx->ev_k[2], ctx->ev_k[3]));
    if (si_ms) *si_ms = a;
    if (di_ms) *di_ms = b;
    return COFFE_B200_OK;
}

/* ---- prepared computations (bench: inputs resident in HBM before the timed region) ---------------- */
int coffe_b200_plan_multipoles(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                               const int *l, int n_l, long item_begin, long item_end, coffe_b200_plan **out)
{
    if (!ctx || !out) return COFFE_B200_EINVAL;
    *out = nullptr;
    COFFE_ENTER(ctx);
    coffe_b200_plan *pl = new coffe_b200_plan();
    const int rc = plan_build(ctx, z_mean, n_z, sep, n_sep, l, n_l, item_begin, item_end, *pl);
    if (rc) { delete pl; return rc; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { delete pl; ctx->err = "upload failed"; return COFFE_B200_ECUDA; }
    ctx->live_plans.push_back(pl);
    *out = pl;
    return COFFE_B200_OK;
}

int coffe_b200_plan_execute(coffe_b200_plan *plan, double *kernel_ms)
{
    if (!plan || !plan->ctx) return COFFE_B200_EINVAL;
    COFFE_ENTER(plan->ctx);
    if (plan->generation != plan->ctx->generation) {
        plan->ctx->err = "the plan was prepared under different tables, options or quadrature settings: prepare it again";
        return COFFE_B200_EINVAL;
    }
    int rc = plan_run(*plan);
    if (rc) return rc;
    rc = plan_sync(*plan);
    if (kernel_ms) *kernel_ms = plan->ctx->last_ms;
    return rc;
}

int coffe_b200_plan_fetch(coffe_b200_plan *plan, double *out)
{
    if (!plan || !out || !plan->ctx) return COFFE_B200_EINVAL;
    COFFE_ENTER(plan->ctx);
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
    if (!plan || !plan->ctx) return COFFE_B200_EINVAL;
    coffe_b200_ctx *ctx = plan->ctx;
    COFFE_ENTER(ctx);
    if (plan->generation != ctx->generation) {
        ctx->err = "the plan was prepared under different tables, options or quadrature settings: prepare it again";
        return COFFE_B200_EINVAL;
    }
    Batch &b = plan->b;
    double si = 0, di = 0;
    if (opt_has_integrated(ctx->o)) {
        CK(cudaSetDevice(ctx->device));
        ctx->o.out_multipoles = 1;
        for (int mode = 1; mode <= 2; ++mode) {
            ItemList &L = mode == 1 ? b.si : b.di;
            if (L.n == 0) continue;
            CK(b.evals.alloc(L.n));
            CK(cudaMemsetAsync(b.evals.p, 0, sizeof(unsigned int) * L.n, ctx->stream));
            const int rc = launch_integrated(ctx, b, mode, b.evals.p);
            if (rc) return rc;
            std::vector<unsigned int> h(L.n);
            CK(cudaMemcpyAsync(h.data(), b.evals.p, h.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            double tot = 0;
            for (int i = 0; i < L.n; ++i) tot += h[i];
            (mode == 1 ? si : di) = tot;
        }
        b.evals.release();
    }
    if (n_items) *n_items = (double)b.si.n + (double)b.di.n;
    if (n_ni_evals) *n_ni_evals = opt_has_ni(ctx->o) ? (double)b.n_points * ctx->qc.n_mu_ni : 0.0;
    if (n_si_evals) *n_si_evals = si;
    if (n_di_evals) *n_di_evals = di;
    return COFFE_B200_OK;
}

void coffe_b200_plan_destroy(coffe_b200_plan *plan)
{
    if (!plan) return;
    if (coffe_b200_ctx *ctx = plan->ctx) {
        COFFE_ENTER(ctx);
        cudaSetDevice(ctx->device);
        for (size_t i = 0; i < ctx->live_plans.size(); ++i)
            if (ctx->live_plans[i] == plan) { ctx->live_plans.erase(ctx->live_plans.begin() + i); break; }
        plan->full.release();
        plan->b.release();          /* device buffers back to the pool in stream order, under the context's allocation scope */
    }
    delete plan;
}

/* achievable FP64 FMA rate of this device [TFLOP/s]: the roofline denominator */
int coffe_b200_fp64_peak(coffe_b200_ctx *ctx, double *tflops)
{
    if (!ctx || !tflops) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    CK(cudaSetDevice(ctx->device));
    const int blocks = ctx->n_sm * 8, threads = 256, iters = 1 << 16;
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

/* ================================================================================================= */
/* ---- cost model and block boundaries ---------------------------------------------------------------- */
int coffe_b200_balance_bounds(const double *cost, long n_items, int n_parts, long *bounds)
{
    if (!bounds || n_parts < 1 || n_items < 0 || (n_items > 0 && !cost)) return COFFE_B200_EINVAL;
    for (int r = 0; r <= n_parts; ++r) bounds[r] = 0;
    if (n_items == 0) return COFFE_B200_OK;
    std::vector<double> c((size_t)n_items + 1, 0.0);
    for (long i = 0; i < n_items; ++i) {
        if (!(cost[i] >= 0.0) || !std::isfinite(cost[i])) return COFFE_B200_EINVAL;
        c[i + 1] = c[i] + cost[i];
    }
    if (!(c[n_items] > 0.0)) return COFFE_B200_EINVAL;
    bounds[n_parts] = n_items;
    for (int r = 1; r < n_parts; ++r) {
        const double target = c[n_items] * r / n_parts;
        long k = (long)(std::lower_bound(c.begin(), c.end(), target) - c.begin());      /* first c[k] >= target */
        if (k > n_items) k = n_items;
        if (k > 0 && std::fabs(c[k] - target) > std::fabs(c[k - 1] - target)) --k;        /* the closer of the two neighbours */
        bounds[r] = std::max(k, bounds[r - 1]);
    }
    return COFFE_B200_OK;
}

int coffe_b200_rebalance_cost(double *cost, long n_items, int n_parts, const long *bounds, const double *part_ms)
{
    if (!cost || !bounds || !part_ms || n_parts < 1 || n_items < 0) return COFFE_B200_EINVAL;
    if (bounds[0] != 0 || bounds[n_parts] != n_items) return COFFE_B200_EINVAL;
    double total_cost = 0, total_ms = 0;
    std::vector<double> block(n_parts, 0.0);
    for (int r = 0; r < n_parts; ++r) {
        if (bounds[r + 1] < bounds[r] || !(part_ms[r] >= 0.0)) return COFFE_B200_EINVAL;
        for (long i = bounds[r]; i < bounds[r + 1]; ++i) block[r] += cost[i];
        total_cost += block[r]; total_ms += part_ms[r];
    }
    if (!(total_cost > 0.0) || !(total_ms > 0.0)) return COFFE_B200_OK;
    for (int r = 0; r < n_parts; ++r) {
        if (!(block[r] > 0.0) || !(part_ms[r] > 0.0)) continue;
        /* time per unit cost of this block relative to the average; damped so that one noisy measurement cannot overshoot */
        double s = (part_ms[r] / block[r]) / (total_ms / total_cost);
        s = std::min(1.25, std::max(0.8, s));
        for (long i = bounds[r]; i < bounds[r + 1]; ++i) cost[i] *= s;
    }
    return COFFE_B200_OK;
}

namespace {

/* cost[iz*n_sep + ir]: evaluations counted exactly on a few rows of the grid (every separation: the rules switch with r), linear in
   z between them (the z dependence is the slow growth of the tail panels with distance) */
int multipoles_cost_core(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep, const int *l, int n_l,
                         std::vector<double> &cost)
{
    cost.assign((size_t)n_z * n_sep, 1.0);
    if (n_z == 0 || n_sep == 0 || !opt_has_integrated(ctx->o)) return COFFE_B200_OK;
    std::vector<int> rows;
    const int n_rows = std::min(n_z, 9);
    for (int k = 0; k < n_rows; ++k) {
        const int iz = n_rows == 1 ? 0 : (int)std::lround((double)k * (n_z - 1) / (n_rows - 1));
        if (rows.empty() || rows.back() != iz) rows.push_back(iz);
    }
    std::vector<double> zs;
    for (int iz : rows) zs.push_back(z_mean[iz]);
    coffe_b200_plan pl;
    int rc = plan_build(ctx, zs.data(), (int)zs.size(), sep, n_sep, l, n_l, 0, (long)zs.size() * n_sep, pl);
    if (rc) return rc;
    Batch &b = pl.b;
    DevBuf<unsigned long long> d_cost;
    CK(d_cost.alloc((size_t)b.n_points));
    CK(cudaMemsetAsync(d_cost.p, 0, sizeof(unsigned long long) * b.n_points, ctx->stream));
    ctx->o.out_multipoles = 1;
    const Rules q = make_rules(ctx);
    /* a local x integrated evaluation costs about two integrated x integrated ones (ncu: 713 against 342 instructions) */
    if (opt_has_si(ctx->o) && b.si.n > 0)
        item_cost_kernel<<<(b.si.n + 127) / 128, 128, 0, ctx->stream>>>(ctx->t, ctx->o, q, 1, b.pz.p, b.psep.p, b.si.point.p, b.si.mu.p, b.si.n, 2u, d_cost.p);
    if (b.di.n > 0)
        item_cost_kernel<<<(b.di.n + 127) / 128, 128, 0, ctx->stream>>>(ctx->t, ctx->o, q, 2, b.pz.p, b.psep.p, b.di.point.p, b.di.mu.p, b.di.n, 1u, d_cost.p);
    CK(cudaGetLastError());
    std::vector<unsigned long long> h((size_t)b.n_points);
    CK(cudaMemcpyAsync(h.data(), d_cost.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int iz = 0, k = 0; iz < n_z; ++iz) {
        while (k + 1 < (int)rows.size() - 1 && rows[k + 1] <= iz) ++k;
        const int a = rows[k], bb = rows[std::min(k + 1, (int)rows.size() - 1)];
        const double t = bb > a ? (double)(iz - a) / (bb - a) : 0.0;
        const int ka = k, kb = std::min(k + 1, (int)rows.size() - 1);
        for (int ir = 0; ir < n_sep; ++ir)
            cost[(size_t)iz * n_sep + ir] = 1.0 + (1.0 - t) * (double)h[(size_t)ka * n_sep + ir] + t * (double)h[(size_t)kb * n_sep + ir];
    }
    return COFFE_B200_OK;
}

}  // namespace

int coffe_b200_multipoles_cost(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                               const int *l, int n_l, double *cost)
{
    if (!ctx || !cost) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    std::vector<double> c;
    const int rc = multipoles_cost_core(ctx, z_mean, n_z, sep, n_sep, l, n_l, c);
    if (rc) return rc;
    std::memcpy(cost, c.data(), c.size() * sizeof(double));
    return COFFE_B200_OK;
}

}  // extern "C"

#include "group.inc"
