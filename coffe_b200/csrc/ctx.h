// Internal state shared by the translation units of libcoffe_b200 (engine.cu, covariance.cu).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "assemble.h"
#include "compact_build.h"
#include "di_fast.cuh"
#include "rules_build.h"

using namespace coffe;

struct coffe_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_k[4] = {nullptr, nullptr, nullptr, nullptr};   /* around the local x integrated / integrated x integrated kernels */
    bool have_kernel_times = false;
    std::string err;
    /* tables */
    bool have_tables = false;
    bool di_mu_symmetric = false;   /* assemble.h: tables_di_symmetric */
    bool pops_identical = false, mag_constant = false;   /* assemble.h: when the local x integrated piece is mirror symmetric too */
    Arena arena;            /* host copy (set-up and range checks only) */
    Layout layout;
    double *d_arena = nullptr;
    size_t d_arena_cap = 0, d_blob_cap = 0;      /* doubles allocated behind d_arena / d_blob */
    Tables t;               /* device view */
    Tables t_host;          /* host view */
    double omega_matter = 0, h = 0;
    /* options / quadrature */
    Options o;
    QuadConfig qc;          /* rules in force: library defaults for the accuracy target, overridden by user_q where it is non-zero */
    coffe_b200_quadrature user_q = {};    /* what coffe_b200_set_quadrature was given (zeros = library default) */
    int target_in_force = 1;              /* COFFE_B200_TARGET_PIECE / _TOTAL behind qc */
    /* the mu rules of the last call, one slot per integrated piece: they depend on the separations and the quadrature
       configuration only, and building them (two rules per separation, ~100 nodes of asin / cos / sin each) was 8 of the 9 ms of
       host work of a cfg5 call -- a sweep over cosmologies or redshifts on a fixed separation grid builds them once */
    struct MuRuleCache {
        std::vector<double> sep;
        QuadConfig qc;
        int piece = 0; bool half_only = false, active = false, valid = false;
        std::vector<int> rule_begin;
        std::vector<double> rule_mu, rule_w;
    } mu_cache[2];
    double *d_gl_x = nullptr, *d_gl_w = nullptr, *d_rpow = nullptr, *d_rpow_di = nullptr;
    /* shared-memory tables of the integrated-pieces kernel */
    CompactBuild compact;
    double *d_blob = nullptr;
    unsigned int *d_counter = nullptr;
    int variant = 1;        /* 1 = shared-memory tables (default), 0 = exact tables through L1/L2 */
    int last_variant = -1;
    int n_sm = 0, smem_optin = 0;
    /* pinned host buffer the results of a call land in (grown on demand): a device-to-host copy into pageable memory goes through the
       driver's own staging buffer in chunks */
    double *h_stage = nullptr;
    size_t h_stage_cap = 0;
    /* timing of the last call */
    double last_ms = 0;
    int last_launches = 0;
    /* bumped by every call that changes what a prepared plan was built from (tables, options, quadrature, kernel variant):
       a plan refuses to run against a different generation (coffe_b200_plan_execute -> EINVAL) */
    /* one process per GPU: NCCL communicator (group.inc), events around the gather */
    void *comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    cudaEvent_t ev_g[2] = {nullptr, nullptr};
    unsigned long generation = 0;
    std::vector<struct coffe_b200_plan *> live_plans;   /* released (device buffers freed, ctx cleared) by coffe_b200_destroy */
    std::mutex mu;
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return COFFE_B200_ECUDA;                                                                 \
        }                                                                                            \
    } while (0)

/* Device buffers of a call come from the stream-ordered allocator (cudaMallocAsync on the context's stream, pool kept
   warm: release threshold = max), so a compute call does not pay cudaMalloc / cudaFree (and their device-wide
   synchronisation) for its ~20 transient arrays.  The stream is published per thread by AllocScope at every entry point. */
inline cudaStream_t &tls_alloc_stream()
{
    static thread_local cudaStream_t s = nullptr;
    return s;
}

struct AllocScope {
    cudaStream_t prev;
    explicit AllocScope(cudaStream_t s) : prev(tls_alloc_stream()) { tls_alloc_stream() = s; }
    ~AllocScope() { tls_alloc_stream() = prev; }
};

#define COFFE_ENTER(c) std::lock_guard<std::mutex> lock((c)->mu); AllocScope alloc_scope_((c)->stream)

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;       /* stream the buffer was allocated on (null: plain cudaMalloc) */
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) { if (s) cudaFreeAsync(p, s); else cudaFree(p); }
        p = nullptr; n = 0;
    }
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        s = tls_alloc_stream();
        const size_t bytes = (count ? count : 1) * sizeof(T);
        return s ? cudaMallocAsync(reinterpret_cast<void **>(&p), bytes, s) : cudaMalloc(reinterpret_cast<void **>(&p), bytes);
    }
    cudaError_t upload(const std::vector<T> &v, cudaStream_t stream)
    {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess) return e;
        if (v.empty()) return cudaSuccess;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream);
    }
};
