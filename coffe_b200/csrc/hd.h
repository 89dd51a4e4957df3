// Host/device portability macros.  Every numerical routine of the hot path is written once as a
// `COFFE_HD` inline function so that nvcc compiles it into the sm_100a kernels; the same headers also
// compile under plain g++ for the CPU-only debugging harness in tests/hostsim (test infrastructure:
// the product library never executes these on the host).
#pragma once
/* branch hints: the persistent kernels stall on instruction fetch, rare paths should leave the main line */
#define COFFE_LIKELY(x) __builtin_expect(!!(x), 1)
#define COFFE_UNLIKELY(x) __builtin_expect(!!(x), 0)
#if defined(__CUDACC__)
#define COFFE_HD __host__ __device__ __forceinline__
#define COFFE_D __device__ __forceinline__
/* rarely executed helpers; kept inline: taken out of line (__noinline__) they measured 1.5 % slower (profiles/README.md) */
#define COFFE_COLD __host__ __device__ __forceinline__
#if defined(__CUDA_ARCH__)
#define COFFE_SYNC_LANES() __syncwarp()
#define COFFE_RSQRT(x) rsqrt(x)
#else
#define COFFE_SYNC_LANES() ((void)0)
#define COFFE_RSQRT(x) (1.0 / sqrt(x))
#endif
#else
#define COFFE_SYNC_LANES() ((void)0)
#define COFFE_RSQRT(x) (1.0 / std::sqrt(x))
#define COFFE_HD inline
#define COFFE_D inline
#define COFFE_COLD inline
#include <cmath>
#include <cstring>
struct alignas(8) int2 { int x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
#endif
