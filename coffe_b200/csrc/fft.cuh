// Hand-written FFT and complex special functions shared by the table-construction (integrals.cu) and 2-D FFTLog covariance
// (covariance.cu) kernels.  Device only.  The FFT is a Stockham autosort over two ping-pong buffers, radices 2, 3 and 5, run
// by one CTA per transform (the batch is the grid): N = 10 000 = 2^4 5^4 for the I^n_l tables, N = 8 000 = 2^6 5^3 for the
// covariance; twiddles from sincospi, so the accuracy does not depend on N.
#pragma once

namespace {

/* ---- complex helpers ------------------------------------------------------------------------------ */
struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(const cplx a, const cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cadd(const cplx a, const cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx clog_(const cplx a) { return {log(hypot(a.re, a.im)), atan2(a.im, a.re)}; }
__device__ __forceinline__ cplx cexp_(const cplx a)
{
    double s, c;
    sincos(a.im, &s, &c);
    const double e = exp(a.re);
    return {e * c, e * s};
}
__device__ __forceinline__ cplx cinv(const cplx a) { const double d = a.re * a.re + a.im * a.im; return {a.re / d, -a.im / d}; }

/* log Gamma(z), Re z > 0: shift to Re z >= 12 with the recurrence, then the Stirling series.  The imaginary part is
   defined modulo 2 pi (only exp of differences is used). */
__device__ cplx clgamma(cplx z)
{
    cplx prod = {1.0, 0.0};
    while (z.re < 12.0) { prod = cmul(prod, z); z.re += 1.0; }
    const cplx lz = clog_(z), zi = cinv(z), zi2 = cmul(zi, zi);
    /* (z - 1/2) ln z - z + ln(2 pi)/2 + sum B_2k / (2k (2k-1) z^(2k-1)) */
    cplx r = cmul({z.re - 0.5, z.im}, lz);
    r.re += -z.re + 0.91893853320467274178; r.im += -z.im;
    const double c[7] = {1.0 / 12.0, -1.0 / 360.0, 1.0 / 1260.0, -1.0 / 1680.0, 1.0 / 1188.0, -691.0 / 360360.0, 1.0 / 156.0};
    cplx acc = {c[6], 0.0};
    for (int i = 5; i >= 0; --i) { acc = cmul(acc, zi2); acc.re += c[i]; }
    r = cadd(r, cmul(acc, zi));
    const cplx lp = clog_(prod);
    return {r.re - lp.re, r.im - lp.im};
}

/* ---- FFT ---------------------------------------------------------------------------------------------- */
/* tw: exp(-2 pi i k / n), k < n, on the device (fft_twiddle_kernel): every twiddle of every pass is one table entry.  (Until round 2 each
   pass called sincospi per element and per butterfly term: the 2-D FFTLog covariance spent its 0.7 s there.) */
struct FftPlan { int n; int n_stage; int radix[24]; const cplx *tw; };

__global__ void fft_twiddle_kernel(const int n, cplx *__restrict__ tw)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s, c;
    sincospi(-2.0 * (double)k / (double)n, &s, &c);
    tw[k] = {c, s};
}

/* roots of unity of the butterflies: cos, sin(2 pi m / R), m < R */
template <int R> struct Roots;
template <> struct Roots<2> { __device__ static double c(int m) { return m ? -1.0 : 1.0; } __device__ static double s(int) { return 0.0; } };
template <> struct Roots<3> {
    __device__ static double c(int m) { return m ? -0.5 : 1.0; }
    __device__ static double s(int m) { return m == 0 ? 0.0 : m == 1 ? 0.86602540378443864676 : -0.86602540378443864676; }
};
template <> struct Roots<5> {
    __device__ static double c(int m) { return m == 0 ? 1.0 : (m == 1 || m == 4) ? 0.30901699437494742410 : -0.80901699437494742410; }
    __device__ static double s(int m)
    {
        return m == 0 ? 0.0 : m == 1 ? 0.95105651629515357212 : m == 2 ? 0.58778525229247312917 : m == 3 ? -0.58778525229247312917 : -0.95105651629515357212;
    }
};

/* one Stockham autosort pass of radix R over the n points of src -> dst; sign = -1 forward, +1 inverse */
template <int R>
__device__ void stockham_pass(const cplx *__restrict__ src, cplx *__restrict__ dst, const int n, const int ns, const double sign,
                              const cplx *__restrict__ tw)
{
    const int m = n / R;
    const int stride = n / (ns * R);            /* exp(-2 pi i k r / (ns R)) = tw[k r stride], k r < ns R */
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const int k = j % ns;
        cplx v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v[r] = src[j + r * m];
            if (r > 0 && k > 0) {
                cplx w = tw[k * r * stride];
                w.im *= -sign;                  /* the table holds the forward sign */
                v[r] = cmul(v[r], w);
            }
        }
        cplx o[R];
#pragma unroll
        for (int a = 0; a < R; ++a) {
            cplx acc = v[0];
#pragma unroll
            for (int r = 1; r < R; ++r) acc = cadd(acc, cmul(v[r], {Roots<R>::c((a * r) % R), sign * Roots<R>::s((a * r) % R)}));
            o[a] = acc;
        }
        const int j0 = (j / ns) * ns * R + k;
#pragma unroll
        for (int a = 0; a < R; ++a) dst[j0 + a * ns] = o[a];
    }
}

/* in-place-looking FFT over two ping-pong buffers; returns the buffer that holds the result */
__device__ cplx *fft_run(const FftPlan &p, cplx *a, cplx *b, const double sign)
{
    int ns = 1;
    for (int s = 0; s < p.n_stage; ++s) {
        const int R = p.radix[s];
        if (R == 2) stockham_pass<2>(a, b, p.n, ns, sign, p.tw);
        else if (R == 3) stockham_pass<3>(a, b, p.n, ns, sign, p.tw);
        else stockham_pass<5>(a, b, p.n, ns, sign, p.tw);
        ns *= R;
        __syncthreads();
        cplx *t = a; a = b; b = t;
    }
    return a;
}

bool make_fft_plan(int n, FftPlan &p)
{
    p.n = n; p.n_stage = 0; p.tw = nullptr;
    for (const int r : {5, 3, 2})
        while (n % r == 0) { if (p.n_stage >= 24) return false; p.radix[p.n_stage++] = r; n /= r; }
    return n == 1;
}

/* fill the twiddle table of a plan into `storage` (2 n doubles, owned by the caller) on `stream` */
inline cudaError_t fft_plan_twiddles(FftPlan &p, double *storage, cudaStream_t stream)
{
    fft_twiddle_kernel<<<(p.n + 255) / 256, 256, 0, stream>>>(p.n, reinterpret_cast<cplx *>(storage));
    p.tw = reinterpret_cast<const cplx *>(storage);
    return cudaGetLastError();
}

}  // namespace
