// Flat-sky Gaussian covariance of the multipoles (coffe_covariance_init, src/covariance.c:477-1103).
//
// Stage A of the reference evaluates, for every pair l1 <= l2 and every pair of separations, the
// oscillatory integrals
//     G^(p)_{l1 l2}(r_m, r_n) = (2l1+1)(2l2+1) (2/pi) int_{kmin}^{kmax} dk k^2 S_p(k) jbar_l1(k r_m) jbar_l2(k r_n) / (p pi)
// (p = 1: S = spline of P; p = 2: S = spline of P^2; jbar = spherical Bessel function or its average over
// a separation bin, src/covariance.c:196-284, 316-351, 651-708) one at a time with QAG.  All of them share
// the same k nodes, so here they are ONE weighted Gram product: with V[k][(l, m)] = jbar_l(k r_m) on a
// fixed composite Gauss-Legendre grid in k (panels no wider than one period of the fastest product, break
// points at the knots of the P(k) spline),
//     G^(p) = V^T diag(w k^2 S_p) V,
// a (n_l n_sep)^2 x N_k FP64 contraction -- tiled in shared memory, split over k across CTAs, reduced in a
// fixed order.  Stage B (bias/growth coefficients, Wigner-3j Legendre integrals, shot noise, volume,
// src/covariance.c:790-1044) is one thread per output entry.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ctx.h"
#include "special.cuh"
#include "fft.cuh"

namespace {

constexpr int kTile = 64;          /* C tile edge */
constexpr int kKc = 16;            /* k nodes per shared-memory step */
constexpr int kMaxCovL = 8;

/* ---- special functions ---------------------------------------------------------------------------- */
__device__ double sine_integral(const double x)
{
    /* power series for x <= 2, modified Lentz continued fraction of E1(ix) beyond (x >= 0) */
    if (x <= 2.0) {
        double term = x, sum = x;
        for (int k = 1; k < 40; ++k) {
            term *= -x * x / ((2.0 * k) * (2.0 * k + 1.0));
            const double add = term / (2.0 * k + 1.0);
            sum += add;
            if (fabs(add) < 1e-18 * fabs(sum)) break;
        }
        return sum;
    }
    double br = 1.0, bi = x, cr = 1e300, ci = 0.0;
    double den = br * br + bi * bi;
    double dr = br / den, di = -bi / den, hr = dr, hi = di;
    for (int i = 1; i < 2000; ++i) {
        const double an = -(double)i * i;
        br += 2.0;
        double tr = an * dr + br, ti = an * di + bi;
        den = tr * tr + ti * ti;
        dr = tr / den; di = -ti / den;
        den = cr * cr + ci * ci;
        tr = br + an * cr / den; ti = bi - an * ci / den;
        cr = tr; ci = ti;
        const double delr = cr * dr - ci * di, deli = cr * di + ci * dr;
        tr = hr * delr - hi * deli; ti = hr * deli + hi * delr;
        hr = tr; hi = ti;
        if (fabs(delr - 1.0) + fabs(deli) < 1e-16) break;
    }
    double sx, cx;
    sincos(x, &sx, &cx);
    return 1.5707963267948966 + (hi * cx - hr * sx);
}

/* int^x s^2 j_l(s) ds, closed forms of src/covariance.c:196-238 (l <= 6) */
__device__ double bessel_power_integral(const int l, const double x)
{
    if (x < 0.75) {
        /* the closed forms cancel like x^-(l+2) at small x; ascending series of s^2 j_l(s) integrated term by term */
        double pre = x * x * x;
        for (int k = 1; k <= l; ++k) pre *= x / (2.0 * k + 1.0);
        const double q = -0.5 * x * x;
        double term = 1.0, sum = 1.0 / (l + 3.0);
        for (int k = 1; k < 40; ++k) {
            term *= q / (k * (2.0 * l + 2.0 * k + 1.0));
            const double add = term / (l + 2.0 * k + 3.0);
            sum += add;
            if (fabs(add) < 1e-18 * fabs(sum)) break;
        }
        /* the closed forms below are antiderivatives with F_l(0) = 0, -2, 0, -8, 0, -16, 0 */
        const double f0 = (l == 1) ? -2.0 : (l == 3) ? -8.0 : (l == 5) ? -16.0 : 0.0;
        return f0 + pre * sum;
    }
    double s, c;
    sincos(x, &s, &c);
    switch (l) {
        case 0: return -x * c + s;
        case 1: return -2. * c - x * s;
        case 2: return x * c - 4. * s + 3. * sine_integral(x);
        case 3: return 7. * c + ((-15. + x * x) * s) / x;
        case 4: return ((105. / x - 2. * x) * c + (22. - 105. / (x * x)) * s + 15. * sine_integral(x)) / 2.;
        case 5: return ((315 * x - 16 * x * x * x) * c - (315 - 105 * x * x + x * x * x * x) * s) / (x * x * x);
        case 6: {
            const double x2 = x * x, x4 = x2 * x2;
            return (x * (20790 - 1575 * x2 + 8 * x4) * c + (-20790 + 8505 * x2 - 176 * x4) * s + 105 * x4 * sine_integral(x)) / (8. * x4);
        }
        default: return 0.0;
    }
}

/* ---- stage A ---------------------------------------------------------------------------------------- */
/* node weights: W[2*i + p-1] = w k^2 S_p(k) (2/pi) / (p pi) */
__global__ void cov_node_weights_kernel(const double *__restrict__ k, const double *__restrict__ w, const int n,
                                        const SplineView s1, const SplineView s2, const int want1, const int want2, double *__restrict__ W)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double kk = k[i], pre = w[i] * kk * kk * 2.0 / 3.14159265358979323846 / 3.14159265358979323846;
    W[2 * i] = want1 ? pre * spline_eval(s1, kk) : 0.0;
    W[2 * i + 1] = want2 ? pre * spline_eval(s2, kk) / 2.0 : 0.0;
}

/* V[i*M + (il*n_sep + m)] = jbar_l(k_i r_m); columns M..Mpad-1 are zero padding */
__global__ void cov_bessel_kernel(const double *__restrict__ k, const int n_nodes, const double *__restrict__ sep, const int n_sep,
                                  const int *__restrict__ ls, const int n_l, const double delta, const int Mpad, double *__restrict__ V)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_nodes * Mpad) return;
    const int col = (int)(idx % Mpad);
    const size_t i = idx / Mpad;
    if (col >= n_l * n_sep) { V[idx] = 0.0; return; }
    const int l = ls[col / n_sep];
    const double r = sep[col % n_sep], kk = k[i];
    double v;
    if (delta == 0.0) v = sph_bessel(l, kk * r);
    else {
        /* average over the separation bin [r - delta/2, r + delta/2]          src/covariance.c:259-284 */
        const double a = r - delta / 2., b = r + delta / 2.;
        const double vol = (b * b * b - a * a * a) / 3.;
        v = (bessel_power_integral(l, kk * b) - bessel_power_integral(l, kk * a)) / (kk * kk * kk) / vol;
    }
    V[idx] = v;
}

/* partial Gram products over a k range: P[split][p][row][col] += sum_k W_p(k) V[k][row] V[k][col].
   The product is symmetric: every pair of tiles is computed in ONE orientation, (ta, tb) with ta <= tb -- rows from tile ta scaled
   by the weight, columns from tile tb -- and written to both places.  sym = 1 (a call that owns every row): one block per pair
   ta <= tb, 15 of the 25 tile pairs of cfg4.  sym = 0 (row ranges of the multi-GPU drivers): one block per tile (tr, tc) of the
   owned rows, which computes the pair in the same orientation and writes the transpose when tc < tr -- so that a row range
   reproduces the entries of the full call bit for bit. */
__global__ void __launch_bounds__(256, 2) cov_gram_kernel(const double *__restrict__ V, const double *__restrict__ W, const int n_nodes,
                                                       const int Mpad, const int n_splits, double *__restrict__ P, const int accumulate,
                                                       const int tr0, const int n_tr /* tile rows [tr0, tr0 + n_tr): the rows this device owns */,
                                                       const int sym)
{
    __shared__ double As[kKc][kTile], Bs[kKc][kTile], Ws[kKc][2];
    const int tiles = Mpad / kTile;
    const int per_split = sym ? tiles * (tiles + 1) / 2 : n_tr * tiles;
    const int tile = blockIdx.x % per_split, split = blockIdx.x / per_split;
    int ta, tb;
    bool write_direct, write_mirror;
    if (sym) {
        /* pair index -> (ta, tb), ta <= tb: row ta of the upper triangle holds tiles - ta pairs */
        int u = tile;
        ta = 0;
        while (u >= tiles - ta) { u -= tiles - ta; ++ta; }
        tb = ta + u;
        write_direct = true; write_mirror = ta != tb;
    }
    else {
        const int tr = tr0 + tile / tiles, tc = tile % tiles;
        ta = tr < tc ? tr : tc; tb = tr < tc ? tc : tr;
        write_direct = tc >= tr; write_mirror = tc < tr;
    }
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int per = (n_nodes + n_splits - 1) / n_splits;
    const int k0 = split * per, k1 = min(n_nodes, k0 + per);
    double acc1[4][4], acc2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc1[i][j] = 0.0; acc2[i][j] = 0.0; }
    /* 16 x 64 doubles per operand tile, 256 threads: 4 each; the next k tile is loaded into registers while the current one is
       multiplied (the loads come from L2 / HBM: without the overlap the kernel waited on them, long scoreboard 3.3 per issue) */
    double ra[4], rb[4], rw = 0.0;
    auto fetch = [&](const int kb) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = threadIdx.x + 256 * u, kk = e / kTile, c = e % kTile;
            const bool ok = kb + kk < k1;
            ra[u] = ok ? V[(size_t)(kb + kk) * Mpad + ta * kTile + c] : 0.0;
            rb[u] = ok ? V[(size_t)(kb + kk) * Mpad + tb * kTile + c] : 0.0;
        }
        if (threadIdx.x < kKc * 2) {
            const int kk = threadIdx.x / 2, p = threadIdx.x % 2;
            rw = kb + kk < k1 ? W[2 * (size_t)(kb + kk) + p] : 0.0;
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = threadIdx.x + 256 * u, kk = e / kTile, c = e % kTile;
            As[kk][c] = ra[u]; Bs[kk][c] = rb[u];
        }
        if (threadIdx.x < kKc * 2) Ws[threadIdx.x / 2][threadIdx.x % 2] = rw;
    };
    static_assert(kKc * kTile == 4 * 256, "four doubles of each operand tile per thread");
    if (k0 < k1) fetch(k0);
    for (int kb = k0; kb < k1; kb += kKc) {
        commit();
        __syncthreads();
        if (kb + kKc < k1) fetch(kb + kKc);
#pragma unroll
        for (int kk = 0; kk < kKc; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
            const double w1 = Ws[kk][0], w2 = Ws[kk][1];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a1 = a[i] * w1, a2 = a[i] * w2;
#pragma unroll
                for (int j = 0; j < 4; ++j) { acc1[i][j] = fma(a1, b[j], acc1[i][j]); acc2[i][j] = fma(a2, b[j], acc2[i][j]); }
            }
        }
        __syncthreads();
    }
    double *P1 = P + ((size_t)split * 2 + 0) * Mpad * Mpad, *P2 = P + ((size_t)split * 2 + 1) * Mpad * Mpad;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = ta * kTile + ty * 4 + i, c = tb * kTile + tx * 4 + j;
            if (write_direct) {
                const size_t o = (size_t)r * Mpad + c;
                if (accumulate) { P1[o] += acc1[i][j]; P2[o] += acc2[i][j]; }
                else { P1[o] = acc1[i][j]; P2[o] = acc2[i][j]; }
            }
            if (write_mirror) {
                const size_t o = (size_t)c * Mpad + r;
                if (accumulate) { P1[o] += acc1[i][j]; P2[o] += acc2[i][j]; }
                else { P1[o] = acc1[i][j]; P2[o] = acc2[i][j]; }
            }
        }
}

/* G[p][row][col] = sum over splits, fixed order */
__global__ void cov_reduce_kernel(const double *__restrict__ P, const int n_splits, const size_t n, double *__restrict__ G)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    const size_t p = i / n, e = i % n;
    double a = 0.0;
    for (int s = 0; s < n_splits; ++s) a += P[((size_t)s * 2 + p) * n + e];
    G[i] = a;
}

/* ---- stage B ---------------------------------------------------------------------------------------- */
struct CovBin {
    double volume, D1z, density1, density2, pixelsize;
};

struct CovAssembly {
    int n_bins, n_l, n_sep, Mpad;
    int l[kMaxCovL];
    int poisson, mixed, cosmic, window;
    int d13, d24, d14, d23;     /* Kronecker deltas of the populations */
    int row_begin, n_rows;      /* rows (i1*n_sep + m) of every bin's matrix this launch assembles */
};

__device__ __forceinline__ int sign_of(const int m) { return (abs(m) % 2 == 0) ? 1 : -1; }

/* out[((ib*n_rows + (row - row_begin))*n_l + i2)*n_sep + n], row = i1*n_sep + m     src/covariance.c:896-1044 */
__global__ void cov_assemble_kernel(const CovAssembly a, const CovBin *__restrict__ bins, const double *__restrict__ coef_cosmic,
                                    const double *__restrict__ coef_mixed, const double *__restrict__ G, const double *__restrict__ sep,
                                    double *__restrict__ out)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_bin = (size_t)a.n_rows * a.n_l * a.n_sep;
    if (idx >= per_bin * a.n_bins) return;
    const int ib = (int)(idx / per_bin);
    size_t r = idx % per_bin;
    const int n = (int)(r % a.n_sep); r /= a.n_sep;
    const int i2 = (int)(r % a.n_l); r /= a.n_l;
    const int row = a.row_begin + (int)r;
    const int m = row % a.n_sep, i1 = row / a.n_sep;
    const int l1 = a.l[i1], l2 = a.l[i2];
    const CovBin b = bins[ib];
    const size_t M2 = (size_t)a.Mpad * a.Mpad;
    /* the reference stores the (l_i <= l_j) integral with (chi1, chi2) = (sep[m], sep[n]) and mirrors the rest
       (src/covariance.c:651-708, 779-788); jbar_l1(k r_m) jbar_l2(k r_n) is the same number either way */
    const size_t g = (size_t)(i1 * a.n_sep + m) * a.Mpad + (i2 * a.n_sep + n);
    const double pre = (2 * l1 + 1) * (2 * l2 + 1);
    const double g1 = pre * G[g], g2 = pre * G[M2 + g];
    double v = 0.0;
    if (a.poisson && l1 == l2 && m == n) {
        const double pops = a.d13 * a.d24 + sign_of(l2) * a.d14 * a.d23;
        double geom;
        if (a.window) {
            const double lo = sep[m] - b.pixelsize / 2., hi = sep[m] + b.pixelsize / 2.;
            geom = 1. / ((hi * hi * hi - lo * lo * lo) / 3.);
        }
        else geom = 1. / sep[n] / sep[m] / b.pixelsize;
        v += (2 * l1 + 1) / 4. / 3.14159265358979323846 * pops * geom / b.density1 / b.density2 / b.volume;
    }
    if (a.mixed)
        v += sign_of((l2 - l1) / 2) * sign_of(l1 + l2) * b.D1z * b.D1z * g1 * coef_mixed[((size_t)ib * a.n_l + i1) * a.n_l + i2] / b.volume;
    if (a.cosmic)
        v += sign_of((l2 - l1) / 2) * b.D1z * b.D1z * b.D1z * b.D1z * g2 * coef_cosmic[((size_t)ib * a.n_l + i1) * a.n_l + i2] / b.volume;
    out[idx] = v;
}

/* ---- stage A, covariance_integration_method = fftlog: 2-D FFTLog (src/covariance.c:357-470, src/twobessel.c:16-113) -----
   The reference transforms the N x N matrix diag(k^3 P^p / dlnk) with a 2-D FFTLog.  A diagonal input makes the forward 2-D
   transform a 1-D one, F[(m + n) mod N]; what remains is the inverse 2-D transform of
       X[m][n] = conj(F[(m+n) mod N] A[m] B[n]),   A, B = windowed (k0 r0)^(-+ i eta) g_l(eta)            (src/twobessel.c:60-90)
   done as N/2+1 column transforms (one CTA each, the column is generated on the fly) followed by the real row transforms of
   the rows the requested separations need (+-50 knots: the reach of the natural-spline derivative estimates of the bicubic
   resampling at double precision). */
struct FftlogGeom { int N; double kmin, ratio, dlnk, nu; };

__device__ double fftlog_window(const long i, const long ncut)       /* src/utils.c:233-236 */
{
    return (double)i / ncut - sin(2. * i * 3.14159265358979323846 / ncut) / (2. * 3.14159265358979323846);
}

/* diag[p][i] = k_i^(3 - 2 nu) S(k_i)^(p+1) / dlnk, as complex numbers ready for the forward transform */
__global__ void cov_fftlog_diag_kernel(const FftlogGeom g, const SplineView pk, cplx *__restrict__ d1, cplx *__restrict__ d2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.N) return;
    const double k = g.kmin * pow(g.ratio, (double)i / g.N);
    const double P = spline_eval(pk, k), pre = pow(k, 3.0 - 2.0 * g.nu) / g.dlnk;
    d1[i] = {pre * P, 0.0};
    d2[i] = {pre * P * P, 0.0};
}

/* one CTA per transform: buf[b*N .. ] -> forward FFT in place (result copied back to buf) */
__global__ void __launch_bounds__(1024) cov_fft1d_kernel(const FftPlan plan, cplx *__restrict__ buf, cplx *__restrict__ work, const double sign)
{
    cplx *a = buf + (size_t)blockIdx.x * plan.n, *b = work + (size_t)blockIdx.x * plan.n;
    cplx *r = fft_run(plan, a, b, sign);
    if (r != a)
        for (int i = threadIdx.x; i < plan.n; i += blockDim.x) a[i] = r[i];
}

/* T[l_index][m], m = 0..N/2: (k0 r0)^(-i eta_m) g_l(eta_m), g_l = 2^z Gamma((l+z)/2) / Gamma((3+l-z)/2), z = nu + i eta   (src/utils.c:196-207) */
__global__ void cov_fftlog_gl_kernel(const FftlogGeom g, const int *__restrict__ ls, const int n_l, cplx *__restrict__ T)
{
    const int half = g.N / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_l * (half + 1)) return;
    const int il = idx / (half + 1), m = idx % (half + 1);
    const double l = ls[il], eta = 2 * 3.14159265358979323846 / g.dlnk / g.N * m;
    const cplx a = clgamma({(l + g.nu) / 2., eta / 2.}), b = clgamma({(3. + l - g.nu) / 2., -eta / 2.});
    const double ln2 = 0.69314718055994530942;
    /* k0 r0 = k[0] / k[N-1] */
    const double lkr = -log(g.ratio) * (double)(g.N - 1) / g.N;
    const cplx e = {g.nu * ln2 + a.re - b.re, eta * ln2 + a.im - b.im - eta * lkr};
    T[idx] = cexp_(e);
}

/* columns: for every n <= N/2 the inverse transform over m of X[m][n]; rows i_lo .. i_lo + n_sub - 1 are kept:
   Y[(i - i_lo) * (N/2+1) + n] */
__global__ void __launch_bounds__(1024) cov_fftlog_columns_kernel(const FftPlan plan, const cplx *__restrict__ F, const cplx *__restrict__ T1,
                                                                  const cplx *__restrict__ T2, const int i_lo, const int n_sub,
                                                                  cplx *__restrict__ work, cplx *__restrict__ Y)
{
    const int N = plan.n, half = N / 2;
    const long ncut = (long)(half * 0.25);
    cplx *a = work + (size_t)blockIdx.x * 2 * N, *b = a + N;
    for (int n = blockIdx.x; n <= half; n += gridDim.x) {
        cplx Bn = T2[n];
        if (n >= half - ncut) { const double w = fftlog_window(half - n, ncut); Bn.re *= w; Bn.im *= w; }
        for (int m = threadIdx.x; m < N; m += blockDim.x) {
            cplx Am;
            double w = 1.0;
            if (m <= half) { Am = T1[m]; if (m >= half - ncut) w = fftlog_window(half - m, ncut); }
            else { Am = T1[N - m]; Am.im = -Am.im; if (m <= half + 1 + ncut) w = fftlog_window(m - half - 1, ncut); }
            const cplx v = cmul(cmul(F[(m + n) % N], Am), Bn);
            a[m] = {v.re * w, -v.im * w};
        }
        __syncthreads();
        const cplx *r = fft_run(plan, a, b, +1.0);
        for (int i = threadIdx.x; i < n_sub; i += blockDim.x) Y[(size_t)i * (half + 1) + n] = r[i_lo + i];
        __syncthreads();
    }
}

/* rows: the real inverse transform over n of row i (Hermitian completion, imaginary parts of the n = 0 and n = N/2 bins dropped
   as FFTW's c2r does), columns i_lo .. kept, scaled by pi / (16 N^2 r_i^nu r_j^nu): R[(i - i_lo) * n_sub + (j - i_lo)] */
__global__ void __launch_bounds__(1024) cov_fftlog_rows_kernel(const FftPlan plan, const FftlogGeom g, const cplx *__restrict__ Y, const int i_lo,
                                                               const int n_sub, cplx *__restrict__ work, double *__restrict__ R)
{
    const int N = plan.n, half = N / 2;
    cplx *a = work + (size_t)blockIdx.x * 2 * N, *b = a + N;
    for (int row = blockIdx.x; row < n_sub; row += gridDim.x) {
        const cplx *y = Y + (size_t)row * (half + 1);
        for (int n = threadIdx.x; n <= half; n += blockDim.x) {
            cplx v = y[n];
            if (n == 0 || n == half) v.im = 0.0;
            a[n] = v;
            if (n > 0 && n < half) a[N - n] = {v.re, -v.im};
        }
        __syncthreads();
        const cplx *r = fft_run(plan, a, b, +1.0);
        /* r_i = 1 / k[N-1-i] */
        const double ri = 1.0 / (g.kmin * pow(g.ratio, (double)(N - 1 - (i_lo + row)) / N));
        for (int j = threadIdx.x; j < n_sub; j += blockDim.x) {
            const double rj = 1.0 / (g.kmin * pow(g.ratio, (double)(N - 1 - (i_lo + j)) / N));
            R[(size_t)row * n_sub + j] = r[i_lo + j].re * 3.14159265358979323846 / (16. * (double)N * (double)N * pow(rj, g.nu) * pow(ri, g.nu));
        }
        __syncthreads();
    }
}

/* ---- host helpers ----------------------------------------------------------------------------------- */
/* bicubic patch on a (possibly non-uniform) grid from node data {z, z_x, z_y, z_xy}, node index j*nx + i (gsl_interp2d_bicubic) */
double covariance_coefficient(const double b1, const double b2, const double f, const int ell)      /* src/covariance.c:40-57 */
{
    switch (ell) {
        case 0: return b1 * b2 + (b1 + b2) * f / 3 + f * f / 5.;
        case 2: return 2 * (b1 + b2) * f / 3. + 4 * f * f / 7.;
        case 4: return 8 * f * f / 35.;
        default: return 0;
    }
}

/* Wigner 3j (a b c; 0 0 0) */
double three_j0(const int a, const int b, const int c)
{
    if (c > a + b || c < std::abs(a - b)) return 0.0;
    const int J = a + b + c;
    if (J % 2) return 0.0;
    const int g = J / 2;
    const double ln = 0.5 * (std::lgamma(J - 2 * a + 1.0) + std::lgamma(J - 2 * b + 1.0) + std::lgamma(J - 2 * c + 1.0) - std::lgamma(J + 2.0))
                      + std::lgamma(g + 1.0) - std::lgamma(g - a + 1.0) - std::lgamma(g - b + 1.0) - std::lgamma(g - c + 1.0);
    const double v = std::exp(ln);
    return (g % 2) ? -v : v;
}

/* int_{-1}^{1} P_n P_m P_a P_b dmu = 2 sum_l (2l+1) (n m l)^2 (a b l)^2          src/common.c:1370-1394 */
double legendre_integral(const int n, const int m, const int a, const int b)
{
    if ((n + m + a + b) % 2 != 0) return 0;
    const int lo = std::max(std::abs(n - m), std::abs(a - b)), hi = std::min(n + m, a + b);
    double r = 0;
    for (int l = lo; l <= hi; ++l) {
        const double x = three_j0(n, m, l), y = three_j0(a, b, l);
        r += (2. * l + 1) * x * x * y * y;
    }
    return 2 * r;
}

/* G^(p)[(i1, m), (i2, n)] of stage B from the 2-D FFTLog: h_G[p * Mpad^2 + row * Mpad + col] */
/* ---- bicubic resampling of a 2-D FFTLog block on the device -------------------------------------------------------------------
   gsl_interp2d_bicubic (src/covariance.c:428-455): node derivatives from natural cubic splines along the grid lines, Hermite patch.
   Until round 2 every block went to the host for this (8 MB down, ~55 ms of spline solves, 12 blocks: the 0.7 s of the method).
   The tridiagonal system depends on the knots only: its elimination factors are computed once on the host (table_build.h:
   natural_c, same arithmetic), a line then costs one forward and one backward sweep.  One thread per line. */
struct SplineLines {
    const double *x, *off, *diagp, *w;      /* knots; off-diagonal, eliminated diagonal and elimination factors of the n - 2 equations */
    int n;
};

/* derivative at every knot of the natural cubic spline through the n values src[line * line_stride + i * elem_stride];
   scratch: n doubles per line, laid out [i * n_lines + line] */
__global__ void cov_spline_derivs_kernel(const SplineLines sp, const double *__restrict__ src, const long line_stride, const long elem_stride,
                                         double *__restrict__ dst, double *__restrict__ scratch, const int n_lines)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_lines) return;
    const int n = sp.n, m = n - 2;
    const double *y = src + (size_t)t * line_stride;
    double *o = dst + (size_t)t * line_stride;
    double *c = scratch + t;                /* c[i] at scratch[i * n_lines + t] */
    auto Y = [&](const int i) { return y[(size_t)i * elem_stride]; };
    auto C = [&](const int i) -> double & { return c[(size_t)i * n_lines]; };
    /* forward sweep: g[i] kept in C(i + 1) */
    double g_prev = 0.0;
    for (int i = 0; i < m; ++i) {
        const double h0 = sp.x[i + 1] - sp.x[i], h1 = sp.x[i + 2] - sp.x[i + 1];
        double g = 3.0 * ((Y(i + 2) - Y(i + 1)) / h1 - (Y(i + 1) - Y(i)) / h0);
        if (i > 0) g -= sp.w[i] * g_prev;
        C(i + 1) = g;
        g_prev = g;
    }
    C(0) = 0.0; C(n - 1) = 0.0;
    C(m) = C(m) / sp.diagp[m - 1];
    for (int i = m - 2; i >= 0; --i) C(i + 1) = (C(i + 1) - sp.off[i] * C(i + 2)) / sp.diagp[i];
    for (int i = 0; i + 1 < n; ++i) {
        const double dx = sp.x[i + 1] - sp.x[i];
        o[(size_t)i * elem_stride] = (Y(i + 1) - Y(i)) / dx - dx * (C(i + 1) + 2.0 * C(i)) / 3.0;
    }
    {
        const int i = n - 2;
        const double dx = sp.x[i + 1] - sp.x[i];
        const double b = (Y(i + 1) - Y(i)) / dx - dx * (C(i + 1) + 2.0 * C(i)) / 3.0;
        const double d = (C(i + 1) - C(i)) / (3.0 * dx);
        o[(size_t)(n - 1) * elem_stride] = b + dx * (2.0 * C(i) + 3.0 * dx * d);
    }
}

/* G[(row0 + m) * Mpad + col0 + n] = norm * patch(sep_m, sep_n) (and the mirrored entry) from the node data A = {z, zx, zy, zxy}[i * n + j],
   i along the first argument */
__global__ void cov_bicubic_eval_kernel(const double *__restrict__ xa, const int n, const double *__restrict__ z, const double *__restrict__ zx,
                                        const double *__restrict__ zy, const double *__restrict__ zxy, const double *__restrict__ sep,
                                        const int n_sep, const double norm, const int Mpad, const int row0, const int col0, const int mirror,
                                        double *__restrict__ G)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_sep * n_sep) return;
    const int m = e / n_sep, q = e % n_sep;
    auto cell = [&](const double v) {
        int lo = 0, hi = n - 1;             /* last knot <= v, clamped to [0, n - 2] */
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (xa[mid] > v) hi = mid; else lo = mid; }
        return lo;
    };
    const double x = sep[m], y = sep[q];
    const int xi = cell(x), yi = cell(y);
    const double dx = xa[xi + 1] - xa[xi], dy = xa[yi + 1] - xa[yi];
    const double t = (x - xa[xi]) / dx, u = (y - xa[yi]) / dy;
    const double t2 = t * t, t3 = t2 * t, u2 = u * u, u3 = u2 * u;
    const double h00t = 2 * t3 - 3 * t2 + 1, h10t = (t3 - 2 * t2 + t) * dx, h01t = -2 * t3 + 3 * t2, h11t = (t3 - t2) * dx;
    const double h00u = 2 * u3 - 3 * u2 + 1, h10u = (u3 - 2 * u2 + u) * dy, h01u = -2 * u3 + 3 * u2, h11u = (u3 - u2) * dy;
    const size_t a00 = (size_t)xi * n + yi, a10 = a00 + n, a01 = a00 + 1, a11 = a10 + 1;       /* a[x-knot][y-knot] */
    const double v = h00t * (h00u * z[a00] + h01u * z[a01] + h10u * zy[a00] + h11u * zy[a01])
                   + h01t * (h00u * z[a10] + h01u * z[a11] + h10u * zy[a10] + h11u * zy[a11])
                   + h10t * (h00u * zx[a00] + h01u * zx[a01] + h10u * zxy[a00] + h11u * zxy[a01])
                   + h11t * (h00u * zx[a10] + h01u * zx[a11] + h10u * zxy[a10] + h11u * zxy[a11]);
    G[(size_t)(row0 + m) * Mpad + col0 + q] = norm * v;
    if (mirror) G[(size_t)(col0 + q) * Mpad + row0 + m] = norm * v;
}

int fill_G_fftlog(coffe_b200_ctx *ctx, const coffe_b200_covariance_params *p, const int Mpad, double *d_G /* [2][Mpad][Mpad] on the device */)
{
    const int n_l = p->n_l, n_sep = p->n_sep;
    const int N = p->integration_bins > 0 ? p->integration_bins : 8000;
    FftPlan plan;
    if (N < 64 || N % 2 || !make_fft_plan(N, plan)) { ctx->err = "covariance_integration_bins must be even with prime factors 2, 3, 5 only"; return COFFE_B200_EINVAL; }
    DevBuf<double> d_tw;
    CK(cudaSetDevice(ctx->device));
    CK(d_tw.alloc((size_t)2 * N));
    CK(fft_plan_twiddles(plan, d_tw.p, ctx->stream));
    FftlogGeom g;
    g.N = N; g.kmin = p->k_min; g.ratio = p->k_max / p->k_min; g.dlnk = std::log(g.ratio) / N; g.nu = 1.01;      /* src/covariance.c:375-377 */
    /* r_i = 1 / k[N-1-i], ascending; rows / columns needed: the cells of the requested separations +- the reach of the spline derivatives */
    std::vector<double> r(N);
    for (int i = 0; i < N; ++i) r[i] = 1.0 / (g.kmin * std::pow(g.ratio, (double)(N - 1 - i) / N));
    double smin = 1e300, smax = 0;
    for (int i = 0; i < n_sep; ++i) { smin = std::min(smin, p->sep[i]); smax = std::max(smax, p->sep[i]); }
    if (!(smin > r[0]) || !(smax < r[N - 1])) { ctx->err = "separations outside the range 1/k_max .. 1/k_min of the 2-D FFTLog"; return COFFE_B200_ERANGE; }
    const int reach = 50;
    int i_lo = (int)(std::upper_bound(r.begin(), r.end(), smin) - r.begin()) - 1 - reach;
    int i_hi = (int)(std::upper_bound(r.begin(), r.end(), smax) - r.begin()) + reach;
    i_lo = std::max(0, i_lo); i_hi = std::min(N - 1, i_hi);
    const int n_sub = i_hi - i_lo + 1, half = N / 2;
    if ((size_t)n_sub * (half + 1) * sizeof(cplx) > ((size_t)8 << 30)) { ctx->err = "separation range too wide for the 2-D FFTLog work space"; return COFFE_B200_EINVAL; }
    std::vector<double> rs(r.begin() + i_lo, r.begin() + i_hi + 1);

    const int kind = ctx->layout.interp_kind;
    std::vector<build::Coef> c1;
    if (!build::spline_coefs(kind, p->pk_k, p->pk_p, p->n_pk, c1)) { ctx->err = "cannot build the P(k) spline"; return COFFE_B200_EINVAL; }
    DevBuf<double> d_pkx, d_c1, d_F, d_fw, d_T, d_work, d_Y, d_R;
    DevBuf<int> d_l;
    CK(d_pkx.upload(std::vector<double>(p->pk_k, p->pk_k + p->n_pk), ctx->stream));
    CK(d_c1.upload(std::vector<double>(reinterpret_cast<double *>(c1.data()), reinterpret_cast<double *>(c1.data()) + 4 * c1.size()), ctx->stream));
    CK(d_l.upload(std::vector<int>(p->l, p->l + n_l), ctx->stream));
    CK(d_F.alloc((size_t)2 * 2 * N));          /* two spectra (P, P^2), complex */
    CK(d_fw.alloc((size_t)2 * 2 * N));
    CK(d_T.alloc((size_t)n_l * (half + 1) * 2));
    const int n_cta = 2 * ctx->n_sm;
    CK(d_work.alloc((size_t)n_cta * 2 * N * 2));
    CK(d_Y.alloc((size_t)n_sub * (half + 1) * 2));
    CK(d_R.alloc((size_t)n_sub * n_sub));
    const SplineView s1{d_pkx.p, reinterpret_cast<const double4 *>(d_c1.p), p->n_pk};
    cplx *F = reinterpret_cast<cplx *>(d_F.p);
    cov_fftlog_diag_kernel<<<(N + 255) / 256, 256, 0, ctx->stream>>>(g, s1, F, F + N);
    CK(cudaGetLastError());
    cov_fft1d_kernel<<<2, 1024, 0, ctx->stream>>>(plan, F, reinterpret_cast<cplx *>(d_fw.p), -1.0);
    CK(cudaGetLastError());
    cov_fftlog_gl_kernel<<<(n_l * (half + 1) + 255) / 256, 256, 0, ctx->stream>>>(g, d_l.p, n_l, reinterpret_cast<cplx *>(d_T.p));
    CK(cudaGetLastError());
    ctx->last_launches += 3;

    CK(cudaMemsetAsync(d_G, 0, sizeof(double) * 2 * Mpad * Mpad, ctx->stream));
    /* the natural-spline system of the knots rs, eliminated once (table_build.h: natural_c) */
    DevBuf<double> d_rs, d_off, d_diagp, d_w, d_zx, d_zy, d_zxy, d_scr, d_seps;
    {
        const int m = n_sub - 2;
        std::vector<double> off(m), diagp(m), w(m, 0.0);
        for (int i = 0; i < m; ++i) {
            const double h0 = rs[i + 1] - rs[i], h1 = rs[i + 2] - rs[i + 1];
            off[i] = h1; diagp[i] = 2.0 * (h1 + h0);
        }
        for (int i = 1; i < m; ++i) { w[i] = off[i - 1] / diagp[i - 1]; diagp[i] -= w[i] * off[i - 1]; }
        CK(d_rs.upload(rs, ctx->stream)); CK(d_off.upload(off, ctx->stream)); CK(d_diagp.upload(diagp, ctx->stream)); CK(d_w.upload(w, ctx->stream));
    }
    CK(d_seps.upload(std::vector<double>(p->sep, p->sep + n_sep), ctx->stream));
    CK(d_zx.alloc((size_t)n_sub * n_sub)); CK(d_zy.alloc((size_t)n_sub * n_sub)); CK(d_zxy.alloc((size_t)n_sub * n_sub)); CK(d_scr.alloc((size_t)n_sub * n_sub));
    const SplineLines sp{d_rs.p, d_off.p, d_diagp.p, d_w.p, n_sub};
    const double pi = 3.14159265358979323846;
    const int lb = (n_sub + 127) / 128;
    for (int i1 = 0; i1 < n_l; ++i1)
        for (int i2 = i1; i2 < n_l; ++i2)
            for (int pw = 0; pw < 2; ++pw) {
                if ((pw == 0 && !p->mixed) || (pw == 1 && !p->cosmic)) continue;
                const cplx *T = reinterpret_cast<const cplx *>(d_T.p);
                cov_fftlog_columns_kernel<<<std::min(n_cta, half + 1), 1024, 0, ctx->stream>>>(plan, F + (size_t)pw * N, T + (size_t)i1 * (half + 1),
                                                                                         T + (size_t)i2 * (half + 1), i_lo, n_sub,
                                                                                         reinterpret_cast<cplx *>(d_work.p), reinterpret_cast<cplx *>(d_Y.p));
                CK(cudaGetLastError());
                cov_fftlog_rows_kernel<<<std::min(n_cta, n_sub), 1024, 0, ctx->stream>>>(plan, g, reinterpret_cast<const cplx *>(d_Y.p), i_lo, n_sub,
                                                                                      reinterpret_cast<cplx *>(d_work.p), d_R.p);
                CK(cudaGetLastError());
                /* bicubic resampling onto the separations, A[i][j] = R[i * n_sub + j] <-> (r1_i, r2_j): z_x along i (stride n_sub), z_y
                   along j, z_xy = (z_y)_x, as gsl_interp2d_bicubic estimates them (src/covariance.c:428-455) */
                cov_spline_derivs_kernel<<<lb, 128, 0, ctx->stream>>>(sp, d_R.p, 1, n_sub, d_zx.p, d_scr.p, n_sub);
                cov_spline_derivs_kernel<<<lb, 128, 0, ctx->stream>>>(sp, d_R.p, n_sub, 1, d_zy.p, d_scr.p, n_sub);
                cov_spline_derivs_kernel<<<lb, 128, 0, ctx->stream>>>(sp, d_zy.p, 1, n_sub, d_zxy.p, d_scr.p, n_sub);
                const double norm = (pw == 0 ? 2.0 : 1.0) / pi / pi;               /* src/covariance.c:760-769 */
                cov_bicubic_eval_kernel<<<(n_sep * n_sep + 127) / 128, 128, 0, ctx->stream>>>(d_rs.p, n_sub, d_R.p, d_zx.p, d_zy.p, d_zxy.p, d_seps.p, n_sep, norm,
                                                                                          Mpad, i1 * n_sep, i2 * n_sep, i2 != i1 /* :779-788 */,
                                                                                          d_G + (size_t)pw * Mpad * Mpad);
                CK(cudaGetLastError());
                ctx->last_launches += 6;
            }
    CK(cudaStreamSynchronize(ctx->stream));
    return COFFE_B200_OK;
}

}  // namespace

/* rows [row_begin, row_end) of every bin (row = il1*n_sep + ir1; row_end < 0: all rows) into the full-size host buffer `out` */
static int covariance_core(coffe_b200_ctx *ctx, const coffe_b200_covariance_params *p, int row_begin, int row_end, double *out)
{
    if (!ctx->have_tables) { ctx->err = "no tables uploaded (the covariance needs the background and galaxy bias tables)"; return COFFE_B200_ENOTABLES; }
    ctx->last_ms = 0; ctx->last_launches = 0;
    const int n_l = p->n_l, n_sep = p->n_sep, n_bins = p->n_bins;
    if (n_l < 1 || n_l > kMaxCovL || n_sep < 0 || n_bins < 0 || !p->l || (n_sep > 0 && !p->sep)) { ctx->err = "bad covariance grid"; return COFFE_B200_EINVAL; }
    if (n_bins == 0 || n_sep == 0) return COFFE_B200_OK;
    if (row_end < 0) { row_begin = 0; row_end = n_l * n_sep; }
    if (row_begin < 0 || row_end > n_l * n_sep || row_begin > row_end) { ctx->err = "row range outside the matrix"; return COFFE_B200_EINVAL; }
    if (row_begin == row_end) return COFFE_B200_OK;
    if (p->integration_method == 2 && (row_begin != 0 || row_end != n_l * n_sep)) { ctx->err = "the 2-D FFTLog covariance is not row-separable"; return COFFE_B200_EINVAL; }
    if (!p->fsky || !p->density1 || !p->density2 || !p->pixelsize || (p->average ? (!p->z_min || !p->z_max) : (!p->z_mean || !p->deltaz))) {
        ctx->err = "missing per-bin covariance arrays"; return COFFE_B200_EINVAL;
    }
    if (!p->pk_k || !p->pk_p || p->n_pk < 5 || !(p->k_max > p->k_min) || !(p->k_min > 0)) { ctx->err = "bad power spectrum"; return COFFE_B200_EINVAL; }
    for (int i = 0; i < n_l; ++i)
        if (p->l[i] < 0 || p->l[i] > (p->window ? 6 : 8)) { ctx->err = "covariance multipole out of range (l <= 6 binned, l <= 8 unbinned)"; return COFFE_B200_EINVAL; }
    for (int i = 0; i < 4; ++i)
        if (p->pop[i] < 1 || p->pop[i] > 4) { ctx->err = "populations must be 1..4"; return COFFE_B200_EINVAL; }
    /* quirk: with pk_type = linear the windowed integrals use pixelsize[0] for every bin (src/covariance.c:527, 668-670) */
    const double delta = p->window ? p->pixelsize[0] : 0.0;
    double r_max = 0, r_min = 1e300;
    for (int i = 0; i < n_sep; ++i) { r_max = std::max(r_max, p->sep[i]); r_min = std::min(r_min, p->sep[i]); }
    if (!(r_min - delta / 2. > 0.0)) { ctx->err = "separations (minus half a pixel) must be positive"; return COFFE_B200_ERANGE; }
    CK(cudaSetDevice(ctx->device));

    /* ---- k nodes: geometric panels below k_t = 4 pi / r_max, then one panel per period pi / r_max of the fastest
       product, break points at the knots of the P(k) spline; n-point Gauss-Legendre per panel ---- */
    const int n_gl = 10;
    std::vector<double> bp;
    {
        const double period = 3.14159265358979323846 / (r_max + delta / 2.);
        const double k_t = std::min(p->k_max, std::max(p->k_min, 4.0 * period));
        double k = k_t;
        while (k > p->k_min) { bp.push_back(k); k /= 1.5; }
        bp.push_back(p->k_min);
        for (k = k_t + period; k < p->k_max; k += period) bp.push_back(k);
        bp.push_back(p->k_max);
        for (int i = 0; i < p->n_pk; ++i)
            if (p->pk_k[i] > p->k_min && p->pk_k[i] < p->k_max) bp.push_back(p->pk_k[i]);
        std::sort(bp.begin(), bp.end());
        bp.erase(std::unique(bp.begin(), bp.end()), bp.end());
    }
    std::vector<double> gx, gw, hk, hw;
    build::gauss_legendre_01(n_gl, gx, gw);
    for (size_t i = 0; i + 1 < bp.size(); ++i) {
        const double a = bp[i], b = bp[i + 1];
        if (!(b - a > 1e-15 * b)) continue;
        for (int j = 0; j < n_gl; ++j) { hk.push_back(a + (b - a) * gx[j]); hw.push_back((b - a) * gw[j]); }
    }
    const int n_nodes = (int)hk.size();

    /* ---- S_1 = spline(P), S_2 = spline(P^2) on the reference's knots (src/covariance.c:599-640) ---- */
    const int kind = ctx->layout.interp_kind;     /* par->interp_method */
    std::vector<double> p2(p->n_pk);
    for (int i = 0; i < p->n_pk; ++i) p2[i] = p->pk_p[i] * p->pk_p[i];
    std::vector<build::Coef> c1, c2;
    if (!build::spline_coefs(kind, p->pk_k, p->pk_p, p->n_pk, c1) || !build::spline_coefs(kind, p->pk_k, p2.data(), p->n_pk, c2)) {
        ctx->err = "cannot build the P(k) splines"; return COFFE_B200_EINVAL;
    }
    DevBuf<double> d_k, d_w, d_pkx, d_c1, d_c2, d_W, d_sep, d_V, d_P, d_G, d_out, d_cc, d_cm;
    DevBuf<int> d_l;
    DevBuf<CovBin> d_bins;
    CK(d_k.upload(hk, ctx->stream));
    CK(d_w.upload(hw, ctx->stream));
    CK(d_pkx.upload(std::vector<double>(p->pk_k, p->pk_k + p->n_pk), ctx->stream));
    CK(d_c1.upload(std::vector<double>(reinterpret_cast<double *>(c1.data()), reinterpret_cast<double *>(c1.data()) + 4 * c1.size()), ctx->stream));
    CK(d_c2.upload(std::vector<double>(reinterpret_cast<double *>(c2.data()), reinterpret_cast<double *>(c2.data()) + 4 * c2.size()), ctx->stream));
    CK(d_sep.upload(std::vector<double>(p->sep, p->sep + n_sep), ctx->stream));
    CK(d_l.upload(std::vector<int>(p->l, p->l + n_l), ctx->stream));
    const int M = n_l * n_sep, Mpad = ((M + kTile - 1) / kTile) * kTile;
    const int tiles = Mpad / kTile;
    const int tr0 = row_begin / kTile, n_tr = (row_end - 1) / kTile - tr0 + 1;
    /* the k split (and with it the order of the partial sums) is fixed by the FULL matrix, not by the rows this call computes: a row
       range must reproduce the entries of the full call bit for bit (multi-GPU sharding over rows) */
    int n_splits = std::max(1, (4 * ctx->n_sm) / (tiles * (tiles + 1) / 2));       /* ~4 blocks per SM over the pairs of tiles ta <= tb */
    const int chunk = std::max(kKc, std::min(n_nodes, (int)std::min<size_t>((size_t)1 << 31, ((size_t)2 << 30) / ((size_t)Mpad * 8))));
    if (n_splits > (chunk + kKc - 1) / kKc) n_splits = (chunk + kKc - 1) / kKc;
    CK(d_W.alloc((size_t)2 * n_nodes));
    CK(d_V.alloc((size_t)chunk * Mpad));
    CK(d_P.alloc((size_t)n_splits * 2 * Mpad * Mpad));
    CK(d_G.alloc((size_t)2 * Mpad * Mpad));

    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const bool fftlog = p->integration_method == 2;        /* COFFE_COVARIANCE_INTEGRATION_FFTLOG, src/common.h:301-302 */
    if (fftlog) {
        /* the reference has no binned variant of this path and falls back to the unbinned one with a warning (src/covariance.c:713-720) */
        const int rc = fill_G_fftlog(ctx, p, Mpad, d_G.p);
        if (rc) return rc;
    }
    SplineView s1{d_pkx.p, reinterpret_cast<const double4 *>(d_c1.p), p->n_pk}, s2{d_pkx.p, reinterpret_cast<const double4 *>(d_c2.p), p->n_pk};
    if (!fftlog) {
    cov_node_weights_kernel<<<(n_nodes + 255) / 256, 256, 0, ctx->stream>>>(d_k.p, d_w.p, n_nodes, s1, s2, p->mixed != 0, p->cosmic != 0, d_W.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    if (p->mixed || p->cosmic) {
        for (int c0 = 0, it = 0; c0 < n_nodes; c0 += chunk, ++it) {
            const int nn = std::min(chunk, n_nodes - c0);
            const size_t tot = (size_t)nn * Mpad;
            cov_bessel_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(d_k.p + c0, nn, d_sep.p, n_sep, d_l.p, n_l, delta, Mpad, d_V.p);
            CK(cudaGetLastError());
            const int sym = tr0 == 0 && n_tr == tiles;
            cov_gram_kernel<<<(sym ? tiles * (tiles + 1) / 2 : n_tr * tiles) * n_splits, 256, 0, ctx->stream>>>(d_V.p, d_W.p + 2 * (size_t)c0, nn, Mpad, n_splits, d_P.p, it > 0, tr0, n_tr, sym);
            CK(cudaGetLastError());
            ctx->last_launches += 2;
        }
        cov_reduce_kernel<<<(unsigned)((2 * (size_t)Mpad * Mpad + 255) / 256), 256, 0, ctx->stream>>>(d_P.p, n_splits, (size_t)Mpad * Mpad, d_G.p);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    else CK(cudaMemsetAsync(d_G.p, 0, sizeof(double) * 2 * Mpad * Mpad, ctx->stream));
    }

    /* ---- stage B inputs: per-bin scalars and coefficient matrices (O(n_bins n_l^2) host work) ---- */
    const Tables &th = ctx->t_host;
    std::vector<CovBin> bins(n_bins);
    std::vector<double> cc((size_t)n_bins * n_l * n_l), cm((size_t)n_bins * n_l * n_l);
    const double zmax_table = th.bg.z[th.bg.n - 1];
    for (int k = 0; k < n_bins; ++k) {
        double z_mean, volume;
        if (!p->average) {
            z_mean = p->z_mean[k];
            const double za = z_mean - p->deltaz[k], zb = z_mean + p->deltaz[k];
            if (!(za >= 0.0) || !(zb < zmax_table)) { ctx->err = "z_mean +- deltaz outside the background table"; return COFFE_B200_ERANGE; }
            const double ra = bg_chi(th.bg, za), rb = bg_chi(th.bg, zb);
            volume = 4 * 3.14159265358979323846 * p->fsky[k] * (rb * rb * rb - ra * ra * ra) / 3.;           /* :799-811 */
        }
        else {
            z_mean = (p->z_min[k] + p->z_max[k]) / 2.;
            if (!(p->z_min[k] > 0.0) || !(p->z_max[k] > p->z_min[k]) || !(p->z_max[k] < zmax_table)) { ctx->err = "bad redshift bin"; return COFFE_B200_ERANGE; }
            /* 4 pi fsky / int dz 1/(Hc chi^2 (1+z)) / (H0 h)^2                                          :812-841 */
            std::vector<double> zx, zw;
            build::gauss_legendre_01(64, zx, zw);
            double integ = 0;
            for (int j = 0; j < 64; ++j) {
                const double z = p->z_min[k] + (p->z_max[k] - p->z_min[k]) * zx[j];
                const int idx = bg_index(th.bg, z);
                const double Hc = cubic_eval(th.bg.Hc[idx], z - th.bg.z[idx]), chi = bg_chi(th.bg, z);
                integ += zw[j] * (p->z_max[k] - p->z_min[k]) / Hc / (chi * chi) / (1 + z);
            }
            volume = 4 * 3.14159265358979323846 * p->fsky[k] / integ / (ctx->o.H0h * ctx->o.H0h);
        }
        if (!(z_mean > 0.0) || !(z_mean < zmax_table)) { ctx->err = "z_mean outside the background table"; return COFFE_B200_ERANGE; }
        double bias[4] = {0, 0, 0, 0}, f = 0;
        if (ctx->o.den)
            for (int i = 0; i < 4; ++i) {
                if (!th.bias[i].x || th.bias[i].n < 2) { ctx->err = "the covariance needs galaxy_bias1..4"; return COFFE_B200_ENOTABLES; }
                bias[i] = spline_eval(th.bias[i], z_mean);
            }
        const BackgroundAtZ b = bg_eval_all(th.bg, z_mean);
        if (ctx->o.rsd) f = b.f;
        bins[k] = CovBin{volume, b.D1, p->density1[k], p->density2[k], p->pixelsize[k]};
        auto cf = [&](int pa, int pb, int ell) { return covariance_coefficient(bias[pa - 1], bias[pb - 1], f, ell); };
        const int p1 = p->pop[0], p2_ = p->pop[1], p3 = p->pop[2], p4 = p->pop[3];
        for (int i = 0; i < n_l; ++i)
            for (int j = 0; j < n_l; ++j) {
                const int l1 = p->l[i], l2 = p->l[j];
                double cs = 0, ms = 0;
                for (int a = 0; a <= 4; a += 2)
                    for (int c = 0; c <= 4; c += 2)
                        cs += (cf(p1, p3, a) * cf(p2_, p4, c) + cf(p1, p4, a) * cf(p2_, p3, c)) * legendre_integral(l1, l2, a, c);   /* sign(c) = +1 for even c */
                cs /= 4.;
                for (int a = 0; a <= 4; a += 2)
                    ms += (cf(p1, p3, a) * (p2_ == p4) / p->density2[k] + cf(p2_, p4, a) * (p1 == p3) / p->density1[k]
                           + cf(p1, p4, a) * (p2_ == p3) / p->density2[k] + cf(p2_, p3, a) * (p1 == p4) / p->density1[k])
                          * legendre_integral(l1, l2, a, 0);
                ms /= 8.;
                cc[((size_t)k * n_l + i) * n_l + j] = cs;
                cm[((size_t)k * n_l + i) * n_l + j] = ms;
            }
    }
    CK(d_bins.upload(bins, ctx->stream));
    CK(d_cc.upload(cc, ctx->stream));
    CK(d_cm.upload(cm, ctx->stream));
    CovAssembly a;
    std::memset(&a, 0, sizeof(a));
    a.n_bins = n_bins; a.n_l = n_l; a.n_sep = n_sep; a.Mpad = Mpad;
    for (int i = 0; i < n_l; ++i) a.l[i] = p->l[i];
    a.poisson = p->poisson != 0; a.mixed = p->mixed != 0; a.cosmic = p->cosmic != 0; a.window = p->window != 0;
    a.d13 = p->pop[0] == p->pop[2]; a.d24 = p->pop[1] == p->pop[3]; a.d14 = p->pop[0] == p->pop[3]; a.d23 = p->pop[1] == p->pop[2];
    a.row_begin = row_begin; a.n_rows = row_end - row_begin;
    const size_t n_out = (size_t)n_bins * a.n_rows * n_l * n_sep;
    CK(d_out.alloc(n_out));
    cov_assemble_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(a, d_bins.p, d_cc.p, d_cm.p, d_G.p, d_sep.p, d_out.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    /* each bin's rows are one contiguous run of the caller's full-size buffer */
    const size_t run = (size_t)a.n_rows * M * sizeof(double);
    CK(cudaMemcpy2DAsync(out + (size_t)row_begin * M, (size_t)M * M * sizeof(double), d_out.p, run, run, n_bins, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    for (int ib = 0; ib < n_bins; ++ib) {
        const double *o = out + ((size_t)ib * M + row_begin) * M;
        for (size_t i = 0; i < (size_t)a.n_rows * M; ++i)
            if (!std::isfinite(o[i])) { ctx->err = "non-finite covariance entry"; return COFFE_B200_ENONFINITE; }
    }
    return COFFE_B200_OK;
}

extern "C" int coffe_b200_covariance(coffe_b200_ctx *ctx, const coffe_b200_covariance_params *p, double *out)
{
    if (!ctx || !p || !out) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    return covariance_core(ctx, p, 0, -1, out);
}

extern "C" int coffe_b200_covariance_rows(coffe_b200_ctx *ctx, const coffe_b200_covariance_params *p, int row_begin, int row_end, double *out)
{
    if (!ctx || !p || !out || row_end < 0) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    return covariance_core(ctx, p, row_begin, row_end, out);
}
