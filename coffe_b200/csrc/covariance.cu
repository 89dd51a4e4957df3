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

/* partial Gram products over a k range: P[split][p][row][col] += sum_k W_p(k) V[k][row] V[k][col] */
__global__ void __launch_bounds__(256) cov_gram_kernel(const double *__restrict__ V, const double *__restrict__ W, const int n_nodes,
                                                       const int Mpad, const int n_splits, double *__restrict__ P, const int accumulate)
{
    __shared__ double As[kKc][kTile], Bs[kKc][kTile], Ws[kKc][2];
    const int tiles = Mpad / kTile;
    const int tile = blockIdx.x % (tiles * tiles), split = blockIdx.x / (tiles * tiles);
    const int tr = tile / tiles, tc = tile % tiles;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int per = (n_nodes + n_splits - 1) / n_splits;
    const int k0 = split * per, k1 = min(n_nodes, k0 + per);
    double acc1[4][4], acc2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc1[i][j] = 0.0; acc2[i][j] = 0.0; }
    for (int kb = k0; kb < k1; kb += kKc) {
        /* 16 x 64 doubles per operand tile, 256 threads: 4 each */
        for (int e = threadIdx.x; e < kKc * kTile; e += 256) {
            const int kk = e / kTile, c = e % kTile;
            const bool ok = kb + kk < k1;
            As[kk][c] = ok ? V[(size_t)(kb + kk) * Mpad + tr * kTile + c] : 0.0;
            Bs[kk][c] = ok ? V[(size_t)(kb + kk) * Mpad + tc * kTile + c] : 0.0;
        }
        if (threadIdx.x < kKc * 2) {
            const int kk = threadIdx.x / 2, p = threadIdx.x % 2;
            Ws[kk][p] = kb + kk < k1 ? W[2 * (size_t)(kb + kk) + p] : 0.0;
        }
        __syncthreads();
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
            const size_t o = (size_t)(tr * kTile + ty * 4 + i) * Mpad + tc * kTile + tx * 4 + j;
            if (accumulate) { P1[o] += acc1[i][j]; P2[o] += acc2[i][j]; }
            else { P1[o] = acc1[i][j]; P2[o] = acc2[i][j]; }
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
};

__device__ __forceinline__ int sign_of(const int m) { return (abs(m) % 2 == 0) ? 1 : -1; }

/* out[(((ib*n_l + i1)*n_sep + m)*n_l + i2)*n_sep + n]                         src/covariance.c:896-1044 */
__global__ void cov_assemble_kernel(const CovAssembly a, const CovBin *__restrict__ bins, const double *__restrict__ coef_cosmic,
                                    const double *__restrict__ coef_mixed, const double *__restrict__ G, const double *__restrict__ sep,
                                    double *__restrict__ out)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_bin = (size_t)a.n_l * a.n_sep * a.n_l * a.n_sep;
    if (idx >= per_bin * a.n_bins) return;
    const int ib = (int)(idx / per_bin);
    size_t r = idx % per_bin;
    const int n = (int)(r % a.n_sep); r /= a.n_sep;
    const int i2 = (int)(r % a.n_l); r /= a.n_l;
    const int m = (int)(r % a.n_sep); r /= a.n_sep;
    const int i1 = (int)r;
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

/* ---- host helpers ----------------------------------------------------------------------------------- */
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

}  // namespace

extern "C" int coffe_b200_covariance(coffe_b200_ctx *ctx, const coffe_b200_covariance_params *p, double *out)
{
    if (!ctx || !p || !out) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    if (!ctx->have_tables) { ctx->err = "no tables uploaded (the covariance needs the background and galaxy bias tables)"; return COFFE_B200_ENOTABLES; }
    ctx->last_ms = 0; ctx->last_launches = 0;
    const int n_l = p->n_l, n_sep = p->n_sep, n_bins = p->n_bins;
    if (n_l < 1 || n_l > kMaxCovL || n_sep < 0 || n_bins < 0 || !p->l || (n_sep > 0 && !p->sep)) { ctx->err = "bad covariance grid"; return COFFE_B200_EINVAL; }
    if (n_bins == 0 || n_sep == 0) return COFFE_B200_OK;
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
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    int n_splits = std::max(1, (4 * prop.multiProcessorCount) / (tiles * tiles));
    const int chunk = std::max(kKc, std::min(n_nodes, (int)std::min<size_t>((size_t)1 << 31, ((size_t)2 << 30) / ((size_t)Mpad * 8))));
    if (n_splits > (chunk + kKc - 1) / kKc) n_splits = (chunk + kKc - 1) / kKc;
    CK(d_W.alloc((size_t)2 * n_nodes));
    CK(d_V.alloc((size_t)chunk * Mpad));
    CK(d_P.alloc((size_t)n_splits * 2 * Mpad * Mpad));
    CK(d_G.alloc((size_t)2 * Mpad * Mpad));

    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    SplineView s1{d_pkx.p, reinterpret_cast<const double4 *>(d_c1.p), p->n_pk}, s2{d_pkx.p, reinterpret_cast<const double4 *>(d_c2.p), p->n_pk};
    cov_node_weights_kernel<<<(n_nodes + 255) / 256, 256, 0, ctx->stream>>>(d_k.p, d_w.p, n_nodes, s1, s2, p->mixed != 0, p->cosmic != 0, d_W.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    if (p->mixed || p->cosmic) {
        for (int c0 = 0, it = 0; c0 < n_nodes; c0 += chunk, ++it) {
            const int nn = std::min(chunk, n_nodes - c0);
            const size_t tot = (size_t)nn * Mpad;
            cov_bessel_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(d_k.p + c0, nn, d_sep.p, n_sep, d_l.p, n_l, delta, Mpad, d_V.p);
            CK(cudaGetLastError());
            cov_gram_kernel<<<tiles * tiles * n_splits, 256, 0, ctx->stream>>>(d_V.p, d_W.p + 2 * (size_t)c0, nn, Mpad, n_splits, d_P.p, it > 0);
            CK(cudaGetLastError());
            ctx->last_launches += 2;
        }
        cov_reduce_kernel<<<(unsigned)((2 * (size_t)Mpad * Mpad + 255) / 256), 256, 0, ctx->stream>>>(d_P.p, n_splits, (size_t)Mpad * Mpad, d_G.p);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    else CK(cudaMemsetAsync(d_G.p, 0, sizeof(double) * 2 * Mpad * Mpad, ctx->stream));

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
    const size_t n_out = (size_t)n_bins * n_l * n_sep * n_l * n_sep;
    CK(d_out.alloc(n_out));
    cov_assemble_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, ctx->stream>>>(a, d_bins.p, d_cc.p, d_cm.p, d_G.p, d_sep.p, d_out.p);
    CK(cudaGetLastError());
    ctx->last_launches += 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaMemcpyAsync(out, d_out.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    for (size_t i = 0; i < n_out; ++i)
        if (!std::isfinite(out[i])) { ctx->err = "non-finite covariance entry"; return COFFE_B200_ENONFINITE; }
    return COFFE_B200_OK;
}
