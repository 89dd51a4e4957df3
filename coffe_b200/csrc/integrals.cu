// Construction of the I^n_l(r) tables (coffe_integrals_init, src/integrals.c:580-1423) on the device.
//
//   K2  fftlog_kernel          one CTA per (n, l) table: the 2-FAST FFTLog of src/twofast.c:252-483 -- sample the natural
//                              cubic spline of P(k) on N log-spaced k with the bias and window of the reference, forward
//                              transform, multiply by the Gamma-function kernel M_ql(t), inverse transform, rescale --
//                              with a hand-written mixed-radix (2, 3, 5) Stockham FFT; all tables run as one batch;
//   K3  direct_integrals_kernel one warp per integral: the k-quadratures the reference does one at a time with QAG-61
//                              (src/integrals.c:270-455): the r -> 0 limits, the 42 knots below 1 Mpc of every table, the
//                              297 knots of the divergent I^4_0 and of its zero-separation renormalisation, and the
//                              201 x 201 renormalisation grid (1/2 pi^2) int dk P(k) j0(k chi1) j0(k chi2) / k^2.
//                              Fixed-node: the break points are the knots of the P(k) spline, every interval is cut into
//                              panels of at most one period of the Bessel factors and each panel gets a 16-point
//                              Gauss-Legendre rule; where the oscillatory envelope makes an interval's contribution
//                              negligible (< 1e-11 of int P/k^2) it is not subdivided.
// The reference's own tables carry the 1e-5 relative tolerance of its QAG calls; parity of these kernels with them is
// tested at that level (tests/test_integrals.py), their own convergence far below it.
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.h"
#include "special.cuh"
#include "fft.cuh"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kSqrtPi = 1.7724538509055160272981674833411451827975494561224;

/* M_ql(t) of src/twofast.c:196-224: alpha^(i t - q) 2^(n-1) sqrt(pi) Gamma((1+l+n)/2) / Gamma((2+l-n)/2), n = q - 1 - i t */
__device__ cplx mql(const double t, const double q, const double l, const double alpha)
{
    const cplx n = {q - 1.0, -t};
    const cplx g1 = clgamma({(1.0 + l + n.re) / 2.0, n.im / 2.0});
    const cplx g2 = clgamma({(2.0 + l - n.re) / 2.0, -n.im / 2.0});
    const double ln2 = 0.69314718055994530942, la = log(alpha);
    /* exponent: (n - 1) ln 2 + lgamma1 - lgamma2 + (i t - q) ln alpha */
    const cplx e = {(n.re - 1.0) * ln2 + g1.re - g2.re - q * la, n.im * ln2 + g1.im - g2.im + t * la};
    const cplx v = cexp_(e);
    return {v.re * kSqrtPi, v.im * kSqrtPi};
}

/* src/twofast.c:54-90 */
__device__ double window_fn(const double v, const double xmin, const double xmax, const double xleft, const double xright)
{
    if (v > xleft && v < xright && v > xmin && v < xmax) return 1.0;
    if (v <= xmin || v >= xmax) return 0.0;
    double r = 0.0;
    if (v < xleft && v > xmin) r = (v - xmin) / (xleft - xmin);
    else if (v > xright && v < xmax) r = (xmax - v) / (xmax - xright);
    return r - sin(2 * kPi * r) / 2. / kPi;
}

/* src/twofast.c:240-250 */
__host__ __device__ double select_q(const double l, const double nu)
{
    const double n1 = 0.9, n2 = 0.9999;
    const double qmin = fmax(n2 - 1.0 - nu, -l), qmax = fmin(n1 + 3.0 - nu, 2.0);
    double q = (2. + n1 + n2) / 2. - nu;
    if (!(qmin < q && q < qmax)) q = (qmin + 2. * qmax) / 3.;
    return q;
}

/* ---- P(k) splines on the device --------------------------------------------------------------------- */
struct PkView {
    const double *x;
    const double4 *k;       /* per-interval cubic {y, b, c, d} */
    int n;
};

__device__ double pk_eval(const PkView &s, const double v)
{
    const int i = knot_search(s.x, s.n, v);
    return cubic_eval(s.k[i], v - s.x[i]);
}

/* ---- K2: FFTLog ------------------------------------------------------------------------------------ */
struct FftlogTable { double l, nu; int power; /* multiply by r^power afterwards (n - l when n >= l) */ };

/* out[table * N + i] = r_i^power I^nu_l(r_i), r_i = r0 (kmax/kmin)^(i/N), r0 = 1      src/twofast.c:362-483, src/integrals.c:496-522 */
__global__ void __launch_bounds__(1024) fftlog_kernel(const PkView pk, const FftPlan plan, const FftlogTable *__restrict__ tables,
                                                      const double kmin, const double kmax, cplx *__restrict__ work, double *__restrict__ out)
{
    const int N = plan.n, N2 = N / 2 + 1;
    const FftlogTable tb = tables[blockIdx.x];
    cplx *a = work + (size_t)blockIdx.x * 2 * N, *b = a + N;
    const double ratio = kmax / kmin, G = log(ratio), k0 = kmin, r0 = 1.0;
    const double q = select_q(tb.l, tb.nu), qq = q + tb.nu;
    const double k_last = k0 * pow(ratio, (double)(N - 1) / N);
    const double xl = exp(0.46) * kmin, xr = exp(-0.46) * k_last;
    /* biased, windowed samples */
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double k = k0 * pow(ratio, (double)i / N);
        a[i] = {pow(ratio, (3. - qq) * i / N) * pk_eval(pk, k) * window_fn(k, kmin, k_last, xl, xr), 0.0};
    }
    __syncthreads();
    cplx *f = fft_run(plan, a, b, -1.0);
    cplx *g = f == a ? b : a;
    /* conj(F_m) W(k_(N/2-1+m)) / L x M_ql(t_m), extended Hermitian-symmetrically for the real inverse transform
       (FFTW's c2r ignores the imaginary parts of the m = 0 and m = N/2 bins) */
    const double L = 2 * kPi * N / G;
    for (int m = threadIdx.x; m < N2; m += blockDim.x) {
        const double km = k0 * pow(ratio, (double)(N2 - 2 + m) / N);
        const double w = window_fn(km, kmin, k_last, xl, xr) / L;
        const cplx fm = {f[m].re * w, -f[m].im * w};
        cplx v = cmul(fm, mql(2 * kPi * m / G, q, tb.l, k0 * r0));
        if (m == 0 || 2 * m == N) v.im = 0.0;
        g[m] = v;
        if (m > 0 && 2 * m < N) g[N - m] = {v.re, -v.im};
    }
    __syncthreads();
    const cplx *res = fft_run(plan, g, f, +1.0);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double r = r0 * pow(ratio, (double)i / N);
        double v = res[i].re * (k0 * k0 * k0 * pow(ratio, -(qq * i / N)) / kPi / pow(r0 * k0, tb.nu) / G);
        if (tb.power != 0) v *= pow(r, (double)tb.power);
        out[(size_t)blockIdx.x * N + i] = v;
    }
}

/* ---- K3: direct k-quadratures ---------------------------------------------------------------------- */
enum { T_BESSEL = 0, T_BESSEL_HALF = 1, T_PREFACTOR = 2, T_RENORM0 = 3, T_RENORM = 4 };

struct DirectTask {
    int type;
    int n, l;           /* integers as stored by the reference; the half-integer family has local n = n/2, l = l/2 */
    int half;
    double r, r2;       /* separation (chi1), chi2 */
    double scale;       /* factor applied to the integral */
};

/* the integrand of src/integrals.c:270-330, 388-425, 236-268 at k */
__device__ double direct_integrand(const DirectTask &t, const double k, const double P)
{
    switch (t.type) {
    case T_BESSEL:
        return k * k * P * sph_bessel(t.l, k * t.r) / pow(k * t.r, (double)t.n);
    case T_BESSEL_HALF:
        return k * k * P * sqrt(kPi / 2.) * jn((t.l + 1) / 2, k * t.r) / pow(k * t.r, 0.5 * t.n + 0.5);
    case T_PREFACTOR: {
        const double ln = t.half ? 0.5 * t.n : (double)t.n, ll = t.half ? 0.5 * t.l : (double)t.l;
        return (t.half ? sqrt(kPi / 2.) : 1.0) * pow(k, 2 + ll - ln) * P;
    }
    case T_RENORM0: {
        const double j = sph_bessel(0, k * t.r);
        return P * (1. - j * j) / k / k;
    }
    default:
        return P * sph_bessel(0, k * t.r) * sph_bessel(0, k * t.r2) / k / k;
    }
}

/* phase rate of the oscillating factor [rad per unit k] and envelope of its oscillating part relative to P / k^2 */
__device__ double task_omega(const DirectTask &t)
{
    if (t.type == T_PREFACTOR) return 0.0;
    if (t.type == T_RENORM0) return 2.0 * t.r;
    if (t.type == T_RENORM) return t.r + t.r2;
    return t.r;
}

__device__ double task_envelope(const DirectTask &t, const double k)
{
    /* bound of |integrand| / (P / k^2) once k r >> 1 */
    if (t.type == T_RENORM) return 1.0 / (fmax(1.0, k * t.r) * fmax(1.0, k * t.r2));
    if (t.type == T_RENORM0) { const double x = fmax(1.0, k * t.r); return 1.0 / (x * x); }
    if (t.type == T_BESSEL && t.n == 4 && t.l == 0) return 1.0 / fmax(1.0, k * t.r);     /* r^4 I^4_0: P j0(k r) / k^2 */
    return 1e300;       /* never negligible: always resolve */
}

/* one warp per task; out[task] = scale * int_{kmin}^{kmax} integrand dk */
__global__ void direct_integrals_kernel(const PkView pk, const DirectTask *__restrict__ tasks, const int n_tasks,
                                        const double kmin, const double kmax, const double s0,
                                        const double *__restrict__ gl_x, const double *__restrict__ gl_w, double *__restrict__ out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_tasks) return;
    const DirectTask t = tasks[warp];
    const double omega = task_omega(t);
    constexpr int NG = 16;
    const double *x = gl_x + rule_offset(NG), *w = gl_w + rule_offset(NG);
    double acc = 0.0;
    for (int i = 0; i + 1 < pk.n; ++i) {
        double a = pk.x[i], b = pk.x[i + 1];
        if (b <= kmin || a >= kmax) continue;
        if (a < kmin) a = kmin;
        if (b > kmax) b = kmax;
        const double4 c = pk.k[i];
        int m = 1;
        if (omega > 0.0) {
            const double pmax = fmax(fabs(c.x), fabs(cubic_eval(c, pk.x[i + 1] - pk.x[i])));
            const double bound = pmax * (b - a) / (a * a) * task_envelope(t, a);
            if (bound > 1e-11 * s0) m = (int)ceil((b - a) * omega / (2 * kPi));
            if (m < 1) m = 1;
        }
        const double h = (b - a) / m;
        /* lanes stride over (panel, node) pairs */
        for (long e = lane; e < (long)m * NG; e += 32) {
            const long p = e / NG;
            const int j = (int)(e - p * NG);
            const double k = a + (p + x[j]) * h;
            acc += h * w[j] * direct_integrand(t, k, cubic_eval(c, k - pk.x[i]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[warp] = acc * t.scale;
}

/* the reference's knot lists (src/integrals.c:76-126) */
const double kSepI40[] = {
    0.000001, 0.000005, 0.0001, 0.005, 0.01, 0.02, 0.05, 0.1, 0.2, 0.4, 0.8, 1., 1.5,
    2., 3., 4., 5., 6., 7., 8., 9., 10., 11., 12., 13., 14., 15., 16., 17., 18., 19., 20.,
    21., 22., 23., 24., 25., 26., 27., 28., 29., 30., 32., 34., 36., 38., 40., 42., 44., 46.,
    48., 50., 52., 54., 56., 58., 60., 62., 64., 66., 68., 70., 72., 74., 76., 78., 80., 81., 82., 83.,
    84., 85., 86., 87., 88., 89., 90., 91., 92., 93., 94., 95., 95.5, 96., 96.5, 97., 97.5, 98.,
    98.5, 99., 99.5, 100., 100.5, 101., 101.5, 102., 102.5, 103., 103.5, 104., 104.5, 105.,
    106., 107., 108., 109., 110., 112., 114., 116., 118., 120., 124., 128., 132., 136., 140.,
    144., 148., 152., 156., 160., 164., 168., 172., 176., 180., 185., 190., 195., 200., 205.,
    210., 215., 220., 225., 230., 235., 240., 250., 260., 270., 280., 290., 300., 320., 340.,
    360., 380., 400., 420., 440., 460., 480., 500., 520., 540., 560., 580., 600., 620., 640.,
    660., 680., 700., 750., 800, 850, 900, 950, 1000, 1050, 1100, 1150, 1200, 1250, 1300, 1350,
    1400, 1450, 1500, 1550, 1600, 1650, 1700, 1750, 1800, 1850, 1900, 1950, 2000, 2050, 2100,
    2150, 2200, 2250, 2300, 2350, 2400, 2450, 2500, 2550, 2600, 2650, 2700, 2750, 2800, 2850,
    2900, 2950, 3000, 3050, 3100, 3150, 3200, 3250, 3300, 3350, 3400, 3450, 3500, 3550, 3600,
    3650, 3700, 3750, 3800, 3850, 3900, 3950, 4000, 4050, 4100, 4150, 4200, 4250, 4300, 4350,
    4400, 4450, 4500, 4550, 4600, 4650, 4700, 4750, 4800, 4850, 4900, 4950, 5000, 5050, 5100,
    5150, 5200, 5250, 5300, 5350, 5400, 5450, 5500, 5550, 5600, 5650, 5700, 5750, 5800, 5850,
    5900, 5950, 6000, 6050, 6100, 6150, 6200, 6250, 6300, 6350, 6400, 6450, 6500, 6550, 6600,
    7000, 8000, 9000, 10000, 11000, 12000, 13000, 14000, 15000, 16000, 17000, 18000, 19000, 20000};
const int kNSepI40 = sizeof(kSepI40) / sizeof(kSepI40[0]);
const double kMinSep[] = {1E-6, 1E-5, 1E-4, 1E-3, 2E-3, 5E-3, 7E-3, 8E-3, 9E-3, 1E-2, 1.1E-2, 1.2E-2, 1.25E-2, 1.3E-2, 1.35E-2,
                          1.5E-2, 2E-2, 2.5E-2, 3E-2, 3.5E-2, 5E-2, 7E-2, 8E-2, 9E-2, 1E-1, 1.1E-1, 1.2E-1, 1.25E-1, 1.3E-1, 1.35E-1,
                          1.5E-1, 1.75E-1, 2E-1, 2.5E-1, 3E-1, 3.5E-1, 4E-1, 5E-1, 6E-1, 7E-1, 8E-1, 9E-1};
const int kNMinSep = sizeof(kMinSep) / sizeof(kMinSep[0]);
const int kRenormBins = 200;

}  // namespace

extern "C" {

int coffe_b200_integrals_sizes(const coffe_b200_integrals_params *p, int *n_r, int *n_r40, int *n_renorm)
{
    if (!p) return COFFE_B200_EINVAL;
    if (n_r) *n_r = p->bessel_bins + kNMinSep + 1;
    if (n_r40) *n_r40 = kNSepI40;
    if (n_renorm) *n_renorm = kRenormBins + 1;
    return COFFE_B200_OK;
}

int coffe_b200_integrals(coffe_b200_ctx *ctx, const coffe_b200_integrals_params *p, double *r, double *tables,
                         double *r40, double *i40, double *renorm0, double *renorm_chi, double *renorm_z)
{
    if (!ctx || !p) return COFFE_B200_EINVAL;
    COFFE_ENTER(ctx);
    if (!p->pk_k || !p->pk_p || p->n_pk < 5 || !(p->k_max > p->k_min) || !(p->k_min > 0)) { ctx->err = "bad power spectrum / k range"; return COFFE_B200_EINVAL; }
    if (p->n_tables < 0 || (p->n_tables > 0 && (!p->n || !p->l || !r || !tables))) { ctx->err = "bad table list"; return COFFE_B200_EINVAL; }
    if (p->divergent && (!r40 || !i40 || !renorm0 || !renorm_chi || !renorm_z || !(p->chi_max > 0))) { ctx->err = "divergent tables need output buffers and chi_max > 0"; return COFFE_B200_EINVAL; }
    const int N = p->bessel_bins;
    FftPlan plan;
    if (N < 16 || N % 2 || !make_fft_plan(N, plan)) { ctx->err = "bessel_bins must be even with prime factors 2, 3, 5 only"; return COFFE_B200_EINVAL; }
    CK(cudaSetDevice(ctx->device));
    ctx->last_ms = 0; ctx->last_launches = 0;
    cudaStream_t st = ctx->stream;
    DevBuf<double> d_tw;
    CK(d_tw.alloc((size_t)2 * N));
    CK(fft_plan_twiddles(plan, d_tw.p, st));

    /* P(k): natural cubic spline for 2-FAST (src/twofast.c:399), the configured interpolation for the direct integrals */
    std::vector<build::Coef> c_nat = build::cspline_coefs(p->pk_k, p->pk_p, p->n_pk), c_int;
    if (!build::spline_coefs(p->interp_method, p->pk_k, p->pk_p, p->n_pk, c_int)) { ctx->err = "unsupported interpolation type for P(k)"; return COFFE_B200_EINVAL; }
    DevBuf<double> d_x, d_cn, d_ci;
    CK(d_x.upload(std::vector<double>(p->pk_k, p->pk_k + p->n_pk), st));
    CK(d_cn.upload(std::vector<double>(reinterpret_cast<const double *>(c_nat.data()), reinterpret_cast<const double *>(c_nat.data()) + 4 * c_nat.size()), st));
    CK(d_ci.upload(std::vector<double>(reinterpret_cast<const double *>(c_int.data()), reinterpret_cast<const double *>(c_int.data()) + 4 * c_int.size()), st));
    const PkView pk_nat{d_x.p, reinterpret_cast<const double4 *>(d_cn.p), p->n_pk};
    const PkView pk_int{d_x.p, reinterpret_cast<const double4 *>(d_ci.p), p->n_pk};

    CK(cudaEventRecord(ctx->ev0, st));
    const int n_r = N + kNMinSep + 1;
    /* ---- K2 ---- */
    DevBuf<double> d_fft;
    DevBuf<FftlogTable> d_tb;
    DevBuf<double> d_work;
    if (p->n_tables > 0) {
        std::vector<FftlogTable> tb(p->n_tables);
        for (int i = 0; i < p->n_tables; ++i) {
            const bool half = p->half && p->half[i];
            const double ln = half ? 0.5 * p->n[i] : p->n[i], ll = half ? 0.5 * p->l[i] : p->l[i];
            tb[i].l = ll; tb[i].nu = ln;
            tb[i].power = ln >= ll ? (int)lround(ln - ll) : 0;
        }
        CK(d_tb.upload(tb, st));
        CK(d_fft.alloc((size_t)p->n_tables * N));
        CK(d_work.alloc((size_t)p->n_tables * 4 * N));
        fftlog_kernel<<<p->n_tables, 1024, 0, st>>>(pk_nat, plan, d_tb.p, p->k_min, p->k_max, reinterpret_cast<cplx *>(d_work.p), d_fft.p);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    /* ---- K3 ---- */
    std::vector<DirectTask> tasks;
    const double inv2pi2 = 1.0 / (2.0 * kPi * kPi);
    for (int i = 0; i < p->n_tables; ++i) {
        const bool half = p->half && p->half[i];
        const double ln = half ? 0.5 * p->n[i] : p->n[i], ll = half ? 0.5 * p->l[i] : p->l[i];
        DirectTask t;
        t.n = p->n[i]; t.l = p->l[i]; t.half = half; t.r2 = 0;
        /* r = 0: int k^(2+l-n) P dk x the small-argument coefficient of the Bessel function (src/integrals.c:212-268, 524-546) */
        t.type = T_PREFACTOR; t.r = 0;
        t.scale = ln >= ll ? inv2pi2 * (half ? 1.0 : std::sqrt(kPi / 2.)) / std::pow(2.0, ll + 0.5) / std::tgamma(ll + 1.5) : 0.0;
        tasks.push_back(t);
        for (int j = 0; j < kNMinSep; ++j) {
            t.type = half ? T_BESSEL_HALF : T_BESSEL; t.r = kMinSep[j];
            t.scale = inv2pi2 * (ln >= ll ? std::pow(t.r, ln - ll) : 1.0);
            tasks.push_back(t);
        }
    }
    const size_t off_div = tasks.size();
    std::vector<double> chi(kRenormBins + 1);
    if (p->divergent) {
        DirectTask t;
        t.n = 4; t.l = 0; t.half = 0; t.r2 = 0;
        for (int j = 0; j < kNSepI40; ++j) {
            t.type = T_BESSEL; t.r = kSepI40[j]; t.scale = inv2pi2 * std::pow(t.r, 4.0);
            tasks.push_back(t);
        }
        for (int j = 0; j < kNSepI40; ++j) {
            /* the reference routes this integral through integrals_integrate_function with (n, l) = (4, 0), which
               multiplies the result by sep^(n-l) (src/integrals.c:365-371, 816-827): the stored zero-separation table is
               r^4 (1/2 pi^2) int P (1 - j0^2(k r)) / k^2 dk -- reproduced as is */
            t.type = T_RENORM0; t.r = kSepI40[j]; t.scale = inv2pi2 * std::pow(t.r, 4.0);
            tasks.push_back(t);
        }
        for (int j = 0; j <= kRenormBins; ++j) chi[j] = (double)j / kRenormBins * p->chi_max;
        /* z[k*(nbins+1) + j] <-> (chi_j, chi_k)      src/integrals.c:880-896 */
        for (int k = 0; k <= kRenormBins; ++k)
            for (int j = 0; j <= kRenormBins; ++j) {
                t.type = T_RENORM; t.r = chi[j]; t.r2 = chi[k]; t.scale = inv2pi2;
                tasks.push_back(t);
            }
    }
    DevBuf<DirectTask> d_tasks;
    DevBuf<double> d_out;
    std::vector<double> h_direct(tasks.size());
    if (!tasks.empty()) {
        /* s0 = int P / k^2 dk: the scale against which oscillatory tails are called negligible */
        double s0 = 0.0;
        {
            std::vector<double> gx, gw;
            build::gauss_legendre_01(16, gx, gw);
            for (int i = 0; i + 1 < p->n_pk; ++i) {
                double a = std::max(p->pk_k[i], p->k_min), b = std::min(p->pk_k[i + 1], p->k_max);
                if (!(b > a)) continue;
                for (int j = 0; j < 16; ++j) {
                    const double k = a + gx[j] * (b - a), dx = k - p->pk_k[i];
                    const build::Coef &c = c_int[i];
                    s0 += (b - a) * gw[j] * (c.y + dx * (c.b + dx * (c.c + dx * c.d))) / (k * k);
                }
            }
        }
        CK(d_tasks.upload(tasks, st));
        CK(d_out.alloc(tasks.size()));
        const int threads = 128, blocks = (int)((tasks.size() * 32 + threads - 1) / threads);
        direct_integrals_kernel<<<blocks, threads, 0, st>>>(pk_int, d_tasks.p, (int)tasks.size(), p->k_min, p->k_max, std::fabs(s0),
                                                            ctx->d_gl_x, ctx->d_gl_w, d_out.p);
        CK(cudaGetLastError());
        ctx->last_launches += 1;
    }
    CK(cudaEventRecord(ctx->ev1, st));
    std::vector<double> h_fft((size_t)p->n_tables * N);
    if (p->n_tables > 0) CK(cudaMemcpyAsync(h_fft.data(), d_fft.p, h_fft.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (!tasks.empty()) CK(cudaMemcpyAsync(h_direct.data(), d_out.p, h_direct.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;

    /* assemble: knots 0, the 42 small separations, then the FFTLog grid r_i = (kmax/kmin)^(i/N)      src/integrals.c:524-562 */
    if (p->n_tables > 0) {
        r[0] = 0.0;
        for (int j = 0; j < kNMinSep; ++j) r[1 + j] = kMinSep[j];
        for (int i = 0; i < N; ++i) r[1 + kNMinSep + i] = std::pow(p->k_max / p->k_min, (double)i / N);
        for (int t = 0; t < p->n_tables; ++t) {
            double *o = tables + (size_t)t * n_r;
            const double *d = h_direct.data() + (size_t)t * (kNMinSep + 1);
            for (int j = 0; j <= kNMinSep; ++j) o[j] = d[j];
            std::memcpy(o + kNMinSep + 1, h_fft.data() + (size_t)t * N, sizeof(double) * N);
        }
    }
    if (p->divergent) {
        const double *d = h_direct.data() + off_div;
        for (int j = 0; j < kNSepI40; ++j) { r40[j] = kSepI40[j]; i40[j] = d[j]; renorm0[j] = d[kNSepI40 + j]; }
        renorm0[0] = 0.0;        /* the zero-separation spline is pinned to (0, 0)      src/integrals.c:838-839 */
        for (int j = 0; j <= kRenormBins; ++j) renorm_chi[j] = chi[j];
        std::memcpy(renorm_z, d + 2 * kNSepI40, sizeof(double) * (kRenormBins + 1) * (kRenormBins + 1));
    }
    const size_t n_check = h_fft.size();
    for (size_t i = 0; i < n_check; ++i)
        if (!std::isfinite(h_fft[i])) { ctx->err = "non-finite FFTLog output"; return COFFE_B200_ENONFINITE; }
    for (size_t i = 0; i < h_direct.size(); ++i)
        if (!std::isfinite(h_direct[i])) { ctx->err = "non-finite k-integral"; return COFFE_B200_ENONFINITE; }
    return COFFE_B200_OK;
}

/* I^n_l(z, r) for the midpoint approximation: one coffe_b200_integrals pass per redshift row of P(k, z) (src/integrals.c:663-767 does
   exactly this, 100 times per table).  Set-up work: ~100 x 5 ms. */
int coffe_b200_integrals2d(coffe_b200_ctx *ctx, const coffe_b200_integrals_params *p, const double *pk2d, int n_z, double *tables2d)
{
    if (!ctx || !p || !pk2d || !tables2d || n_z < 1) return COFFE_B200_EINVAL;
    int n_r = 0, n_r40 = 0, n_ren = 0;
    int rc = coffe_b200_integrals_sizes(p, &n_r, &n_r40, &n_ren);
    if (rc) return rc;
    coffe_b200_integrals_params q = *p;
    q.divergent = 0;
    std::vector<double> r((size_t)n_r), tab((size_t)n_r * p->n_tables);
    double total_ms = 0;
    for (int i = 0; i < n_z; ++i) {
        q.pk_p = pk2d + (size_t)i * p->n_pk;
        rc = coffe_b200_integrals(ctx, &q, r.data(), tab.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
        if (rc) return rc;
        total_ms += ctx->last_ms;
        for (int t = 0; t < p->n_tables; ++t)
            for (int j = 0; j < n_r; ++j) tables2d[((size_t)t * n_r + j) * n_z + i] = tab[(size_t)t * n_r + j];
    }
    ctx->last_ms = total_ms;
    return COFFE_B200_OK;
}

}  // extern "C"
