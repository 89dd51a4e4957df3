/*
 * TEST INFRASTRUCTURE ONLY -- parity oracle; see fftw3.h in this directory.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <complex.h>
#include "fftw3.h"

typedef double _Complex cplx;

struct fftw_shim_plan_s {
    int kind; /* 1: r2c 1d, 2: c2r 1d, 3: r2c 2d, 4: c2r 2d */
    int n0, n1;
    double *re;
    cplx *co;
};

void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }

/* recursive mixed-radix DFT: out[k] = sum_j in[j*stride] w^(jk), w = exp(sign*2*pi*i/n) */
static void fft_rec(const cplx *in, cplx *out, int n, int stride, int sign, const cplx *tw, int twstride, cplx *scratch)
{
    if (n == 1) { out[0] = in[0]; return; }
    int p = 0;
    if (n % 4 == 0) p = 4;
    else if (n % 2 == 0) p = 2;
    else { for (int q = 3; q * q <= n; q += 2) if (n % q == 0) { p = q; break; } if (!p) p = n; }
    const int m = n / p;
    if (m == 1) {
        /* plain DFT of prime length p */
        for (int k = 0; k < p; ++k) {
            cplx s = 0;
            for (int j = 0; j < p; ++j) s += in[j * stride] * tw[((long)j * k % n) * twstride];
            out[k] = s;
        }
        return;
    }
    /* p sub-transforms of length m over decimated inputs */
    for (int r = 0; r < p; ++r)
        fft_rec(in + r * stride, out + r * m, m, stride * p, sign, tw, twstride * p, scratch);
    /* butterflies: X[k + q*m] = sum_r w_n^{r(k+qm)} Y_r[k] */
    for (int k = 0; k < m; ++k) {
        cplx t[64];
        cplx *tt = (p <= 64) ? t : scratch;
        for (int r = 0; r < p; ++r) tt[r] = out[r * m + k] * tw[((long)r * k % n) * twstride];
        for (int q = 0; q < p; ++q) {
            cplx s = 0;
            for (int r = 0; r < p; ++r) s += tt[r] * tw[(((long)r * q * m) % n) * twstride];
            scratch[p + q] = s;
        }
        for (int q = 0; q < p; ++q) out[q * m + k] = scratch[p + q];
    }
}

static void fft_c2c(const cplx *in, cplx *out, int n, int sign)
{
    cplx *tw = (cplx *)malloc(sizeof(cplx) * n);
    for (int j = 0; j < n; ++j) {
        const long double a = 2.0L * 3.141592653589793238462643383279502884L * j / n;
        tw[j] = (double)cosl(a) + sign * I * (double)sinl(a);
    }
    cplx *scratch = (cplx *)malloc(sizeof(cplx) * (2 * (size_t)n + 128));
    fft_rec(in, out, n, 1, sign, tw, 1, scratch);
    free(scratch);
    free(tw);
}

static fftw_plan mk(int kind, int n0, int n1, double *re, cplx *co)
{
    fftw_plan p = (fftw_plan)malloc(sizeof(*p));
    p->kind = kind; p->n0 = n0; p->n1 = n1; p->re = re; p->co = co;
    return p;
}

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned f) { (void)f; return mk(1, n, 1, in, (cplx *)out); }
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned f) { (void)f; return mk(2, n, 1, out, (cplx *)in); }
fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double *in, fftw_complex *out, unsigned f) { (void)f; return mk(3, n0, n1, in, (cplx *)out); }
fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex *in, double *out, unsigned f) { (void)f; return mk(4, n0, n1, out, (cplx *)in); }
void fftw_destroy_plan(fftw_plan p) { free(p); }

void fftw_execute(const fftw_plan p)
{
    if (p->kind == 1) {
        const int n = p->n0;
        cplx *a = (cplx *)malloc(sizeof(cplx) * n), *b = (cplx *)malloc(sizeof(cplx) * n);
        for (int j = 0; j < n; ++j) a[j] = p->re[j];
        fft_c2c(a, b, n, -1);
        for (int k = 0; k <= n / 2; ++k) p->co[k] = b[k];
        free(a); free(b);
    }
    else if (p->kind == 2) {
        const int n = p->n0;
        cplx *a = (cplx *)malloc(sizeof(cplx) * n), *b = (cplx *)malloc(sizeof(cplx) * n);
        for (int k = 0; k <= n / 2; ++k) a[k] = p->co[k];
        for (int k = n / 2 + 1; k < n; ++k) a[k] = conj(p->co[n - k]);
        /* FFTW c2r ignores the imaginary parts of the DC and Nyquist bins */
        a[0] = creal(a[0]);
        if (n % 2 == 0) a[n / 2] = creal(a[n / 2]);
        fft_c2c(a, b, n, +1);
        for (int j = 0; j < n; ++j) p->re[j] = creal(b[j]);
        free(a); free(b);
    }
    else if (p->kind == 3) {
        const int n0 = p->n0, n1 = p->n1, h = n1 / 2 + 1;
        cplx *row = (cplx *)malloc(sizeof(cplx) * (n0 > n1 ? n0 : n1));
        cplx *res = (cplx *)malloc(sizeof(cplx) * (n0 > n1 ? n0 : n1));
        for (int i = 0; i < n0; ++i) {
            for (int j = 0; j < n1; ++j) row[j] = p->re[(size_t)i * n1 + j];
            fft_c2c(row, res, n1, -1);
            for (int k = 0; k < h; ++k) p->co[(size_t)i * h + k] = res[k];
        }
        for (int k = 0; k < h; ++k) {
            for (int i = 0; i < n0; ++i) row[i] = p->co[(size_t)i * h + k];
            fft_c2c(row, res, n0, -1);
            for (int i = 0; i < n0; ++i) p->co[(size_t)i * h + k] = res[i];
        }
        free(row); free(res);
    }
    else if (p->kind == 4) {
        const int n0 = p->n0, n1 = p->n1, h = n1 / 2 + 1;
        cplx *row = (cplx *)malloc(sizeof(cplx) * (n0 > n1 ? n0 : n1));
        cplx *res = (cplx *)malloc(sizeof(cplx) * (n0 > n1 ? n0 : n1));
        cplx *tmp = (cplx *)malloc(sizeof(cplx) * (size_t)n0 * h);
        for (int k = 0; k < h; ++k) {
            for (int i = 0; i < n0; ++i) row[i] = p->co[(size_t)i * h + k];
            fft_c2c(row, res, n0, +1);
            for (int i = 0; i < n0; ++i) tmp[(size_t)i * h + k] = res[i];
        }
        for (int i = 0; i < n0; ++i) {
            for (int k = 0; k < h; ++k) row[k] = tmp[(size_t)i * h + k];
            for (int k = h; k < n1; ++k) row[k] = conj(tmp[(size_t)i * h + (n1 - k)]);
            row[0] = creal(row[0]);
            if (n1 % 2 == 0) row[n1 / 2] = creal(row[n1 / 2]);
            fft_c2c(row, res, n1, +1);
            for (int j = 0; j < n1; ++j) p->re[(size_t)i * n1 + j] = creal(res[j]);
        }
        free(row); free(res); free(tmp);
    }
}
