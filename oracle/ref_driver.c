/*
 * TEST INFRASTRUCTURE ONLY -- parity oracle; never linked into libcoffe_b200.so.
 *
 * Flat C entry points (for ctypes) over the UNMODIFIED reference sources compiled in place from
 * /root/reference/src (see oracle/Makefile).  Each setter mirrors what the corresponding property
 * setter of the reference's Cython class does to `coffe_parameters_t` (python/coffe/coffe.pyx), and
 * each compute call runs the reference's own driver:
 *   ref_background   -> coffe_background_init          (src/background.c:301)
 *   ref_integrals    -> coffe_integrals_init           (src/integrals.c:580)
 *   ref_multipoles   -> coffe_multipoles_init          (src/multipoles.c:50)
 *   ref_corrfunc     -> coffe_corrfunc_init            (src/corrfunc.c:56)
 *   ref_average_multipoles -> coffe_average_multipoles_init (src/average_multipoles.c:72)
 *   ref_covariance   -> coffe_covariance_init          (src/covariance.c:477)
 * plus "converged" variants that integrate the reference's own integrands
 * (functions_nonintegrated / _single_integrated / _double_integrated, src/functions.c) with a
 * composite Gauss-Legendre tensor rule of caller-chosen size, which defines truth for the 1e-5 gate
 * (the stock Cuhre path is budget-truncated).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "common.h"
#include "parser.h"
#include "background.h"
#include "integrals.h"
#ifndef COFFE_DROPIN_BUILD
#include "functions.h"
#endif
#include "signal.h"
#include "corrfunc.h"
#include "multipoles.h"
#include "average_multipoles.h"
#include "covariance.h"

typedef struct ref_state {
    coffe_parameters_t par;
    coffe_background_t bg;
    coffe_integral_array_t integral;
    coffe_corrfunc_array_t cf;
    coffe_multipoles_array_t mp;
    coffe_average_multipoles_array_t ramp;
    coffe_covariance_array_t cov;
    coffe_average_covariance_array_t cov_ramp;
} ref_state;

void *ref_new(void)
{
    ref_state *s = (ref_state *)calloc(1, sizeof(ref_state));
    coffe_parse_default_parameters(&s->par);
    s->bg.flag = 0;
    s->integral.size = 0; s->integral.array = NULL;
    return s;
}

static void free_outputs(ref_state *s)
{
    coffe_corrfunc_free(&s->cf);
    coffe_multipoles_free(&s->mp);
    coffe_average_multipoles_free(&s->ramp);
    coffe_covariance_free(&s->cov);
    coffe_average_covariance_free(&s->cov_ramp);
}

void ref_invalidate(void *h)
{
    ref_state *s = (ref_state *)h;
    coffe_background_free(&s->bg);
    coffe_integrals_free(&s->integral);
    free_outputs(s);
}

void ref_free(void *h)
{
    ref_state *s = (ref_state *)h;
    ref_invalidate(h);
    coffe_parameters_free(&s->par);
    free(s);
}

int ref_nthreads(void *h) { return ((ref_state *)h)->par.nthreads; }
void ref_set_nthreads(void *h, int n) { ((ref_state *)h)->par.nthreads = n; }

/* coffe.pyx:2929-2995 (set_power_spectrum_linear) */
void ref_set_power_spectrum(void *h, const double *k, const double *pk, int n)
{
    ref_state *s = (ref_state *)h;
    coffe_free_spline(&s->par.power_spectrum);
    coffe_free_spline(&s->par.power_spectrum_norm);
    s->par.k_min = k[0]; s->par.k_max = k[n - 1];
    s->par.k_min_norm = k[0]; s->par.k_max_norm = k[n - 1];
    coffe_init_spline(&s->par.power_spectrum, k, pk, n, s->par.interp_method);
    coffe_init_spline(&s->par.power_spectrum_norm, k, pk, n, s->par.interp_method);
    coffe_integrals_free(&s->integral);
    free_outputs(s);
}

/* coffe.pyx:1171-1300: omega_m/omega_baryon/omega_gamma/h/w0/wa setters + _balance_content */
void ref_set_cosmology(void *h, double omega_m, double omega_baryon, double omega_gamma, double hh, double w0, double wa)
{
    ref_state *s = (ref_state *)h;
    s->par.Omega0_m = omega_m;
    s->par.Omega0_baryon = omega_baryon;
    s->par.Omega0_gamma = omega_gamma;
    s->par.Omega0_cdm = omega_m - omega_baryon - s->par.Omega0_nu;
    s->par.h = hh;
    s->par.w0 = w0; s->par.wa = wa;
    s->par.Omega0_de = 1 - s->par.Omega0_m - s->par.Omega0_gamma;
    ref_invalidate(h);
}

/* order: den rsd len d1 d2 g1 g2 g3 g4 g5 */
void ref_set_contributions(void *h, const int *c)
{
    ref_state *s = (ref_state *)h;
    s->par.correlation_contrib.den = c[0]; s->par.correlation_contrib.rsd = c[1];
    s->par.correlation_contrib.len = c[2]; s->par.correlation_contrib.d1 = c[3];
    s->par.correlation_contrib.d2 = c[4]; s->par.correlation_contrib.g1 = c[5];
    s->par.correlation_contrib.g2 = c[6]; s->par.correlation_contrib.g3 = c[7];
    s->par.correlation_contrib.g4 = c[8]; s->par.correlation_contrib.g5 = c[9];
    /* coffe.pyx:1108-1109 (_check_contributions) */
    const int old_divergent = s->par.divergent;
    if (c[4] || c[5] || c[6] || c[7] || c[8] || c[9]) s->par.divergent = 1;
    /* the I^n_l tables depend on the contributions only through `divergent` */
    if (s->par.divergent != old_divergent) coffe_integrals_free(&s->integral);
    free_outputs(s);
}

void ref_set_divergent(void *h, int d) { ((ref_state *)h)->par.divergent = d; coffe_integrals_free(&((ref_state *)h)->integral); }

void ref_set_flags(void *h, int flatsky_local, int flatsky_local_nonlocal, int flatsky_nonlocal, int only_cross)
{
    ref_state *s = (ref_state *)h;
    s->par.flatsky_local = flatsky_local;
    s->par.flatsky_local_nonlocal = flatsky_local_nonlocal;
    s->par.flatsky_nonlocal = flatsky_nonlocal;
    s->par.only_cross_correlations = only_cross;
    ref_invalidate(h);
}

static void set_darray(double **dst, size_t *len, const double *src, int n)
{
    if (*dst) free(*dst);
    *dst = (double *)coffe_malloc(sizeof(double) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) (*dst)[i] = src[i];
    *len = n;
}

void ref_set_sep(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.sep, &s->par.sep_len, v, n); ref_invalidate(h); }
void ref_set_mu(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.mu, &s->par.mu_len, v, n); free_outputs(s); }
void ref_set_z_mean(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.z_mean, &s->par.z_mean_len, v, n); ref_invalidate(h); }
void ref_set_deltaz(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.deltaz, &s->par.deltaz_len, v, n); ref_invalidate(h); }
void ref_set_z_min(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.zmin, &s->par.zmin_len, v, n); ref_invalidate(h); }
void ref_set_z_max(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.zmax, &s->par.zmax_len, v, n); ref_invalidate(h); }
void ref_set_fsky(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.fsky, &s->par.fsky_len, v, n); free_outputs(s); }
void ref_set_pixelsize(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.pixelsize, &s->par.pixelsize_len, v, n); free_outputs(s); }
void ref_set_density1(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.density1, &s->par.density1_len, v, n); free_outputs(s); }
void ref_set_density2(void *h, const double *v, int n) { ref_state *s = h; set_darray(&s->par.density2, &s->par.density2_len, v, n); free_outputs(s); }

void ref_set_l(void *h, const int *v, int n)
{
    ref_state *s = h;
    if (s->par.multipole_values) free(s->par.multipole_values);
    s->par.multipole_values = (int *)coffe_malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) s->par.multipole_values[i] = v[i];
    s->par.multipole_values_len = n;
    ref_invalidate(h);
}

/* which: 0..3 galaxy_bias1..4, 4,5 magnification_bias1,2, 6,7 evolution_bias1,2 (coffe.pyx:1728-1932) */
void ref_set_bias(void *h, int which, const double *z, const double *v, int n)
{
    ref_state *s = h;
    coffe_interpolation *t[8] = {
        &s->par.galaxy_bias1, &s->par.galaxy_bias2, &s->par.galaxy_bias3, &s->par.galaxy_bias4,
        &s->par.magnification_bias1, &s->par.magnification_bias2,
        &s->par.evolution_bias1, &s->par.evolution_bias2};
    coffe_free_spline(t[which]);
    coffe_init_spline(t[which], z, v, n, s->par.interp_method);
    ref_invalidate(h);
}

void ref_set_integration(void *h, int integration_bins, double prec_1d, int type_1d)
{
    ref_state *s = h;
    if (integration_bins > 0) s->par.integration_bins = integration_bins;
    if (prec_1d > 0) s->par.integration_1d_prec = prec_1d;
    if (type_1d > 0) s->par.integration_1d_type = type_1d;
    free_outputs(s);
}

void ref_set_sampling(void *h, int background_bins, int bessel_bins)
{
    ref_state *s = h;
    if (background_bins > 0) s->par.background_bins = background_bins;
    if (bessel_bins > 0) s->par.bessel_bins = bessel_bins;
    ref_invalidate(h);
}

/* covariance knobs: populations, term switches, windowed (binned) flag, integration method */
void ref_set_covariance(void *h, const int *pops, int cosmic, int mixed, int poisson, int window, int method)
{
    ref_state *s = h;
    s->par.covariance_pop1 = pops[0]; s->par.covariance_pop2 = pops[1];
    s->par.covariance_pop3 = pops[2]; s->par.covariance_pop4 = pops[3];
    s->par.covariance_cosmic = cosmic; s->par.covariance_mixed = mixed; s->par.covariance_poisson = poisson;
    s->par.covariance_window = window;
    if (method > 0) s->par.covariance_integration_method = method;
    free_outputs(s);
}

/* par->covariance_integration_bins: sampling of the 2-D FFTLog covariance (src/parser.c:809) */
void ref_set_covariance_bins(void *h, int bins)
{
    ref_state *s = h;
    s->par.covariance_integration_bins = bins;
    free_outputs(s);
}

int ref_background(void *h)
{
    ref_state *s = h;
    if (!s->bg.flag) coffe_background_init(&s->par, &s->bg);
    return 0;
}

int ref_integrals(void *h)
{
    ref_state *s = h;
    ref_background(h);
    if (!s->integral.size) coffe_integrals_init(&s->par, &s->bg, &s->integral);
    return 0;
}

/* which: 0 z_as_chi(chi) 1 a 2 Hz 3 conformal_Hz 4 conformal_Hz_prime 5 D1 6 f 7 G1 8 G2 9 comoving_distance */
static coffe_interpolation *bg_pick(ref_state *s, int which)
{
    switch (which) {
        case 0: return &s->bg.z_as_chi; case 1: return &s->bg.a; case 2: return &s->bg.Hz;
        case 3: return &s->bg.conformal_Hz; case 4: return &s->bg.conformal_Hz_prime;
        case 5: return &s->bg.D1; case 6: return &s->bg.f; case 7: return &s->bg.G1;
        case 8: return &s->bg.G2; default: return &s->bg.comoving_distance;
    }
}

int ref_background_size(void *h) { ref_state *s = h; ref_background(h); return (int)s->bg.a.spline->size; }

void ref_background_table(void *h, int which, double *x, double *y)
{
    ref_state *s = h;
    ref_background(h);
    const gsl_spline *sp = bg_pick(s, which)->spline;
    memcpy(x, sp->x, sizeof(double) * sp->size);
    memcpy(y, sp->y, sizeof(double) * sp->size);
}

double ref_background_eval(void *h, int which, double x)
{
    ref_state *s = h;
    ref_background(h);
    return coffe_interp_spline(bg_pick(s, which), x);
}

/* I^n_l tables; half = 1 selects the half-integer (flat-sky) family */
int ref_integral_size(void *h, int n, int l, int half)
{
    ref_state *s = h;
    ref_integrals(h);
    const enum coffe_integer_state st = half ? COFFE_HALF_INTEGER : COFFE_INTEGER;
    coffe_integral_t *it = coffe_find_integral(&s->integral, n, l, st, st);
    if (!it) return 0;
    return (int)it->result.spline->size;
}

void ref_integral_table(void *h, int n, int l, int half, double *x, double *y)
{
    ref_state *s = h;
    ref_integrals(h);
    const enum coffe_integer_state st = half ? COFFE_HALF_INTEGER : COFFE_INTEGER;
    coffe_integral_t *it = coffe_find_integral(&s->integral, n, l, st, st);
    memcpy(x, it->result.spline->x, sizeof(double) * it->result.spline->size);
    memcpy(y, it->result.spline->y, sizeof(double) * it->result.spline->size);
}

double ref_integral_eval(void *h, int n, int l, int half, double r)
{
    ref_state *s = h;
    ref_integrals(h);
    const enum coffe_integer_state st = half ? COFFE_HALF_INTEGER : COFFE_INTEGER;
    coffe_integral_t *it = coffe_find_integral(&s->integral, n, l, st, st);
    return coffe_interp_spline(&it->result, r);
}

/* divergent I^4_0 extras: 201x201 renormalisation grid (z[j*nx+i]) and the zero-separation spline */
int ref_renorm_size(void *h)
{
    ref_state *s = h;
    ref_integrals(h);
    coffe_integral_t *it = coffe_find_integral(&s->integral, 4, 0, COFFE_INTEGER, COFFE_INTEGER);
    if (!it || !it->renormalization.spline) return 0;
    return (int)it->renormalization.spline->xsize;
}

void ref_renorm_table(void *h, double *chi, double *z)
{
    ref_state *s = h;
    coffe_integral_t *it = coffe_find_integral(&s->integral, 4, 0, COFFE_INTEGER, COFFE_INTEGER);
    const gsl_spline2d *sp = it->renormalization.spline;
    memcpy(chi, sp->xarr, sizeof(double) * sp->xsize);
    memcpy(z, sp->zarr, sizeof(double) * sp->xsize * sp->ysize);
}

double ref_renorm_eval(void *h, double chi1, double chi2)
{
    ref_state *s = h;
    ref_integrals(h);
    coffe_integral_t *it = coffe_find_integral(&s->integral, 4, 0, COFFE_INTEGER, COFFE_INTEGER);
    return coffe_interp_spline2d(&it->renormalization, chi1, chi2);
}

int ref_renorm0_size(void *h)
{
    ref_state *s = h;
    ref_integrals(h);
    coffe_integral_t *it = coffe_find_integral(&s->integral, 4, 0, COFFE_INTEGER, COFFE_INTEGER);
    if (!it || !it->renormalization_zero_separation.spline) return 0;
    return (int)it->renormalization_zero_separation.spline->size;
}

void ref_renorm0_table(void *h, double *x, double *y)
{
    ref_state *s = h;
    coffe_integral_t *it = coffe_find_integral(&s->integral, 4, 0, COFFE_INTEGER, COFFE_INTEGER);
    const gsl_spline *sp = it->renormalization_zero_separation.spline;
    memcpy(x, sp->x, sizeof(double) * sp->size);
    memcpy(y, sp->y, sizeof(double) * sp->size);
}

/* ---- stock drivers; `out` is dense in the reference's own loop order ---- */
int ref_multipoles(void *h, double *out)
{
    ref_state *s = h;
    s->par.output_type = MULTIPOLES;
    ref_integrals(h);
    coffe_multipoles_init(&s->par, &s->bg, &s->integral, &s->mp);
    for (size_t i = 0; i < s->mp.size; ++i) out[i] = s->mp.array[i].value;
    return (int)s->mp.size;
}

int ref_corrfunc(void *h, double *out)
{
    ref_state *s = h;
    /* from Python output_type stays MULTIPOLES */
    s->par.output_type = MULTIPOLES;
    ref_integrals(h);
    coffe_corrfunc_init(&s->par, &s->bg, &s->integral, &s->cf);
    for (size_t i = 0; i < s->cf.size; ++i) out[i] = s->cf.array[i].value;
    return (int)s->cf.size;
}

int ref_average_multipoles(void *h, double *out)
{
    ref_state *s = h;
    s->par.output_type = MULTIPOLES;
    ref_integrals(h);
    coffe_average_multipoles_init(&s->par, &s->bg, &s->integral, &s->ramp);
    for (size_t i = 0; i < s->ramp.size; ++i) out[i] = s->ramp.array[i].value;
    return (int)s->ramp.size;
}

/* average = 0: covariance of multipoles; 1: of redshift-averaged multipoles (coffe.pyx:3465-3469, 3519-3523) */
int ref_covariance(void *h, int average, double *out)
{
    ref_state *s = h;
    ref_background(h);
    const int old = s->par.output_type;
    s->par.output_type = average ? COVARIANCE_AVERAGE_MULTIPOLES : COVARIANCE_MULTIPOLES;
    coffe_covariance_init(&s->par, &s->bg, &s->cov, &s->cov_ramp);
    s->par.output_type = old;
    size_t n = 0;
    if (average) { n = s->cov_ramp.size; for (size_t i = 0; i < n; ++i) out[i] = s->cov_ramp.array[i].value; }
    else { n = s->cov.size; for (size_t i = 0; i < n; ++i) out[i] = s->cov.array[i].value; }
    return (int)n;
}

/* single-point access to the reference dispatcher (src/signal.c:472) */
double ref_integrate(void *h, double z_mean, double z_min, double z_max, double sep, double mu, int l,
                     int flag_integral, int flag_output)
{
    ref_state *s = h;
    ref_integrals(h);
    return coffe_integrate(&s->par, &s->bg, &s->integral, z_mean, z_min, z_max, sep, mu, l,
                           (enum coffe_integral_type)flag_integral, (enum coffe_output_type)flag_output);
}

#ifndef COFFE_DROPIN_BUILD   /* everything below needs the reference's own integrands / Cuhre; the drop-in harness stops here */
/* raw integrands, for kernel-level parity of the physics (src/functions.c:89,1486,3460) */
double ref_f_nonintegrated(void *h, double z, double mu, double sep)
{ ref_state *s = h; ref_integrals(h); return functions_nonintegrated(&s->par, &s->bg, &s->integral, z, mu, sep); }
double ref_f_single(void *h, double z, double mu, double sep, double x)
{ ref_state *s = h; ref_integrals(h); return functions_single_integrated(&s->par, &s->bg, &s->integral, z, mu, sep, x); }
double ref_f_double(void *h, double z, double mu, double sep, double x1, double x2)
{ ref_state *s = h; ref_integrals(h); return functions_double_integrated(&s->par, &s->bg, &s->integral, z, mu, sep, x1, x2); }

void ref_f_nonintegrated_v(void *h, int n, const double *z, const double *mu, const double *sep, double *out)
{
    ref_state *s = h; ref_integrals(h);
    #pragma omp parallel for
    for (int i = 0; i < n; ++i) out[i] = functions_nonintegrated(&s->par, &s->bg, &s->integral, z[i], mu[i], sep[i]);
}
void ref_f_single_v(void *h, int n, const double *z, const double *mu, const double *sep, const double *x, double *out)
{
    ref_state *s = h; ref_integrals(h);
    #pragma omp parallel for
    for (int i = 0; i < n; ++i) out[i] = functions_single_integrated(&s->par, &s->bg, &s->integral, z[i], mu[i], sep[i], x[i]);
}
void ref_f_double_v(void *h, int n, const double *z, const double *mu, const double *sep, const double *x1, const double *x2, double *out)
{
    ref_state *s = h; ref_integrals(h);
    #pragma omp parallel for
    for (int i = 0; i < n; ++i) out[i] = functions_double_integrated(&s->par, &s->bg, &s->integral, z[i], mu[i], sep[i], x1[i], x2[i]);
}

/* ---- converged oracle: reference integrands x caller-supplied 1-D rules on [0,1] ----
 * out[3 pieces][n_l]: NONINTEGRATED, SINGLE_INTEGRATED, DOUBLE_INTEGRATED pieces of the multipoles
 * xi_l = (2l+1) int_0^1 du P_l(2u-1) int dx... (src/signal.c:83-237, 521-752), full-sky branches only.
 * (umu,wmu): rule in u for mu = 2u-1; (ux,wx): rule for each line-of-sight variable. */
void ref_multipoles_converged(void *h, double z_mean, double sep, const int *ls, int n_l,
                              const double *umu, const double *wmu, int nmu,
                              const double *ux, const double *wx, int nx, double *out)
{
    ref_state *s = h;
    ref_integrals(h);
    const coffe_correlation_contributions c = s->par.correlation_contrib;
    const int need_ni = c.den || c.rsd || c.d1 || c.d2 || c.g1 || c.g2 || c.g3;
    const int need_si = (c.len || c.g4 || c.g5) && need_ni;
    const int need_di = c.len || c.g4 || c.g5;
    for (int i = 0; i < 3 * n_l; ++i) out[i] = 0.0;
    double *acc = (double *)calloc((size_t)nmu * 3, sizeof(double));
    #pragma omp parallel for schedule(dynamic, 1)
    for (int im = 0; im < nmu; ++im) {
        const double mu = 2 * umu[im] - 1;
        double ni = 0, si = 0, di = 0;
        if (need_ni) ni = functions_nonintegrated(&s->par, &s->bg, &s->integral, z_mean, mu, sep);
        if (need_si)
            for (int i = 0; i < nx; ++i)
                si += wx[i] * functions_single_integrated(&s->par, &s->bg, &s->integral, z_mean, mu, sep, ux[i]);
        if (need_di)
            for (int i = 0; i < nx; ++i) {
                double row = 0;
                for (int j = 0; j < nx; ++j)
                    row += wx[j] * functions_double_integrated(&s->par, &s->bg, &s->integral, z_mean, mu, sep, ux[i], ux[j]);
                di += wx[i] * row;
            }
        acc[3 * im + 0] = ni; acc[3 * im + 1] = si; acc[3 * im + 2] = di;
    }
    for (int il = 0; il < n_l; ++il) {
        const int l = ls[il];
        double a[3] = {0, 0, 0};
        for (int im = 0; im < nmu; ++im) {
            const double mu = 2 * umu[im] - 1;
            double p0 = 1.0, p1 = mu, pl = (l == 0) ? 1.0 : mu;
            for (int k = 1; k < l; ++k) { pl = ((2.0 * k + 1.0) * mu * p1 - k * p0) / (k + 1.0); p0 = p1; p1 = pl; }
            for (int p = 0; p < 3; ++p) a[p] += wmu[im] * pl * acc[3 * im + p];
        }
        for (int p = 0; p < 3; ++p) out[p * n_l + il] = (2 * l + 1) * a[p];
    }
    free(acc);
}

/* converged 2PCF at fixed mu: out[3] = NI, SI, DI pieces (src/signal.c:515-520, 584-606, 695-703) */
void ref_corrfunc_converged(void *h, double z_mean, double sep, double mu,
                            const double *ux, const double *wx, int nx, double *out)
{
    ref_state *s = h;
    ref_integrals(h);
    const coffe_correlation_contributions c = s->par.correlation_contrib;
    const int need_ni = c.den || c.rsd || c.d1 || c.d2 || c.g1 || c.g2 || c.g3;
    const int need_si = (c.len || c.g4 || c.g5) && need_ni;
    const int need_di = c.len || c.g4 || c.g5;
    out[0] = out[1] = out[2] = 0;
    if (need_ni) out[0] = functions_nonintegrated(&s->par, &s->bg, &s->integral, z_mean, mu, sep);
    if (need_si)
        for (int i = 0; i < nx; ++i)
            out[1] += wx[i] * functions_single_integrated(&s->par, &s->bg, &s->integral, z_mean, mu, sep, ux[i]);
    if (need_di) {
        double tot = 0;
        #pragma omp parallel for reduction(+:tot)
        for (int i = 0; i < nx; ++i) {
            double row = 0;
            for (int j = 0; j < nx; ++j)
                row += wx[j] * functions_double_integrated(&s->par, &s->bg, &s->integral, z_mean, mu, sep, ux[i], ux[j]);
            tot += wx[i] * row;
        }
        out[2] = tot;
    }
}

/* ---- scheme-independent truth: the vendored Cuhre called DIRECTLY on the reference's integrands ----------------------
 * The stock path hard-codes epsrel = 5e-4 and maxeval = integration_bins (src/common.c:1098-1118, result used whatever `fail`
 * says).  Here the same library routine runs on the same integrand functions with caller-chosen epsrel / maxeval, so that where
 * it converges (2-D: the local x integrated multipole in (mu, x), src/signal.c:141-187, and the integrated x integrated 2PCF in
 * (x1, x2), src/signal.c:44-80) it is an adaptive, panel-free arbiter for both oracle/converged.py and the kernels.
 *   which = 0: (2l+1) int du dx  P_l(2u-1) functions_single_integrated(z, 2u-1, sep, x)          [2-D]
 *   which = 1:        int dx1 dx2 functions_double_integrated(z, mu, sep, x1, x2)                  [2-D]
 *   which = 2: (2l+1) int du dx1 dx2 P_l(2u-1) functions_double_integrated(z, 2u-1, sep, x1, x2)  [3-D]
 * Cuhre needs >= 2 dimensions, so the 1-D local x integrated 2PCF is not offered.  Returns the integral; error, neval, fail out. */
typedef struct { ref_state *s; double z, sep, mu; int l, which; } cuhre_job;

static int cuhre_direct_integrand(const int *ndim, const cubareal v[], const int *ncomp, cubareal f[], void *p)
{
    const cuhre_job *j = (const cuhre_job *)p;
    ref_state *s = j->s;
    double val;
    if (j->which == 1) val = functions_double_integrated(&s->par, &s->bg, &s->integral, j->z, j->mu, j->sep, v[0], v[1]);
    else {
        const double mu = 2 * v[0] - 1;
        val = j->which == 0 ? functions_single_integrated(&s->par, &s->bg, &s->integral, j->z, mu, j->sep, v[1])
                            : functions_double_integrated(&s->par, &s->bg, &s->integral, j->z, mu, j->sep, v[1], v[2]);
        if (j->l) val *= gsl_sf_legendre_Pl(j->l, mu);
    }
    f[0] = val;
    return 0;
}

double ref_cuhre_direct(void *h, int which, double z, double sep, double mu, int l, double epsrel, int maxeval,
                        double *error, int *neval_out, int *fail_out)
{
    ref_state *s = h;
    ref_integrals(h);
    cuhre_job j = {s, z, sep, mu, l, which};
    int nregions = 0, neval = 0, fail = 0;
    double result[1] = {0}, err[1] = {0}, prob[1] = {0};
    Cuhre(which == 2 ? 3 : 2, 1, cuhre_direct_integrand, &j, 1, epsrel, 0, 0, 1, maxeval, 0, NULL, NULL,
          &nregions, &neval, &fail, result, err, prob);
    const double pre = which == 1 ? 1.0 : 2 * l + 1;
    if (error) *error = pre * err[0];
    if (neval_out) *neval_out = neval;
    if (fail_out) *fail_out = fail;
    return pre * result[0];
}
#endif /* COFFE_DROPIN_BUILD */

#ifdef REF_HAVE_CLASS
/* ---- third flavour only (oracle/Makefile: class): the reference compiled with -DHAVE_CLASS ------------------------------------
 * coffe.pyx:2760-2928 (set_power_spectrum with pk_type / midpoint_approximation): runs CLASS inside the reference
 * (parse_external_power_spectrum, src/parser.c:283-790), which also replaces par->power_spectrum by CLASS's linear P(k, z = 0) and
 * fills the 2-D P(k, z) splines; coffe_integrals_init then adds the 2-D tables I^n_l(z, r) (src/integrals.c:663-767). */
int ref_run_class(void *h, int pk_type, int midpoint)
{
    ref_state *s = h;
    s->par.pk_type = (enum coffe_pk_type)pk_type;
    s->par.midpoint_approximation = midpoint;
    s->par.have_class = 1;
    ref_invalidate(h);
    return parse_external_power_spectrum(&s->par);
}

int ref_power_spectrum_size(void *h) { ref_state *s = h; return (int)s->par.power_spectrum.spline->size; }
void ref_power_spectrum_table(void *h, double *k, double *pk)
{
    ref_state *s = h;
    const gsl_spline *sp = s->par.power_spectrum.spline;
    memcpy(k, sp->x, sizeof(double) * sp->size);
    memcpy(pk, sp->y, sizeof(double) * sp->size);
}
double ref_k_min(void *h) { return ((ref_state *)h)->par.k_min_norm; }
double ref_k_max(void *h) { return ((ref_state *)h)->par.k_max_norm; }

/* the rows P(k_j, z_i) of par->power_spectrum2d (src/parser.c:655-709: what CLASS returned at the 100 redshifts z_i = 3 i/100, the same
   rows coffe_integrals_init transforms into the 2-D tables): rows[i*n_k + j] */
int ref_pk2d_nz(void *h) { ref_state *s = h; return s->par.power_spectrum2d.spline ? (int)s->par.power_spectrum2d.spline->xsize : 0; }
void ref_pk2d_table(void *h, double *z, double *rows)
{
    ref_state *s = h;
    const gsl_spline2d *sp = s->par.power_spectrum2d.spline;
    memcpy(z, sp->xarr, sizeof(double) * sp->xsize);
    for (size_t i = 0; i < sp->xsize; ++i)
        for (size_t j = 0; j < sp->ysize; ++j) rows[i * sp->ysize + j] = sp->zarr[j * sp->xsize + i];
}

/* 2-D table of one (n, l): sizes, then z[nz], r[nr], values[j*nz + i] <-> (z_i, r_j)  (src/integrals.c:745) */
int ref_integral2d_sizes(void *h, int n, int l, int *nz, int *nr)
{
    ref_state *s = h;
    ref_integrals(h);
    coffe_integral_t *it = coffe_find_integral(&s->integral, n, l, COFFE_INTEGER, COFFE_INTEGER);
    if (!it || !it->result2d.spline) return 0;
    *nz = (int)it->result2d.spline->xsize; *nr = (int)it->result2d.spline->ysize;
    return 1;
}
void ref_integral2d_table(void *h, int n, int l, double *z, double *r, double *values)
{
    ref_state *s = h;
    coffe_integral_t *it = coffe_find_integral(&s->integral, n, l, COFFE_INTEGER, COFFE_INTEGER);
    const gsl_spline2d *sp = it->result2d.spline;
    memcpy(z, sp->xarr, sizeof(double) * sp->xsize);
    memcpy(r, sp->yarr, sizeof(double) * sp->ysize);
    memcpy(values, sp->zarr, sizeof(double) * sp->xsize * sp->ysize);
}
double ref_integral2d_eval(void *h, int n, int l, double z, double r)
{
    ref_state *s = h;
    coffe_integral_t *it = coffe_find_integral(&s->integral, n, l, COFFE_INTEGER, COFFE_INTEGER);
    return coffe_interp_spline2d(&it->result2d, z, r);
}
#endif /* REF_HAVE_CLASS */
