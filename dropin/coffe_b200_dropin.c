/*
 * libcoffe_dropin -- COFFE's own driver symbols on top of libcoffe_b200.
 *
 * This file is what replaces src/integrals.c, signal.c, functions.c, corrfunc.c, multipoles.c, average_multipoles.c,
 * covariance.c, twofast.c and twobessel.c in a COFFE build: it exports the symbols the Cython extension binds
 * (python/coffe/coffe.pxd:399-558) and the C CLI calls (src/main.c:125-182),
 *
 *   coffe_integrals_init / _free, coffe_integral_new, coffe_find_integral        src/integrals.h:42-92
 *   coffe_integrate                                                              src/signal.h:17-29
 *   coffe_corrfunc_init / _free                                                  src/corrfunc.h:36-45
 *   coffe_multipoles_init / _free                                                src/multipoles.h:47-56
 *   coffe_average_multipoles_init / _free                                        src/average_multipoles.h:36-45
 *   coffe_covariance_init / _free, coffe_average_covariance_free, coffe_covariance_find   src/covariance.h:48-70
 *
 * with the reference's signatures, struct layouts (it is compiled against the reference's own headers), result order and
 * ownership rules (the callee frees a previous result and coffe_malloc's `array`; size == 0 means "not computed";
 * src/multipoles.c:57, 69-73, 135-144), and forwards the arithmetic to the sm_100a kernels through the C ABI of
 * include/coffe_b200.h.  Everything else of COFFE (common.c, parser.c, background.c, errors.c: parameters, host splines,
 * background ODE) is compiled unchanged next to it (dropin/Makefile).  There is no CPU fallback: without a GPU the drivers
 * print the library's message and return EXIT_FAILURE.
 *
 * All visible GPUs are used by every call (coffe_b200_group_*), like the reference uses every host core
 * (par->nthreads = omp_get_max_threads(), src/parser.c:737-741).  COFFE_B200_DEVICES=n limits the group to n devices.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "background.h"
#include "integrals.h"
#include "signal.h"
#include "corrfunc.h"
#include "multipoles.h"
#include "average_multipoles.h"
#include "covariance.h"

#include "coffe_b200.h"

static coffe_b200_group *g_group;

static int fail(const char *what)
{
    fprintf(stderr, "ERROR: libcoffe_b200: %s: %s\n", what, coffe_b200_group_last_error(g_group));
    return EXIT_FAILURE;
}

static int ensure_group(void)
{
    if (g_group) return 0;
    const char *env = getenv("COFFE_B200_DEVICES");
    const int n = env && atoi(env) > 0 ? atoi(env) : -1;
    if (coffe_b200_group_create(n, &g_group)) {
        fprintf(stderr, "ERROR: libcoffe_b200: %s\n", coffe_b200_group_last_error(NULL));
        g_group = NULL;
        return EXIT_FAILURE;
    }
    return 0;
}

/* the knots of a reference spline (gsl_spline exposes x, y, size; python/coffe/coffe.pxd:20-23) */
static coffe_b200_spline view(const coffe_interpolation *s)
{
    coffe_b200_spline v = {NULL, NULL, 0};
    if (s && s->spline) { v.x = s->spline->x; v.y = s->spline->y; v.n = (int)s->spline->size; }
    return v;
}

static coffe_b200_options options_of(const coffe_parameters_t *par)
{
    coffe_b200_options o;
    memset(&o, 0, sizeof(o));
    o.den = par->correlation_contrib.den; o.rsd = par->correlation_contrib.rsd; o.len = par->correlation_contrib.len;
    o.d1 = par->correlation_contrib.d1; o.d2 = par->correlation_contrib.d2;
    o.g1 = par->correlation_contrib.g1; o.g2 = par->correlation_contrib.g2; o.g3 = par->correlation_contrib.g3;
    o.g4 = par->correlation_contrib.g4; o.g5 = par->correlation_contrib.g5;
    o.only_cross_correlations = par->only_cross_correlations;
    o.flatsky_local = par->flatsky_local; o.flatsky_local_nonlocal = par->flatsky_local_nonlocal; o.flatsky_nonlocal = par->flatsky_nonlocal;
    o.midpoint = par->pk_type != COFFE_PK_LINEAR && par->midpoint_approximation;      /* src/functions.c:136-155 */
    return o;
}

/* hand the state the reference drivers read (par, bg, integral) to every device; `integral` may be NULL (covariance) */
static int upload(const coffe_parameters_t *par, const coffe_background_t *bg, const coffe_integral_array_t *integral)
{
    coffe_b200_tables T;
    memset(&T, 0, sizeof(T));
    T.background[COFFE_B200_BG_Z_AS_CHI] = view(&bg->z_as_chi);
    T.background[COFFE_B200_BG_A] = view(&bg->a);
    T.background[COFFE_B200_BG_HZ] = view(&bg->Hz);
    T.background[COFFE_B200_BG_CONFORMAL_HZ] = view(&bg->conformal_Hz);
    T.background[COFFE_B200_BG_CONFORMAL_HZ_PRIME] = view(&bg->conformal_Hz_prime);
    T.background[COFFE_B200_BG_D1] = view(&bg->D1);
    T.background[COFFE_B200_BG_F] = view(&bg->f);
    T.background[COFFE_B200_BG_G1] = view(&bg->G1);
    T.background[COFFE_B200_BG_G2] = view(&bg->G2);
    T.background[COFFE_B200_BG_COMOVING_DISTANCE] = view(&bg->comoving_distance);
    T.bias[COFFE_B200_BIAS_GALAXY1] = view(&par->galaxy_bias1);
    T.bias[COFFE_B200_BIAS_GALAXY2] = view(&par->galaxy_bias2);
    T.bias[COFFE_B200_BIAS_GALAXY3] = view(&par->galaxy_bias3);
    T.bias[COFFE_B200_BIAS_GALAXY4] = view(&par->galaxy_bias4);
    T.bias[COFFE_B200_BIAS_MAGNIFICATION1] = view(&par->magnification_bias1);
    T.bias[COFFE_B200_BIAS_MAGNIFICATION2] = view(&par->magnification_bias2);
    T.bias[COFFE_B200_BIAS_EVOLUTION1] = view(&par->evolution_bias1);
    T.bias[COFFE_B200_BIAS_EVOLUTION2] = view(&par->evolution_bias2);
    T.interp_method = par->interp_method;
    const size_t n = integral ? integral->size : 0;
    coffe_b200_integral *I = (coffe_b200_integral *)coffe_malloc(sizeof(coffe_b200_integral) * (n ? n : 1));
    memset(I, 0, sizeof(coffe_b200_integral) * (n ? n : 1));
    int used = 0;
    for (size_t i = 0; i < n; ++i) {
        const coffe_integral_t *it = &integral->array[i];
        if (!it->result.spline) continue;
        I[used].n = it->n; I[used].l = it->l;
        I[used].half_integer = it->state_n == COFFE_HALF_INTEGER;
        I[used].result = view(&it->result);
        if (it->result2d.spline) {                 /* I^n_l(z, r) of the midpoint approximation (HAVE_CLASS builds, src/integrals.c:663-767) */
            const gsl_spline2d *r2 = it->result2d.spline;
            I[used].z2d = r2->xarr; I[used].values2d = r2->zarr; I[used].n_z2d = (int)r2->xsize;
        }
        if (it->n == 4 && it->l == 0 && it->state_n == COFFE_INTEGER && it->renormalization.spline) {      /* src/integrals.c:773-914 */
            const gsl_spline2d *r = it->renormalization.spline;
            T.renorm_chi = r->xarr; T.renorm_z = r->zarr; T.n_renorm = (int)r->xsize;
            T.renorm_zero_separation = view(&it->renormalization_zero_separation);
        }
        ++used;
    }
    T.integrals = I; T.n_integrals = used;
    T.omega_matter = par->Omega0_cdm + par->Omega0_baryon;
    T.h = par->h;
    int rc = coffe_b200_group_upload_tables(g_group, &T);
    free(I);
    if (rc) return rc;
    const coffe_b200_options o = options_of(par);
    return coffe_b200_group_set_options(g_group, &o);
}

static int *int_copy(const int *v, size_t n)
{
    int *o = (int *)coffe_malloc(sizeof(int) * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) o[i] = v[i];
    return o;
}

/* ---- coffe_integrals_init (src/integrals.c:580-1423) -------------------------------------------------------------------- */
int coffe_integral_new(coffe_integral_t *integral)
{
    coffe_new_spline(&integral->result);
    coffe_new_spline(&integral->renormalization_zero_separation);
    coffe_new_spline2d(&integral->renormalization);
    coffe_new_spline2d(&integral->result2d);
    return EXIT_SUCCESS;
}

coffe_integral_t *coffe_find_integral(const coffe_integral_array_t *integral, const int n, const int l,
                                      const enum coffe_integer_state state_n, const enum coffe_integer_state state_l)
{
    if (!integral) return NULL;
    for (size_t i = 0; i < integral->size; ++i)
        if (integral->array[i].n == n && integral->array[i].l == l && integral->array[i].state_n == state_n && integral->array[i].state_l == state_l)
            return &integral->array[i];
    return NULL;
}

int coffe_integrals_free(coffe_integral_array_t *integral)
{
    if (integral->size) {
        for (size_t i = 0; i < integral->size; ++i) {
            coffe_free_spline(&integral->array[i].result);
            coffe_free_spline(&integral->array[i].renormalization_zero_separation);
            coffe_free_spline2d(&integral->array[i].renormalization);
            coffe_free_spline2d(&integral->array[i].result2d);
        }
        free(integral->array);
    }
    integral->array = NULL;
    integral->size = 0;
    return EXIT_SUCCESS;
}

#define MAX_TABLES 64

static int add_table(int *n, int *l, int *half, int count, int nn, int ll, int hh)
{
    for (int i = 0; i < count; ++i)
        if (n[i] == nn && l[i] == ll && half[i] == hh) return count;      /* the reference appends duplicates; lookups return the first */
    if (count >= MAX_TABLES) return count;
    n[count] = nn; l[count] = ll; half[count] = hh;
    return count + 1;
}

int coffe_integrals_init(const coffe_parameters_t *par, const coffe_background_t *bg, coffe_integral_array_t *integral)
{
    coffe_integrals_free(integral);
    if (ensure_group()) return EXIT_FAILURE;
    /* the (n, l) list of the reference for the current switches: the eight full-sky pairs (src/integrals.c:597-608), the
       half-integer (1, -1) and, for the multipoles, I^1_l / the half-integer family of the flat-sky closed forms (:916-1413) */
    static const int fam[8][2] = {{0, 0}, {0, 2}, {0, 4}, {1, 1}, {1, 3}, {2, 0}, {2, 2}, {3, 1}};
    int n[MAX_TABLES], l[MAX_TABLES], half[MAX_TABLES], nt = 0;
    for (int i = 0; i < 8; ++i) nt = add_table(n, l, half, nt, fam[i][0], fam[i][1], 0);
    if (par->flatsky_local_nonlocal || par->flatsky_nonlocal) nt = add_table(n, l, half, nt, 1, -1, 1);
    if (par->flatsky_nonlocal && par->output_type == MULTIPOLES)
        for (size_t i = 0; i < par->multipole_values_len; ++i) nt = add_table(n, l, half, nt, 1, par->multipole_values[i], 0);
    if (par->flatsky_local_nonlocal && par->output_type == MULTIPOLES)
        for (size_t i = 0; i < par->multipole_values_len; ++i)
            for (int k = 0; k <= par->multipole_values[i] / 2; ++k)
                nt = add_table(n, l, half, nt, par->multipole_values[i] - 2 * k + 3, par->multipole_values[i] - 2 * k + 1, 1);

    coffe_b200_integrals_params p;
    memset(&p, 0, sizeof(p));
    p.pk_k = par->power_spectrum_norm.spline->x; p.pk_p = par->power_spectrum_norm.spline->y;
    p.n_pk = (int)par->power_spectrum_norm.spline->size;
    p.interp_method = par->interp_method;
    p.k_min = par->k_min_norm; p.k_max = par->k_max_norm;
    p.bessel_bins = (int)par->bessel_bins;
    p.n = n; p.l = l; p.half = half; p.n_tables = nt;
    p.divergent = par->divergent;
    if (par->divergent) {
        /* end of the renormalisation grid (src/integrals.c:858-881) */
        double zmax = 0;
        if (par->output_type == CORRFUNC || par->output_type == MULTIPOLES)
            for (size_t i = 0; i < par->z_mean_len; ++i) { const double v = par->z_mean[i] + par->deltaz[i]; if (i == 0 || v > zmax) zmax = v; }
        else if (par->output_type == AVERAGE_MULTIPOLES)
            for (size_t i = 0; i < par->zmax_len; ++i) if (i == 0 || par->zmax[i] > zmax) zmax = par->zmax[i];
        p.chi_max = zmax > 0 ? coffe_interp_spline(&bg->comoving_distance, zmax) : 0.0;
    }
    int n_r = 0, n_r40 = 0, n_ren = 0;
    if (coffe_b200_integrals_sizes(&p, &n_r, &n_r40, &n_ren)) { fprintf(stderr, "ERROR: libcoffe_b200: bad table request\n"); return EXIT_FAILURE; }
    double *r = (double *)coffe_malloc(sizeof(double) * n_r), *tab = (double *)coffe_malloc(sizeof(double) * (size_t)n_r * nt);
    double *r40 = (double *)coffe_malloc(sizeof(double) * (n_r40 + 1)), *i40 = (double *)coffe_malloc(sizeof(double) * (n_r40 + 1));
    double *ren0 = (double *)coffe_malloc(sizeof(double) * (n_r40 + 1));
    double *chi = (double *)coffe_malloc(sizeof(double) * (n_ren + 1)), *z = (double *)coffe_malloc(sizeof(double) * ((size_t)n_ren * n_ren + 1));
    coffe_b200_ctx *ctx = coffe_b200_group_member(g_group, 0);         /* a 10 ms set-up job: one device */
    const int rc = coffe_b200_integrals(ctx, &p, r, tab, par->divergent ? r40 : NULL, par->divergent ? i40 : NULL, par->divergent ? ren0 : NULL,
                                        par->divergent ? chi : NULL, par->divergent ? z : NULL);
    if (rc) {
        fprintf(stderr, "ERROR: libcoffe_b200: coffe_integrals_init: %s\n", coffe_b200_last_error(ctx));
        free(r); free(tab); free(r40); free(i40); free(ren0); free(chi); free(z);
        return EXIT_FAILURE;
    }
    /* same order as the reference: the eight pairs, then the divergent (4, 0), then the flat-sky tables */
    integral->array = (coffe_integral_t *)coffe_malloc(sizeof(coffe_integral_t) * (nt + 1));
    size_t out = 0;
    for (int t = 0; t < nt; ++t) {
        if (t == 8 && par->divergent) {
            coffe_integral_t *d = &integral->array[out++];
            coffe_integral_new(d);
            d->n = 4; d->l = 0; d->state_n = d->state_l = COFFE_INTEGER;
            coffe_init_spline(&d->result, r40, i40, n_r40, par->interp_method);
            r40[0] = 0.0;                                                          /* src/integrals.c:838-839 */
            coffe_init_spline(&d->renormalization_zero_separation, r40, ren0, n_r40, par->interp_method);
            coffe_init_spline2d(&d->renormalization, chi, chi, z, n_ren, n_ren, par->covariance_interpolation_method);
        }
        coffe_integral_t *it = &integral->array[out++];
        coffe_integral_new(it);
        it->n = n[t]; it->l = l[t];
        it->state_n = it->state_l = half[t] ? COFFE_HALF_INTEGER : COFFE_INTEGER;
        double *col = tab + (size_t)t * n_r;
        if (half[t] && n[t] == 1 && l[t] == -1)                                    /* src/integrals.c:955-961 */
            for (int i = 0; i < n_r; ++i) col[i] *= sqrt(2. / M_PI) * 2 * M_PI * M_PI;
        coffe_init_spline(&it->result, r, col, n_r, par->interp_method);
    }
    if (nt <= 8 && par->divergent) {
        coffe_integral_t *d = &integral->array[out++];
        coffe_integral_new(d);
        d->n = 4; d->l = 0; d->state_n = d->state_l = COFFE_INTEGER;
        coffe_init_spline(&d->result, r40, i40, n_r40, par->interp_method);
        r40[0] = 0.0;
        coffe_init_spline(&d->renormalization_zero_separation, r40, ren0, n_r40, par->interp_method);
        coffe_init_spline2d(&d->renormalization, chi, chi, z, n_ren, n_ren, par->covariance_interpolation_method);
    }
    integral->size = out;
    free(r); free(tab); free(r40); free(i40); free(ren0); free(chi); free(z);
    return EXIT_SUCCESS;
}

/* ---- coffe_multipoles_init (src/multipoles.c:50-133) ------------------------------------------------------------------- */
int coffe_multipoles_free(coffe_multipoles_array_t *mp)
{
    if (mp->size) free(mp->array);
    mp->array = NULL; mp->size = 0;
    return EXIT_SUCCESS;
}

int coffe_multipoles_init(coffe_parameters_t *par, coffe_background_t *bg, coffe_integral_array_t *integral, coffe_multipoles_array_t *mp)
{
    coffe_multipoles_free(mp);
    if (ensure_group()) return EXIT_FAILURE;
    if (upload(par, bg, integral)) return fail("coffe_multipoles_init");
    const size_t nz = par->z_mean_len, nl = par->multipole_values_len, nr = par->sep_len;
    if (nz * nl * nr == 0) return EXIT_SUCCESS;
    double *dense = (double *)coffe_malloc(sizeof(double) * nz * nl * nr);
    if (coffe_b200_group_multipoles(g_group, par->z_mean, (int)nz, par->sep, (int)nr, par->multipole_values, (int)nl, dense)) { free(dense); return fail("coffe_multipoles_init"); }
    mp->size = nz * nl * nr;
    mp->array = (coffe_multipoles_t *)coffe_malloc(sizeof(coffe_multipoles_t) * mp->size);
    size_t c = 0;
    for (size_t i = 0; i < nz; ++i) for (size_t j = 0; j < nl; ++j) for (size_t k = 0; k < nr; ++k, ++c) {      /* src/multipoles.c:75-86 */
        mp->array[c].coords.z_mean = par->z_mean[i];
        mp->array[c].coords.l = par->multipole_values[j];
        mp->array[c].coords.separation = par->sep[k];
        mp->array[c].value = dense[c];
    }
    free(dense);
    return EXIT_SUCCESS;
}

/* ---- coffe_corrfunc_init (src/corrfunc.c:56-139) ------------------------------------------------------------------------ */
int coffe_corrfunc_free(coffe_corrfunc_array_t *cf)
{
    if (cf->size) free(cf->array);
    cf->array = NULL; cf->size = 0;
    return EXIT_SUCCESS;
}

int coffe_corrfunc_init(coffe_parameters_t *par, coffe_background_t *bg, coffe_integral_array_t *integral, coffe_corrfunc_array_t *cf)
{
    coffe_corrfunc_free(cf);
    if (ensure_group()) return EXIT_FAILURE;
    if (upload(par, bg, integral)) return fail("coffe_corrfunc_init");
    const size_t nz = par->z_mean_len, nm = par->mu_len, nr = par->sep_len;
    if (nz * nm * nr == 0) return EXIT_SUCCESS;
    double *dense = (double *)coffe_malloc(sizeof(double) * nz * nm * nr);
    if (coffe_b200_group_corrfunc(g_group, par->z_mean, (int)nz, par->mu, (int)nm, par->sep, (int)nr, dense)) { free(dense); return fail("coffe_corrfunc_init"); }
    cf->size = nz * nm * nr;
    cf->array = (coffe_corrfunc_t *)coffe_malloc(sizeof(coffe_corrfunc_t) * cf->size);
    size_t c = 0;
    for (size_t i = 0; i < nz; ++i) for (size_t j = 0; j < nm; ++j) for (size_t k = 0; k < nr; ++k, ++c) {      /* src/corrfunc.c:81-93 */
        cf->array[c].coords.z_mean = par->z_mean[i];
        cf->array[c].coords.mu = par->mu[j];
        cf->array[c].coords.separation = par->sep[k];
        cf->array[c].value = dense[c];
    }
    free(dense);
    return EXIT_SUCCESS;
}

/* ---- coffe_average_multipoles_init (src/average_multipoles.c:72-156) ---------------------------------------------------- */
int coffe_average_multipoles_free(coffe_average_multipoles_array_t *ramp)
{
    if (ramp->size) free(ramp->array);
    ramp->array = NULL; ramp->size = 0;
    return EXIT_SUCCESS;
}

int coffe_average_multipoles_init(coffe_parameters_t *par, coffe_background_t *bg, coffe_integral_array_t *integral,
                                  coffe_average_multipoles_array_t *ramp)
{
    coffe_average_multipoles_free(ramp);
    if (ensure_group()) return EXIT_FAILURE;
    if (upload(par, bg, integral)) return fail("coffe_average_multipoles_init");
    const size_t nb = par->zmin_len, nl = par->multipole_values_len, nr = par->sep_len;
    if (nb * nl * nr == 0) return EXIT_SUCCESS;
    double *dense = (double *)coffe_malloc(sizeof(double) * nb * nl * nr);
    if (coffe_b200_group_average_multipoles(g_group, par->zmin, par->zmax, (int)nb, par->sep, (int)nr, par->multipole_values, (int)nl, dense)) {
        free(dense); return fail("coffe_average_multipoles_init");
    }
    ramp->size = nb * nl * nr;
    ramp->array = (coffe_average_multipoles_t *)coffe_malloc(sizeof(coffe_average_multipoles_t) * ramp->size);
    size_t c = 0;
    for (size_t i = 0; i < nb; ++i) for (size_t j = 0; j < nl; ++j) for (size_t k = 0; k < nr; ++k, ++c) {      /* src/average_multipoles.c:97-110 */
        ramp->array[c].coords.z_min = par->zmin[i];
        ramp->array[c].coords.z_max = par->zmax[i];
        ramp->array[c].coords.l = par->multipole_values[j];
        ramp->array[c].coords.separation = par->sep[k];
        ramp->array[c].value = dense[c];
    }
    free(dense);
    return EXIT_SUCCESS;
}

/* ---- coffe_integrate (src/signal.c:472-771): one point, one piece ------------------------------------------------------- */
double coffe_integrate(coffe_parameters_t *par, coffe_background_t *bg, coffe_integral_array_t *integral,
                       const double z_mean, const double z_min, const double z_max, const double sep, const double mu, const int l,
                       const enum coffe_integral_type flag_integral, const enum coffe_output_type flag_output)
{
    if (ensure_group() || upload(par, bg, integral)) { fail("coffe_integrate"); exit(EXIT_FAILURE); }     /* the reference aborts too (src/functions.c:1470-1483) */
    coffe_b200_ctx *ctx = coffe_b200_group_member(g_group, 0);
    const int piece = flag_integral == NONINTEGRATED ? 0 : flag_integral == SINGLE_INTEGRATED ? 1 : 2;
    double out[3] = {0, 0, 0};
    int rc;
    switch (flag_output) {
        case CORRFUNC:
            rc = coffe_b200_corrfunc_range_pieces(ctx, &z_mean, 1, &mu, 1, &sep, 1, 0, 1, out);
            break;
        case MULTIPOLES:
            rc = coffe_b200_multipoles_pieces(ctx, &z_mean, 1, &sep, 1, &l, 1, out);
            break;
        case AVERAGE_MULTIPOLES:
            rc = coffe_b200_average_multipoles_range_pieces(ctx, &z_min, &z_max, 1, &sep, 1, &l, 1, 0, 1, out);
            break;
        default:
            fprintf(stderr, "ERROR: coffe_integrate: unsupported output type\n");
            exit(EXIT_FAILURE);
    }
    if (rc) { fprintf(stderr, "ERROR: libcoffe_b200: coffe_integrate: %s\n", coffe_b200_last_error(ctx)); exit(EXIT_FAILURE); }
    return out[piece];
}

/* ---- coffe_covariance_init (src/covariance.c:477-1103) ------------------------------------------------------------------ */
int coffe_covariance_free(coffe_covariance_array_t *cov)
{
    if (cov->size) free(cov->array);
    cov->array = NULL; cov->size = 0;
    return EXIT_SUCCESS;
}

int coffe_average_covariance_free(coffe_average_covariance_array_t *cov)
{
    if (cov->size) free(cov->array);
    cov->array = NULL; cov->size = 0;
    return EXIT_SUCCESS;
}

coffe_covariance_t coffe_covariance_find(const coffe_covariance_array_t *cov, const double z_mean, const int l1, const int l2,
                                         const double sep1, const double sep2)
{
    for (size_t i = 0; i < cov->size; ++i)
        if (cov->array[i].coords.l1 == l1 && cov->array[i].coords.l2 == l2 && coffe_approx_equal(cov->array[i].coords.separation1, sep1, 1e-3, 0) &&
            coffe_approx_equal(cov->array[i].coords.separation2, sep2, 1e-3, 0) && coffe_approx_equal(cov->array[i].coords.z_mean, z_mean, 1e-3, 0))
            return cov->array[i];
    fprintf(stderr, "ERROR: cannot find the appropriate element in the covariance array.\n");
    exit(EXIT_FAILURE);
}

int coffe_covariance_init(coffe_parameters_t *par, coffe_background_t *bg, coffe_covariance_array_t *cov_mp,
                          coffe_average_covariance_array_t *cov_ramp)
{
    if (par->output_type != COVARIANCE_MULTIPOLES && par->output_type != COVARIANCE_AVERAGE_MULTIPOLES) return EXIT_SUCCESS;   /* src/covariance.c:484-487 */
    coffe_covariance_free(cov_mp);
    coffe_average_covariance_free(cov_ramp);
    if (ensure_group()) return EXIT_FAILURE;
    if (upload(par, bg, NULL)) return fail("coffe_covariance_init");
    const int average = par->output_type == COVARIANCE_AVERAGE_MULTIPOLES;
    const size_t nb = average ? par->zmin_len : par->z_mean_len, nl = par->multipole_values_len, nr = par->sep_len;
    coffe_b200_covariance_params p;
    memset(&p, 0, sizeof(p));
    p.pk_k = par->power_spectrum.spline->x; p.pk_p = par->power_spectrum.spline->y; p.n_pk = (int)par->power_spectrum.spline->size;
    p.k_min = par->k_min; p.k_max = par->k_max;
    p.n_bins = (int)nb;
    p.z_mean = par->z_mean; p.deltaz = par->deltaz; p.z_min = par->zmin; p.z_max = par->zmax;
    p.fsky = par->fsky; p.density1 = par->density1; p.density2 = par->density2; p.pixelsize = par->pixelsize;
    p.sep = par->sep; p.n_sep = (int)nr;
    p.l = par->multipole_values; p.n_l = (int)nl;
    p.pop[0] = par->covariance_pop1; p.pop[1] = par->covariance_pop2; p.pop[2] = par->covariance_pop3; p.pop[3] = par->covariance_pop4;
    p.cosmic = par->covariance_cosmic; p.mixed = par->covariance_mixed; p.poisson = par->covariance_poisson;
    p.window = par->covariance_window;
    p.average = average;
    p.integration_method = par->covariance_integration_method;
    p.integration_bins = par->covariance_integration_bins;
    const size_t total = nb * nl * nr * nl * nr;
    if (total == 0) return EXIT_SUCCESS;
    double *dense = (double *)coffe_malloc(sizeof(double) * total);
    if (coffe_b200_group_covariance(g_group, &p, dense)) { free(dense); return fail("coffe_covariance_init"); }
    size_t c = 0;
    if (!average) {
        cov_mp->size = total;
        cov_mp->array = (coffe_covariance_t *)coffe_malloc(sizeof(coffe_covariance_t) * total);
        for (size_t zi = 0; zi < nb; ++zi) for (size_t l1 = 0; l1 < nl; ++l1) for (size_t r1 = 0; r1 < nr; ++r1)
        for (size_t l2 = 0; l2 < nl; ++l2) for (size_t r2 = 0; r2 < nr; ++r2, ++c) {                            /* src/covariance.c:1046-1059 */
            cov_mp->array[c].coords.z_mean = par->z_mean[zi];
            cov_mp->array[c].coords.l1 = par->multipole_values[l1]; cov_mp->array[c].coords.l2 = par->multipole_values[l2];
            cov_mp->array[c].coords.separation1 = par->sep[r1]; cov_mp->array[c].coords.separation2 = par->sep[r2];
            cov_mp->array[c].value = dense[c];
        }
    }
    else {
        cov_ramp->size = total;
        cov_ramp->array = (coffe_average_covariance_t *)coffe_malloc(sizeof(coffe_average_covariance_t) * total);
        for (size_t zi = 0; zi < nb; ++zi) for (size_t l1 = 0; l1 < nl; ++l1) for (size_t r1 = 0; r1 < nr; ++r1)
        for (size_t l2 = 0; l2 < nl; ++l2) for (size_t r2 = 0; r2 < nr; ++r2, ++c) {                            /* src/covariance.c:1062-1077 */
            cov_ramp->array[c].coords.z_min = par->zmin[zi]; cov_ramp->array[c].coords.z_max = par->zmax[zi];
            cov_ramp->array[c].coords.l1 = par->multipole_values[l1]; cov_ramp->array[c].coords.l2 = par->multipole_values[l2];
            cov_ramp->array[c].coords.separation1 = par->sep[r1]; cov_ramp->array[c].coords.separation2 = par->sep[r2];
            cov_ramp->array[c].value = dense[c];
        }
    }
    free(dense);
    return EXIT_SUCCESS;
}

/* for tests and tools: how many devices the drivers are using (0 before the first call) */
int coffe_dropin_devices(void) { return g_group ? coffe_b200_group_size(g_group) : 0; }
