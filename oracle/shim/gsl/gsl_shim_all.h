/*
 * TEST INFRASTRUCTURE ONLY -- part of the parity oracle, never linked into the product library.
 *
 * API-compatible stand-in for the subset of the GNU Scientific Library that the COFFE reference sources
 * use (inventory: `grep gsl_ /root/reference/src/*.c`).  GSL itself is NOT present in
 * this image and is not vendored under /root/reference, so the algorithms are restated from their
 * published definitions (Akima 1970 with GSL's end conditions; natural cubic spline; bicubic Hermite
 * patch with spline-estimated derivatives; QUADPACK QAG with the 61-point Gauss-Kronrod rule;
 * embedded RK2(3) with GSL's standard step control).  Parity below ~1e-6 against "real GSL" is
 * therefore unpinned; it is anchored end-to-end on the reference's golden files (tests/golden).
 *
 * Every gsl header name the reference includes forwards to this file.
 */
#ifndef COFFE_B200_ORACLE_GSL_SHIM_ALL_H
#define COFFE_B200_ORACLE_GSL_SHIM_ALL_H

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSL_MAJOR_VERSION 2
#define GSL_MINOR_VERSION 7
#define GSL_VERSION "2.7-coffe-b200-oracle-shim"

#define GSL_SUCCESS 0
#define GSL_FAILURE (-1)
#define GSL_EDOM 1
#define GSL_NAN (NAN)
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_DBL_MIN 2.2250738585072014e-308

/* ---- errno ---- */
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler_off(void);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *new_handler);

/* ---- math ---- */
int gsl_finite(const double x);

/* ---- 1-D interpolation ---- */
typedef struct gsl_interp_type {
    const char *name;
    unsigned int min_size;
    int kind; /* shim-private discriminator */
} gsl_interp_type;

extern const gsl_interp_type *gsl_interp_linear;
extern const gsl_interp_type *gsl_interp_polynomial;
extern const gsl_interp_type *gsl_interp_cspline;
extern const gsl_interp_type *gsl_interp_cspline_periodic;
extern const gsl_interp_type *gsl_interp_akima;
extern const gsl_interp_type *gsl_interp_akima_periodic;
extern const gsl_interp_type *gsl_interp_steffen;

typedef struct gsl_interp_accel {
    size_t cache;
} gsl_interp_accel;

gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);

typedef struct gsl_interp {
    const gsl_interp_type *type;
    double xmin, xmax;
    size_t size;
    /* per-interval cubic coefficients y = y_i + d*(b + d*(c + d*dd)) */
    double *b, *c, *d;
} gsl_interp;

typedef struct gsl_spline {
    gsl_interp *interp;
    double *x;
    double *y;
    size_t size;
} gsl_spline;

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *spline, const double xa[], const double ya[], size_t size);
double gsl_spline_eval(const gsl_spline *spline, double x, gsl_interp_accel *a);
double gsl_spline_eval_deriv(const gsl_spline *spline, double x, gsl_interp_accel *a);
void gsl_spline_free(gsl_spline *spline);

/* ---- 2-D interpolation ---- */
typedef struct gsl_interp2d_type {
    const char *name;
    unsigned int min_size;
    int kind;
} gsl_interp2d_type;

extern const gsl_interp2d_type *gsl_interp2d_bilinear;
extern const gsl_interp2d_type *gsl_interp2d_bicubic;

typedef struct gsl_spline2d {
    const gsl_interp2d_type *type;
    double *xarr, *yarr, *zarr;
    size_t xsize, ysize;
    double *zx, *zy, *zxy;
} gsl_spline2d;

gsl_spline2d *gsl_spline2d_alloc(const gsl_interp2d_type *T, size_t xsize, size_t ysize);
int gsl_spline2d_init(gsl_spline2d *interp, const double xa[], const double ya[], const double za[],
                      size_t xsize, size_t ysize);
double gsl_spline2d_eval(const gsl_spline2d *interp, const double x, const double y,
                         gsl_interp_accel *xa, gsl_interp_accel *ya);
void gsl_spline2d_free(gsl_spline2d *interp);

/* ---- integration ---- */
typedef struct gsl_function {
    double (*function)(double x, void *params);
    void *params;
} gsl_function;

enum { GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21 = 2, GSL_INTEG_GAUSS31 = 3,
       GSL_INTEG_GAUSS41 = 4, GSL_INTEG_GAUSS51 = 5, GSL_INTEG_GAUSS61 = 6 };

typedef struct gsl_integration_workspace {
    size_t limit;
    size_t size;
    double *alist, *blist, *rlist, *elist;
    size_t *heap; /* max-heap of interval indices keyed by elist */
} gsl_integration_workspace;

gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *workspace,
                        double *result, double *abserr);

/* ---- special functions ---- */
double gsl_sf_bessel_j0(const double x);
double gsl_sf_bessel_j1(const double x);
double gsl_sf_bessel_jl(const int l, const double x);
double gsl_sf_bessel_Jn(const int n, const double x);
double gsl_sf_legendre_P1(double x);
double gsl_sf_legendre_P2(double x);
double gsl_sf_legendre_Pl(const int l, const double x);
double gsl_sf_gamma(const double x);
double gsl_sf_fact(const unsigned int n);
double gsl_sf_choose(unsigned int n, unsigned int m);
double gsl_sf_Si(const double x);
double gsl_sf_coupling_3j(int two_ja, int two_jb, int two_jc, int two_ma, int two_mb, int two_mc);

/* ---- vectors / matrices / least squares (only what background.c touches) ---- */
typedef struct gsl_vector { size_t size; double *data; } gsl_vector;
typedef struct gsl_matrix { size_t size1, size2; double *data; int owner; } gsl_matrix;
typedef struct gsl_matrix_view { gsl_matrix matrix; } gsl_matrix_view;

gsl_vector *gsl_vector_alloc(const size_t n);
void gsl_vector_free(gsl_vector *v);
void gsl_vector_set(gsl_vector *v, const size_t i, double x);
double gsl_vector_get(const gsl_vector *v, const size_t i);
gsl_matrix *gsl_matrix_alloc(const size_t n1, const size_t n2);
void gsl_matrix_free(gsl_matrix *m);
void gsl_matrix_set(gsl_matrix *m, const size_t i, const size_t j, const double x);
double gsl_matrix_get(const gsl_matrix *m, const size_t i, const size_t j);
gsl_matrix_view gsl_matrix_view_array(double *base, const size_t n1, const size_t n2);

typedef struct gsl_multifit_linear_workspace { size_t n, p; } gsl_multifit_linear_workspace;
gsl_multifit_linear_workspace *gsl_multifit_linear_alloc(const size_t n, const size_t p);
void gsl_multifit_linear_free(gsl_multifit_linear_workspace *w);
int gsl_multifit_linear(const gsl_matrix *X, const gsl_vector *y, gsl_vector *c, gsl_matrix *cov,
                        double *chisq, gsl_multifit_linear_workspace *work);

/* ---- ODE (odeiv2: rk2 stepper, standard y control, evolve) ---- */
typedef struct gsl_odeiv2_system {
    int (*function)(double t, const double y[], double dydt[], void *params);
    int (*jacobian)(double t, const double y[], double *dfdy, double dfdt[], void *params);
    size_t dimension;
    void *params;
} gsl_odeiv2_system;

typedef struct gsl_odeiv2_step_type { const char *name; int can_use_dydt_in; int gives_exact_dydt_out; unsigned int order; } gsl_odeiv2_step_type;
extern const gsl_odeiv2_step_type *gsl_odeiv2_step_rk2;

typedef struct gsl_odeiv2_step {
    const gsl_odeiv2_step_type *type;
    size_t dimension;
    double *k1, *k2, *k3, *ytmp;
} gsl_odeiv2_step;

typedef struct gsl_odeiv2_control {
    double eps_abs, eps_rel, a_y, a_dydt;
} gsl_odeiv2_control;

typedef struct gsl_odeiv2_evolve {
    size_t dimension;
    double *y0, *yerr, *dydt_in, *dydt_out;
    double last_step;
    unsigned long int count;
    unsigned long int failed_steps;
} gsl_odeiv2_evolve;

gsl_odeiv2_step *gsl_odeiv2_step_alloc(const gsl_odeiv2_step_type *T, size_t dim);
void gsl_odeiv2_step_free(gsl_odeiv2_step *s);
gsl_odeiv2_control *gsl_odeiv2_control_y_new(double eps_abs, double eps_rel);
void gsl_odeiv2_control_free(gsl_odeiv2_control *c);
gsl_odeiv2_evolve *gsl_odeiv2_evolve_alloc(size_t dim);
void gsl_odeiv2_evolve_free(gsl_odeiv2_evolve *e);
int gsl_odeiv2_evolve_apply(gsl_odeiv2_evolve *e, gsl_odeiv2_control *con, gsl_odeiv2_step *step,
                            const gsl_odeiv2_system *dydt, double *t, double t1, double *h, double y[]);

/* ---- Monte-Carlo integration: declarations only, so that src/common.c compiles WITHOUT -DHAVE_CUBA (the drop-in build of
   dropin/Makefile keeps coffe_integrate_multidimensional out of every call path; the stubs in gsl_shim.c abort if reached) ---- */
typedef struct gsl_monte_function { double (*f)(double *x, size_t dim, void *params); size_t dim; void *params; } gsl_monte_function;
typedef struct gsl_rng_type { const char *name; } gsl_rng_type;
typedef struct gsl_rng { const gsl_rng_type *type; } gsl_rng;
extern const gsl_rng_type *gsl_rng_default;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
typedef struct gsl_monte_plain_state { size_t dim; } gsl_monte_plain_state;
typedef struct gsl_monte_miser_state { size_t dim; } gsl_monte_miser_state;
typedef struct gsl_monte_vegas_state { size_t dim; } gsl_monte_vegas_state;
gsl_monte_plain_state *gsl_monte_plain_alloc(size_t dim);
gsl_monte_miser_state *gsl_monte_miser_alloc(size_t dim);
gsl_monte_vegas_state *gsl_monte_vegas_alloc(size_t dim);
void gsl_monte_plain_free(gsl_monte_plain_state *s);
void gsl_monte_miser_free(gsl_monte_miser_state *s);
void gsl_monte_vegas_free(gsl_monte_vegas_state *s);
int gsl_monte_plain_integrate(const gsl_monte_function *f, const double xl[], const double xu[], size_t dim, size_t calls, gsl_rng *r,
                              gsl_monte_plain_state *state, double *result, double *abserr);
int gsl_monte_miser_integrate(gsl_monte_function *f, const double xl[], const double xu[], size_t dim, size_t calls, gsl_rng *r,
                              gsl_monte_miser_state *state, double *result, double *abserr);
int gsl_monte_vegas_integrate(gsl_monte_function *f, double xl[], double xu[], size_t dim, size_t calls, gsl_rng *r,
                              gsl_monte_vegas_state *state, double *result, double *abserr);

#ifdef __cplusplus
}
#endif

#endif
