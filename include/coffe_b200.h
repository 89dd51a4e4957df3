/*
 * libcoffe_b200 -- C ABI of the B200-native signal-evaluation path of COFFE.
 *
 * Plain C, no CUDA or torch types in any signature.  Every entry point returns 0 on success and a
 * non-zero COFFE_B200_E* code on failure (never exits, never throws); coffe_b200_last_error() gives the
 * message.  The reference's drivers call exit(EXIT_FAILURE) on non-finite integrands
 * (src/functions.c:1470-1483); here that condition is reported as COFFE_B200_ENONFINITE instead.
 *
 * Which reference interface each entry point replaces (paths relative to the reference root):
 *   coffe_b200_upload_tables      the state the drivers read through coffe_parameters_t /
 *                                 coffe_background_t / coffe_integral_array_t
 *                                 (src/common.h:340-550, src/background.h:23-45, src/integrals.h:25-52)
 *   coffe_b200_multipoles         coffe_multipoles_init            src/multipoles.h:47, src/multipoles.c:50-133
 *   coffe_b200_corrfunc           coffe_corrfunc_init              src/corrfunc.h:36,   src/corrfunc.c:56-139
 *   coffe_b200_average_multipoles coffe_average_multipoles_init    src/average_multipoles.h:36, src/average_multipoles.c:72-156
 *   coffe_b200_multipoles_pieces, coffe_b200_corrfunc_range_pieces, coffe_b200_average_multipoles_range_pieces
 *                                 coffe_integrate (per piece)      src/signal.h:17-29, src/signal.c:472-771
 *   coffe_b200_group_*            the `omp parallel for num_threads(par->nthreads)` of the drivers: one call -> all devices
 *   coffe_b200_covariance         coffe_covariance_init            src/covariance.h:48, src/covariance.c:477-1103
 *   coffe_b200_integrals          coffe_integrals_init             src/integrals.h:84, src/integrals.c:580-1423
 * Output buffers are caller-owned, dense, and in the reference's own loop order so that a patched
 * driver can scatter them into its AoS result arrays.
 */
#ifndef COFFE_B200_H
#define COFFE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct coffe_b200_ctx coffe_b200_ctx;

enum {
    COFFE_B200_OK = 0,
    COFFE_B200_EINVAL = 1,      /* bad argument / unsupported option */
    COFFE_B200_ECUDA = 2,       /* CUDA runtime failure (no device, out of memory, launch error) */
    COFFE_B200_ENOTABLES = 3,   /* compute call before coffe_b200_upload_tables */
    COFFE_B200_ENONFINITE = 4,  /* a result is NaN/inf (argument outside a table's range) */
    COFFE_B200_ERANGE = 5       /* a requested point lies outside the uploaded tables */
};

/* one 1-D table: knots of a reference spline (gsl_spline->x, ->y, ->size; python/coffe/coffe.pxd:20-23) */
typedef struct coffe_b200_spline {
    const double *x;
    const double *y;
    int n;
} coffe_b200_spline;

/* index into coffe_b200_tables.background[] -- the members of coffe_background_t (src/background.h:23-45) */
enum {
    COFFE_B200_BG_Z_AS_CHI = 0,     /* x = chi [Mpc], y = z */
    COFFE_B200_BG_A, COFFE_B200_BG_HZ, COFFE_B200_BG_CONFORMAL_HZ, COFFE_B200_BG_CONFORMAL_HZ_PRIME,
    COFFE_B200_BG_D1, COFFE_B200_BG_F, COFFE_B200_BG_G1, COFFE_B200_BG_G2,
    COFFE_B200_BG_COMOVING_DISTANCE,  /* all of these: x = z, uniform knots (src/background.c:446) */
    COFFE_B200_BG_COUNT
};

/* index into coffe_b200_tables.bias[] (src/common.h: galaxy_bias1..4, magnification_bias1,2, evolution_bias1,2) */
enum {
    COFFE_B200_BIAS_GALAXY1 = 0, COFFE_B200_BIAS_GALAXY2, COFFE_B200_BIAS_GALAXY3, COFFE_B200_BIAS_GALAXY4,
    COFFE_B200_BIAS_MAGNIFICATION1, COFFE_B200_BIAS_MAGNIFICATION2,
    COFFE_B200_BIAS_EVOLUTION1, COFFE_B200_BIAS_EVOLUTION2,
    COFFE_B200_BIAS_COUNT
};

/* one coffe_integral_t (src/integrals.h:25-39): the stored quantity is r^(n-l) I^n_l(r) when n >= l */
typedef struct coffe_b200_integral {
    int n, l;
    int half_integer;            /* COFFE_HALF_INTEGER family (flat-sky closed forms) */
    coffe_b200_spline result;
    /* optional coffe_integral_t.result2d (src/integrals.h:29-30), the table I^n_l(z, r) the reference builds from P(k, z) when it is
       compiled with CLASS and pk_type != linear with midpoint_approximation (src/integrals.c:663-767): n_z2d redshift knots
       (gsl_spline2d->xarr), values2d[j*n_z2d + i] <-> (z2d[i], result.x[j]) (->zarr); n_z2d = 0 when absent */
    const double *z2d;
    const double *values2d;
    int n_z2d;
} coffe_b200_integral;

typedef struct coffe_b200_tables {
    coffe_b200_spline background[COFFE_B200_BG_COUNT];
    coffe_b200_spline bias[COFFE_B200_BIAS_COUNT];
    int interp_method;           /* enum coffe_interp1d_type (src/common.h:271-280); 5 = Akima */
    const coffe_b200_integral *integrals;
    int n_integrals;
    /* divergent I^4_0 extras (src/integrals.c:773-914); n_renorm = 0 if par->divergent is off */
    const double *renorm_chi;    /* n_renorm knots, same on both axes */
    const double *renorm_z;      /* n_renorm^2 values, z[j*n + i] */
    int n_renorm;
    coffe_b200_spline renorm_zero_separation;
    /* scalars of coffe_parameters_t */
    double omega_matter;         /* Omega0_cdm + Omega0_baryon */
    double h;
} coffe_b200_tables;

/* which terms to compute: coffe_correlation_contributions + the flat-sky / cross-correlation switches */
typedef struct coffe_b200_options {
    int den, rsd, len, d1, d2, g1, g2, g3, g4, g5;
    int only_cross_correlations;
    int flatsky_local, flatsky_local_nonlocal, flatsky_nonlocal;
    /* par->pk_type != COFFE_PK_LINEAR && par->midpoint_approximation (src/functions.c:136-155 and the same pattern in every standard
       term): den, rsd, d1 and len auto- and cross-terms read I^n_l((z_a + z_b)/2, r) from the 2-D tables instead of
       D1(z_a) D1(z_b) I^n_l(r); needs result2d of the eight family tables.  Not combined with the flat-sky switches. */
    int midpoint;
} coffe_b200_options;

/* fixed-node quadrature sizes (0 = keep the library default); DESIGN.md holds the convergence evidence.
   See coffe_b200/csrc/quadrature.cuh and rules_build.h for the meaning of each knob. */
typedef struct coffe_b200_quadrature {
    int n_mu_ni;                 /* Gauss-Legendre nodes in mu, local x local piece */
    int n_mu_panel;              /* nodes per r_perp panel of the mu rule of the integrated pieces */
    double mu_rho0, mu_panel_width, mu_uniform_extent;   /* r_perp panel plan [Mpc] */
    int n_outer;                 /* outer line-of-sight nodes (integrated x integrated); -1 = by separation */
    int n_panel_si, n_panel_di;  /* nodes per line-of-sight panel (at most 32: a panel is walked by one group of lanes of a warp) */
    int n_z;                     /* nodes of the redshift average */
    double panel_width, panel_uniform_extent, eps_min, tail_ratio;   /* line-of-sight panel plan [Mpc] */
    /* the integrated x integrated piece has its own mu rule and inner panel plan (it is smoother and dominates the cost) */
    int n_mu_panel_di;
    double panel_width_di, panel_uniform_extent_di;
    double mu_panel_width_di;
    double tail_ratio_di;
    int geometric_shift_di;      /* log2 of the growth factor of its geometric panels: 1 = doubling, 2 = x4 */
    /* What the 1e-5 gate is measured against, which picks the library defaults the zeros above stand for:
       0 = automatic: COFFE_B200_TARGET_TOTAL when density is among the contributions, the uploaded biases keep the integrated terms
           small against it (max |2 - 5 s(z)| / b(z) at most twice its default of 2) and the call returns mu-integrated totals
           (multipoles / average multipoles / their group and sharded forms; the 2-point function at a fixed mu always gets the piece rules), COFFE_B200_TARGET_PIECE otherwise and
           always for the *_pieces entry points (a caller of coffe_integrate(flag_integral) gets each piece to 1e-5 of itself);
       1 = COFFE_B200_TARGET_PIECE: every piece (local x local, local x integrated, integrated x integrated) within 1e-5 of its own
           l = 0..4 scale -- what a lensing-only or cross-only configuration needs;
       2 = COFFE_B200_TARGET_TOTAL: the returned multipole within 1e-5 of the l = 0..4 scale of the point; the integrated pieces are
           then at most a few per cent of the total and carry coarser rules (a third of the evaluations). */
    int accuracy_target;
    /* Geometric panels of the line-of-sight plans under the sinh map d = eps sinh(t) (eps = impact parameter of the two lines of
       sight, d = distance from their closest approach): the region |d| < panel_width is cut into equal panels in t with a growth
       factor of at most this per panel, Gauss-Legendre nodes in t (quadrature.cuh: panel_plan).  0 = library default, (0, 1] = off
       (doubling panels with nodes in d, the round-1 scheme), > 1 = growth factor.  First: local x integrated, second: integrated
       x integrated. */
    double geometric_growth, geometric_growth_di;
    /* nodes of the inner rule of renormalization(lambda_1, lambda_2) in the integrated x integrated piece: two Gauss-Legendre panels
       split at the diagonal, this many nodes together (even; default 32 = one per lane).  I^4_0(R) stays on the ridge panels. */
    int n_renorm;
    /* variable of the Gauss-Legendre nodes on the tail panels beyond the uniform extent: 0 = library default, 1 = d, 2 = log(d)
       (power-law tails are entire functions of log(d): wider panels at the same node count).  First: local x integrated. */
    int tail_map, tail_map_di;
    /* mu rules of the integrated pieces: growth factor of the geometric r_perp panels between mu_rho0 and the panel width, with
       the nodes Gauss-Legendre in log(r_perp) (0 = default, (0, 1] = doubling panels with nodes in phi); and whether the rule of
       the integrated x integrated piece keeps the break at |mu| = 0.999 (0 = default, 1 = yes, 2 = no: only the local x integrated
       integrand switches formulas there, src/functions.c:2199) */
    double mu_geometric_growth, mu_geometric_growth_di;
    int mu_switch_break_di;
    /* outer nodes by separation (n_outer = -1 / default): max(n_outer_min, 10 + sep / n_outer_step [Mpc]) */
    int n_outer_min;
    double n_outer_step;
    double mu_tail_ratio;        /* growth factor of the r_perp panels of the mu rules beyond mu_uniform_extent (default 1.5) */
} coffe_b200_quadrature;
enum { COFFE_B200_TARGET_AUTO = 0, COFFE_B200_TARGET_PIECE = 1, COFFE_B200_TARGET_TOTAL = 2 };

/* ---- life cycle ---------------------------------------------------------------------------------- */
int coffe_b200_create(int device, coffe_b200_ctx **out);
void coffe_b200_destroy(coffe_b200_ctx *ctx);
const char *coffe_b200_last_error(const coffe_b200_ctx *ctx);   /* ctx may be NULL: last create() error */
int coffe_b200_upload_tables(coffe_b200_ctx *ctx, const coffe_b200_tables *tables);
int coffe_b200_set_options(coffe_b200_ctx *ctx, const coffe_b200_options *opt);
int coffe_b200_set_quadrature(coffe_b200_ctx *ctx, const coffe_b200_quadrature *q);
int coffe_b200_get_quadrature(const coffe_b200_ctx *ctx, coffe_b200_quadrature *q);

/* The line-of-sight kernel reads the I^n_l family either from tables resampled into shared memory
   (variant 1, default; resampling error reported by coffe_b200_table_info) or from the
   exact per-interval cubics of the uploaded knots through L1/L2 (variant 0; cross-check). */
int coffe_b200_set_kernel_variant(coffe_b200_ctx *ctx, int variant);
/* variant the last compute call ran, and size / achieved resampling error of the shared-memory tables */
int coffe_b200_table_info(const coffe_b200_ctx *ctx, int *variant_used, int *compact_knots, int *compact_bytes,
                          double *compact_max_rel_error);

/* ---- signal -------------------------------------------------------------------------------------- */
/* out[(iz*n_l + il)*n_sep + ir]   (order of src/multipoles.c:75-86) */
int coffe_b200_multipoles(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                          const int *l, int n_l, double *out);
/* same grid, restricted to the flattened (iz, ir) work items [item_begin, item_end) (iz*n_sep + ir): the
   unit of multi-GPU sharding.  out has n_l*(item_end-item_begin) entries: out[il*n_items + (item-item_begin)] */
int coffe_b200_multipoles_range(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                                const int *l, int n_l, long item_begin, long item_end, double *out);
/* the three pieces coffe_integrate returns for flag_integral = NONINTEGRATED / SINGLE_INTEGRATED /
   DOUBLE_INTEGRATED: out[((piece*n_z + iz)*n_l + il)*n_sep + ir] */
int coffe_b200_multipoles_pieces(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                                 const int *l, int n_l, double *out);
/* out[(iz*n_mu + imu)*n_sep + ir]   (order of src/corrfunc.c:81-93) */
int coffe_b200_corrfunc(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *mu, int n_mu,
                        const double *sep, int n_sep, double *out);
/* out[(ibin*n_l + il)*n_sep + ir]   (order of src/average_multipoles.c:97-110) */
int coffe_b200_average_multipoles(coffe_b200_ctx *ctx, const double *z_min, const double *z_max, int n_bins,
                                  const double *sep, int n_sep, const int *l, int n_l, double *out);

/* Range and piece variants of the two calls above (same role as coffe_b200_multipoles_range / _pieces): the unit of multi-GPU
   sharding and what a single-point coffe_integrate(flag_integral, flag_output) maps to (src/signal.c:472-771).
   corrfunc: flattened output entries (iz*n_mu + imu)*n_sep + ir in [begin, end); out has 3*(end-begin) entries,
             out[piece*(end-begin) + k], pieces NONINTEGRATED / SINGLE_INTEGRATED / DOUBLE_INTEGRATED.
   average multipoles: flattened (ibin*n_sep + ir) groups in [begin, end); out[(piece*n_l + il)*(end-begin) + k]. */
int coffe_b200_corrfunc_range_pieces(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *mu, int n_mu,
                                     const double *sep, int n_sep, long begin, long end, double *out);
int coffe_b200_average_multipoles_range_pieces(coffe_b200_ctx *ctx, const double *z_min, const double *z_max, int n_bins,
                                               const double *sep, int n_sep, const int *l, int n_l, long begin, long end, double *out);

/* ---- covariance ----------------------------------------------------------------------------------- */
/* the members of coffe_parameters_t that coffe_covariance_init reads (src/covariance.c:477-1103) */
typedef struct coffe_b200_covariance_params {
    const double *pk_k, *pk_p;   /* knots of par->power_spectrum (k [1/Mpc], P [Mpc^3]) */
    int n_pk;
    double k_min, k_max;         /* par->k_min, par->k_max */
    int n_bins;
    const double *z_mean, *deltaz;       /* covariance of multipoles   (average = 0) */
    const double *z_min, *z_max;         /* covariance of redshift-averaged multipoles (average = 1) */
    const double *fsky, *density1, *density2, *pixelsize;   /* per bin */
    const double *sep;
    int n_sep;
    const int *l;
    int n_l;
    int pop[4];                  /* covariance_pop1..4, 1-based */
    int cosmic, mixed, poisson;  /* covariance_cosmic / _mixed / _poisson */
    int window;                  /* covariance_window (bin-averaged Bessel functions) */
    int average;                 /* output_type == COVARIANCE_AVERAGE_MULTIPOLES */
    int integration_method;      /* par->covariance_integration_method: 0 or 1 = standard (k-quadrature), 2 = fftlog (2-D FFTLog,
                                    src/covariance.c:357-470, 710-770; always unbinned, like the reference) */
    int integration_bins;        /* par->covariance_integration_bins (fftlog; 0 = 8000); even, prime factors 2, 3, 5 */
} coffe_b200_covariance_params;

/* out[(((ibin*n_l + il1)*n_sep + ir1)*n_l + il2)*n_sep + ir2]   (order of src/covariance.c:1046-1059);
   density / rsd switches come from coffe_b200_set_options, background and galaxy_bias1..4 from the uploaded
   tables */
int coffe_b200_covariance(coffe_b200_ctx *ctx, const coffe_b200_covariance_params *p, double *out);

/* rows [row_begin, row_end) of every bin's matrix, row = il1*n_sep + ir1 (the unit of multi-GPU sharding of the covariance,
   src/covariance.c:653, 680 loop over the same rows): only out[((ibin*n_rows + row)*n_l + il2)*n_sep + ir2] of those rows is
   written, `out` is the full-size buffer of coffe_b200_covariance (standard integration method only) */
int coffe_b200_covariance_rows(coffe_b200_ctx *ctx, const coffe_b200_covariance_params *p, int row_begin, int row_end, double *out);

/* ---- I^n_l tables ----------------------------------------------------------------------------------
   What coffe_integrals_init builds (src/integrals.c:580-1423): for every requested (n, l) the table of
   r^(n-l) I^n_l(r) (plain I^n_l(r) when n < l) on the knots r = 0, the reference's 42 separations below 1 Mpc
   (src/integrals.c:111-126) and the 2-FAST grid r_i = (k_max/k_min)^(i/N), i < N = bessel_bins (src/twofast.c:
   362-483); with `divergent`, also r^4 I^4_0 and the zero-separation renormalisation (which the reference also
   stores multiplied by r^4, src/integrals.c:365-371, 816-827) on the reference's 297 knots
   (src/integrals.c:76-96, 797-849) and the 201 x 201 grid renormalisation(chi_j, chi_k), chi_j = j chi_max / 200
   (src/integrals.c:851-907).  The driver wraps the arrays into coffe_init_spline / coffe_init_spline2d. */
typedef struct coffe_b200_integrals_params {
    const double *pk_k, *pk_p;   /* knots of par->power_spectrum_norm (k [1/Mpc], P [Mpc^3]) */
    int n_pk;
    int interp_method;           /* par->interp_method: interpolation of P(k) inside the k-integrals (2-FAST always
                                    uses the natural cubic spline, src/twofast.c:399) */
    double k_min, k_max;         /* par->k_min_norm, par->k_max_norm */
    int bessel_bins;             /* par->bessel_bins; even, prime factors 2, 3, 5 */
    const int *n, *l;            /* n_tables (n, l) pairs as stored in coffe_integral_t */
    const int *half;             /* per table: state_n = state_l = COFFE_HALF_INTEGER (may be NULL: all integer) */
    int n_tables;
    int divergent;               /* par->divergent */
    double chi_max;              /* end of the renormalisation grid: chi(max(z_mean + deltaz)), src/integrals.c:858-871 */
} coffe_b200_integrals_params;

/* the 2-D tables I^n_l(z, r) of the same (n, l) list from P(k, z_i) rows on the knots pk_k (what coffe_integrals_init adds per table under
   HAVE_CLASS, src/integrals.c:663-767: one 2-FAST transform + small-r integrals per redshift): pk2d[i*n_pk + j] = P(k_j, z_i);
   tables2d[(t*n_r + j)*n_z + i] <-> table t at (z_i, r_j), the layout of coffe_b200_integral.values2d.  divergent is ignored. */
int coffe_b200_integrals2d(coffe_b200_ctx *ctx, const coffe_b200_integrals_params *p, const double *pk2d, int n_z, double *tables2d);

/* array sizes: knots per table, knots of I^4_0, knots per axis of the renormalisation grid */
int coffe_b200_integrals_sizes(const coffe_b200_integrals_params *p, int *n_r, int *n_r40, int *n_renorm);
/* r[n_r]; tables[t*n_r + i]; r40[n_r40], i40[n_r40]; renorm0[n_r40] (its knots are r40 with r40[0] replaced by 0);
   renorm_chi[n_renorm]; renorm_z[k*n_renorm + j] <-> (chi_j, chi_k).  The divergent outputs may be NULL when
   p->divergent == 0. */
int coffe_b200_integrals(coffe_b200_ctx *ctx, const coffe_b200_integrals_params *p, double *r, double *tables,
                         double *r40, double *i40, double *renorm0, double *renorm_chi, double *renorm_z);

/* ---- background: growth factor and growth rate --------------------------------------------------------
   D1(z_i) (normalised to 1 at z = 0) and f(z_i) on the background grid z_i = 15 i/(bins - 1) of coffe_background_init
   (src/background.c:427-480, 536-541), by the reference's own scheme: embedded RK2(3) at eps_rel = 1e-6, restarted at a = 0.05 for
   every knot with the step size and the stepper state carried over (coffe_b200/csrc/growth_host.cpp says why the scheme matters).
   Host arithmetic, no GPU needed.  omega_matter = Omega0_cdm + Omega0_baryon. */
int coffe_b200_background_growth(double omega_matter, double w0, double wa, int bins, double *D1, double *f);

/* Tracing: with COFFE_B200_TRACE set in the environment the multipole entry points print the host phases of a call on
   stderr (range checks, uploads, mu-rule construction, launches queued, result fetched).  The mu rules of the last call are
   kept in the context and reused when the separations and the quadrature configuration are unchanged. */

/* device time of the kernels of the last compute call [ms] and their launch count (bench / roofline) */
int coffe_b200_last_timing(const coffe_b200_ctx *ctx, double *kernel_ms, int *n_launches);

/* device time [ms] of the two line-of-sight kernels of the last multipole computation, CUDA events on the launching
   stream: local x integrated (si_items_smem_kernel) and integrated x integrated (di_items_smem_kernel, the dominant
   kernel: roofline numerator of bench.py) */
int coffe_b200_last_kernel_times(coffe_b200_ctx *ctx, double *si_ms, double *di_ms);

/* ---- prepared computations ------------------------------------------------------------------------
   coffe_b200_plan_multipoles uploads the grid and its quadrature items for the (iz, ir) items
   [item_begin, item_end) once; coffe_b200_plan_execute then only launches kernels on inputs that are
   already resident in HBM (what bench.py times as `value`); coffe_b200_plan_fetch copies the result
   back in the layout of coffe_b200_multipoles_range.  coffe_b200_plan_count reports the number of
   (point, mu) work items and of integrand evaluations per piece (roofline accounting). */
typedef struct coffe_b200_plan coffe_b200_plan;
int coffe_b200_plan_multipoles(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                               const int *l, int n_l, long item_begin, long item_end, coffe_b200_plan **out);
int coffe_b200_plan_execute(coffe_b200_plan *plan, double *kernel_ms);
int coffe_b200_plan_fetch(coffe_b200_plan *plan, double *out);
int coffe_b200_plan_count(coffe_b200_plan *plan, double *n_items, double *n_ni_evals, double *n_si_evals, double *n_di_evals);
void coffe_b200_plan_destroy(coffe_b200_plan *plan);

/* ---- multi-GPU -------------------------------------------------------------------------------------
   The reference's drivers use every host core from ONE call (omp parallel for num_threads(par->nthreads), src/multipoles.c:88,
   par->nthreads = omp_get_max_threads(), src/parser.c:737-741).  A group does the same with the GPUs of the box: one host call
   cuts the output grid into contiguous cost-balanced blocks (integrand evaluations counted per item by the library), drives every
   device from its own host thread and stream with the tables replicated, and each block lands by D2H straight in the caller's
   buffer.  No data-path collective; results are bit-identical to the single-device calls whatever the device count. */
typedef struct coffe_b200_group coffe_b200_group;
int coffe_b200_group_create(int n_devices /* -1: every visible device */, coffe_b200_group **out);
void coffe_b200_group_destroy(coffe_b200_group *g);
int coffe_b200_group_size(const coffe_b200_group *g);
coffe_b200_ctx *coffe_b200_group_member(coffe_b200_group *g, int i);      /* the i-th device's context (owned by the group) */
const char *coffe_b200_group_last_error(const coffe_b200_group *g);
int coffe_b200_group_upload_tables(coffe_b200_group *g, const coffe_b200_tables *tables);
int coffe_b200_group_set_options(coffe_b200_group *g, const coffe_b200_options *opt);
int coffe_b200_group_set_quadrature(coffe_b200_group *g, const coffe_b200_quadrature *q);
/* same arguments and output order as the single-device calls */
int coffe_b200_group_multipoles(coffe_b200_group *g, const double *z_mean, int n_z, const double *sep, int n_sep,
                                const int *l, int n_l, double *out);
int coffe_b200_group_corrfunc(coffe_b200_group *g, const double *z_mean, int n_z, const double *mu, int n_mu,
                              const double *sep, int n_sep, double *out);
int coffe_b200_group_average_multipoles(coffe_b200_group *g, const double *z_min, const double *z_max, int n_bins,
                                        const double *sep, int n_sep, const int *l, int n_l, double *out);
int coffe_b200_group_covariance(coffe_b200_group *g, const coffe_b200_covariance_params *p, double *out);
/* the block boundaries the last group call used (n_parts + 1 entries) and the device time of each member [ms] */
int coffe_b200_group_last_blocks(const coffe_b200_group *g, long *bounds, double *member_ms);

/* Estimated cost (integrand evaluations: integrated x integrated + 2 x local x integrated) of every (iz, ir) item of a multipole grid,
   cost[iz*n_sep + ir], counted by the library on the device, and the boundaries of n_parts contiguous blocks of (nearly) equal
   cost (n_parts + 1 entries, deterministic: every process derives the same ones).  coffe_b200_balance_bounds is host arithmetic. */
int coffe_b200_multipoles_cost(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                               const int *l, int n_l, double *cost);
int coffe_b200_balance_bounds(const double *cost, long n_items, int n_parts, long *bounds);

/* Feedback for repeated calls on the same grid (a sweep over cosmologies): given the device time every block took, scale the cost of its
   items by the block's measured time per unit cost (normalised so the total is unchanged), in place; the next coffe_b200_balance_bounds
   then evens out what the evaluation count does not see (per-item and per-outer-node set-up).  Host arithmetic.  The group calls and
   coffe_b200_multipoles_sharded(bounds = NULL) apply it themselves whenever the grid of the previous call repeats. */
int coffe_b200_rebalance_cost(double *cost, long n_items, int n_parts, const long *bounds, const double *part_ms);

/* One process per GPU (the launch torchrun / mpirun give): every rank computes its block, and the blocks are gathered device to
   device with NCCL (libnccl.so.2 is loaded on first use; grouped ncclBroadcast of the variable-size blocks, issued on the compute
   stream right behind the kernels), then copied to the host buffer of every rank that passes one.  The unique id comes from
   rank 0 (coffe_b200_comm_unique_id) and travels to the other ranks by whatever the launcher offers (MPI_Bcast, a file, a
   torch.distributed broadcast): 128 bytes. */
int coffe_b200_comm_unique_id(unsigned char id[128]);
int coffe_b200_comm_init(coffe_b200_ctx *ctx, int rank, int world, const unsigned char id[128]);
void coffe_b200_comm_destroy(coffe_b200_ctx *ctx);
/* bounds: world + 1 item boundaries (NULL: cost-balanced, computed identically on every rank); out: full grid in the order of
   coffe_b200_multipoles, or NULL on ranks that do not need the result; gather_ms: device time of the collective alone */
int coffe_b200_multipoles_sharded(coffe_b200_ctx *ctx, const double *z_mean, int n_z, const double *sep, int n_sep,
                                  const int *l, int n_l, const long *bounds, double *out, double *kernel_ms, double *gather_ms);
/* prepared variant (inputs resident in HBM): execute launches this rank's kernels and the gather, returns their device times */
int coffe_b200_plan_execute_gather(coffe_b200_plan *plan, const long *bounds, long n_items_total, double *kernel_ms, double *gather_ms);
int coffe_b200_plan_fetch_gathered(coffe_b200_plan *plan, double *out /* [n_l][n_items_total], layout of coffe_b200_multipoles_range */);

/* measured FP64 FMA throughput of the device [TFLOP/s] (dependent-chain-free DFMA probe kernel) */
int coffe_b200_fp64_peak(coffe_b200_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* COFFE_B200_H */
