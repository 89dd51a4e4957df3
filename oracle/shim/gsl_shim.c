/*
 * TEST INFRASTRUCTURE ONLY -- parity oracle; never linked into libcoffe_b200.so.
 *
 * Implementation of the GSL subset declared in gsl/gsl_shim_all.h.  GSL is absent from this image and
 * from /root/reference (it is an un-vendored, un-pinned ">= 2.1" dependency, reference CMakeLists.txt:59),
 * so each routine restates the published algorithm that the GSL routine of the same name implements:
 *   - Akima (1970) spline with the non-periodic end slopes m[-1]=2m0-m1, m[-2]=3m0-2m1
 *   - natural cubic spline (tridiagonal solve for the second derivatives)
 *   - bicubic Hermite patch with node derivatives estimated by 1-D natural cubic splines
 *   - QUADPACK QAG (bisect worst panel; 61-point Gauss-Kronrod; QUADPACK error heuristic)
 *   - embedded RK2(3) stepper + standard step-size control + evolve loop (odeiv2 semantics)
 *   - elementary special functions (spherical Bessel, Legendre, Si, 3j with m=0)
 * Reference call sites: src/common.c:524-719,1004-1040; src/background.c:427-480;
 * src/integrals.c:299-310,397,414; src/covariance.c:187-238; src/common.c:1389-1390.
 */
#include <stdio.h>
#include <string.h>
#include <float.h>
#include "gsl/gsl_shim_all.h"
#include "gk61_table.h"

/* ------------------------------------------------------------------ errno / math */
gsl_error_handler_t *gsl_set_error_handler_off(void) { return NULL; }
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h) { (void)h; return NULL; }
int gsl_finite(const double x) { return isfinite(x) ? 1 : 0; }

/* ------------------------------------------------------------------ 1-D interpolation */
enum { K_LINEAR = 1, K_POLY, K_CSPLINE, K_CSPLINE_P, K_AKIMA, K_AKIMA_P, K_STEFFEN };
static const gsl_interp_type t_linear = {"linear", 2, K_LINEAR};
static const gsl_interp_type t_poly = {"polynomial", 3, K_POLY};
static const gsl_interp_type t_cspline = {"cspline", 3, K_CSPLINE};
static const gsl_interp_type t_cspline_p = {"cspline-periodic", 2, K_CSPLINE_P};
static const gsl_interp_type t_akima = {"akima", 5, K_AKIMA};
static const gsl_interp_type t_akima_p = {"akima-periodic", 5, K_AKIMA_P};
static const gsl_interp_type t_steffen = {"steffen", 3, K_STEFFEN};
const gsl_interp_type *gsl_interp_linear = &t_linear;
const gsl_interp_type *gsl_interp_polynomial = &t_poly;
const gsl_interp_type *gsl_interp_cspline = &t_cspline;
const gsl_interp_type *gsl_interp_cspline_periodic = &t_cspline_p;
const gsl_interp_type *gsl_interp_akima = &t_akima;
const gsl_interp_type *gsl_interp_akima_periodic = &t_akima_p;
const gsl_interp_type *gsl_interp_steffen = &t_steffen;

gsl_interp_accel *gsl_interp_accel_alloc(void)
{
    gsl_interp_accel *a = (gsl_interp_accel *)malloc(sizeof(gsl_interp_accel));
    a->cache = 0;
    return a;
}
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

/* index i with xa[i] <= x < xa[i+1], clamped to [0, n-2]; stateless so concurrent use is safe */
static size_t shim_bsearch(const double *xa, size_t n, double x)
{
    size_t lo = 0, hi = n - 1;
    while (hi > lo + 1) {
        size_t mid = (hi + lo) / 2;
        if (xa[mid] > x) hi = mid; else lo = mid;
    }
    return lo;
}

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size)
{
    gsl_spline *s = (gsl_spline *)malloc(sizeof(gsl_spline));
    s->interp = (gsl_interp *)malloc(sizeof(gsl_interp));
    s->interp->type = T;
    s->interp->size = size;
    s->interp->b = (double *)calloc(size, sizeof(double));
    s->interp->c = (double *)calloc(size, sizeof(double));
    s->interp->d = (double *)calloc(size, sizeof(double));
    s->x = (double *)malloc(sizeof(double) * size);
    s->y = (double *)malloc(sizeof(double) * size);
    s->size = size;
    return s;
}

void gsl_spline_free(gsl_spline *s)
{
    if (!s) return;
    free(s->interp->b); free(s->interp->c); free(s->interp->d);
    free(s->interp);
    free(s->x); free(s->y);
    free(s);
}

/* natural cubic spline: returns c[i] = y''(x_i)/2 for i = 0..n-1 */
static void natural_cspline_c(const double *xa, const double *ya, size_t n, double *c)
{
    c[0] = 0.0; c[n - 1] = 0.0;
    if (n < 3) return;
    const size_t m = n - 2; /* unknowns c[1..n-2] */
    double *diag = (double *)malloc(sizeof(double) * m);
    double *off = (double *)malloc(sizeof(double) * m);
    double *g = (double *)malloc(sizeof(double) * m);
    for (size_t i = 0; i < m; ++i) {
        const double h_i = xa[i + 1] - xa[i];
        const double h_ip1 = xa[i + 2] - xa[i + 1];
        const double yd_i = ya[i + 1] - ya[i];
        const double yd_ip1 = ya[i + 2] - ya[i + 1];
        off[i] = h_ip1;
        diag[i] = 2.0 * (h_ip1 + h_i);
        g[i] = 3.0 * (yd_ip1 / h_ip1 - yd_i / h_i);
    }
    /* symmetric tridiagonal solve (Thomas) */
    for (size_t i = 1; i < m; ++i) {
        const double w = off[i - 1] / diag[i - 1];
        diag[i] -= w * off[i - 1];
        g[i] -= w * g[i - 1];
    }
    c[m] = g[m - 1] / diag[m - 1];
    for (size_t i = m - 1; i-- > 0;) c[i + 1] = (g[i] - off[i] * c[i + 2]) / diag[i];
    free(diag); free(off); free(g);
}

int gsl_spline_init(gsl_spline *s, const double xa[], const double ya[], size_t size)
{
    if (size != s->size) { fprintf(stderr, "gsl shim: spline size mismatch\n"); abort(); }
    memcpy(s->x, xa, sizeof(double) * size);
    memcpy(s->y, ya, sizeof(double) * size);
    gsl_interp *in = s->interp;
    in->xmin = xa[0];
    in->xmax = xa[size - 1];
    double *b = in->b, *c = in->c, *d = in->d;
    const int kind = in->type->kind;
    if (kind == K_LINEAR) {
        for (size_t i = 0; i + 1 < size; ++i) {
            b[i] = (ya[i + 1] - ya[i]) / (xa[i + 1] - xa[i]);
            c[i] = 0; d[i] = 0;
        }
    }
    else if (kind == K_CSPLINE) {
        double *cc = (double *)malloc(sizeof(double) * size);
        natural_cspline_c(xa, ya, size, cc);
        for (size_t i = 0; i + 1 < size; ++i) {
            const double dx = xa[i + 1] - xa[i];
            const double dy = ya[i + 1] - ya[i];
            b[i] = (dy / dx) - dx * (cc[i + 1] + 2.0 * cc[i]) / 3.0;
            c[i] = cc[i];
            d[i] = (cc[i + 1] - cc[i]) / (3.0 * dx);
        }
        free(cc);
    }
    else if (kind == K_AKIMA) {
        double *mm = (double *)malloc(sizeof(double) * (size + 4));
        double *m = mm + 2;
        for (size_t i = 0; i + 1 < size; ++i) m[i] = (ya[i + 1] - ya[i]) / (xa[i + 1] - xa[i]);
        m[-2] = 3.0 * m[0] - 2.0 * m[1];
        m[-1] = 2.0 * m[0] - m[1];
        m[size - 1] = 2.0 * m[size - 2] - m[size - 3];
        m[size] = 3.0 * m[size - 2] - 2.0 * m[size - 3];
        for (long i = 0; i < (long)size - 1; ++i) {
            const double NE = fabs(m[i + 1] - m[i]) + fabs(m[i - 1] - m[i - 2]);
            if (NE == 0.0) {
                b[i] = m[i]; c[i] = 0.0; d[i] = 0.0;
            }
            else {
                const double h_i = xa[i + 1] - xa[i];
                const double NE_next = fabs(m[i + 2] - m[i + 1]) + fabs(m[i] - m[i - 1]);
                const double alpha_i = fabs(m[i - 1] - m[i - 2]) / NE;
                double tL_ip1;
                if (NE_next == 0.0) tL_ip1 = m[i];
                else {
                    const double alpha_ip1 = fabs(m[i] - m[i - 1]) / NE_next;
                    tL_ip1 = (1.0 - alpha_ip1) * m[i] + alpha_ip1 * m[i + 1];
                }
                b[i] = (1.0 - alpha_i) * m[i - 1] + alpha_i * m[i];
                c[i] = (3.0 * m[i] - 2.0 * b[i] - tL_ip1) / h_i;
                d[i] = (b[i] + tL_ip1 - 2.0 * m[i]) / (h_i * h_i);
            }
        }
        free(mm);
    }
    else {
        fprintf(stderr, "gsl shim: interpolation type '%s' is not implemented in the oracle shim\n", in->type->name);
        abort();
    }
    return GSL_SUCCESS;
}

double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a)
{
    (void)a;
    const gsl_interp *in = s->interp;
    if (!(x >= in->xmin && x <= in->xmax)) return GSL_NAN;
    const size_t i = shim_bsearch(s->x, s->size, x);
    const double dx = x - s->x[i];
    return s->y[i] + dx * (in->b[i] + dx * (in->c[i] + dx * in->d[i]));
}

double gsl_spline_eval_deriv(const gsl_spline *s, double x, gsl_interp_accel *a)
{
    (void)a;
    const gsl_interp *in = s->interp;
    if (!(x >= in->xmin && x <= in->xmax)) return GSL_NAN;
    const size_t i = shim_bsearch(s->x, s->size, x);
    const double dx = x - s->x[i];
    return in->b[i] + dx * (2.0 * in->c[i] + 3.0 * dx * in->d[i]);
}

/* ------------------------------------------------------------------ 2-D interpolation */
enum { K2_BILINEAR = 1, K2_BICUBIC };
static const gsl_interp2d_type t2_bilinear = {"bilinear", 2, K2_BILINEAR};
static const gsl_interp2d_type t2_bicubic = {"bicubic", 4, K2_BICUBIC};
const gsl_interp2d_type *gsl_interp2d_bilinear = &t2_bilinear;
const gsl_interp2d_type *gsl_interp2d_bicubic = &t2_bicubic;

gsl_spline2d *gsl_spline2d_alloc(const gsl_interp2d_type *T, size_t xsize, size_t ysize)
{
    gsl_spline2d *s = (gsl_spline2d *)calloc(1, sizeof(gsl_spline2d));
    s->type = T; s->xsize = xsize; s->ysize = ysize;
    s->xarr = (double *)malloc(sizeof(double) * xsize);
    s->yarr = (double *)malloc(sizeof(double) * ysize);
    s->zarr = (double *)malloc(sizeof(double) * xsize * ysize);
    if (T->kind == K2_BICUBIC) {
        s->zx = (double *)malloc(sizeof(double) * xsize * ysize);
        s->zy = (double *)malloc(sizeof(double) * xsize * ysize);
        s->zxy = (double *)malloc(sizeof(double) * xsize * ysize);
    }
    return s;
}

void gsl_spline2d_free(gsl_spline2d *s)
{
    if (!s) return;
    free(s->xarr); free(s->yarr); free(s->zarr); free(s->zx); free(s->zy); free(s->zxy);
    free(s);
}

/* derivative at the knots of the natural cubic spline through (xa, ya) */
static void natural_spline_node_derivs(const double *xa, const double *ya, size_t n, double *out)
{
    double *c = (double *)malloc(sizeof(double) * n);
    natural_cspline_c(xa, ya, n, c);
    for (size_t i = 0; i + 1 < n; ++i) {
        const double dx = xa[i + 1] - xa[i];
        out[i] = (ya[i + 1] - ya[i]) / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    }
    {   /* last knot: derivative of the last interval's cubic at its right end */
        const size_t i = n - 2;
        const double dx = xa[i + 1] - xa[i];
        const double b = (ya[i + 1] - ya[i]) / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
        const double d = (c[i + 1] - c[i]) / (3.0 * dx);
        out[n - 1] = b + dx * (2.0 * c[i] + 3.0 * dx * d);
    }
    free(c);
}

#define IDX2D(i, j, s) ((j) * (s)->xsize + (i))

int gsl_spline2d_init(gsl_spline2d *s, const double xa[], const double ya[], const double za[],
                      size_t xsize, size_t ysize)
{
    memcpy(s->xarr, xa, sizeof(double) * xsize);
    memcpy(s->yarr, ya, sizeof(double) * ysize);
    memcpy(s->zarr, za, sizeof(double) * xsize * ysize);
    if (s->type->kind != K2_BICUBIC) return GSL_SUCCESS;
    double *tx = (double *)malloc(sizeof(double) * (xsize > ysize ? xsize : ysize));
    double *td = (double *)malloc(sizeof(double) * (xsize > ysize ? xsize : ysize));
    /* zx: splines along x at fixed j */
    for (size_t j = 0; j < ysize; ++j) {
        for (size_t i = 0; i < xsize; ++i) tx[i] = za[IDX2D(i, j, s)];
        natural_spline_node_derivs(xa, tx, xsize, td);
        for (size_t i = 0; i < xsize; ++i) s->zx[IDX2D(i, j, s)] = td[i];
    }
    /* zy: splines along y at fixed i */
    for (size_t i = 0; i < xsize; ++i) {
        for (size_t j = 0; j < ysize; ++j) tx[j] = za[IDX2D(i, j, s)];
        natural_spline_node_derivs(ya, tx, ysize, td);
        for (size_t j = 0; j < ysize; ++j) s->zy[IDX2D(i, j, s)] = td[j];
    }
    /* zxy: splines along x through zy */
    for (size_t j = 0; j < ysize; ++j) {
        for (size_t i = 0; i < xsize; ++i) tx[i] = s->zy[IDX2D(i, j, s)];
        natural_spline_node_derivs(xa, tx, xsize, td);
        for (size_t i = 0; i < xsize; ++i) s->zxy[IDX2D(i, j, s)] = td[i];
    }
    free(tx); free(td);
    return GSL_SUCCESS;
}

double gsl_spline2d_eval(const gsl_spline2d *s, const double x, const double y,
                         gsl_interp_accel *xacc, gsl_interp_accel *yacc)
{
    (void)xacc; (void)yacc;
    if (!(x >= s->xarr[0] && x <= s->xarr[s->xsize - 1])) return GSL_NAN;
    if (!(y >= s->yarr[0] && y <= s->yarr[s->ysize - 1])) return GSL_NAN;
    const size_t xi = shim_bsearch(s->xarr, s->xsize, x);
    const size_t yi = shim_bsearch(s->yarr, s->ysize, y);
    const double xmin = s->xarr[xi], xmax = s->xarr[xi + 1];
    const double ymin = s->yarr[yi], ymax = s->yarr[yi + 1];
    const double z00 = s->zarr[IDX2D(xi, yi, s)], z01 = s->zarr[IDX2D(xi, yi + 1, s)];
    const double z10 = s->zarr[IDX2D(xi + 1, yi, s)], z11 = s->zarr[IDX2D(xi + 1, yi + 1, s)];
    const double dx = xmax - xmin, dy = ymax - ymin;
    const double t = (x - xmin) / dx, u = (y - ymin) / dy;
    if (s->type->kind == K2_BILINEAR)
        return (1 - t) * (1 - u) * z00 + t * (1 - u) * z10 + (1 - t) * u * z01 + t * u * z11;
    /* Hermite data scaled to the unit square */
    const double zx00 = s->zx[IDX2D(xi, yi, s)] * dx, zx01 = s->zx[IDX2D(xi, yi + 1, s)] * dx;
    const double zx10 = s->zx[IDX2D(xi + 1, yi, s)] * dx, zx11 = s->zx[IDX2D(xi + 1, yi + 1, s)] * dx;
    const double zy00 = s->zy[IDX2D(xi, yi, s)] * dy, zy01 = s->zy[IDX2D(xi, yi + 1, s)] * dy;
    const double zy10 = s->zy[IDX2D(xi + 1, yi, s)] * dy, zy11 = s->zy[IDX2D(xi + 1, yi + 1, s)] * dy;
    const double zxy00 = s->zxy[IDX2D(xi, yi, s)] * dx * dy, zxy01 = s->zxy[IDX2D(xi, yi + 1, s)] * dx * dy;
    const double zxy10 = s->zxy[IDX2D(xi + 1, yi, s)] * dx * dy, zxy11 = s->zxy[IDX2D(xi + 1, yi + 1, s)] * dx * dy;
    /* cubic Hermite basis */
    const double t2 = t * t, t3 = t2 * t, u2 = u * u, u3 = u2 * u;
    const double h00t = 2 * t3 - 3 * t2 + 1, h10t = t3 - 2 * t2 + t, h01t = -2 * t3 + 3 * t2, h11t = t3 - t2;
    const double h00u = 2 * u3 - 3 * u2 + 1, h10u = u3 - 2 * u2 + u, h01u = -2 * u3 + 3 * u2, h11u = u3 - u2;
    return
        h00t * (h00u * z00 + h01u * z01 + h10u * zy00 + h11u * zy01)
      + h01t * (h00u * z10 + h01u * z11 + h10u * zy10 + h11u * zy11)
      + h10t * (h00u * zx00 + h01u * zx01 + h10u * zxy00 + h11u * zxy01)
      + h11t * (h00u * zx10 + h01u * zx11 + h10u * zxy10 + h11u * zxy11);
}

/* ------------------------------------------------------------------ QAG (QUADPACK) */
gsl_integration_workspace *gsl_integration_workspace_alloc(const size_t n)
{
    gsl_integration_workspace *w = (gsl_integration_workspace *)malloc(sizeof(*w));
    w->limit = n;
    w->size = 0;
    /* grown on demand: the reference allocates 50000 panels per call, almost never used */
    w->alist = w->blist = w->rlist = w->elist = NULL;
    w->heap = NULL;
    return w;
}

void gsl_integration_workspace_free(gsl_integration_workspace *w)
{
    if (!w) return;
    free(w->alist); free(w->blist); free(w->rlist); free(w->elist); free(w->heap);
    free(w);
}

static double rescale_error(double err, const double result_abs, const double result_asc)
{
    err = fabs(err);
    if (result_asc != 0 && err != 0) {
        const double scale = pow((200 * err / result_asc), 1.5);
        if (scale < 1) err = result_asc * scale; else err = result_asc;
    }
    if (result_abs > GSL_DBL_MIN / (50 * GSL_DBL_EPSILON)) {
        const double min_err = 50 * GSL_DBL_EPSILON * result_abs;
        if (min_err > err) err = min_err;
    }
    return err;
}

static void qk61(const gsl_function *f, double a, double b, double *result, double *abserr,
                 double *resabs, double *resasc)
{
    const int n = 31;
    double fv1[31], fv2[31];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
    const double abs_half_length = fabs(half_length);
    const double f_center = f->function(center, f->params);
    double result_gauss = 0;
    double result_kronrod = f_center * gk61_wgk[n - 1];
    double result_abs = fabs(result_kronrod);
    double result_asc = 0;
    for (int j = 0; j < (n - 1) / 2; j++) {
        const int jtw = j * 2 + 1;
        const double abscissa = half_length * gk61_xgk[jtw];
        const double fval1 = f->function(center - abscissa, f->params);
        const double fval2 = f->function(center + abscissa, f->params);
        const double fsum = fval1 + fval2;
        fv1[jtw] = fval1; fv2[jtw] = fval2;
        result_gauss += gk61_wg[j] * fsum;
        result_kronrod += gk61_wgk[jtw] * fsum;
        result_abs += gk61_wgk[jtw] * (fabs(fval1) + fabs(fval2));
    }
    for (int j = 0; j < n / 2; j++) {
        const int jtwm1 = j * 2;
        const double abscissa = half_length * gk61_xgk[jtwm1];
        const double fval1 = f->function(center - abscissa, f->params);
        const double fval2 = f->function(center + abscissa, f->params);
        fv1[jtwm1] = fval1; fv2[jtwm1] = fval2;
        result_kronrod += gk61_wgk[jtwm1] * (fval1 + fval2);
        result_abs += gk61_wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
    }
    const double mean = result_kronrod * 0.5;
    result_asc = gk61_wgk[n - 1] * fabs(f_center - mean);
    for (int j = 0; j < n - 1; j++)
        result_asc += gk61_wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
    const double err = (result_kronrod - result_gauss) * half_length;
    result_kronrod *= half_length;
    result_abs *= abs_half_length;
    result_asc *= abs_half_length;
    *result = result_kronrod;
    *resabs = result_abs;
    *resasc = result_asc;
    *abserr = rescale_error(err, result_abs, result_asc);
}

static void ws_grow(gsl_integration_workspace *w, size_t *cap, size_t need)
{
    if (need <= *cap) return;
    size_t nc = *cap ? *cap * 2 : 64;
    while (nc < need) nc *= 2;
    if (nc > w->limit) nc = w->limit;
    w->alist = (double *)realloc(w->alist, sizeof(double) * nc);
    w->blist = (double *)realloc(w->blist, sizeof(double) * nc);
    w->rlist = (double *)realloc(w->rlist, sizeof(double) * nc);
    w->elist = (double *)realloc(w->elist, sizeof(double) * nc);
    w->heap = (size_t *)realloc(w->heap, sizeof(size_t) * nc);
    *cap = nc;
}

static void heap_sift_up(gsl_integration_workspace *w, size_t pos)
{
    size_t *h = w->heap;
    while (pos > 0) {
        size_t parent = (pos - 1) / 2;
        if (w->elist[h[parent]] >= w->elist[h[pos]]) break;
        size_t t = h[parent]; h[parent] = h[pos]; h[pos] = t;
        pos = parent;
    }
}

static void heap_sift_down(gsl_integration_workspace *w, size_t pos, size_t n)
{
    size_t *h = w->heap;
    for (;;) {
        size_t l = 2 * pos + 1, r = l + 1, m = pos;
        if (l < n && w->elist[h[l]] > w->elist[h[m]]) m = l;
        if (r < n && w->elist[h[r]] > w->elist[h[m]]) m = r;
        if (m == pos) break;
        size_t t = h[m]; h[m] = h[pos]; h[pos] = t;
        pos = m;
    }
}

static int subinterval_too_small(double a1, double a2, double b2)
{
    const double e = GSL_DBL_EPSILON;
    const double u = GSL_DBL_MIN;
    const double tmp = (1 + 100 * e) * (fabs(a2) + 1000 * u);
    return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w,
                        double *result, double *abserr)
{
    if (key != GSL_INTEG_GAUSS61) { fprintf(stderr, "gsl shim: only GSL_INTEG_GAUSS61 is implemented\n"); abort(); }
    if (limit > w->limit) limit = w->limit;
    size_t cap = 0;
    free(w->alist); free(w->blist); free(w->rlist); free(w->elist); free(w->heap);
    w->alist = w->blist = w->rlist = w->elist = NULL; w->heap = NULL;
    ws_grow(w, &cap, 2);

    double result0, abserr0, resabs0, resasc0;
    qk61(f, a, b, &result0, &abserr0, &resabs0, &resasc0);
    w->size = 1;
    w->alist[0] = a; w->blist[0] = b; w->rlist[0] = result0; w->elist[0] = abserr0;
    w->heap[0] = 0;

    double tolerance = fmax(epsabs, epsrel * fabs(result0));
    const double round_off = 50 * GSL_DBL_EPSILON * resabs0;
    *result = result0; *abserr = abserr0;
    if (abserr0 <= round_off && abserr0 > tolerance) return GSL_FAILURE;
    if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) return GSL_SUCCESS;
    if (limit == 1) return GSL_FAILURE;

    double area = result0, errsum = abserr0;
    size_t iteration = 1;
    int roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0;
    do {
        const size_t imax = w->heap[0];
        const double a_i = w->alist[imax], b_i = w->blist[imax];
        const double r_i = w->rlist[imax], e_i = w->elist[imax];
        const double a1 = a_i, b1 = 0.5 * (a_i + b_i), a2 = b1, b2 = b_i;
        double area1, area2, error1, error2, resabs1, resabs2, resasc1, resasc2;
        qk61(f, a1, b1, &area1, &error1, &resabs1, &resasc1);
        qk61(f, a2, b2, &area2, &error2, &resabs2, &resasc2);
        const double area12 = area1 + area2;
        const double error12 = error1 + error2;
        errsum += (error12 - e_i);
        area += area12 - r_i;
        if (resasc1 != error1 && resasc2 != error2) {
            const double delta = r_i - area12;
            if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
            if (iteration >= 10 && error12 > e_i) roundoff_type2++;
        }
        tolerance = fmax(epsabs, epsrel * fabs(area));
        if (errsum > tolerance) {
            if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
            if (subinterval_too_small(a1, a2, b2)) error_type = 3;
        }
        /* replace the worst panel by its left half, append the right half */
        ws_grow(w, &cap, w->size + 1);
        const size_t inew = w->size;
        w->alist[imax] = a1; w->blist[imax] = b1; w->rlist[imax] = area1; w->elist[imax] = error1;
        w->alist[inew] = a2; w->blist[inew] = b2; w->rlist[inew] = area2; w->elist[inew] = error2;
        heap_sift_down(w, 0, w->size);
        w->heap[inew] = inew;
        w->size++;
        heap_sift_up(w, inew);
        iteration++;
    } while (iteration < limit && !error_type && errsum > tolerance);

    double sum = 0;
    for (size_t k = 0; k < w->size; ++k) sum += w->rlist[k];
    *result = sum;
    *abserr = errsum;
    return (errsum <= tolerance) ? GSL_SUCCESS : GSL_FAILURE;
}

/* ------------------------------------------------------------------ special functions */
double gsl_sf_bessel_j0(const double x)
{
    const double ax = fabs(x);
    if (ax < 0.5) {
        const double y = x * x;
        /* sin(x)/x Taylor series to x^12 */
        return 1.0 + y * (-1.0 / 6.0 + y * (1.0 / 120.0 + y * (-1.0 / 5040.0 + y * (1.0 / 362880.0
             + y * (-1.0 / 39916800.0 + y * (1.0 / 6227020800.0))))));
    }
    return sin(x) / x;
}

/* ascending series j_l(x) = x^l/(2l+1)!! * sum_k (-x^2/2)^k / (k! (2l+3)(2l+5)...(2l+2k+1)) */
static double sph_bessel_series(int l, double x)
{
    double pre = 1.0;
    for (int k = 1; k <= l; ++k) pre *= x / (2.0 * k + 1.0);
    const double q = -0.5 * x * x;
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 60; ++k) {
        term *= q / (k * (2.0 * l + 2.0 * k + 1.0));
        sum += term;
        if (fabs(term) < 1e-18 * fabs(sum)) break;
    }
    return pre * sum;
}

double gsl_sf_bessel_j1(const double x)
{
    if (fabs(x) < 1.0) return sph_bessel_series(1, x);
    return (sin(x) / x - cos(x)) / x;
}

double gsl_sf_bessel_jl(const int l, const double x)
{
    if (l < 0 || x < 0.0) return GSL_NAN;
    if (x == 0.0) return (l == 0) ? 1.0 : 0.0;
    if (l == 0) return gsl_sf_bessel_j0(x);
    if (l == 1) return gsl_sf_bessel_j1(x);
    /* series where upward recurrence would lose digits */
    if (x < l + 1.0 || x < 2.0) {
        if (x * x < 10.0 * (l + 0.5)) return sph_bessel_series(l, x);
        /* downward recurrence (Miller) normalised with j0 */
        const int lstart = l + 30 + (int)x;
        double jp1 = 0.0, j = 1e-300, jl = 0.0;
        for (int k = lstart; k > 0; --k) {
            const double jm1 = (2.0 * k + 1.0) / x * j - jp1;
            jp1 = j; j = jm1;
            if (k - 1 == l) jl = j;
            if (fabs(j) > 1e250) { j *= 1e-250; jp1 *= 1e-250; jl *= 1e-250; }
        }
        return jl * (gsl_sf_bessel_j0(x) / j);
    }
    double jm = gsl_sf_bessel_j0(x), j = gsl_sf_bessel_j1(x);
    for (int k = 1; k < l; ++k) {
        const double jn = (2.0 * k + 1.0) / x * j - jm;
        jm = j; j = jn;
    }
    return j;
}

double gsl_sf_bessel_Jn(const int n, const double x) { return jn(n, x); }

double gsl_sf_legendre_P1(double x) { return x; }
double gsl_sf_legendre_P2(double x) { return 0.5 * (3.0 * x * x - 1.0); }
double gsl_sf_legendre_Pl(const int l, const double x)
{
    if (l == 0) return 1.0;
    if (l == 1) return x;
    double p0 = 1.0, p1 = x;
    for (int k = 1; k < l; ++k) {
        const double p2 = ((2.0 * k + 1.0) * x * p1 - k * p0) / (k + 1.0);
        p0 = p1; p1 = p2;
    }
    return p1;
}

double gsl_sf_gamma(const double x) { return tgamma(x); }
double gsl_sf_fact(const unsigned int n)
{
    double r = 1.0;
    for (unsigned int k = 2; k <= n; ++k) r *= k;
    return r;
}
double gsl_sf_choose(unsigned int n, unsigned int m)
{
    if (m > n) return 0.0;
    if (m > n - m) m = n - m;
    double r = 1.0;
    for (unsigned int k = 1; k <= m; ++k) r = r * (n - m + k) / k;
    return floor(r + 0.5);
}

/* sine integral: power series for |x| <= 2, Lentz continued fraction of E1(ix) beyond */
double gsl_sf_Si(const double x)
{
    const double ax = fabs(x);
    double si;
    if (ax <= 2.0) {
        double term = ax, sum = ax;
        for (int k = 1; k < 40; ++k) {
            term *= -ax * ax / ((2.0 * k) * (2.0 * k + 1.0));
            const double add = term / (2.0 * k + 1.0);
            sum += add;
            if (fabs(add) < 1e-18 * fabs(sum)) break;
        }
        si = sum;
    }
    else {
        const double tiny = 1e-300;
        double br = 1.0, bi = ax;               /* b = 1 + i x */
        double cr = 1.0 / tiny, ci = 0.0;        /* c */
        double den = br * br + bi * bi;
        double dr = br / den, di = -bi / den;    /* d = 1/b */
        double hr = dr, hi = di;
        for (int i = 1; i < 10000; ++i) {
            const double an = -(double)i * i;
            br += 2.0;
            /* d = 1/(a*d + b) */
            double tr = an * dr + br, ti = an * di + bi;
            den = tr * tr + ti * ti;
            dr = tr / den; di = -ti / den;
            /* c = b + a/c */
            den = cr * cr + ci * ci;
            tr = br + an * cr / den; ti = bi - an * ci / den;
            cr = tr; ci = ti;
            /* del = c*d */
            const double delr = cr * dr - ci * di, deli = cr * di + ci * dr;
            tr = hr * delr - hi * deli; ti = hr * deli + hi * delr;
            hr = tr; hi = ti;
            if (fabs(delr - 1.0) + fabs(deli) < 1e-16) break;
        }
        /* h *= exp(-i x) */
        const double cx = cos(ax), sx = sin(ax);
        const double him = hi * cx - hr * sx;
        si = M_PI_2 + him;
    }
    return (x < 0) ? -si : si;
}

double gsl_sf_coupling_3j(int two_ja, int two_jb, int two_jc, int two_ma, int two_mb, int two_mc)
{
    if (two_ma != 0 || two_mb != 0 || two_mc != 0 || (two_ja | two_jb | two_jc) & 1) {
        fprintf(stderr, "gsl shim: 3j symbol only implemented for integer j and m = 0\n");
        abort();
    }
    const int a = two_ja / 2, b = two_jb / 2, c = two_jc / 2;
    if (a < 0 || b < 0 || c < 0) return GSL_NAN;
    if (c > a + b || c < abs(a - b)) return 0.0;
    const int J = a + b + c;
    if (J % 2) return 0.0;
    const int g = J / 2;
    /* (-1)^g sqrt[(J-2a)!(J-2b)!(J-2c)!/(J+1)!] g!/((g-a)!(g-b)!(g-c)!)  via lgamma */
    const double lnum = 0.5 * (lgamma(J - 2 * a + 1.0) + lgamma(J - 2 * b + 1.0) + lgamma(J - 2 * c + 1.0)
                               - lgamma(J + 2.0))
                        + lgamma(g + 1.0) - lgamma(g - a + 1.0) - lgamma(g - b + 1.0) - lgamma(g - c + 1.0);
    const double v = exp(lnum);
    return (g % 2) ? -v : v;
}

/* ------------------------------------------------------------------ vectors / matrices / least squares */
gsl_vector *gsl_vector_alloc(const size_t n)
{
    gsl_vector *v = (gsl_vector *)malloc(sizeof(*v));
    v->size = n; v->data = (double *)calloc(n, sizeof(double));
    return v;
}
void gsl_vector_free(gsl_vector *v) { if (v) { free(v->data); free(v); } }
void gsl_vector_set(gsl_vector *v, const size_t i, double x) { v->data[i] = x; }
double gsl_vector_get(const gsl_vector *v, const size_t i) { return v->data[i]; }
gsl_matrix *gsl_matrix_alloc(const size_t n1, const size_t n2)
{
    gsl_matrix *m = (gsl_matrix *)malloc(sizeof(*m));
    m->size1 = n1; m->size2 = n2; m->data = (double *)calloc(n1 * n2, sizeof(double)); m->owner = 1;
    return m;
}
void gsl_matrix_free(gsl_matrix *m) { if (m) { if (m->owner) free(m->data); free(m); } }
void gsl_matrix_set(gsl_matrix *m, const size_t i, const size_t j, const double x) { m->data[i * m->size2 + j] = x; }
double gsl_matrix_get(const gsl_matrix *m, const size_t i, const size_t j) { return m->data[i * m->size2 + j]; }
gsl_matrix_view gsl_matrix_view_array(double *base, const size_t n1, const size_t n2)
{
    gsl_matrix_view v;
    v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.data = base; v.matrix.owner = 0;
    return v;
}

gsl_multifit_linear_workspace *gsl_multifit_linear_alloc(const size_t n, const size_t p)
{
    gsl_multifit_linear_workspace *w = (gsl_multifit_linear_workspace *)malloc(sizeof(*w));
    w->n = n; w->p = p;
    return w;
}
void gsl_multifit_linear_free(gsl_multifit_linear_workspace *w) { free(w); }

/* ordinary least squares by Householder QR in long double (GSL uses an SVD; same minimiser) */
int gsl_multifit_linear(const gsl_matrix *X, const gsl_vector *y, gsl_vector *c, gsl_matrix *cov,
                        double *chisq, gsl_multifit_linear_workspace *work)
{
    (void)work;
    const size_t n = X->size1, p = X->size2;
    long double *A = (long double *)malloc(sizeof(long double) * n * p);
    long double *r = (long double *)malloc(sizeof(long double) * n);
    for (size_t i = 0; i < n; ++i) {
        r[i] = y->data[i];
        for (size_t j = 0; j < p; ++j) A[i * p + j] = X->data[i * p + j];
    }
    for (size_t k = 0; k < p; ++k) {
        long double norm = 0;
        for (size_t i = k; i < n; ++i) norm += A[i * p + k] * A[i * p + k];
        norm = sqrtl(norm);
        if (norm == 0) continue;
        const long double alpha = (A[k * p + k] > 0) ? -norm : norm;
        long double *v = (long double *)malloc(sizeof(long double) * n);
        long double vnorm = 0;
        for (size_t i = k; i < n; ++i) { v[i] = A[i * p + k]; if (i == k) v[i] -= alpha; vnorm += v[i] * v[i]; }
        if (vnorm > 0) {
            for (size_t j = k; j < p; ++j) {
                long double dot = 0;
                for (size_t i = k; i < n; ++i) dot += v[i] * A[i * p + j];
                dot = 2 * dot / vnorm;
                for (size_t i = k; i < n; ++i) A[i * p + j] -= dot * v[i];
            }
            long double dot = 0;
            for (size_t i = k; i < n; ++i) dot += v[i] * r[i];
            dot = 2 * dot / vnorm;
            for (size_t i = k; i < n; ++i) r[i] -= dot * v[i];
        }
        free(v);
    }
    for (size_t k = p; k-- > 0;) {
        long double s = r[k];
        for (size_t j = k + 1; j < p; ++j) s -= A[k * p + j] * (long double)c->data[j];
        c->data[k] = (double)(s / A[k * p + k]);
    }
    long double chi = 0;
    for (size_t i = p; i < n; ++i) chi += r[i] * r[i];
    *chisq = (double)chi;
    if (cov) for (size_t i = 0; i < cov->size1 * cov->size2; ++i) cov->data[i] = 0.0;
    free(A); free(r);
    return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ ODE */
static const gsl_odeiv2_step_type rk2_type = {"rk2", 1, 1, 2};
const gsl_odeiv2_step_type *gsl_odeiv2_step_rk2 = &rk2_type;

gsl_odeiv2_step *gsl_odeiv2_step_alloc(const gsl_odeiv2_step_type *T, size_t dim)
{
    gsl_odeiv2_step *s = (gsl_odeiv2_step *)malloc(sizeof(*s));
    s->type = T; s->dimension = dim;
    s->k1 = (double *)calloc(dim, sizeof(double));
    s->k2 = (double *)calloc(dim, sizeof(double));
    s->k3 = (double *)calloc(dim, sizeof(double));
    s->ytmp = (double *)calloc(dim, sizeof(double));
    return s;
}
void gsl_odeiv2_step_free(gsl_odeiv2_step *s) { if (s) { free(s->k1); free(s->k2); free(s->k3); free(s->ytmp); free(s); } }

gsl_odeiv2_control *gsl_odeiv2_control_y_new(double eps_abs, double eps_rel)
{
    gsl_odeiv2_control *c = (gsl_odeiv2_control *)malloc(sizeof(*c));
    c->eps_abs = eps_abs; c->eps_rel = eps_rel; c->a_y = 1.0; c->a_dydt = 0.0;
    return c;
}
void gsl_odeiv2_control_free(gsl_odeiv2_control *c) { free(c); }

gsl_odeiv2_evolve *gsl_odeiv2_evolve_alloc(size_t dim)
{
    gsl_odeiv2_evolve *e = (gsl_odeiv2_evolve *)malloc(sizeof(*e));
    e->dimension = dim;
    e->y0 = (double *)calloc(dim, sizeof(double));
    e->yerr = (double *)calloc(dim, sizeof(double));
    e->dydt_in = (double *)calloc(dim, sizeof(double));
    e->dydt_out = (double *)calloc(dim, sizeof(double));
    e->last_step = 0.0; e->count = 0; e->failed_steps = 0;
    return e;
}
void gsl_odeiv2_evolve_free(gsl_odeiv2_evolve *e) { if (e) { free(e->y0); free(e->yerr); free(e->dydt_in); free(e->dydt_out); free(e); } }

/* embedded RK2(3): k1, k2 (midpoint), k3; 3rd-order solution, error = h*(k2 - (k1+4k2+k3)/6) */
static int rk2_apply(gsl_odeiv2_step *s, double t, double h, double y[], double yerr[],
                     const double dydt_in[], double dydt_out[], const gsl_odeiv2_system *sys)
{
    const size_t dim = s->dimension;
    double *k1 = s->k1, *k2 = s->k2, *k3 = s->k3, *ytmp = s->ytmp;
    if (dydt_in) memcpy(k1, dydt_in, sizeof(double) * dim);
    else if (sys->function(t, y, k1, sys->params)) return GSL_FAILURE;
    for (size_t i = 0; i < dim; i++) ytmp[i] = y[i] + 0.5 * h * k1[i];
    if (sys->function(t + 0.5 * h, ytmp, k2, sys->params)) return GSL_FAILURE;
    for (size_t i = 0; i < dim; i++) ytmp[i] = y[i] + h * (-k1[i] + 2.0 * k2[i]);
    if (sys->function(t + h, ytmp, k3, sys->params)) return GSL_FAILURE;
    for (size_t i = 0; i < dim; i++) {
        ytmp[i] = y[i];
        const double ksum3 = (k1[i] + 4.0 * k2[i] + k3[i]) / 6.0;
        y[i] += h * ksum3;
    }
    if (dydt_out) {
        if (sys->function(t + h, y, dydt_out, sys->params)) {
            memcpy(y, ytmp, sizeof(double) * dim);
            return GSL_FAILURE;
        }
    }
    for (size_t i = 0; i < dim; i++) {
        const double ksum3 = (k1[i] + 4.0 * k2[i] + k3[i]) / 6.0;
        yerr[i] = h * (k2[i] - ksum3);
    }
    return GSL_SUCCESS;
}

enum { HADJ_DEC = -1, HADJ_NIL = 0, HADJ_INC = 1 };

static int std_control_hadjust(const gsl_odeiv2_control *c, size_t dim, unsigned int ord, const double y[],
                               const double yerr[], const double yp[], double *h)
{
    const double S = 0.9;
    const double h_old = *h;
    double rmax = DBL_MIN;
    for (size_t i = 0; i < dim; i++) {
        const double D0 = c->eps_rel * (c->a_y * fabs(y[i]) + c->a_dydt * fabs(h_old * yp[i])) + c->eps_abs;
        const double r = fabs(yerr[i]) / fabs(D0);
        if (r > rmax) rmax = r;
    }
    if (rmax > 1.1) {
        double r = S / pow(rmax, 1.0 / ord);
        if (r < 0.2) r = 0.2;
        *h = r * h_old;
        return HADJ_DEC;
    }
    else if (rmax < 0.5) {
        double r = S / pow(rmax, 1.0 / (ord + 1.0));
        if (r > 5.0) r = 5.0;
        if (r < 1.0) r = 1.0;
        *h = r * h_old;
        return HADJ_INC;
    }
    return HADJ_NIL;
}

int gsl_odeiv2_evolve_apply(gsl_odeiv2_evolve *e, gsl_odeiv2_control *con, gsl_odeiv2_step *step,
                            const gsl_odeiv2_system *dydt, double *t, double t1, double *h, double y[])
{
    const double t0 = *t;
    double h0 = *h;
    int step_status;
    int final_step = 0;
    const double dt = t1 - t0;
    if ((dt < 0.0 && h0 > 0.0) || (dt > 0.0 && h0 < 0.0)) return GSL_FAILURE;
    memcpy(e->y0, y, sizeof(double) * e->dimension);
    if (step->type->can_use_dydt_in) {
        if (e->count == 0) {
            if (dydt->function(t0, y, e->dydt_in, dydt->params)) return GSL_FAILURE;
        }
        else memcpy(e->dydt_in, e->dydt_out, sizeof(double) * e->dimension);
    }
    for (;;) {
        if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) { h0 = dt; final_step = 1; }
        else final_step = 0;
        step_status = rk2_apply(step, t0, h0, y, e->yerr, e->dydt_in, e->dydt_out, dydt);
        if (step_status != GSL_SUCCESS) {
            /* stepper failure: halve the step and retry, as the odeiv2 driver does */
            const double h_old = h0;
            h0 *= 0.5;
            if (fabs(h0) < fabs(h_old) && (t0 + h0) != t0) {
                memcpy(y, e->y0, sizeof(double) * e->dimension);
                e->failed_steps++;
                continue;
            }
            *h = h0; *t = t0;
            return step_status;
        }
        e->count++;
        e->last_step = h0;
        if (final_step) *t = t1; else *t = t0 + h0;
        if (con != NULL) {
            const double h_old = h0;
            const int hadj = std_control_hadjust(con, e->dimension, step->type->order, y, e->yerr, e->dydt_out, &h0);
            if (hadj == HADJ_DEC) {
                const double t_curr = *t;
                const double t_next = (*t) + h0;
                if (fabs(h0) < fabs(h_old) && t_next != t_curr) {
                    memcpy(y, e->y0, sizeof(double) * e->dimension);
                    e->failed_steps++;
                    continue;
                }
                else {
                    *h = h0; *t = t0;
                    return GSL_FAILURE;
                }
            }
        }
        break;
    }
    if (final_step == 0) *h = h0;
    return step_status;
}


/* ---- Monte-Carlo stubs (see gsl_shim_all.h): never on a call path of the oracle (it builds with -DHAVE_CUBA) or of the drop-in */
static void monte_unavailable(void)
{
    fprintf(stderr, "gsl shim: Monte-Carlo integration is not provided (the reference uses it only without CUBA)\n");
    abort();
}
static const gsl_rng_type shim_rng_type = {"shim"};
const gsl_rng_type *gsl_rng_default = &shim_rng_type;
const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) { gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng)); r->type = T; return r; }
void gsl_rng_free(gsl_rng *r) { free(r); }
gsl_monte_plain_state *gsl_monte_plain_alloc(size_t dim) { (void)dim; monte_unavailable(); return NULL; }
gsl_monte_miser_state *gsl_monte_miser_alloc(size_t dim) { (void)dim; monte_unavailable(); return NULL; }
gsl_monte_vegas_state *gsl_monte_vegas_alloc(size_t dim) { (void)dim; monte_unavailable(); return NULL; }
void gsl_monte_plain_free(gsl_monte_plain_state *s) { (void)s; }
void gsl_monte_miser_free(gsl_monte_miser_state *s) { (void)s; }
void gsl_monte_vegas_free(gsl_monte_vegas_state *s) { (void)s; }
int gsl_monte_plain_integrate(const gsl_monte_function *f, const double xl[], const double xu[], size_t dim, size_t calls, gsl_rng *r,
                              gsl_monte_plain_state *state, double *result, double *abserr)
{ (void)f; (void)xl; (void)xu; (void)dim; (void)calls; (void)r; (void)state; (void)result; (void)abserr; monte_unavailable(); return 1; }
int gsl_monte_miser_integrate(gsl_monte_function *f, const double xl[], const double xu[], size_t dim, size_t calls, gsl_rng *r,
                              gsl_monte_miser_state *state, double *result, double *abserr)
{ (void)f; (void)xl; (void)xu; (void)dim; (void)calls; (void)r; (void)state; (void)result; (void)abserr; monte_unavailable(); return 1; }
int gsl_monte_vegas_integrate(gsl_monte_function *f, double xl[], double xu[], size_t dim, size_t calls, gsl_rng *r,
                              gsl_monte_vegas_state *state, double *result, double *abserr)
{ (void)f; (void)xl; (void)xu; (void)dim; (void)calls; (void)r; (void)state; (void)result; (void)abserr; monte_unavailable(); return 1; }
