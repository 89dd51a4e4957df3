// Host-side flattening of spline knots into per-interval cubic coefficients.
//
// The reference hands every table to GSL (`coffe_init_spline`, src/common.c:524-613; default type
// Akima, src/parser.c:807; bicubic for 2-D, src/parser.c:810) and GSL's coefficient state is opaque, so
// the upload path recomputes the coefficients from the raw (x, y) knots with the published algorithms
// (SURVEY.md Appendix B): Akima 1970 with GSL's non-periodic end conditions, natural cubic spline, and
// bicubic Hermite patches whose node derivatives come from natural cubic splines along rows/columns.
// This is one-off O(n) set-up work on the host; evaluation happens only on the device.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace coffe {
namespace build {

struct Coef { double y, b, c, d; };

enum InterpKind { INTERP_LINEAR = 1, INTERP_POLYNOMIAL = 2, INTERP_CSPLINE = 3, INTERP_CSPLINE_PERIODIC = 4,
                  INTERP_AKIMA = 5, INTERP_AKIMA_PERIODIC = 6, INTERP_STEFFEN = 7 };   /* src/common.h enum coffe_interp1d_type */

/* c[i] = y''(x_i)/2 of the natural cubic spline */
inline std::vector<double> natural_c(const double *x, const double *y, const int n)
{
    std::vector<double> c(n, 0.0);
    if (n < 3) return c;
    const int m = n - 2;
    std::vector<double> diag(m), off(m), g(m);
    for (int i = 0; i < m; ++i) {
        const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
        off[i] = h1;
        diag[i] = 2.0 * (h1 + h0);
        g[i] = 3.0 * ((y[i + 2] - y[i + 1]) / h1 - (y[i + 1] - y[i]) / h0);
    }
    for (int i = 1; i < m; ++i) {
        const double w = off[i - 1] / diag[i - 1];
        diag[i] -= w * off[i - 1];
        g[i] -= w * g[i - 1];
    }
    c[m] = g[m - 1] / diag[m - 1];
    for (int i = m - 2; i >= 0; --i) c[i + 1] = (g[i] - off[i] * c[i + 2]) / diag[i];
    return c;
}

inline std::vector<Coef> cspline_coefs(const double *x, const double *y, const int n)
{
    std::vector<Coef> k(n > 1 ? n - 1 : 0);
    const std::vector<double> c = natural_c(x, y, n);
    for (int i = 0; i + 1 < n; ++i) {
        const double dx = x[i + 1] - x[i], dy = y[i + 1] - y[i];
        k[i].y = y[i];
        k[i].b = dy / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
        k[i].c = c[i];
        k[i].d = (c[i + 1] - c[i]) / (3.0 * dx);
    }
    return k;
}

inline std::vector<Coef> linear_coefs(const double *x, const double *y, const int n)
{
    std::vector<Coef> k(n > 1 ? n - 1 : 0);
    for (int i = 0; i + 1 < n; ++i) k[i] = Coef{y[i], (y[i + 1] - y[i]) / (x[i + 1] - x[i]), 0.0, 0.0};
    return k;
}

inline std::vector<Coef> akima_coefs(const double *x, const double *y, const int n)
{
    std::vector<Coef> k(n - 1);
    std::vector<double> mm(n + 4);
    double *m = mm.data() + 2;
    for (int i = 0; i + 1 < n; ++i) m[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    m[-2] = 3.0 * m[0] - 2.0 * m[1];
    m[-1] = 2.0 * m[0] - m[1];
    m[n - 1] = 2.0 * m[n - 2] - m[n - 3];
    m[n] = 3.0 * m[n - 2] - 2.0 * m[n - 3];
    for (int i = 0; i < n - 1; ++i) {
        const double NE = std::fabs(m[i + 1] - m[i]) + std::fabs(m[i - 1] - m[i - 2]);
        k[i].y = y[i];
        if (NE == 0.0) { k[i].b = m[i]; k[i].c = 0.0; k[i].d = 0.0; continue; }
        const double h = x[i + 1] - x[i];
        const double NE_next = std::fabs(m[i + 2] - m[i + 1]) + std::fabs(m[i] - m[i - 1]);
        const double alpha = std::fabs(m[i - 1] - m[i - 2]) / NE;
        double t_next;
        if (NE_next == 0.0) t_next = m[i];
        else {
            const double alpha_next = std::fabs(m[i] - m[i - 1]) / NE_next;
            t_next = (1.0 - alpha_next) * m[i] + alpha_next * m[i + 1];
        }
        k[i].b = (1.0 - alpha) * m[i - 1] + alpha * m[i];
        k[i].c = (3.0 * m[i] - 2.0 * k[i].b - t_next) / h;
        k[i].d = (k[i].b + t_next - 2.0 * m[i]) / (h * h);
    }
    return k;
}

/* returns false for interpolation kinds the device path does not implement */
inline bool spline_coefs(const int kind, const double *x, const double *y, const int n, std::vector<Coef> &out)
{
    if (kind == INTERP_AKIMA && n >= 5) { out = akima_coefs(x, y, n); return true; }
    if (kind == INTERP_CSPLINE && n >= 3) { out = cspline_coefs(x, y, n); return true; }
    if (kind == INTERP_LINEAR && n >= 2) { out = linear_coefs(x, y, n); return true; }
    return false;
}

/* derivative of the natural cubic spline through (x, y) at every knot */
inline void natural_node_derivs(const double *x, const double *y, const int n, double *out)
{
    const std::vector<double> c = natural_c(x, y, n);
    for (int i = 0; i + 1 < n; ++i) {
        const double dx = x[i + 1] - x[i];
        out[i] = (y[i + 1] - y[i]) / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    }
    const int i = n - 2;
    const double dx = x[i + 1] - x[i];
    const double b = (y[i + 1] - y[i]) / dx - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    const double d = (c[i + 1] - c[i]) / (3.0 * dx);
    out[n - 1] = b + dx * (2.0 * c[i] + 3.0 * dx * d);
}

/* bicubic node data {z, zx, zy, zxy}; z is stored z[j*nx + i] (src/integrals.c:889) */
inline std::vector<Coef> bicubic_nodes(const double *xa, const double *ya, const double *za, const int nx, const int ny)
{
    std::vector<Coef> node((size_t)nx * ny);
    std::vector<double> zx((size_t)nx * ny), zy((size_t)nx * ny), zxy((size_t)nx * ny);
    std::vector<double> tx(nx > ny ? nx : ny), td(nx > ny ? nx : ny);
    for (int j = 0; j < ny; ++j) {
        for (int i = 0; i < nx; ++i) tx[i] = za[(size_t)j * nx + i];
        natural_node_derivs(xa, tx.data(), nx, td.data());
        for (int i = 0; i < nx; ++i) zx[(size_t)j * nx + i] = td[i];
    }
    for (int i = 0; i < nx; ++i) {
        for (int j = 0; j < ny; ++j) tx[j] = za[(size_t)j * nx + i];
        natural_node_derivs(ya, tx.data(), ny, td.data());
        for (int j = 0; j < ny; ++j) zy[(size_t)j * nx + i] = td[j];
    }
    for (int j = 0; j < ny; ++j) {
        for (int i = 0; i < nx; ++i) tx[i] = zy[(size_t)j * nx + i];
        natural_node_derivs(xa, tx.data(), nx, td.data());
        for (int i = 0; i < nx; ++i) zxy[(size_t)j * nx + i] = td[i];
    }
    for (size_t q = 0; q < node.size(); ++q) node[q] = Coef{za[q], zx[q], zy[q], zxy[q]};
    return node;
}

/* Gauss-Legendre nodes/weights on [0, 1] */
inline void gauss_legendre_01(const int n, std::vector<double> &x, std::vector<double> &w)
{
    x.assign(n, 0.0); w.assign(n, 0.0);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < (n + 1) / 2; ++i) {
        double z = std::cos(pi * (i + 0.75) / (n + 0.5)), pp = 0.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            const double z1 = z;
            z = z1 - p1 / pp;
            if (std::fabs(z - z1) < 1e-16) break;
        }
        x[i] = 0.5 * (1.0 - z); x[n - 1 - i] = 0.5 * (1.0 + z);
        w[i] = w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
    }
}

}  // namespace build
}  // namespace coffe
