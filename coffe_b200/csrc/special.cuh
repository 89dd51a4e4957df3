// Special functions shared by the table-construction and covariance kernels (device only).
#pragma once

namespace {

/* spherical Bessel function j_l(x), x >= 0, 0 <= l (small: the upward recurrence is used for x > l) */
__device__ double sph_bessel(const int l, const double x)
{
    if (x == 0.0) return l == 0 ? 1.0 : 0.0;
    if (x * x < 10.0 * (l + 0.5) && (l > 0 || x < 0.5)) {
        /* ascending series x^l/(2l+1)!! sum_k (-x^2/2)^k / (k! (2l+3)...(2l+2k+1)) */
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
    double sx, cx;
    sincos(x, &sx, &cx);
    double jm = sx / x;
    if (l == 0) return jm;
    double j = (sx / x - cx) / x;
    if (x > l) {
        for (int k = 1; k < l; ++k) { const double jn = (2.0 * k + 1.0) / x * j - jm; jm = j; j = jn; }
        return j;
    }
    /* x <= l but beyond the series range: downward (Miller) recurrence normalised with j0 */
    const int lstart = l + 30 + (int)x;
    double jp1 = 0.0, jj = 1e-300, jl = 0.0;
    for (int k = lstart; k > 0; --k) {
        const double jm1 = (2.0 * k + 1.0) / x * jj - jp1;
        jp1 = jj; jj = jm1;
        if (k - 1 == l) jl = jj;
        if (fabs(jj) > 1e250) { jj *= 1e-250; jp1 *= 1e-250; jl *= 1e-250; }
    }
    return jl * (jm / jj);
}


}  // namespace
