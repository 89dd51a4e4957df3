/*
 * TEST INFRASTRUCTURE ONLY -- parity oracle.
 * Minimal FFTW3-API stand-in (FFTW is an un-vendored dependency of the reference, CMakeLists.txt:60):
 * unnormalised r2c / c2r transforms in 1-D and 2-D, the only plans the reference creates
 * (src/twofast.c:272,453; src/twobessel.c:60,100).  Mixed-radix Cooley-Tukey with an O(n^2) DFT for
 * prime factors; same mathematical definition as FFTW (forward sign -1, backward +1, no scaling).
 */
#ifndef COFFE_B200_ORACLE_FFTW3_SHIM_H
#define COFFE_B200_ORACLE_FFTW3_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(_Complex_I) && defined(complex) && defined(I) && !defined(__cplusplus)
typedef double _Complex fftw_complex;
#else
typedef double fftw_complex[2];
#endif

#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

typedef struct fftw_shim_plan_s *fftw_plan;

void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags);
fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex *in, double *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif
#endif
