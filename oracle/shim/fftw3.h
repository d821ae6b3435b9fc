/*
 * oracle/shim/fftw3.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Minimal stand-in for the FFTW 3 API, only so that the UNMODIFIED reference
 * sources under /root/reference/src compile in a container where FFTW cannot
 * be installed.  Declares exactly the entry points the reference calls
 * (src/fftw.cc:42-57, src/sfft.cc:267-295,435-475, execute sites in
 * src/computefourier-1.0-2.0.cc:69,270,273 and src/computefourier-3.0.cc:123,208,293).
 * Backed by oracle/fft_ref.c (radix-2 DIT / Bluestein), not by FFTW.
 */
#ifndef ORC_FFTW3_SHIM_H
#define ORC_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Like the real header: if <complex.h> came first, fftw_complex is the native
 * complex type, otherwise double[2]. */
#if defined(_Complex_I) && defined(complex) && defined(I)
typedef double _Complex fftw_complex;
#else
typedef double fftw_complex[2];
#endif

typedef struct orc_fftw_plan_s *fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out,
                           int sign, unsigned flags);
fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed,
                             int istride, int idist,
                             fftw_complex *out, const int *onembed,
                             int ostride, int odist, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void *fftw_malloc(size_t n);
void fftw_free(void *p);
void fftw_flops(const fftw_plan p, double *add, double *mul, double *fmas);

#ifdef __cplusplus
}
#endif
#endif
