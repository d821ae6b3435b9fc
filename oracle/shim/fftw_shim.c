/*
 * oracle/shim/fftw_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * FFTW-API stand-in backed by oracle/fft_ref.c; see shim/fftw3.h.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "../fft_ref.h"

typedef struct orc_fftw_plan_s {
  long n;
  int howmany;
  double *in, *out;          /* interleaved re,im */
  long istride, idist, ostride, odist;
  int sign;
  ocplx *tw;                 /* twiddle table when n is a power of two */
} orc_fftw_plan;

typedef orc_fftw_plan *fftw_plan_t;

static fftw_plan_t make(long n, int howmany, void *in, long istride, long idist,
                        void *out, long ostride, long odist, int sign)
{
  fftw_plan_t p = (fftw_plan_t)calloc(1, sizeof(*p));
  p->n = n; p->howmany = howmany;
  p->in = (double *)in; p->out = (double *)out;
  p->istride = istride; p->idist = idist;
  p->ostride = ostride; p->odist = odist;
  p->sign = sign;
  p->tw = orc_is_pow2(n) ? orc_twiddle_table(n) : NULL;
  return p;
}

void *fftw_plan_dft_1d(int n, void *in, void *out, int sign, unsigned flags)
{
  (void)flags;
  return make(n, 1, in, 1, n, out, 1, n, sign);
}

void *fftw_plan_many_dft(int rank, const int *n, int howmany, void *in,
                         const int *inembed, int istride, int idist, void *out,
                         const int *onembed, int ostride, int odist, int sign,
                         unsigned flags)
{
  (void)inembed; (void)onembed; (void)flags;
  if (rank != 1) return NULL;
  return make(n[0], howmany, in, istride, idist, out, ostride, odist, sign);
}

void fftw_execute(void *vp)
{
  fftw_plan_t p = (fftw_plan_t)vp;
  const long n = p->n;
  ocplx *buf = (ocplx *)malloc((size_t)n * sizeof(ocplx));
  for (int b = 0; b < p->howmany; b++) {
    const double *src = p->in + 2 * (long)b * p->idist;
    double *dst = p->out + 2 * (long)b * p->odist;
    for (long i = 0; i < n; i++) {
      buf[i].re = src[2 * i * p->istride];
      buf[i].im = src[2 * i * p->istride + 1];
    }
    if (p->tw) orc_fft_pow2(buf, n, p->sign, p->tw, n);
    else orc_fft_any(buf, n, p->sign);
    for (long i = 0; i < n; i++) {
      dst[2 * i * p->ostride] = buf[i].re;
      dst[2 * i * p->ostride + 1] = buf[i].im;
    }
  }
  free(buf);
}

void fftw_destroy_plan(void *vp)
{
  fftw_plan_t p = (fftw_plan_t)vp;
  if (!p) return;
  free(p->tw);
  free(p);
}

void *fftw_malloc(size_t n)
{
  void *p = NULL;
  if (posix_memalign(&p, 64, n ? n : 64)) return NULL;
  return p;
}

void fftw_free(void *p) { free(p); }

void fftw_flops(void *vp, double *add, double *mul, double *fmas)
{
  fftw_plan_t p = (fftw_plan_t)vp;
  double lg = log2((double)p->n);
  *add = 3.0 * p->n * lg * p->howmany; *mul = 2.0 * p->n * lg * p->howmany; *fmas = 0;
}
