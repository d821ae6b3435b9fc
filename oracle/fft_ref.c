/*
 * oracle/fft_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See fft_ref.h.
 * Stands in for FFTW 3 (absent here; call sites /root/reference/src/fftw.cc:42-57,
 * src/sfft.cc:267-295,435-475).  Build with -ffp-contract=off and without
 * -ffast-math so that every operation is a single IEEE-754 rounding.
 */
#include "fft_ref.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

int orc_is_pow2(long n) { return n > 0 && (n & (n - 1)) == 0; }

void orc_twiddle(long k, long n, double *re, double *im)
{
  /* angle = 2 pi k / n, 0 <= k < n/2 */
  const long n4 = n / 4, n8 = n / 8, n2 = n / 2;
  const double unit = M_PI / (double)(n2 > 0 ? n2 : 1); /* 2 pi / n, exact scaling of M_PI */
  double c, s;
  if (k <= n8) {
    double a = (double)k * unit;
    c = cos(a); s = sin(a);
  } else if (k <= n4) {
    double a = (double)(n4 - k) * unit;
    c = sin(a); s = cos(a);
  } else if (k <= n4 + n8) {
    double a = (double)(k - n4) * unit;
    c = -sin(a); s = cos(a);
  } else {
    double a = (double)(n2 - k) * unit;
    c = -cos(a); s = sin(a);
  }
  *re = c;
  *im = -s;
}

ocplx *orc_twiddle_table(long n)
{
  long h = n / 2 > 0 ? n / 2 : 1;
  ocplx *t = (ocplx *)malloc((size_t)h * sizeof(ocplx));
  if (n <= ORC_TW_DIRECT_MAX) {
    for (long k = 0; k < n / 2; k++)
      orc_twiddle(k, n, &t[k].re, &t[k].im);
  } else {
    /* Large transforms (the plan builder's n-point and Bluestein FFTs): a twiddle is
     * DEFINED as the rounded product of a coarse and a fine factor,
     *     W_n^k = A[k >> 14] * F[k & 16383],  A[m] = W_n^(16384 m),  F[l] = W_n^l,
     * each factor by the octant rule.  Two small tables instead of n/2 libm calls, and a
     * definition the CUDA plan builder can evaluate bit-identically on the fly. */
    const long lo = ORC_TW_FINE, nhi = (n / 2 + lo - 1) / lo;
    ocplx *A = (ocplx *)malloc((size_t)nhi * sizeof(ocplx));
    ocplx *F = (ocplx *)malloc((size_t)lo * sizeof(ocplx));
    for (long m = 0; m < nhi; m++) orc_twiddle(m * lo, n, &A[m].re, &A[m].im);
    for (long l = 0; l < lo; l++) orc_twiddle(l, n, &F[l].re, &F[l].im);
    for (long k = 0; k < n / 2; k++) {
      const ocplx a = A[k / lo], f = F[k % lo];
      const double p0 = a.re * f.re, p1 = a.im * f.im, p2 = a.re * f.im, p3 = a.im * f.re;
      t[k].re = p0 - p1;
      t[k].im = p2 + p3;
    }
    free(A); free(F);
  }
  if (n < 2) { t[0].re = 1.0; t[0].im = -0.0; }
  return t;
}

static void bit_reverse_permute(ocplx *x, long n)
{
  long j = 0;
  for (long i = 0; i < n - 1; i++) {
    if (i < j) { ocplx t = x[i]; x[i] = x[j]; x[j] = t; }
    long m = n >> 1;
    while (m >= 1 && (j & m)) { j ^= m; m >>= 1; }
    j |= m;
  }
}

void orc_fft_pow2(ocplx *x, long n, int sign, const ocplx *tw, long tw_n)
{
  if (n < 2) return;
  bit_reverse_permute(x, n);
  for (long h = 1; h < n; h <<= 1) {
    const long tstep = tw_n / (2 * h);          /* W_{2h}^k = tw[k * tstep] */
    for (long base = 0; base < n; base += 2 * h) {
      for (long k = 0; k < h; k++) {
        const double wr = tw[k * tstep].re;
        const double wi = sign < 0 ? tw[k * tstep].im : -tw[k * tstep].im;
        ocplx *u = &x[base + k], *v = &x[base + k + h];
        const double p0 = wr * v->re, p1 = wi * v->im;
        const double p2 = wr * v->im, p3 = wi * v->re;
        const double tr = p0 - p1, ti = p2 + p3;
        const double ur = u->re, ui = u->im;
        u->re = ur + tr; u->im = ui + ti;
        v->re = ur - tr; v->im = ui - ti;
      }
    }
  }
}

/* Bluestein: X[k] = conj-chirp[k] * sum_j (x[j] chirp[j]) * conj... with
 * chirp[j] = e^{sign * pi i j^2 / n}. */
static void bluestein(ocplx *x, long n, int sign)
{
  long m = 1;
  while (m < 2 * n - 1) m <<= 1;
  ocplx *tw = orc_twiddle_table(m);
  ocplx *a = (ocplx *)calloc((size_t)m, sizeof(ocplx));
  ocplx *b = (ocplx *)calloc((size_t)m, sizeof(ocplx));
  ocplx *ch = (ocplx *)malloc((size_t)n * sizeof(ocplx));
  for (long j = 0; j < n; j++) {
    long long q = ((long long)j * j) % (2 * (long long)n);
    double ang = M_PI * (double)q / (double)n;
    ch[j].re = cos(ang);
    ch[j].im = (sign < 0 ? -1.0 : 1.0) * sin(ang);   /* e^{sign pi i j^2/n} */
  }
  for (long j = 0; j < n; j++) {
    a[j].re = x[j].re * ch[j].re - x[j].im * ch[j].im;
    a[j].im = x[j].re * ch[j].im + x[j].im * ch[j].re;
  }
  b[0].re = ch[0].re; b[0].im = -ch[0].im;
  for (long j = 1; j < n; j++) {
    b[j].re = ch[j].re;  b[j].im = -ch[j].im;
    b[m - j] = b[j];
  }
  orc_fft_pow2(a, m, -1, tw, m);
  orc_fft_pow2(b, m, -1, tw, m);
  for (long j = 0; j < m; j++) {
    double r = a[j].re * b[j].re - a[j].im * b[j].im;
    double i = a[j].re * b[j].im + a[j].im * b[j].re;
    a[j].re = r; a[j].im = i;
  }
  orc_fft_pow2(a, m, +1, tw, m);
  const double inv = 1.0 / (double)m;
  for (long k = 0; k < n; k++) {
    double r = a[k].re * inv, i = a[k].im * inv;
    x[k].re = r * ch[k].re - i * ch[k].im;
    x[k].im = r * ch[k].im + i * ch[k].re;
  }
  free(tw); free(a); free(b); free(ch);
}

void orc_fft_any(ocplx *x, long n, int sign)
{
  if (n <= 1) return;
  if (orc_is_pow2(n)) {
    ocplx *tw = orc_twiddle_table(n);
    orc_fft_pow2(x, n, sign, tw, n);
    free(tw);
  } else {
    bluestein(x, n, sign);
  }
}
