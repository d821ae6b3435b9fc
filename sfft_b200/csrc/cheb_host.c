/*
 * cheb_host.c -- host-side seeds of the plan builder that must come from the
 * host libm to stay bit-compatible with a CPU build of the reference:
 *
 *  - the w frequency samples of the Dolph-Chebyshev window
 *    (reference src/filters.cc:62-80).  A handful of main-lobe samples carry a
 *    relative sensitivity of ~1e-6 to the rounding of t0*cos(pi*i/w) (cosh of a
 *    large multiple of acosh(1+eps)), so they are evaluated with the very same
 *    libm calls; the DFTs that follow run on the device.
 *  - the phase-ramp step e^{-2 pi i (w/2)/n} (src/filters.cc:134), whose rounding
 *    the reference amplifies over n running products.
 *
 * Plain C because it needs C99 <complex.h> (ccosh/cacosh/cexp).
 */
#include <complex.h>
#include <math.h>

static double cheb_poly(double m, double x)
{
  if (fabs(x) <= 1) return cos(m * acos(x));
  return creal(ccosh(m * cacosh(x)));
}

int sfftb_host_dolph_width(double lobefrac, double tolerance)
{
  /* src/filters.cc:72-74 */
  int w = (int)((1 / M_PI) * (1 / lobefrac) * acosh(1. / tolerance));
  if (!(w % 2)) w--;
  return w;
}

void sfftb_host_cheb_samples(double tolerance, int w, double *out)
{
  /* src/filters.cc:76-80 */
  double t0 = cosh(acosh(1 / tolerance) / (w - 1));
  for (int i = 0; i < w; i++)
    out[i] = cheb_poly(w - 1, t0 * cos(M_PI * i / w)) * tolerance;
}

void sfftb_host_ramp_step(int w, int n, double *re, double *im)
{
  /* src/filters.cc:134 */
  double complex step = cexp(-2 * M_PI * I * (w / 2) / n);
  *re = creal(step);
  *im = cimag(step);
}

/* Bluestein chirp e^{sign*pi*i*j^2/n}, j < n, with the argument reduced exactly (j^2 mod 2n)
 * before it reaches libm: the definition the oracle pins in fft_ref.c:bluestein. */
void sfftb_host_chirp(int n, int sign, double *out_re_im)
{
  for (long j = 0; j < n; j++) {
    long long q = ((long long)j * j) % (2 * (long long)n);
    double ang = M_PI * (double)q / (double)n;
    out_re_im[2 * j] = cos(ang);
    out_re_im[2 * j + 1] = (sign < 0 ? -1.0 : 1.0) * sin(ang);
  }
}

/* |re + i*im| as the reference takes it (cabs, src/filters.cc:128) */
double sfftb_host_cabs(double re, double im)
{
  return cabs(CMPLX(re, im));
}
