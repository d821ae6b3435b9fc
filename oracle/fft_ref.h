/*
 * oracle/fft_ref.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Fully specified double-precision complex DFT used (a) by the FFTW-API shim
 * that lets the reference sources under /root/reference compile in a container
 * without FFTW, and (b) by the C restatement in sfft_oracle.c.
 *
 * The reference delegates every DFT to FFTW 3 (un-vendored, unpinned;
 * /root/reference/src/fftw.cc:29-59, src/sfft.cc:255-296,422-477).  FFTW's
 * rounding is therefore not part of the reference's own sources; this file
 * pins ONE arithmetic for it:
 *
 *   power-of-two n : iterative radix-2 decimation-in-time, explicit bit
 *                    reversal, twiddles from orc_twiddle() below, every
 *                    product and sum individually rounded (build with
 *                    -ffp-contract=off, no -ffast-math);
 *   any other n    : Bluestein chirp-z over the power-of-two transform.
 *
 * The CUDA engine evaluates the same butterfly graph with the same twiddle
 * table for the bucket FFTs, so bucket spectra are bit-identical.
 */
#ifndef ORC_FFT_REF_H
#define ORC_FFT_REF_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } ocplx;

/* transforms up to this size take every twiddle straight from the octant rule; larger
 * ones use the two-factor definition in orc_twiddle_table() */
#define ORC_TW_DIRECT_MAX (1L << 17)
#define ORC_TW_FINE (1L << 14)

/* e^{-2 pi i k / n} for n a power of two, 0 <= k < n/2, by octant reduction:
 * only angles in [0, pi/4] are handed to libm, so values at multiples of
 * pi/4 are exact and the table is symmetric. */
void orc_twiddle(long k, long n, double *re, double *im);

/* table of n/2 twiddles for a power-of-two n (caller frees) */
ocplx *orc_twiddle_table(long n);

/* in-place power-of-two transform; sign=-1 forward, +1 backward; unnormalised.
 * tw is orc_twiddle_table(tw_n) for any power of two tw_n >= n. */
void orc_fft_pow2(ocplx *x, long n, int sign, const ocplx *tw, long tw_n);

/* in-place transform of any length (allocates its own tables) */
void orc_fft_any(ocplx *x, long n, int sign);

int orc_is_pow2(long n);

#ifdef __cplusplus
}
#endif
#endif
