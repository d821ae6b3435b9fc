// fft.cuh -- batched power-of-two complex-double FFTs for sm_100a.
//
// Replaces FFTW's role in the reference (fftw_execute at cf12.cc:69,270,273 and
// computefourier-3.0.cc:123,208,293; plans at sfft.cc:255-296,422-477).
//
// One arithmetic is used everywhere: the radix-2 decimation-in-time butterfly
// graph over bit-reversed input,
//     t = W_{2h}^k * v ;  (u, v) <- (u + t, u - t),   h = 1, 2, 4, ..., N/2
// with each product and sum individually rounded.  The kernels walk that graph
// in shared-memory tiles of 2048 points (up to 11 stages per pass), so the
// result does not depend on the tiling.  With a twiddle TABLE (built on the host
// by the octant rule below) the bucket spectra are bit-identical to the oracle's;
// the plan builder's n-point transforms use sincospi() on the fly instead.
#pragma once

#include <vector>

#include "common.cuh"

namespace sfftb {

// e^{-2 pi i k / n}, n a power of two, 0 <= k < n/2, by octant reduction so that
// only angles in [0, pi/4] reach libm (documented in DESIGN.md "twiddle rule").
void host_twiddle(long k, long n, double *re, double *im);
void host_twiddle_table(long n, cplx *out /* n/2 entries (>=1) */);
// level-ordered copy of the same values, n-1 entries: level s (half-size h = 2^s)
// at offset h-1 holds W_{2h}^k, k < h.  This is the layout the kernels read.
void host_twiddle_levels(long n, cplx *out /* max(n-1, 1) entries */);

// In-place DIT FFT over BIT-REVERSED input, natural-order output.
//   element e of transform f of signal s lives at base[s*sig_stride + f*fft_stride + e]
//   tw == nullptr -> twiddles from sincospi(); else LEVEL-ORDERED table for size 2^log_twN >= N
//   sign = -1 forward, +1 backward (unnormalised)
int fft_dit_inplace(cplx *base, int logN, int nfft, long long fft_stride, int nsig,
                    long long sig_stride, const cplx *tw, int log_twN, int sign,
                    cudaStream_t st);

// Transforms larger than 2^17 points (the plan builder's) take their twiddles from the
// oracle's two-factor definition W_N^K = coarse[K >> 14] * fine[K & 16383] (one rounded
// complex product), both factors by the octant rule at size N = 2^logN.
constexpr int kTwDirectMaxLog = 17;
void host_twiddle_factors(long n, std::vector<cplx> &coarse, std::vector<cplx> &fine);
int fft_dit_inplace_ex(cplx *base, int logN, int nfft, long long fft_stride, int nsig,
                       long long sig_stride, const cplx *tw, const cplx *tw_fine, int log_twN, int sign,
                       cudaStream_t st);

// out[bitrev(i)] = in[i] (out-of-place)
int bitrev_permute(const cplx *in, cplx *out, int logN, cudaStream_t st);

}  // namespace sfftb
