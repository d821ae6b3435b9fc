// fft.cu -- shared-memory tiled radix-2 DIT FFT passes (see fft.cuh).
#include "fft.cuh"

#include <math.h>

#include <vector>

namespace sfftb {

// ---------------------------------------------------------------------------
// host twiddles (octant rule; same definition the oracle pins in fft_ref.c)
// ---------------------------------------------------------------------------
void host_twiddle(long k, long n, double *re, double *im)
{
  const long n2 = n / 2, n4 = n / 4, n8 = n / 8;
  const double unit = M_PI / (double)(n2 > 0 ? n2 : 1);   // 2*pi/n
  double c, s;
  if (k <= n8) {
    const double a = (double)k * unit;
    c = cos(a); s = sin(a);
  } else if (k <= n4) {
    const double a = (double)(n4 - k) * unit;
    c = sin(a); s = cos(a);
  } else if (k <= n4 + n8) {
    const double a = (double)(k - n4) * unit;
    c = -sin(a); s = cos(a);
  } else {
    const double a = (double)(n2 - k) * unit;
    c = -cos(a); s = sin(a);
  }
  *re = c;
  *im = -s;
}

void host_twiddle_table(long n, cplx *out)
{
  if (n < 2) { out[0] = make_double2(1.0, -0.0); return; }
  for (long k = 0; k < n / 2; k++) host_twiddle(k, n, &out[k].x, &out[k].y);
}

void host_twiddle_levels(long n, cplx *out)
{
  // level s (0-based, half-size h = 2^s) holds W_{2h}^k, k < h, at offset h - 1.
  // W_{2h}^k = e^{-2 pi i k / 2h} evaluated by the octant rule at size n: the same
  // double whatever n >= 2h is (the scale factor is a power of two).
  for (long h = 1; h < n; h <<= 1)
    for (long k = 0; k < h; k++) host_twiddle(k * (n / (2 * h)), n, &out[h - 1 + k].x, &out[h - 1 + k].y);
}

void host_twiddle_factors(long n, std::vector<cplx> &coarse, std::vector<cplx> &fine)
{
  const long lo = 1L << 14, nhi = (n / 2 + lo - 1) / lo;
  coarse.resize((size_t)nhi);
  fine.resize((size_t)lo);
  for (long m = 0; m < nhi; m++) host_twiddle(m * lo, n, &coarse[(size_t)m].x, &coarse[(size_t)m].y);
  for (long l = 0; l < lo; l++) host_twiddle(l, n, &fine[(size_t)l].x, &fine[(size_t)l].y);
}

// ---------------------------------------------------------------------------
// one pass = stages [s0, s0+ns) of the DIT graph on a tile of 2^ns rows (stride
// 2^s0 elements) by 2^logT adjacent columns, staged in shared memory
// ---------------------------------------------------------------------------
constexpr int kFftThreads = 256;
constexpr int kTwFineLog = 14;         // == log2(ORC_TW_FINE) of the oracle's definition
constexpr int kMaxTileLog = 11;        // 2048 points = 32 KB of shared memory.  (Round 2 tried ONE 8192-point
                                       // pass for C1's location rows -- 144 KB tile, twiddles from L1/L2, three
                                       // CTAs of work -- and measured it slower than this plus a second pass: 55 vs 28 us.)
constexpr int kLaterPassStages = 8;    // keeps >= 8 adjacent columns (128 B) per row

// TABLE: `tw` is the level-ordered twiddle table (host_twiddle_levels).  The first
// pass (s0 == 0) copies levels [0, ns) next to the tile in shared memory; later
// passes read their (column-contiguous) twiddles through L1/L2.
// MODE 0: sincospi on the fly (accuracy reference only)
// MODE 1: level-ordered table `tw`
// MODE 2: two-factor twiddles for large transforms, W_N^K = A[K >> 14] * F[K & 16383]
//         (`tw` = A, `tw2` = F, N = 2^log_twN): the oracle's definition for N > 2^17
template <int MODE>
__device__ __forceinline__ cplx twiddle_at(int s, unsigned long long k, bool first_pass, const cplx *stw,
                                           const cplx *__restrict__ tw, const cplx *__restrict__ tw2,
                                           int log_twN, int sign)
{
  cplx w;
  if (MODE == 1) {
    w = first_pass ? stw[(1 << s) - 1 + (int)k] : __ldg(&tw[(1ull << s) - 1ull + k]);
  } else if (MODE == 2) {
    const unsigned long long K = k << (log_twN - s - 1);
    w = cmul_rn(__ldg(&tw[K >> kTwFineLog]), __ldg(&tw2[K & ((1ull << kTwFineLog) - 1ull)]));
  } else {
    double sn, cs;
    sincospi((double)k / (double)(1ull << s), &sn, &cs);
    w = make_double2(cs, -sn);
  }
  if (sign > 0) w.y = -w.y;
  return w;
}

__device__ __forceinline__ void butterfly(cplx &u, cplx &v, cplx w)
{
  const cplx t = cmul_rn(w, v);
  const cplx a = u;
  u = cadd_rn(a, t);
  v = csub_rn(a, t);
}

// The first pass works on one contiguous run of points: its threads read 8 neighbours
// each, which would land on the same banks; one pad slot per 8 points spreads them.
__device__ __forceinline__ int tile_slot(int i, bool padded) { return padded ? i + (i >> 3) : i; }

// Each round takes up to three consecutive stages of the radix-2 graph in registers:
// a thread owns the 8 points that differ in the three stage bits, so the arithmetic (and
// its rounding) is exactly the radix-2 butterflies, with a third of the barriers and
// shared-memory traffic.
template <int MODE>
__global__ void __launch_bounds__(kFftThreads)
fft_pass_kernel(cplx *base, int s0, int ns, int logT, long long fft_stride,
                long long sig_stride, const cplx *__restrict__ tw, const cplx *__restrict__ tw2,
                int log_twN, int sign)
{
  extern __shared__ cplx tile[];
  const bool first_pass = s0 == 0;
  const bool padded = logT == 0;
  const int T = 1 << logT;
  const int elems = 1 << (ns + logT);
  cplx *stw = tile + tile_slot(elems, padded) + 1;      // only used when MODE == 1 && first_pass
  const int lhi_bits = s0 - logT;
  const unsigned tile_id = blockIdx.x;
  const unsigned Lhi = tile_id & ((1u << lhi_bits) - 1u);
  const unsigned long long H = (unsigned long long)(tile_id >> lhi_bits);
  cplx *fft = base + (long long)blockIdx.z * sig_stride + (long long)blockIdx.y * fft_stride;
  const unsigned long long base_idx =
      (H << (s0 + ns)) | ((unsigned long long)Lhi << logT);

  for (int e = threadIdx.x; e < elems; e += kFftThreads) {
    const int r = e >> logT, c = e & (T - 1);
    tile[tile_slot(e, padded)] = fft[base_idx + ((unsigned long long)r << s0) + c];
  }
  if (MODE == 1 && first_pass)
    for (int e = threadIdx.x; e < (1 << ns) - 1; e += kFftThreads) stw[e] = __ldg(&tw[e]);
  __syncthreads();

  const unsigned long long Lfixed = (unsigned long long)Lhi << logT;
  int st = 0;
  while (st < ns) {
    const int r = ns - st >= 3 ? 3 : ns - st;          // stages this round
    const int units = elems >> r;
    for (int u = threadIdx.x; u < units; u += kFftThreads) {
      const int c = u & (T - 1);
      const int ru = u >> logT;
      const int r_lo = ru & ((1 << st) - 1);
      const int r_hi = ru >> st;
      const int row0 = (r_hi << (st + r)) | r_lo;
      const unsigned long long kbase = ((unsigned long long)r_lo << s0) | Lfixed | (unsigned)c;
      const unsigned long long kstep = 1ull << (s0 + st);
      cplx v[8];
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j < (1 << r)) v[j] = tile[tile_slot(((row0 + (j << st)) << logT) | c, padded)];
      // stage s0+st: pairs (j, j+1)
      {
        const cplx w = twiddle_at<MODE>(s0 + st, kbase, first_pass, stw, tw, tw2, log_twN, sign);
#pragma unroll
        for (int j = 0; j < 8; j += 2)
          if (j < (1 << r)) butterfly(v[j], v[j + 1], w);
      }
      if (r >= 2) {   // stage s0+st+1: pairs (j, j+2)
        const cplx w0 = twiddle_at<MODE>(s0 + st + 1, kbase, first_pass, stw, tw, tw2, log_twN, sign);
        const cplx w1 = twiddle_at<MODE>(s0 + st + 1, kbase + kstep, first_pass, stw, tw, tw2, log_twN, sign);
#pragma unroll
        for (int j = 0; j < 8; j += 4)
          if (j < (1 << r)) {
            butterfly(v[j], v[j + 2], w0);
            butterfly(v[j + 1], v[j + 3], w1);
          }
      }
      if (r >= 3) {   // stage s0+st+2: pairs (j, j+4)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const cplx w = twiddle_at<MODE>(s0 + st + 2, kbase + (unsigned long long)j * kstep, first_pass, stw, tw,
                                          tw2, log_twN, sign);
          butterfly(v[j], v[j + 4], w);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; j++)
        if (j < (1 << r)) tile[tile_slot(((row0 + (j << st)) << logT) | c, padded)] = v[j];
    }
    __syncthreads();
    st += r;
  }

  for (int e = threadIdx.x; e < elems; e += kFftThreads) {
    const int r = e >> logT, c = e & (T - 1);
    fft[base_idx + ((unsigned long long)r << s0) + c] = tile[tile_slot(e, padded)];
  }
}

int fft_dit_inplace(cplx *base, int logN, int nfft, long long fft_stride, int nsig,
                    long long sig_stride, const cplx *tw, int log_twN, int sign,
                    cudaStream_t st)
{
  return fft_dit_inplace_ex(base, logN, nfft, fft_stride, nsig, sig_stride, tw, nullptr, log_twN, sign, st);
}

int fft_dit_inplace_ex(cplx *base, int logN, int nfft, long long fft_stride, int nsig,
                       long long sig_stride, const cplx *tw, const cplx *tw_fine, int log_twN, int sign,
                       cudaStream_t st)
{
  if (logN <= 0 || nfft <= 0 || nsig <= 0) return 0;
  if (tw && !tw_fine && log_twN < logN) {
    set_error("fft_dit_inplace: twiddle table smaller than the transform");
    return -1;
  }
  if (tw_fine && log_twN != logN) {
    set_error("fft_dit_inplace: two-factor twiddles are per transform size");
    return -1;
  }
  int s0 = 0;
  while (s0 < logN) {
    int ns, logT;
    if (s0 == 0) {
      ns = logN < kMaxTileLog ? logN : kMaxTileLog;
      logT = 0;
    } else {
      const int rem = logN - s0;
      ns = rem < kLaterPassStages ? rem : kLaterPassStages;
      logT = kMaxTileLog - ns;
      if (logT > s0) logT = s0;
    }
    const long long tiles = 1ll << (logN - ns - logT);
    dim3 grid((unsigned)tiles, (unsigned)nfft, (unsigned)nsig);
    const size_t tile_elems = (size_t)1 << (ns + logT);
    size_t smem = sizeof(cplx) * (tile_elems + (logT == 0 ? (tile_elems >> 3) : 0) + 2);
    if (tw && !tw_fine && s0 == 0) smem += sizeof(cplx) << ns;
    SFFTB_ONCE_PER_DEVICE({
      const int max_smem = (int)(sizeof(cplx) * ((2u << kMaxTileLog) + (1u << (kMaxTileLog - 3)) + 4));
      SFFTB_CUDA(cudaFuncSetAttribute(fft_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      SFFTB_CUDA(cudaFuncSetAttribute(fft_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      SFFTB_CUDA(cudaFuncSetAttribute(fft_pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    });
    if (tw && tw_fine)
      fft_pass_kernel<2><<<grid, kFftThreads, smem, st>>>(base, s0, ns, logT, fft_stride, sig_stride, tw,
                                                         tw_fine, log_twN, sign);
    else if (tw)
      fft_pass_kernel<1><<<grid, kFftThreads, smem, st>>>(base, s0, ns, logT, fft_stride, sig_stride, tw,
                                                         nullptr, log_twN, sign);
    else
      fft_pass_kernel<0><<<grid, kFftThreads, smem, st>>>(base, s0, ns, logT, fft_stride, sig_stride, tw,
                                                         nullptr, log_twN, sign);
    SFFTB_LAUNCH_CHECK();
    s0 += ns;
  }
  return 0;
}

__global__ void bitrev_permute_kernel(const cplx *__restrict__ in, cplx *__restrict__ out,
                                      int logN)
{
  const long long n = 1ll << logN;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[bitrev((unsigned)i, logN)] = in[i];
}

int bitrev_permute(const cplx *in, cplx *out, int logN, cudaStream_t st)
{
  const long long n = 1ll << logN;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bitrev_permute_kernel<<<blocks, 256, 0, st>>>(in, out, logN);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sfftb
