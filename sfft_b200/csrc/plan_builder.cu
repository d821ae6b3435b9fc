// plan_builder.cu -- window/filter construction on the device, BIT-IDENTICAL to the
// reference's arithmetic over the pinned DFT.
//
// Reference: make_dolphchebyshev_t (src/filters.cc:70-86) and make_multiple_t
// (src/filters.cc:109-160): one odd-length w-point DFT and two n-point DFTs per filter
// (seconds to minutes on a CPU at n >= 2^26).
//
// Why bit-identical and not merely close: for exact-sparse inputs the top-2k cutoff picks
// among noise-floor buckets that are ~1e-8 of the signal buckets, so a 1e-12 relative
// difference in the window taps is a 1e-4 relative difference there and flips a
// leakage-level location every few dozen transforms (measured).  Only identical taps make
// "locations bit-exact" hold by construction.  Hence:
//   * libm-dependent seeds come from the host's libm through the same calls as the
//     reference (cheb_host.c): Chebyshev samples, Bluestein chirp, ramp step, and the peak
//     magnitude (hypot) of a handful of candidates;
//   * the DFTs use the oracle's twiddle definition (octant rule; two-factor product above
//     2^17 points) and butterfly graph, all products/sums individually rounded;
//   * the two recurrences whose rounding depends on history -- the boxcar running sum
//     (filters.cc:125-130) and the phase-ramp running product (:134-140) -- are run as
//     genuinely sequential chains (one thread each, fed through shared memory by the rest
//     of its CTA, on two streams); everything else is data-parallel.
#include "plan_builder.cuh"

#include <math.h>
#include <stdlib.h>
#include <vector>

#include "fft.cuh"
#include "v12_kernels.cuh"

extern "C" {
int sfftb_host_dolph_width(double lobefrac, double tolerance);
void sfftb_host_cheb_samples(double tolerance, int w, double *out);
void sfftb_host_ramp_step(int w, int n, double *re, double *im);
void sfftb_host_chirp(int n, int sign, double *out_re_im);
double sfftb_host_cabs(double re, double im);
}

#include <chrono>
#include <string.h>
#include <thread>

namespace sfftb {

namespace {

// SFFTB_PLAN_TIMING=1: wall-clock of the builder's steps on stderr (device drained at each mark)
struct StepClock {
  bool on;
  std::chrono::steady_clock::time_point t0;
  StepClock() : on(getenv("SFFTB_PLAN_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void mark(const char *what)
  {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[libsfft plan] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

constexpr int kT = 256;
inline int grid_for(long long n)
{
  long long b = (n + kT - 1) / kT;
  if (b > 148 * 32) b = 148 * 32;
  return (int)(b < 1 ? 1 : b);
}
#define GRID_STRIDE(i, n)                                                   \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); \
       i += (long long)gridDim.x * blockDim.x)

// ---- twiddles of one transform size, in the form the FFT passes want ---------
struct Twiddles {
  int logN = 0;
  cplx *levels = nullptr;              // N <= 2^17: level-ordered table
  cplx *coarse = nullptr, *fine = nullptr;   // N > 2^17: two-factor definition
};

int make_twiddles(int logN, Twiddles *t, cudaStream_t st)
{
  t->logN = logN;
  const long n = 1L << logN;
  if (logN <= kTwDirectMaxLog) {
    std::vector<cplx> lv((size_t)(n > 1 ? n - 1 : 1));
    host_twiddle_levels(n, lv.data());
    SFFTB_CUDA(cudaMalloc(&t->levels, sizeof(cplx) * lv.size()));
    SFFTB_CUDA(cudaMemcpyAsync(t->levels, lv.data(), sizeof(cplx) * lv.size(), cudaMemcpyHostToDevice, st));
    SFFTB_CUDA(cudaStreamSynchronize(st));
  } else {
    std::vector<cplx> c, f;
    host_twiddle_factors(n, c, f);
    SFFTB_CUDA(cudaMalloc(&t->coarse, sizeof(cplx) * c.size()));
    SFFTB_CUDA(cudaMalloc(&t->fine, sizeof(cplx) * f.size()));
    SFFTB_CUDA(cudaMemcpyAsync(t->coarse, c.data(), sizeof(cplx) * c.size(), cudaMemcpyHostToDevice, st));
    SFFTB_CUDA(cudaMemcpyAsync(t->fine, f.data(), sizeof(cplx) * f.size(), cudaMemcpyHostToDevice, st));
    SFFTB_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

void free_twiddles(Twiddles *t)
{
  cudaFree(t->levels); cudaFree(t->coarse); cudaFree(t->fine);
  t->levels = t->coarse = t->fine = nullptr;
}

int fft_with(const Twiddles &t, cplx *data, int sign, cudaStream_t st)
{
  const long long n = 1ll << t.logN;
  if (t.levels) return fft_dit_inplace_ex(data, t.logN, 1, n, 1, n, t.levels, nullptr, t.logN, sign, st);
  return fft_dit_inplace_ex(data, t.logN, 1, n, 1, n, t.coarse, t.fine, t.logN, sign, st);
}

// ---- Bluestein pieces, op for op as oracle/fft_ref.c:bluestein ---------------
__global__ void bluestein_prep_kernel(const cplx *__restrict__ x, const cplx *__restrict__ ch, int w, int logM,
                                      cplx *A, cplx *Bk)
{
  const long long M = 1ll << logM;
  GRID_STRIDE(j, M)
  {
    cplx a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
    if (j < w) {
      const cplx c = ch[j];
      a = cmul_rn(x[j], c);                  // (x.re*c.re - x.im*c.im, x.re*c.im + x.im*c.re)
      b = make_double2(c.x, -c.y);
    } else if (M - j < w) {
      const cplx c = ch[M - j];
      b = make_double2(c.x, -c.y);
    }
    const unsigned r = bitrev((unsigned)j, logM);
    A[r] = a;
    Bk[r] = b;
  }
}

__global__ void pointwise_mul_bitrev_kernel(const cplx *__restrict__ A, const cplx *__restrict__ Bk,
                                            int logM, cplx *out)
{
  const long long M = 1ll << logM;
  GRID_STRIDE(j, M) { out[bitrev((unsigned)j, logM)] = cmul_rn(A[j], Bk[j]); }
}

// X[k] = (C[k] * (1/M)) * chirp[k]
__global__ void bluestein_finish_kernel(const cplx *__restrict__ C, const cplx *__restrict__ ch, int w,
                                        int logM, cplx *X)
{
  const double inv = 1.0 / (double)(1ll << logM);
  GRID_STRIDE(k, w)
  {
    const cplx v = make_double2(__dmul_rn(C[k].x, inv), __dmul_rn(C[k].y, inv));
    X[k] = cmul_rn(v, ch[k]);
  }
}

// taps0[(k + w/2) % w] = Re X[k]       (filters.cc:82-84, utils.cc:28-38)
__global__ void rotate_real_kernel(const cplx *__restrict__ X, int w, double *taps0)
{
  GRID_STRIDE(k, w) { taps0[(k + w / 2) % w] = X[k].x; }
}

// g[(i - w/2) mod n] = taps0[i], written at its bit-reversed place (filters.cc:113-114)
__global__ void centre_scatter_kernel(const double *__restrict__ taps0, int w, int logn, cplx *G)
{
  const long long n = 1ll << logn;
  GRID_STRIDE(i, w)
  {
    const long long idx = (i - w / 2 + n) & (n - 1);
    G[bitrev((unsigned)idx, logn)] = make_double2(taps0[i], 0.0);
  }
}

// ---- the two history-dependent recurrences -----------------------------------
constexpr int kSeqChunk = 2048;
constexpr int kSeqThreads = 256;

// filters.cc:119-130:  s = sum_{i<b} g[i];  for i: h[(i+b/2)%n] = s;  s = s + (g[(i+b)%n] - g[i])
// One CTA: all threads form the differences of a chunk (each a single rounding, as in the
// reference), thread 0 runs the additions in order, all threads store the chunk.
__global__ void __launch_bounds__(kSeqThreads)
boxcar_sequential_kernel(const cplx *__restrict__ g, int logn, int b, cplx *H, unsigned long long *max_q_bits)
{
  __shared__ cplx dbuf[kSeqChunk];        // differences in, running sums out (in place)
  __shared__ double s_re, s_im;
  const long long n = 1ll << logn;
  const int tid = threadIdx.x;
  if (tid == 0) {
    double sr = 0.0, si = 0.0;
    for (int i = 0; i < b; i++) {           // :121-124
      sr = __dadd_rn(sr, g[i].x);
      si = __dadd_rn(si, g[i].y);
    }
    s_re = sr; s_im = si;
  }
  __syncthreads();
  const long long off = b / 2;
  double qmax = 0.0;
  for (long long base = 0; base < n; base += kSeqChunk) {
    for (int e = tid; e < kSeqChunk; e += kSeqThreads) {
      const long long i = base + e;
      if (i < n) {
        const cplx in = g[(i + b) & (n - 1)], out = g[i];
        dbuf[e] = make_double2(__dsub_rn(in.x, out.x), __dsub_rn(in.y, out.y));
      }
    }
    __syncthreads();
    if (tid == 0) {
      double sr = s_re, si = s_im;
      const int cnt = (int)((n - base) < kSeqChunk ? (n - base) : kSeqChunk);
#pragma unroll 8
      for (int e = 0; e < cnt; e++) {
        const cplx d = dbuf[e];
        dbuf[e] = make_double2(sr, si);
        sr = __dadd_rn(sr, d.x);
        si = __dadd_rn(si, d.y);
      }
      s_re = sr; s_im = si;
    }
    __syncthreads();
    for (int e = tid; e < kSeqChunk; e += kSeqThreads) {
      const long long i = base + e;
      if (i < n) {
        const cplx s = dbuf[e];
        H[(i + n + off) & (n - 1)] = s;
        const double q = s.x * s.x + s.y * s.y;
        qmax = q > qmax ? q : qmax;
      }
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, qmax, o);
    qmax = other > qmax ? other : qmax;
  }
  if ((tid & 31) == 0) atomicMax(max_q_bits, (unsigned long long)__double_as_longlong(qmax));
}

// candidates for the peak: every entry whose |.|^2 is within 1e-12 of the largest
__global__ void peak_candidates_kernel(const cplx *__restrict__ H, int logn, const unsigned long long *max_q_bits,
                                       cplx *cand, int *ncand, int cap)
{
  const long long n = 1ll << logn;
  const double thr = __longlong_as_double((long long)*max_q_bits) * (1.0 - 1e-12);
  GRID_STRIDE(i, n)
  {
    const cplx s = H[i];
    if (s.x * s.x + s.y * s.y >= thr) {
      const int pos = atomicAdd(ncand, 1);
      if (pos < cap) cand[pos] = s;
    }
  }
}

// ---- the same two recurrences on HOST threads ---------------------------------------------
// A dependent chain of n double-precision additions is ~30 ns a step on one GPU thread and
// ~1.5 ns on a host core; at n = 2^27 the device chains were 96 % of the plan time.  For
// large n each chain runs on its own host thread, streaming through two pinned staging
// buffers (download window -> add in order -> upload), so copies hide under the arithmetic.
// Same operations, same order, one rounding each (host code is built with -ffp-contract=off).
constexpr long long kHostChunk = 1ll << 21;            // elements per staging buffer (32 MiB)

struct HostChainResult {
  int rc = 0;
  double qmax = 0.0;
};

static bool host_chains_wanted(int logn)
{
  if (const char *e = getenv("SFFTB_HOST_CHAINS")) return atoi(e) != 0;
  return logn >= 24;
}

#define CHAIN_CUDA(call) do { if ((call) != cudaSuccess) { res->rc = -1; goto done; } } while (0)

// filters.cc:134-140
static void ramp_host_chain(int device, int logn, double step_re, double step_im, cplx *d_R, HostChainResult *res)
{
  const long long n = 1ll << logn;
  const long long C = n < kHostChunk ? n : kHostChunk;
  cplx *h[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaStream_t st = nullptr;
  double cr = 1.0, ci = 0.0;
  CHAIN_CUDA(cudaSetDevice(device));
  CHAIN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (int k = 0; k < 2; k++) {
    CHAIN_CUDA(cudaHostAlloc(&h[k], sizeof(cplx) * C, cudaHostAllocDefault));
    CHAIN_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
  }
  for (long long base = 0, k = 0; base < n; base += C, k ^= 1) {
    CHAIN_CUDA(cudaEventSynchronize(ev[k]));
    cplx *o = h[k];
    const long long cnt = n - base < C ? n - base : C;
    for (long long e = 0; e < cnt; e++) {
      o[e].x = cr; o[e].y = ci;
      const double nr = cr * step_re - ci * step_im;
      const double ni = cr * step_im + ci * step_re;
      cr = nr; ci = ni;
    }
    CHAIN_CUDA(cudaMemcpyAsync(d_R + base, o, sizeof(cplx) * cnt, cudaMemcpyHostToDevice, st));
    CHAIN_CUDA(cudaEventRecord(ev[k], st));
  }
  CHAIN_CUDA(cudaStreamSynchronize(st));
done:
  for (int k = 0; k < 2; k++) {
    if (h[k]) cudaFreeHost(h[k]);
    if (ev[k]) cudaEventDestroy(ev[k]);
  }
  if (st) cudaStreamDestroy(st);
}

// copy `cnt` elements starting at index `start` (mod n) of a device array of n elements
static cudaError_t copy_ring(cplx *host, cplx *dev, long long n, long long start, long long cnt, bool to_host,
                             cudaStream_t st)
{
  start &= n - 1;
  const long long first = cnt < n - start ? cnt : n - start;
  cudaError_t e = to_host ? cudaMemcpyAsync(host, dev + start, sizeof(cplx) * first, cudaMemcpyDeviceToHost, st)
                          : cudaMemcpyAsync(dev + start, host, sizeof(cplx) * first, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess || first == cnt) return e;
  return to_host ? cudaMemcpyAsync(host + first, dev, sizeof(cplx) * (cnt - first), cudaMemcpyDeviceToHost, st)
                 : cudaMemcpyAsync(dev, host + first, sizeof(cplx) * (cnt - first), cudaMemcpyHostToDevice, st);
}

// filters.cc:119-130:  s = sum_{i<b} g[i];  for i: h[(i+b/2)%n] = s;  s = s + (g[(i+b)%n] - g[i])
static void boxcar_host_chain(int device, int logn, int b, cplx *d_G, cplx *d_H, HostChainResult *res)
{
  const long long n = 1ll << logn;
  const long long C = n < kHostChunk ? n : kHostChunk;
  const long long win = C + b;                          // chunk i needs g[base .. base + C + b)
  const long long off = b / 2;
  cplx *hin[2] = {nullptr, nullptr}, *hout[2] = {nullptr, nullptr};
  cudaEvent_t evd[2] = {nullptr, nullptr}, evu[2] = {nullptr, nullptr};
  cudaStream_t sd = nullptr, su = nullptr;
  double sr = 0.0, si = 0.0, qmax = 0.0;
  const long long nchunks = (n + C - 1) / C;
  CHAIN_CUDA(cudaSetDevice(device));
  CHAIN_CUDA(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
  CHAIN_CUDA(cudaStreamCreateWithFlags(&su, cudaStreamNonBlocking));
  for (int k = 0; k < 2; k++) {
    CHAIN_CUDA(cudaHostAlloc(&hin[k], sizeof(cplx) * win, cudaHostAllocDefault));
    CHAIN_CUDA(cudaHostAlloc(&hout[k], sizeof(cplx) * C, cudaHostAllocDefault));
    CHAIN_CUDA(cudaEventCreateWithFlags(&evd[k], cudaEventDisableTiming));
    CHAIN_CUDA(cudaEventCreateWithFlags(&evu[k], cudaEventDisableTiming));
  }
  // :121-124, the first b elements in order (b can exceed a staging buffer)
  for (long long base = 0; base < b; base += C) {
    const long long cnt = b - base < C ? b - base : C;
    CHAIN_CUDA(copy_ring(hin[0], d_G, n, base, cnt, true, sd));
    CHAIN_CUDA(cudaStreamSynchronize(sd));
    for (long long e = 0; e < cnt; e++) { sr = sr + hin[0][e].x; si = si + hin[0][e].y; }
  }
  CHAIN_CUDA(copy_ring(hin[0], d_G, n, 0, win < n ? win : n, true, sd));
  CHAIN_CUDA(cudaEventRecord(evd[0], sd));
  for (long long c = 0; c < nchunks; c++) {
    const int k = (int)(c & 1);
    const long long base = c * C;
    const long long cnt = n - base < C ? n - base : C;
    if (c + 1 < nchunks) {
      // the next window goes into the other buffer, whose chunk has been consumed already
      CHAIN_CUDA(copy_ring(hin[k ^ 1], d_G, n, base + C, win < n ? win : n, true, sd));
      CHAIN_CUDA(cudaEventRecord(evd[k ^ 1], sd));
    }
    CHAIN_CUDA(cudaEventSynchronize(evd[k]));
    CHAIN_CUDA(cudaEventSynchronize(evu[k]));
    const cplx *g = hin[k];
    cplx *o = hout[k];
    if (win <= n) {
      for (long long e = 0; e < cnt; e++) {
        o[e].x = sr; o[e].y = si;
        const double q = sr * sr + si * si;
        qmax = q > qmax ? q : qmax;
        const double dr = g[e + b].x - g[e].x, di = g[e + b].y - g[e].y;
        sr = sr + dr; si = si + di;
      }
    } else {
      // tiny n (only reachable through SFFTB_HOST_CHAINS=1): the window wraps inside the buffer
      for (long long e = 0; e < cnt; e++) {
        o[e].x = sr; o[e].y = si;
        const double q = sr * sr + si * si;
        qmax = q > qmax ? q : qmax;
        const cplx in = g[(e + b) & (n - 1)], out = g[e];
        sr = sr + (in.x - out.x); si = si + (in.y - out.y);
      }
    }
    CHAIN_CUDA(copy_ring(o, d_H, n, base + off, cnt, false, su));
    CHAIN_CUDA(cudaEventRecord(evu[k], su));
  }
  CHAIN_CUDA(cudaStreamSynchronize(su));
  res->qmax = qmax;
done:
  for (int k = 0; k < 2; k++) {
    if (hin[k]) cudaFreeHost(hin[k]);
    if (hout[k]) cudaFreeHost(hout[k]);
    if (evd[k]) cudaEventDestroy(evd[k]);
    if (evu[k]) cudaEventDestroy(evu[k]);
  }
  if (sd) cudaStreamDestroy(sd);
  if (su) cudaStreamDestroy(su);
}
#undef CHAIN_CUDA

// filters.cc:134-140:  offsetc = 1;  for i: ramp[i] = offsetc;  offsetc *= step
__global__ void __launch_bounds__(kSeqThreads)
ramp_sequential_kernel(int logn, double step_re, double step_im, cplx *R)
{
  __shared__ cplx rbuf[kSeqChunk];
  __shared__ double c_re, c_im;
  const long long n = 1ll << logn;
  const int tid = threadIdx.x;
  if (tid == 0) { c_re = 1.0; c_im = 0.0; }
  __syncthreads();
  for (long long base = 0; base < n; base += kSeqChunk) {
    if (tid == 0) {
      double cr = c_re, ci = c_im;
      const int cnt = (int)((n - base) < kSeqChunk ? (n - base) : kSeqChunk);
#pragma unroll 8
      for (int e = 0; e < cnt; e++) {
        rbuf[e] = make_double2(cr, ci);
        const double nr = __dsub_rn(__dmul_rn(cr, step_re), __dmul_rn(ci, step_im));
        const double ni = __dadd_rn(__dmul_rn(cr, step_im), __dmul_rn(ci, step_re));
        cr = nr; ci = ni;
      }
      c_re = cr; c_im = ci;
    }
    __syncthreads();
    for (int e = tid; e < kSeqChunk; e += kSeqThreads) {
      const long long i = base + e;
      if (i < n) R[i] = rbuf[e];
    }
    __syncthreads();
  }
}

// h[i] = (h[i] / peak) * ramp[i]   (filters.cc:131-138), written bit-reversed for the inverse FFT,
// plus the response window fwin[m] = h[(m - half) mod n]
__global__ void normalise_ramp_kernel(const cplx *__restrict__ H, const cplx *__restrict__ R, int logn,
                                      double peak, cplx *G, int half, cplx *fwin)
{
  const long long n = 1ll << logn;
  GRID_STRIDE(i, n)
  {
    const cplx h0 = H[i];
    const cplx h = cmul_rn(make_double2(__ddiv_rn(h0.x, peak), __ddiv_rn(h0.y, peak)), R[i]);
    G[bitrev((unsigned)i, logn)] = h;
    if (i <= half) fwin[half + i] = h;
    if (i >= n - half) fwin[half - (n - i)] = h;
  }
}

__global__ void extract_taps_kernel(const cplx *__restrict__ G, int w, int logn, cplx *taps)
{
  const double nn = (double)(1ll << logn);
  GRID_STRIDE(i, w) { taps[i] = make_double2(__ddiv_rn(G[i].x, nn), __ddiv_rn(G[i].y, nn)); }
}

}  // namespace

// forward DFT of arbitrary length w (device in/out): Bluestein over a 2^logM-point FFT,
// the chirp e^{-pi i j^2 / w} evaluated by the host's libm exactly as the oracle does
int bluestein_forward(const cplx *d_x, int w, cplx *d_out, cudaStream_t st)
{
  int logM = 0;
  while ((1ll << logM) < 2ll * w - 1) logM++;
  const long long M = 1ll << logM;
  std::vector<cplx> chirp((size_t)w);
  sfftb_host_chirp(w, -1, reinterpret_cast<double *>(chirp.data()));
  Twiddles tw;
  if (make_twiddles(logM, &tw, st)) return -1;
  cplx *d_A = nullptr, *d_B = nullptr, *d_C = nullptr, *d_ch = nullptr;
  ScratchGuard guard;
  guard.track(tw.levels); guard.track(tw.coarse); guard.track(tw.fine);
  SFFTB_CUDA(cudaMalloc(&d_A, sizeof(cplx) * M));
  guard.track(d_A);
  SFFTB_CUDA(cudaMalloc(&d_B, sizeof(cplx) * M));
  guard.track(d_B);
  SFFTB_CUDA(cudaMalloc(&d_C, sizeof(cplx) * M));
  guard.track(d_C);
  SFFTB_CUDA(cudaMalloc(&d_ch, sizeof(cplx) * w));
  guard.track(d_ch);
  SFFTB_CUDA(cudaMemcpyAsync(d_ch, chirp.data(), sizeof(cplx) * w, cudaMemcpyHostToDevice, st));
  bluestein_prep_kernel<<<grid_for(M), kT, 0, st>>>(d_x, d_ch, w, logM, d_A, d_B);
  SFFTB_LAUNCH_CHECK();
  if (fft_with(tw, d_A, -1, st)) return -1;
  if (fft_with(tw, d_B, -1, st)) return -1;
  pointwise_mul_bitrev_kernel<<<grid_for(M), kT, 0, st>>>(d_A, d_B, logM, d_C);
  SFFTB_LAUNCH_CHECK();
  if (fft_with(tw, d_C, +1, st)) return -1;
  bluestein_finish_kernel<<<grid_for(w), kT, 0, st>>>(d_C, d_ch, w, logM, d_out);
  SFFTB_LAUNCH_CHECK();
  SFFTB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_A); cudaFree(d_B); cudaFree(d_C); cudaFree(d_ch);
  free_twiddles(&tw);
  guard.dismiss();
  return 0;
}

int filter_width(double lobefrac, double tolerance)
{
  return sfftb_host_dolph_width(lobefrac, tolerance);
}

namespace {

// one distinct Dolph-Chebyshev window: its width, spectrum G (n points) and phase ramp R
struct WindowWork {
  double lobefrac = 0, tolerance = 0;
  int w = 0;
  cplx *d_G = nullptr;      // FFT_n of the centred window
  cplx *d_R = nullptr;      // ramp step^i
  cudaStream_t ramp_stream = nullptr;
  std::thread *ramp_thread = nullptr;      // host chain (large n)
  HostChainResult ramp_res;
};

int build_window(int logn, WindowWork *ww, const Twiddles &twn, cudaStream_t st)
{
  const long long n = 1ll << logn;
  const int w = ww->w;
  std::vector<double> samples_re((size_t)w);
  sfftb_host_cheb_samples(ww->tolerance, w, samples_re.data());
  std::vector<cplx> samples((size_t)w);
  for (int i = 0; i < w; i++) samples[(size_t)i] = make_double2(samples_re[(size_t)i], 0.0);
  double step_re, step_im;
  sfftb_host_ramp_step(w, (int)n, &step_re, &step_im);

  cplx *d_samples = nullptr, *d_X = nullptr;
  double *d_taps0 = nullptr;
  ScratchGuard guard;
  SFFTB_CUDA(cudaMalloc(&d_samples, sizeof(cplx) * w));
  guard.track(d_samples);
  SFFTB_CUDA(cudaMalloc(&d_X, sizeof(cplx) * w));
  guard.track(d_X);
  SFFTB_CUDA(cudaMalloc(&d_taps0, sizeof(double) * w));
  guard.track(d_taps0);
  SFFTB_CUDA(cudaMalloc(&ww->d_G, sizeof(cplx) * n));
  SFFTB_CUDA(cudaMalloc(&ww->d_R, sizeof(cplx) * n));
  // the phase ramp does not depend on the data: start its chain on its own stream now
  SFFTB_CUDA(cudaStreamCreateWithFlags(&ww->ramp_stream, cudaStreamNonBlocking));
  if (host_chains_wanted(logn)) {
    int device = 0;
    SFFTB_CUDA(cudaGetDevice(&device));
    ww->ramp_thread = new std::thread(ramp_host_chain, device, logn, step_re, step_im, ww->d_R, &ww->ramp_res);
  } else {
    ramp_sequential_kernel<<<1, kSeqThreads, 0, ww->ramp_stream>>>(logn, step_re, step_im, ww->d_R);
    SFFTB_LAUNCH_CHECK();
  }

  SFFTB_CUDA(cudaMemcpyAsync(d_samples, samples.data(), sizeof(cplx) * w, cudaMemcpyHostToDevice, st));
  // w-point DFT by Bluestein (filters.cc:81), rotate, keep the real part
  if (bluestein_forward(d_samples, w, d_X, st)) return -1;
  rotate_real_kernel<<<grid_for(w), kT, 0, st>>>(d_X, w, d_taps0);
  SFFTB_LAUNCH_CHECK();
  // filters.cc:113-115
  SFFTB_CUDA(cudaMemsetAsync(ww->d_G, 0, sizeof(cplx) * n, st));
  centre_scatter_kernel<<<grid_for(w), kT, 0, st>>>(d_taps0, w, logn, ww->d_G);
  SFFTB_LAUNCH_CHECK();
  if (fft_with(twn, ww->d_G, -1, st)) return -1;
  SFFTB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_samples); cudaFree(d_X); cudaFree(d_taps0);
  guard.dismiss();
  return 0;
}

}  // namespace

// Build `count` filters of one plan together: filters that share (lobefrac, tolerance)
// share the window, its n-point spectrum and the phase ramp (only the boxcar width differs,
// the default for k > 50: src/sfft.cc:306-314), and all sequential chains run concurrently.
static thread_local PresetFilters g_preset;
void set_preset_filters(const PresetFilters *preset) { g_preset = preset ? *preset : PresetFilters(); }

// upload cached filters (sfftb_load_plan) after checking them against the plan's geometry
static int upload_preset(int count, const FilterSpec *specs, DeviceFilter **outs, cudaStream_t st)
{
  const PresetFilters pre = g_preset;
  g_preset = PresetFilters();
  if (pre.count != count) { set_error("plan cache: filter count does not match this plan"); return -1; }
  for (int f = 0; f < count; f++) {
    const int w = sfftb_host_dolph_width(specs[f].lobefrac, specs[f].tolerance);
    if (pre.w[f] != w || pre.fw_half[f] != specs[f].fw_half) {
      set_error("plan cache: filter sizes do not match the plan derived from (n, k, version, flags)");
      return -1;
    }
    DeviceFilter *o = outs[f];
    o->w = w;
    o->fw_half = specs[f].fw_half;
    const long long len = 2ll * o->fw_half + 1;
    SFFTB_CUDA(cudaMalloc(&o->time, sizeof(cplx) * w));
    SFFTB_CUDA(cudaMalloc(&o->fwin, sizeof(cplx) * len));
    SFFTB_CUDA(cudaMemcpyAsync(o->time, pre.time[f], sizeof(cplx) * w, cudaMemcpyHostToDevice, st));
    SFFTB_CUDA(cudaMemcpyAsync(o->fwin, pre.fwin[f], sizeof(cplx) * len, cudaMemcpyHostToDevice, st));
    if (filter_refresh(o, st)) return -1;
  }
  SFFTB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int build_filters(int logn, int count, const FilterSpec *specs, DeviceFilter **outs, cudaStream_t st)
{
  const long long n = 1ll << logn;
  if (count < 1 || count > 2) { set_error("build_filters: one or two filters per plan"); return -1; }
  if (g_preset.count) return upload_preset(count, specs, outs, st);
  WindowWork win[2];
  int which_win[2] = {0, 0}, nwin = 0;
  for (int f = 0; f < count; f++) {
    const int w = sfftb_host_dolph_width(specs[f].lobefrac, specs[f].tolerance);
    if (w < 1 || w > n || specs[f].b > n || specs[f].b < 1 || specs[f].fw_half >= n / 2) {
      set_error("build_filter: window does not fit the signal length (reference asserts, filters.cc:111-112)");
      return -1;
    }
    outs[f]->w = w;
    outs[f]->fw_half = specs[f].fw_half;
    int found = -1;
    for (int q = 0; q < nwin; q++)
      if (win[q].lobefrac == specs[f].lobefrac && win[q].tolerance == specs[f].tolerance) found = q;
    if (found < 0) {
      found = nwin++;
      win[found].lobefrac = specs[f].lobefrac;
      win[found].tolerance = specs[f].tolerance;
      win[found].w = w;
    }
    which_win[f] = found;
  }
  StepClock clk;
  // every device temporary below is registered here: an early `return -1` frees them (the
  // filters' own arrays belong to the plan and are released by free_filter)
  ScratchGuard guard;
  // declared after the guard, hence destroyed before it: a host ramp chain still running on an
  // early return is joined before its target buffer is freed
  struct RampJoiner {
    WindowWork *w;
    ~RampJoiner()
    {
      for (int q = 0; q < 2; q++)
        if (w[q].ramp_thread) { w[q].ramp_thread->join(); delete w[q].ramp_thread; w[q].ramp_thread = nullptr; }
    }
  } ramp_joiner{win};
  Twiddles twn;
  if (make_twiddles(logn, &twn, st)) return -1;
  guard.track(twn.levels); guard.track(twn.coarse); guard.track(twn.fine);
  clk.mark("twiddles");
  for (int q = 0; q < nwin; q++) {
    const int rc = build_window(logn, &win[q], twn, st);
    guard.track(win[q].d_G); guard.track(win[q].d_R); guard.track_stream(win[q].ramp_stream);
    if (rc) return -1;
  }
  clk.mark("window + its spectrum");

  // boxcar chains, one stream per filter
  const int cand_cap = 4096;
  cudaStream_t fs[2] = {nullptr, nullptr};
  cplx *d_H[2] = {nullptr, nullptr}, *d_cand[2] = {nullptr, nullptr};
  unsigned long long *d_maxq[2] = {nullptr, nullptr};
  int *d_ncand[2] = {nullptr, nullptr};
  // allocate for every filter first: a cudaMalloc between the two launches made the second
  // chain wait for the first (measured: 2 x 1.4 s back to back at n = 2^27)
  for (int f = 0; f < count; f++) {
    SFFTB_CUDA(cudaStreamCreateWithFlags(&fs[f], cudaStreamNonBlocking));
    guard.track_stream(fs[f]);
    SFFTB_CUDA(cudaMalloc(&d_H[f], sizeof(cplx) * n));
    guard.track(d_H[f]);
    SFFTB_CUDA(cudaMalloc(&d_cand[f], sizeof(cplx) * cand_cap));
    guard.track(d_cand[f]);
    SFFTB_CUDA(cudaMalloc(&d_maxq[f], sizeof(unsigned long long)));
    guard.track(d_maxq[f]);
    SFFTB_CUDA(cudaMalloc(&d_ncand[f], sizeof(int)));
    guard.track(d_ncand[f]);
    SFFTB_CUDA(cudaMalloc(&outs[f]->time, sizeof(cplx) * outs[f]->w));
    SFFTB_CUDA(cudaMalloc(&outs[f]->fwin, sizeof(cplx) * (2ll * specs[f].fw_half + 1)));
  }
  const bool host_chains = host_chains_wanted(logn);
  std::thread *box_thread[2] = {nullptr, nullptr};
  HostChainResult box_res[2];
  if (host_chains) {
    int device = 0;
    SFFTB_CUDA(cudaGetDevice(&device));
    for (int f = 0; f < count; f++)
      box_thread[f] = new std::thread(boxcar_host_chain, device, logn, specs[f].b, win[which_win[f]].d_G, d_H[f], &box_res[f]);
  }
  for (int f = 0; f < count; f++) {
    SFFTB_CUDA(cudaMemsetAsync(d_ncand[f], 0, sizeof(int), fs[f]));
    if (host_chains) {
      box_thread[f]->join();
      delete box_thread[f];
      box_thread[f] = nullptr;
      if (box_res[f].rc) {
        for (int q = f + 1; q < count; q++) { box_thread[q]->join(); delete box_thread[q]; }
        set_error("build_filter: the host boxcar chain failed (CUDA error in its staging copies)");
        return -1;
      }
      unsigned long long bits;
      memcpy(&bits, &box_res[f].qmax, sizeof bits);
      SFFTB_CUDA(cudaMemcpyAsync(d_maxq[f], &bits, sizeof bits, cudaMemcpyHostToDevice, fs[f]));
      SFFTB_CUDA(cudaStreamSynchronize(fs[f]));
    } else {
      SFFTB_CUDA(cudaMemsetAsync(d_maxq[f], 0, sizeof(unsigned long long), fs[f]));
      boxcar_sequential_kernel<<<1, kSeqThreads, 0, fs[f]>>>(win[which_win[f]].d_G, logn, specs[f].b, d_H[f], d_maxq[f]);
      SFFTB_LAUNCH_CHECK();
    }
    peak_candidates_kernel<<<grid_for(n), kT, 0, fs[f]>>>(d_H[f], logn, d_maxq[f], d_cand[f], d_ncand[f], cand_cap);
    SFFTB_LAUNCH_CHECK();
  }
  clk.mark("boxcar + ramp chains");
  for (int f = 0; f < count; f++) {
    const WindowWork &ww = win[which_win[f]];
    int ncand = 0;
    SFFTB_CUDA(cudaMemcpyAsync(&ncand, d_ncand[f], sizeof(int), cudaMemcpyDeviceToHost, fs[f]));
    SFFTB_CUDA(cudaStreamSynchronize(fs[f]));
    if (ncand < 1) { set_error("build_filter: empty filter response"); return -1; }
    if (ncand > cand_cap) ncand = cand_cap;
    std::vector<cplx> cand((size_t)ncand);
    SFFTB_CUDA(cudaMemcpy(cand.data(), d_cand[f], sizeof(cplx) * ncand, cudaMemcpyDeviceToHost));
    double peak = 0.0;                                   // max over cabs(s), filters.cc:128 (host hypot)
    for (int i = 0; i < ncand; i++) {
      const double m = sfftb_host_cabs(cand[(size_t)i].x, cand[(size_t)i].y);
      if (m > peak) peak = m;
    }
    if (win[which_win[f]].ramp_thread) {
      WindowWork &wr = win[which_win[f]];
      wr.ramp_thread->join();
      delete wr.ramp_thread;
      wr.ramp_thread = nullptr;
      if (wr.ramp_res.rc) { set_error("build_filter: the host ramp chain failed"); return -1; }
    }
    SFFTB_CUDA(cudaStreamSynchronize(ww.ramp_stream));
    // the spectrum G of the window is no longer needed once every boxcar over it has run;
    // the inverse transform gets its own buffer
    cplx *d_T = nullptr;
    SFFTB_CUDA(cudaMalloc(&d_T, sizeof(cplx) * n));
    guard.track(d_T);
    normalise_ramp_kernel<<<grid_for(n), kT, 0, fs[f]>>>(d_H[f], ww.d_R, logn, peak, d_T, specs[f].fw_half,
                                                        outs[f]->fwin);
    SFFTB_LAUNCH_CHECK();
    if (fft_with(twn, d_T, +1, fs[f])) return -1;
    extract_taps_kernel<<<grid_for(outs[f]->w), kT, 0, fs[f]>>>(d_T, outs[f]->w, logn, outs[f]->time);
    SFFTB_LAUNCH_CHECK();
    if (filter_refresh(outs[f], fs[f])) return -1;
    SFFTB_CUDA(cudaStreamSynchronize(fs[f]));
    cudaFree(d_T);
    guard.forget(d_T);
  }
  clk.mark("normalise, inverse FFT, taps");
  for (int f = 0; f < count; f++) {
    cudaStreamDestroy(fs[f]);
    cudaFree(d_H[f]); cudaFree(d_cand[f]); cudaFree(d_maxq[f]); cudaFree(d_ncand[f]);
  }
  for (int q = 0; q < nwin; q++) {
    cudaStreamDestroy(win[q].ramp_stream);
    cudaFree(win[q].d_G); cudaFree(win[q].d_R);
  }
  free_twiddles(&twn);
  guard.dismiss();
  clk.mark("free");
  return 0;
}

int build_filter(int logn, double lobefrac, double tolerance, int b, int fw_half, DeviceFilter *out,
                 cudaStream_t st)
{
  FilterSpec spec;
  spec.lobefrac = lobefrac; spec.tolerance = tolerance; spec.b = b; spec.fw_half = fw_half;
  DeviceFilter *outs[1] = {out};
  return build_filters(logn, 1, &spec, outs, st);
}

int filter_refresh(DeviceFilter *f, cudaStream_t st)
{
  const int len = 2 * f->fw_half + 1;
  if (!f->fdr) SFFTB_CUDA(cudaMalloc(&f->fdr, sizeof(double2) * len));
  return launch_filter_den(f->fwin, len, f->fdr, st);
}

void free_filter(DeviceFilter *f)
{
  if (f->time) cudaFree(f->time);
  if (f->fwin) cudaFree(f->fwin);
  if (f->fdr) cudaFree(f->fdr);
  f->time = nullptr;
  f->fwin = nullptr;
  f->fdr = nullptr;
}

}  // namespace sfftb
