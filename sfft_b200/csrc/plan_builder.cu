// plan_builder.cu -- window/filter construction on the device.
//
// Reference: make_dolphchebyshev_t (src/filters.cc:70-86) and make_multiple_t
// (src/filters.cc:109-160), which spend their time in one odd-length w-point DFT
// and two n-point DFTs per filter (seconds to minutes on a CPU at n >= 2^26).
// Here the host only evaluates the O(w) libm seeds (cheb_host.c); the w-point DFT
// (Bluestein over power-of-two FFTs), both n-point FFTs, the boxcar (prefix scan),
// the peak search, the phase ramp and the tap extraction run on the GPU.
#include "plan_builder.cuh"

#include <math.h>
#include <vector>

#include "fft.cuh"
#include "v12_kernels.cuh"

extern "C" {
int sfftb_host_dolph_width(double lobefrac, double tolerance);
void sfftb_host_cheb_samples(double tolerance, int w, double *out);
void sfftb_host_ramp_step(int w, int n, double *re, double *im);
}

namespace sfftb {

namespace {

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

// ---- Bluestein pieces (forward transform, sign -1) -------------------------
__device__ __forceinline__ cplx chirp(long long j, int w)
{
  // e^{-pi i j^2 / w}
  const long long q = (j * j) % (2ll * w);
  double sn, cs;
  sincospi((double)q / (double)w, &sn, &cs);
  return make_double2(cs, -sn);
}

__global__ void bluestein_prep_kernel(const cplx *__restrict__ x, int w, int logM, cplx *A,
                                      cplx *Bk)
{
  const long long M = 1ll << logM;
  GRID_STRIDE(j, M)
  {
    cplx a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
    if (j < w) {
      const cplx ch = chirp(j, w);
      const cplx xv = x[j];
      a = make_double2(xv.x * ch.x - xv.y * ch.y, xv.x * ch.y + xv.y * ch.x);
      b = make_double2(ch.x, -ch.y);
    } else if (M - j < w) {
      const cplx ch = chirp(M - j, w);
      b = make_double2(ch.x, -ch.y);
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
  GRID_STRIDE(j, M)
  {
    const cplx a = A[j], b = Bk[j];
    out[bitrev((unsigned)j, logM)] = make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
  }
}

// X[k] = (C[k]/M) * chirp[k]
__global__ void bluestein_finish_kernel(const cplx *__restrict__ C, int w, int logM, cplx *X)
{
  const double inv = 1.0 / (double)(1ll << logM);
  GRID_STRIDE(k, w)
  {
    const cplx ch = chirp(k, w);
    const double r = C[k].x * inv, i = C[k].y * inv;
    X[k] = make_double2(r * ch.x - i * ch.y, r * ch.y + i * ch.x);
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

// ---- prefix sums of a complex array: P[i] = sum_{j<i} g[j], P[n] = total ----
constexpr int kScanTile = 2048;

__global__ void __launch_bounds__(256) scan_block_sums_kernel(const cplx *__restrict__ g, long long n,
                                                              cplx *bs)
{
  __shared__ double sr[256], si[256];
  const long long base = (long long)blockIdx.x * kScanTile;
  double ar = 0, ai = 0;
  for (int e = threadIdx.x; e < kScanTile; e += 256) {
    const long long i = base + e;
    if (i < n) { ar += g[i].x; ai += g[i].y; }
  }
  sr[threadIdx.x] = ar; si[threadIdx.x] = ai;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sr[threadIdx.x] += sr[threadIdx.x + o]; si[threadIdx.x] += si[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) bs[blockIdx.x] = make_double2(sr[0], si[0]);
}

// exclusive scan of the block sums, one CTA
__global__ void __launch_bounds__(1024) scan_of_sums_kernel(cplx *bs, int nb)
{
  __shared__ double sr[1024], si[1024];
  const int per = (nb + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(lo + per, nb);
  double ar = 0, ai = 0;
  for (int i = lo; i < hi; i++) { ar += bs[i].x; ai += bs[i].y; }
  sr[threadIdx.x] = ar; si[threadIdx.x] = ai;
  __syncthreads();
  // Hillis-Steele inclusive scan
  for (int o = 1; o < 1024; o <<= 1) {
    double tr = 0, ti = 0;
    if (threadIdx.x >= o) { tr = sr[threadIdx.x - o]; ti = si[threadIdx.x - o]; }
    __syncthreads();
    sr[threadIdx.x] += tr; si[threadIdx.x] += ti;
    __syncthreads();
  }
  double br = sr[threadIdx.x] - ar, bi = si[threadIdx.x] - ai;   // exclusive
  for (int i = lo; i < hi; i++) {
    const cplx v = bs[i];
    bs[i] = make_double2(br, bi);
    br += v.x; bi += v.y;
  }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(const cplx *__restrict__ g, long long n,
                                                         const cplx *__restrict__ bs, cplx *P)
{
  // each thread owns 8 consecutive elements of the 2048-tile
  __shared__ double sr[256], si[256];
  const long long base = (long long)blockIdx.x * kScanTile + threadIdx.x * 8;
  cplx v[8];
  double ar = 0, ai = 0;
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const long long i = base + e;
    v[e] = i < n ? g[i] : make_double2(0.0, 0.0);
    ar += v[e].x; ai += v[e].y;
  }
  sr[threadIdx.x] = ar; si[threadIdx.x] = ai;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {
    double tr = 0, ti = 0;
    if (threadIdx.x >= o) { tr = sr[threadIdx.x - o]; ti = si[threadIdx.x - o]; }
    __syncthreads();
    sr[threadIdx.x] += tr; si[threadIdx.x] += ti;
    __syncthreads();
  }
  double pr = bs[blockIdx.x].x + (sr[threadIdx.x] - ar);
  double pi = bs[blockIdx.x].y + (si[threadIdx.x] - ai);
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const long long i = base + e;
    if (i < n) P[i] = make_double2(pr, pi);
    pr += v[e].x; pi += v[e].y;
    if (i == n - 1) P[n] = make_double2(pr, pi);
  }
}

// hraw[(i + b/2) % n] = sum_{j=i}^{i+b-1} g[j mod n]; peak = max |hraw|   (filters.cc:116-130)
__global__ void boxcar_kernel(const cplx *__restrict__ P, int logn, int b, cplx *H,
                              unsigned long long *peak_bits)
{
  const long long n = 1ll << logn;
  double local = 0.0;
  GRID_STRIDE(i, n)
  {
    const long long e = i + b;
    cplx s;
    if (e <= n) {
      s = make_double2(P[e].x - P[i].x, P[e].y - P[i].y);
    } else {
      s = make_double2((P[n].x - P[i].x) + P[e - n].x, (P[n].y - P[i].y) + P[e - n].y);
    }
    H[(i + b / 2) & (n - 1)] = s;
    const double m = hypot(s.x, s.y);
    local = m > local ? m : local;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double other = __shfl_xor_sync(0xffffffffu, local, o);
    local = other > local ? other : local;
  }
  if ((threadIdx.x & 31) == 0)
    atomicMax(peak_bits, (unsigned long long)__double_as_longlong(local));
}

// h[i] = (h[i]/peak) * step^i with step^i = T_hi[i >> lo_bits] * T_lo[i & mask]  (filters.cc:131-140)
__global__ void normalise_ramp_kernel(cplx *H, int logn, const unsigned long long *peak_bits,
                                      const cplx *__restrict__ t_lo, const cplx *__restrict__ t_hi,
                                      int lo_bits)
{
  const long long n = 1ll << logn;
  const double peak = __longlong_as_double((long long)*peak_bits);
  GRID_STRIDE(i, n)
  {
    cplx h = H[i];
    h = make_double2(__ddiv_rn(h.x, peak), __ddiv_rn(h.y, peak));
    const cplx lo = t_lo[i & ((1ll << lo_bits) - 1)];
    const cplx hi = t_hi[i >> lo_bits];
    const cplx ramp = cmul_rn(hi, lo);
    H[i] = cmul_rn(h, ramp);
  }
}

// fwin[m] = h[(m - half) mod n], m in [0, 2*half]
__global__ void freq_window_kernel(const cplx *__restrict__ H, int logn, int half, cplx *fwin)
{
  const long long n = 1ll << logn;
  GRID_STRIDE(m, 2ll * half + 1) { fwin[m] = H[(m - half + n) & (n - 1)]; }
}

__global__ void extract_taps_kernel(const cplx *__restrict__ G, int w, int logn, cplx *taps)
{
  const double nn = (double)(1ll << logn);
  GRID_STRIDE(i, w) { taps[i] = make_double2(__ddiv_rn(G[i].x, nn), __ddiv_rn(G[i].y, nn)); }
}

}  // namespace

// forward DFT of arbitrary length w (device in/out), Bluestein over 2^logM-point FFTs
int bluestein_forward(const cplx *d_x, int w, cplx *d_out, cudaStream_t st)
{
  int logM = 0;
  while ((1ll << logM) < 2ll * w - 1) logM++;
  const long long M = 1ll << logM;
  cplx *d_A = nullptr, *d_B = nullptr, *d_C = nullptr;
  SFFTB_CUDA(cudaMalloc(&d_A, sizeof(cplx) * M));
  SFFTB_CUDA(cudaMalloc(&d_B, sizeof(cplx) * M));
  SFFTB_CUDA(cudaMalloc(&d_C, sizeof(cplx) * M));
  bluestein_prep_kernel<<<grid_for(M), kT, 0, st>>>(d_x, w, logM, d_A, d_B);
  SFFTB_LAUNCH_CHECK();
  if (fft_dit_inplace(d_A, logM, 1, M, 1, M, nullptr, 0, -1, st)) return -1;
  if (fft_dit_inplace(d_B, logM, 1, M, 1, M, nullptr, 0, -1, st)) return -1;
  pointwise_mul_bitrev_kernel<<<grid_for(M), kT, 0, st>>>(d_A, d_B, logM, d_C);
  SFFTB_LAUNCH_CHECK();
  if (fft_dit_inplace(d_C, logM, 1, M, 1, M, nullptr, 0, +1, st)) return -1;
  bluestein_finish_kernel<<<grid_for(w), kT, 0, st>>>(d_C, w, logM, d_out);
  SFFTB_LAUNCH_CHECK();
  SFFTB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_A); cudaFree(d_B); cudaFree(d_C);
  return 0;
}

int filter_width(double lobefrac, double tolerance)
{
  return sfftb_host_dolph_width(lobefrac, tolerance);
}

int build_filter(int logn, double lobefrac, double tolerance, int b, int fw_half, DeviceFilter *out,
                 cudaStream_t st)
{
  const long long n = 1ll << logn;
  const int w = sfftb_host_dolph_width(lobefrac, tolerance);
  if (w < 1 || w > n || b > n || b < 1) {
    set_error("build_filter: window does not fit the signal length (reference asserts, filters.cc:111-112)");
    return -1;
  }
  out->w = w;
  out->fw_half = fw_half;

  // ---- host seeds ----
  std::vector<double> samples_re((size_t)w);
  sfftb_host_cheb_samples(tolerance, w, samples_re.data());
  std::vector<cplx> samples((size_t)w);
  for (int i = 0; i < w; i++) samples[(size_t)i] = make_double2(samples_re[(size_t)i], 0.0);
  double step_re, step_im;
  sfftb_host_ramp_step(w, (int)n, &step_re, &step_im);
  const int lo_bits = logn < 14 ? logn : 14;
  const long long n_lo = 1ll << lo_bits, n_hi = n >> lo_bits;
  std::vector<cplx> t_lo((size_t)n_lo), t_hi((size_t)n_hi);
  {
    // running product exactly as filters.cc:136-139 for the first 2^lo_bits steps,
    // then the same recurrence on the 2^lo_bits-th power
    double cr = 1, ci = 0;
    for (long long i = 0; i < n_lo; i++) {
      t_lo[(size_t)i] = make_double2(cr, ci);
      const double nr = cr * step_re - ci * step_im;
      const double ni = cr * step_im + ci * step_re;
      cr = nr; ci = ni;
    }
    const double br = cr, bi = ci;   // step^(2^lo_bits)
    cr = 1; ci = 0;
    for (long long m = 0; m < n_hi; m++) {
      t_hi[(size_t)m] = make_double2(cr, ci);
      const double nr = cr * br - ci * bi;
      const double ni = cr * bi + ci * br;
      cr = nr; ci = ni;
    }
  }

  // ---- device buffers ----
  cplx *d_samples = nullptr, *d_X = nullptr;
  double *d_taps0 = nullptr;
  cplx *d_G = nullptr, *d_P = nullptr, *d_H = nullptr, *d_bs = nullptr, *d_tlo = nullptr, *d_thi = nullptr;
  unsigned long long *d_peak = nullptr;
  const int nb = (int)((n + kScanTile - 1) / kScanTile);
  SFFTB_CUDA(cudaMalloc(&d_samples, sizeof(cplx) * w));
  SFFTB_CUDA(cudaMalloc(&d_X, sizeof(cplx) * w));
  SFFTB_CUDA(cudaMalloc(&d_taps0, sizeof(double) * w));
  SFFTB_CUDA(cudaMalloc(&d_G, sizeof(cplx) * n));
  SFFTB_CUDA(cudaMalloc(&d_P, sizeof(cplx) * (n + 1)));
  SFFTB_CUDA(cudaMalloc(&d_H, sizeof(cplx) * n));
  SFFTB_CUDA(cudaMalloc(&d_bs, sizeof(cplx) * nb));
  SFFTB_CUDA(cudaMalloc(&d_tlo, sizeof(cplx) * n_lo));
  SFFTB_CUDA(cudaMalloc(&d_thi, sizeof(cplx) * n_hi));
  SFFTB_CUDA(cudaMalloc(&d_peak, sizeof(unsigned long long)));
  SFFTB_CUDA(cudaMalloc(&out->time, sizeof(cplx) * w));
  SFFTB_CUDA(cudaMalloc(&out->fwin, sizeof(cplx) * (2ll * fw_half + 1)));
  SFFTB_CUDA(cudaMemcpyAsync(d_samples, samples.data(), sizeof(cplx) * w, cudaMemcpyHostToDevice, st));
  SFFTB_CUDA(cudaMemcpyAsync(d_tlo, t_lo.data(), sizeof(cplx) * n_lo, cudaMemcpyHostToDevice, st));
  SFFTB_CUDA(cudaMemcpyAsync(d_thi, t_hi.data(), sizeof(cplx) * n_hi, cudaMemcpyHostToDevice, st));
  SFFTB_CUDA(cudaMemsetAsync(d_peak, 0, sizeof(unsigned long long), st));

  // ---- w-point DFT by Bluestein (filters.cc:81), rotate, keep the real part ----
  if (bluestein_forward(d_samples, w, d_X, st)) return -1;
  rotate_real_kernel<<<grid_for(w), kT, 0, st>>>(d_X, w, d_taps0);
  SFFTB_LAUNCH_CHECK();

  // ---- make_multiple_t (filters.cc:109-160) ----
  SFFTB_CUDA(cudaMemsetAsync(d_G, 0, sizeof(cplx) * n, st));
  centre_scatter_kernel<<<grid_for(w), kT, 0, st>>>(d_taps0, w, logn, d_G);
  SFFTB_LAUNCH_CHECK();
  if (fft_dit_inplace(d_G, logn, 1, n, 1, n, nullptr, 0, -1, st)) return -1;
  scan_block_sums_kernel<<<nb, 256, 0, st>>>(d_G, n, d_bs);
  SFFTB_LAUNCH_CHECK();
  scan_of_sums_kernel<<<1, 1024, 0, st>>>(d_bs, nb);
  SFFTB_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, 256, 0, st>>>(d_G, n, d_bs, d_P);
  SFFTB_LAUNCH_CHECK();
  boxcar_kernel<<<grid_for(n), kT, 0, st>>>(d_P, logn, b, d_H, d_peak);
  SFFTB_LAUNCH_CHECK();
  normalise_ramp_kernel<<<grid_for(n), kT, 0, st>>>(d_H, logn, d_peak, d_tlo, d_thi, lo_bits);
  SFFTB_LAUNCH_CHECK();
  freq_window_kernel<<<grid_for(2ll * fw_half + 1), kT, 0, st>>>(d_H, logn, fw_half, out->fwin);
  SFFTB_LAUNCH_CHECK();
  if (bitrev_permute(d_H, d_G, logn, st)) return -1;
  if (fft_dit_inplace(d_G, logn, 1, n, 1, n, nullptr, 0, +1, st)) return -1;
  extract_taps_kernel<<<grid_for(w), kT, 0, st>>>(d_G, w, logn, out->time);
  SFFTB_LAUNCH_CHECK();
  if (filter_refresh(out, st)) return -1;
  SFFTB_CUDA(cudaStreamSynchronize(st));

  cudaFree(d_samples); cudaFree(d_taps0); cudaFree(d_X);
  cudaFree(d_G); cudaFree(d_P); cudaFree(d_H); cudaFree(d_bs); cudaFree(d_tlo); cudaFree(d_thi);
  cudaFree(d_peak);
  return 0;
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
