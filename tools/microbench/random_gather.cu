// random_gather.cu -- what HBM bandwidth can a 16-byte random gather reach on B200?
//
// The permuted-filter gather of the sparse FFT reads x[(i*ai) mod n] for i < w: one
// 16-byte sample per 32-byte sector, at an odd random stride.  This microbenchmark
// measures, for footprints n*16 B of 64 MiB .. 2 GiB:
//   stream   : plain coalesced read of the whole array          (GB/s)
//   gather   : `count` reads at (i*a) mod n, U independent loads per thread in flight
//   sorted   : the same number of reads at ascending addresses, equally spaced
//              (what a position-sorted traversal of the same sample set would do)
// Output: one JSON line per (footprint, mode).  "useful" counts 16 B per read,
// "sector" counts 32 B per read.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double2 ldg_stream(const double2 *p)
{
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

template <int U>
__global__ void gather_kernel(const double2 *__restrict__ x, unsigned mask, unsigned a, long long count,
                              double2 *sink)
{
  double sr = 0, si = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride * U) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long ii = i + u * stride;
      const unsigned idx = (unsigned)(((unsigned long long)(ii < count ? ii : i) * a) & mask);
      v[u] = ldg_stream(x + idx);
    }
#pragma unroll
    for (int u = 0; u < U; u++) { sr += v[u].x; si += v[u].y; }
  }
  if (sr == 1.2345 && si == 5.4321) sink[0] = make_double2(sr, si);
}

template <int U>
__global__ void sorted_kernel(const double2 *__restrict__ x, unsigned mask, unsigned spacing, long long count,
                              double2 *sink)
{
  double sr = 0, si = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += stride * U) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long long ii = i + u * stride;
      const unsigned idx = (unsigned)(((unsigned long long)(ii < count ? ii : i) * spacing) & mask);
      v[u] = ldg_stream(x + idx);
    }
#pragma unroll
    for (int u = 0; u < U; u++) { sr += v[u].x; si += v[u].y; }
  }
  if (sr == 1.2345 && si == 5.4321) sink[0] = make_double2(sr, si);
}

__global__ void stream_kernel(const double2 *__restrict__ x, long long n, double2 *sink)
{
  double sr = 0, si = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double2 v = ldg_stream(x + i);
    sr += v.x; si += v.y;
  }
  if (sr == 1.2345 && si == 5.4321) sink[0] = make_double2(sr, si);
}

__global__ void fill_kernel(double2 *x, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = make_double2((double)(i & 1023), 1.0);
}

int main(int argc, char **argv)
{
  // optional: L2 fetch granularity hint in bytes (32, 64 or 128)
  if (argc > 1) {
    size_t g = (size_t)atoi(argv[1]);
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g));
  }
  size_t gran = 0;
  CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
  printf("{\"l2_fetch_granularity\": %zu}\n", gran);
  const int logs[] = {22, 24, 26, 27};
  double2 *sink; CK(cudaMalloc(&sink, 64));
  char *flush; const size_t flush_bytes = 512ull << 20; CK(cudaMalloc(&flush, flush_bytes));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int li = 0; li < 4; li++) {
    const int logn = logs[li];
    const long long n = 1ll << logn;
    double2 *x; CK(cudaMalloc(&x, sizeof(double2) * n));
    fill_kernel<<<148 * 8, 256>>>(x, n);
    const long long count = 12ll << 20;            // ~ config 4's 12.1 M samples
    const unsigned a = 0x9E3779B1u | 1u;           // odd multiplier
    const unsigned spacing = (unsigned)(n / count > 0 ? n / count : 1) | 1u;
    for (int mode = 0; mode < 5; mode++) {
      float best = 1e30f;
      for (int rep = 0; rep < 5; rep++) {
        CK(cudaMemsetAsync(flush, rep, flush_bytes));
        CK(cudaEventRecord(e0));
        if (mode == 0) stream_kernel<<<148 * 16, 256>>>(x, n, sink);
        else if (mode == 1) gather_kernel<1><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink);
        else if (mode == 2) gather_kernel<8><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink);
        else if (mode == 3) gather_kernel<16><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink);
        else sorted_kernel<8><<<148 * 8, 256>>>(x, (unsigned)(n - 1), spacing, count, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
      }
      const char *names[] = {"stream", "gather_u1", "gather_u8", "gather_u16", "sorted_u8"};
      const double reads = mode == 0 ? (double)n : (double)count;
      const double useful = reads * 16 / (best * 1e-3) / 1e9;
      const double sector = mode == 0 ? useful : reads * 32 / (best * 1e-3) / 1e9;
      printf("{\"footprint_MiB\": %lld, \"mode\": \"%s\", \"reads\": %.0f, \"ms\": %.4f, \"useful_GBs\": %.1f, \"sector_GBs\": %.1f}\n",
             (n * 16) >> 20, names[mode], reads, best, useful, sector);
    }
    CK(cudaFree(x));
  }
  return 0;
}
