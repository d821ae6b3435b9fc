// ld_variants.cu -- how many DRAM bytes does one random 16-byte read cost on B200,
// per flavour of the load instruction?  Run under
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./ld_variants <log2 n> <count>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int V> __device__ __forceinline__ double2 ld(const double2 *p)
{
  double2 r;
  if (V == 0) asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  if (V == 1) asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  if (V == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  if (V == 3) asm volatile("ld.global.cv.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  if (V == 4) asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  if (V == 5) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  if (V == 6) {  // two 8-byte halves
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r.x) : "l"(p));
    asm volatile("ld.global.nc.f64 %0, [%1+8];" : "=d"(r.y) : "l"(p));
  }
  return r;
}

template <int V>
__global__ void gather(const double2 *__restrict__ x, unsigned mask, unsigned a, long long count, double2 *sink)
{
  double sr = 0, si = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const unsigned idx = (unsigned)(((unsigned long long)i * a) & mask);
    const double2 v = ld<V>(x + idx);
    sr += v.x; si += v.y;
  }
  if (sr == 1.2345 && si == 5.4321) sink[0] = make_double2(sr, si);
}

__global__ void fill(double2 *x, long long n)
{
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = make_double2((double)(i & 1023), 1.0);
}

int main(int argc, char **argv)
{
  const int logn = argc > 1 ? atoi(argv[1]) : 27;
  const long long count = argc > 2 ? atoll(argv[2]) : (12ll << 20);
  const long long n = 1ll << logn;
  double2 *x, *sink; CK(cudaMalloc(&x, sizeof(double2) * n)); CK(cudaMalloc(&sink, 64));
  char *flush; CK(cudaMalloc(&flush, 512ull << 20));
  fill<<<148 * 8, 256>>>(x, n);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const unsigned a = 0x9E3779B1u;
  for (int v = 0; v < 7; v++) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaMemsetAsync(flush, rep, 512ull << 20));
      CK(cudaEventRecord(e0));
      switch (v) {
        case 0: gather<0><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink); break;
        case 1: gather<1><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink); break;
        case 2: gather<2><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink); break;
        case 3: gather<3><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink); break;
        case 4: gather<4><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink); break;
        case 5: gather<5><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink); break;
        case 6: gather<6><<<148 * 8, 256>>>(x, (unsigned)(n - 1), a, count, sink); break;
      }
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    printf("{\"variant\": %d, \"footprint_MiB\": %lld, \"reads\": %lld, \"ms\": %.4f, \"sector_GBs\": %.1f}\n", v,
           (n * 16) >> 20, count, best, count * 32.0 / (best * 1e-3) / 1e9);
  }
  return 0;
}
