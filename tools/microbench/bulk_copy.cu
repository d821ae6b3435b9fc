// bulk_copy.cu -- how fast can an SM pull L2-resident runs into shared memory?
//
// The v2 estimation kernel stages, per tile, L contiguous runs of a 5 MB (L2-resident)
// array into shared memory.  This measures that data path in isolation, for run sizes
// 1..16 KB, with (a) cp.async.bulk (TMA engine, one elected thread per run) and
// (b) plain LDG.128 + STS.128 by all threads, at 1 or 2 CTAs per SM.
//
//   ./bulk_copy            -> one JSON line per configuration
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash(unsigned x)
{
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// every CTA: `iters` rounds of `runs` copies of `bytes` each, double-buffered
__global__ void tma_kernel(const char *src, unsigned src_runs, int bytes, int runs, int iters, unsigned long long *sink)
{
  extern __shared__ __align__(128) char smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  const unsigned b0 = (unsigned)__cvta_generic_to_shared(&bar[0]);
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0 + 8 * i), "r"(runs));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned seed = blockIdx.x * 7919u;
  auto issue = [&](int it) {
    const int st = it & 1;
    if ((int)threadIdx.x < runs) {
      const unsigned run = hash(seed + it * 131u + threadIdx.x) % src_runs;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0 + 8 * st), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(s0 + (unsigned)((st * runs + threadIdx.x) * bytes)), "l"(src + (size_t)run * bytes), "r"(bytes), "r"(b0 + 8 * st) : "memory");
    }
  };
  issue(0);
  issue(1);
  unsigned long long acc = 0;
  for (int it = 0; it < iters; it++) {
    const int st = it & 1;
    const unsigned parity = (it >> 1) & 1;
    unsigned done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b0 + 8 * st), "r"(parity) : "memory");
    } while (!done);
    acc += *reinterpret_cast<unsigned long long *>(smem + (size_t)st * runs * bytes + (threadIdx.x * 8) % (runs * bytes));
    __syncthreads();
    if (it + 2 < iters) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(it + 2);
    }
  }
  if (acc == 0x1234567ull) sink[0] = acc;
}

__global__ void ldg_kernel(const char *src, unsigned src_runs, int bytes, int runs, int iters, unsigned long long *sink)
{
  extern __shared__ __align__(128) char smem[];
  unsigned seed = blockIdx.x * 7919u;
  unsigned long long acc = 0;
  const int per_run = bytes / 16;
  for (int it = 0; it < iters; it++) {
    for (int e = threadIdx.x; e < runs * per_run; e += blockDim.x) {
      const int r = e / per_run, k = e - r * per_run;
      const unsigned run = hash(seed + it * 131u + r) % src_runs;
      const int4 v = *reinterpret_cast<const int4 *>(src + (size_t)run * bytes + (size_t)k * 16);
      *reinterpret_cast<int4 *>(smem + (size_t)e * 16) = v;
    }
    __syncthreads();
    acc += *reinterpret_cast<unsigned long long *>(smem + (threadIdx.x * 8) % (runs * bytes));
    __syncthreads();
  }
  if (acc == 0x1234567ull) sink[0] = acc;
}

int main()
{
  const size_t src_bytes = 5u << 20;
  char *src;
  unsigned long long *sink;
  CK(cudaMalloc(&src, src_bytes));
  CK(cudaMemset(src, 1, src_bytes));
  CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(ldg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int mode = 0; mode < 2; mode++)
    for (int per_sm = 1; per_sm <= 2; per_sm++)
      for (int bytes = 1024; bytes <= 16384; bytes *= 2) {
        // keep ~40 KB (x2 stages for TMA) in flight per CTA, like 20 runs of a T=128 tile
        int runs = 40960 / bytes;
        if (runs < 2) runs = 2;
        if (runs > 32) runs = 32;
        const size_t smem = (size_t)(mode == 0 ? 2 : 1) * runs * bytes;
        if (smem * per_sm > 200 * 1024) continue;
        const int iters = 2000;
        const int threads = 256;
        const dim3 grid(sms * per_sm);
        for (int rep = 0; rep < 2; rep++) {
          CK(cudaEventRecord(e0));
          if (mode == 0) tma_kernel<<<grid, threads, smem>>>(src, (unsigned)(src_bytes / bytes), bytes, runs, iters, sink);
          else ldg_kernel<<<grid, threads, smem>>>(src, (unsigned)(src_bytes / bytes), bytes, runs, iters, sink);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
        }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double total = (double)grid.x * iters * runs * bytes;
        printf("{\"mode\": \"%s\", \"ctas_per_sm\": %d, \"run_bytes\": %d, \"runs_per_round\": %d, \"ms\": %.3f, \"TB_per_s\": %.3f, "
               "\"bytes_per_clk_per_sm\": %.1f, \"cycles_per_run_per_sm\": %.1f}\n",
               mode == 0 ? "cp.async.bulk" : "ldg+sts", per_sm, bytes, runs, ms, total / ms * 1e-9,
               total / sms / (ms * 1e-3 * 1.9e9), (ms * 1e-3 * 1.9e9) / ((double)per_sm * iters * runs));
      }
  return 0;
}
