// cluster_barrier.cu -- what does one team barrier of the v3 peeling kernel cost?
// A thread-block cluster of C CTAs x 1024 threads runs `iters` barriers:
//   acqrel : barrier.cluster.arrive.release + wait.acquire (what v3.cu's Team::sync uses; ptxas
//            adds MEMBAR.ALL.GPU before and CCTL.IVALL after)
//   relaxed: barrier.cluster.arrive.relaxed + wait (no memory ordering)
// each with and without one global store + one dependent global load per thread per iteration.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <bool ACQREL, bool TRAFFIC>
__global__ void __launch_bounds__(1024) bar(int iters, int *buf, long long *cycles)
{
  const int gt = blockIdx.x * 1024 + threadIdx.x, nt = gridDim.x * 1024;
  int acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (TRAFFIC) buf[gt] = it + acc;
    if (ACQREL) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
    if (TRAFFIC) acc += buf[(gt + 1024) % nt];
  }
  const long long t1 = clock64();
  if (acc == 0x7fffffff) buf[0] = acc;
  if (gt == 0) cycles[0] = t1 - t0;
}

template <bool ACQREL, bool TRAFFIC>
void run(const char *name, int C, int *buf, long long *cyc)
{
  const int iters = 2000;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(C); cfg.blockDim = dim3(1024);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (C > 8) CK(cudaFuncSetAttribute(bar<ACQREL, TRAFFIC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int rep = 0; rep < 2; rep++) CK(cudaLaunchKernelEx(&cfg, bar<ACQREL, TRAFFIC>, iters, buf, cyc));
  CK(cudaDeviceSynchronize());
  long long hc; CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
  printf("{\"barrier\": \"%s\", \"ctas\": %d, \"cycles_per_barrier\": %.1f, \"us_at_1965MHz\": %.3f}\n", name, C,
         (double)hc / iters, (double)hc / iters / 1965.0);
}

int main()
{
  int *buf; long long *cyc;
  CK(cudaMalloc(&buf, 4 * 16 * 1024)); CK(cudaMalloc(&cyc, 64));
  CK(cudaMemset(buf, 0, 4 * 16 * 1024));
  const int sizes[4] = {1, 4, 8, 16};
  for (int i = 0; i < 4; i++) {
    run<true, false>("acqrel", sizes[i], buf, cyc);
    run<false, false>("relaxed", sizes[i], buf, cyc);
    run<true, true>("acqrel+store+load", sizes[i], buf, cyc);
    run<false, true>("relaxed+store+load (racy)", sizes[i], buf, cyc);
  }
  return 0;
}
