// pipe_overlap.cu -- do the ALU pipe (FSEL/LOP3/IADD3), the FP64 pipe (DFMA/DSETP) and the FMA pipe
// (FFMA/IMAD) of an sm_100a SM sub-partition issue concurrently, or do they share issue bandwidth?
// The v2 estimation kernel executes ~540 ALU selects and ~470 FP64 instructions per recovered
// coefficient; ncu shows ALU 56 % + FP64 45 % busy and no speed-up from running two CTAs in
// anti-phase.  This measures the three pipes alone and pairwise: 16 warps per SM (4 per
// sub-partition, as in that kernel), 8 independent dependency chains per thread per pipe.
// Output: one JSON line per mix with cycles per warp-instruction per sub-partition.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kChains = 8;
constexpr int kIters = 4096;

template <int ALU, int F64, int FMA>
__global__ void __launch_bounds__(512) mix(unsigned *out, unsigned seed, long long *cycles)
{
  unsigned a[kChains];
  double d[kChains];
  float f[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) {
    a[c] = seed + threadIdx.x * 7 + c;
    d[c] = 1.0 + 1e-9 * (threadIdx.x + c);
    f[c] = 1.0f + 1e-3f * (threadIdx.x + c);
  }
  const unsigned m = seed | 1u;
  const double dm = 1.0000001;
  const float fm = 1.0001f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int c = 0; c < kChains; c++) {
      if (ALU) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(m), "r"(seed));     // ALU pipe
      if (F64) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(dm));                  // FP64 pipe
      if (FMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[c]) : "f"(fm));                  // FMA pipe
    }
  }
  const long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int c = 0; c < kChains; c++) acc += a[c] + (unsigned)d[c] + (unsigned)f[c];
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// Same instruction counts as mix<1,1,0>, but the pipes are split between warps: even warps run
// only the ALU chains, odd warps only the FP64 chains (each twice as long), as when two CTAs
// of the estimation kernel sit in different phases on one SM.
__global__ void __launch_bounds__(512) split(unsigned *out, unsigned seed, long long *cycles)
{
  unsigned a[kChains];
  double d[kChains];
#pragma unroll
  for (int c = 0; c < kChains; c++) {
    a[c] = seed + threadIdx.x * 7 + c;
    d[c] = 1.0 + 1e-9 * (threadIdx.x + c);
  }
  const unsigned m = seed | 1u;
  const double dm = 1.0000001;
  const bool fp = (threadIdx.x >> 7) & 1;      // warps 0-3 ALU, 4-7 FP64, ...: every sub-partition (warp % 4) gets two of each
  const long long t0 = clock64();
  if (fp) {
#pragma unroll 1
    for (int it = 0; it < 2 * kIters; it++) {
#pragma unroll
      for (int c = 0; c < kChains; c++) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(dm));
    }
  } else {
#pragma unroll 1
    for (int it = 0; it < 2 * kIters; it++) {
#pragma unroll
      for (int c = 0; c < kChains; c++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(m), "r"(seed));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int c = 0; c < kChains; c++) acc += a[c] + (unsigned)d[c];
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int ALU, int F64, int FMA>
void run(const char *name, unsigned *out, long long *cyc)
{
  float best = 1e30f;
  long long hc = 0;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    mix<ALU, F64, FMA><<<148, 512>>>(out, 12345u, cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
  // warp-instructions issued per sub-partition: 4 warps x kIters x kChains x (pipes in the mix)
  const double winst = 4.0 * kIters * kChains * (ALU + F64 + FMA);
  printf("{\"mix\": \"%s\", \"ms\": %.4f, \"cycles\": %lld, \"warp_inst_per_smsp\": %.0f, \"cycles_per_warp_inst\": %.3f}\n",
         name, best, hc, winst, (double)hc / winst);
}

int main()
{
  unsigned *out; long long *cyc;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 64));
  run<1, 0, 0>("alu", out, cyc);
  run<0, 1, 0>("fp64", out, cyc);
  run<0, 0, 1>("fma", out, cyc);
  run<1, 1, 0>("alu+fp64", out, cyc);
  run<1, 0, 1>("alu+fma", out, cyc);
  run<0, 1, 1>("fp64+fma", out, cyc);
  run<1, 1, 1>("alu+fp64+fma", out, cyc);
  {
    long long hc = 0;
    for (int rep = 0; rep < 3; rep++) split<<<148, 512>>>(out, 12345u, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
    const double winst = 4.0 * kIters * kChains * 2;
    printf("{\"mix\": \"alu warps | fp64 warps\", \"cycles\": %lld, \"warp_inst_per_smsp\": %.0f, \"cycles_per_warp_inst\": %.3f}\n",
           hc, winst, (double)hc / winst);
  }
  return 0;
}
