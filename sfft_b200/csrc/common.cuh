// common.cuh -- shared device/host helpers for the sm_100a sparse-FFT engine.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

namespace sfftb {

typedef double2 cplx;   // (x = re, y = im), 16-byte aligned -> LDG.E.128 / STG.E.128

// ---- error plumbing --------------------------------------------------------
void set_error(const std::string &msg);
extern std::atomic<long long> g_launches;   // kernels launched by this library (any thread)

#define SFFTB_CUDA(call)                                                            \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      char buf__[512];                                                              \
      snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,   \
               cudaGetErrorString(e__));                                            \
      ::sfftb::set_error(buf__);                                                    \
      return -1;                                                                    \
    }                                                                               \
  } while (0)

#define SFFTB_LAUNCH_CHECK()                                                        \
  do {                                                                              \
    ::sfftb::g_launches++;                                                          \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      char buf__[512];                                                              \
      snprintf(buf__, sizeof buf__, "%s:%d: kernel launch -> %s", __FILE__,         \
               __LINE__, cudaGetErrorString(e__));                                  \
      ::sfftb::set_error(buf__);                                                    \
      return -1;                                                                    \
    }                                                                               \
  } while (0)

// Run `body` (kernel attribute setup) the first time this call site is reached on each
// device: function attributes are per device, and a process may hold plans on several.
#define SFFTB_ONCE_PER_DEVICE(body)                                                 \
  do {                                                                              \
    static std::atomic<unsigned long long> done__{0};                               \
    int dev__ = 0;                                                                  \
    cudaGetDevice(&dev__);                                                          \
    const unsigned long long bit__ = 1ull << (dev__ & 63);                          \
    if (!(done__.load(std::memory_order_acquire) & bit__)) {                        \
      body;                                                                         \
      done__.fetch_or(bit__, std::memory_order_release);                            \
    }                                                                               \
  } while (0)

// Device temporaries of a host function that can return early: everything `track`ed is freed /
// destroyed when the guard goes out of scope, unless the normal path (which frees in its own
// order) has called dismiss().
struct ScratchGuard {
  void *dev[32];
  cudaStream_t streams[8];
  int ndev = 0, nstreams = 0;
  bool armed = true;
  template <class T> T *track(T *p) { if (p && ndev < 32) dev[ndev++] = (void *)p; return p; }
  void track_stream(cudaStream_t s) { if (s && nstreams < 8) streams[nstreams++] = s; }
  void forget(const void *p) { for (int i = 0; i < ndev; i++) if (dev[i] == p) dev[i] = nullptr; }
  void dismiss() { armed = false; }
  ~ScratchGuard()
  {
    if (!armed) return;
    for (int i = 0; i < nstreams; i++) cudaStreamSynchronize(streams[i]);
    for (int i = 0; i < ndev; i++) if (dev[i]) cudaFree(dev[i]);
    for (int i = 0; i < nstreams; i++) cudaStreamDestroy(streams[i]);
  }
};

// ---- exactly-rounded complex arithmetic ------------------------------------
// The parity contract is "one IEEE rounding per product and per sum", the same
// as the reference's SSE2 mul/hadd sequences (cf12.cc:243-256).  The intrinsics
// below are never contracted into FMAs, whatever -fmad says.
__device__ __forceinline__ cplx cmul_rn(cplx a, cplx b)
{
  const double p0 = __dmul_rn(a.x, b.x), p1 = __dmul_rn(a.y, b.y);
  const double p2 = __dmul_rn(a.x, b.y), p3 = __dmul_rn(a.y, b.x);
  return make_double2(__dsub_rn(p0, p1), __dadd_rn(p2, p3));
}
__device__ __forceinline__ cplx cadd_rn(cplx a, cplx b)
{
  return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));
}
__device__ __forceinline__ cplx csub_rn(cplx a, cplx b)
{
  return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y));
}
__device__ __forceinline__ double cabs2_rn(cplx a)
{
  return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y));
}

// read-only 16-byte load that does not pollute L1 (random gathers of the signal)
__device__ __forceinline__ cplx ldg_stream(const cplx *p)
{
  cplx r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
               : "=d"(r.x), "=d"(r.y)
               : "l"(p));
  return r;
}

// same, asking L2 to fill only 64 bytes around the sample instead of the whole 128-byte
// line: halves DRAM traffic when neighbouring samples will not be wanted soon
// (tools/microbench/ld_variants: 128 B -> 64 B of DRAM per random 16-byte read)
__device__ __forceinline__ cplx ldg_stream64(const cplx *p)
{
  cplx r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.f64 {%0, %1}, [%2];"
               : "=d"(r.x), "=d"(r.y)
               : "l"(p));
  return r;
}

__host__ __device__ __forceinline__ int ilog2(unsigned long long v)
{
  int l = 0;
  while (v > 1) { v >>= 1; l++; }
  return l;
}

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits)
{
  return bits == 0 ? 0u : (__brev(v) >> (32 - bits));
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace sfftb
