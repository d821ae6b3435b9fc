// v3.cu -- sFFT v3 (exact-sparse) on the device.
//
// Reference: alternate_fft and helpers, src/computefourier-3.0.cc:66-1088; plan in
// src/sfft.cc:506-579.  The reference is a strictly sequential peeling loop on one
// CPU thread.  Here:
//   * the three bucketisations (aliasing "Mansour" subsample, contiguous window,
//     permuted window) are grids over (bucket, shift), followed by the shared
//     tiled FFT passes;
//   * everything after that -- decode every bucket, append to the result, peel the
//     found coefficients out of all three bucket arrays, repeat until the occupied
//     bucket counts stop changing -- is ONE persistent CTA per signal: each round
//     decodes all buckets in parallel (ordered compaction keeps the reference's
//     ascending-bucket order), turns every found coefficient into its 14 bucket
//     deltas in parallel, and applies the deltas per bucket in item order (short
//     linked lists), so the result is deterministic and independent of thread timing.
//
// Bucket arrays are planar here, S[shift][bucket]; the reference interleaves them,
// S[2*bucket + shift] (FFTW stride-2 plans, sfft.cc:434-475).
//
// libm (atan2, sincos, sqrt) is CUDA's, not glibc's: v3 parity is "same locations,
// values to 1e-9", not bit-identity.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fft.cuh"
#include "plan.cuh"

namespace sfftb {

int mod_inverse_pub(int a, int n);
int gcd_pub(int a, int b);

struct PlanV3 {
  int W_Man = 0, B_g1 = 0, B_g2 = 0;
  int logW = 0, logB1 = 0, logB2 = 0;
  DeviceFilter filt[2];            // [0] first (contiguous) window, [1] second (permuted) window
  cplx *d_tw = nullptr;
  int log_twN = 0;
  int cap = 0;                     // signals
  // per-signal scratch
  long long nslots = 0;            // 2*B2 + 2*B1 + 2*W
  int est_cap = 0, ans_cap = 0, tgt_cap = 0, hash_size = 0, log_hash = 0;
  cplx *d_samp = nullptr;          // [cap][nslots]: G2 planes, G1 planes, MAN planes
  int *d_head = nullptr;           // [cap][nslots]
  int *d_est_key = nullptr;        // [cap][est_cap]
  cplx *d_est_val = nullptr;       // [cap][est_cap]
  int *d_t_slot = nullptr;         // [cap][tgt_cap*14], tgt_cap = max(est_cap, ans_cap)
  int *d_t_next = nullptr;
  cplx *d_t_delta = nullptr;
  int *d_hkey = nullptr, *d_hidx = nullptr;   // [cap][hash_size]
  int *d_ans_key = nullptr;        // [cap][ans_cap]
  cplx *d_ans_val = nullptr;       // [cap][ans_cap]
  int *d_count = nullptr;          // [cap]
  int *d_rounds = nullptr;         // [cap]
  int *d_draw = nullptr;           // [cap][8]
  long long *d_prof = nullptr;     // [cap][8] cycle counters of the peeling phases
  int *d_aux = nullptr;            // [cap][aux_ints] team scratch (v3_peel_kernel)
  int flag_words = 0, aux_ints = 0;
  int team = 1;                    // CTAs per signal in the peeling kernel (a thread-block cluster)
  bool team_checked = false;       // the device has been asked whether it co-schedules `team` CTAs
  // CUDA-graph replay of the single-signal transform (as plan_v12.cu:v12_exec_graph)
  cudaGraphExec_t graph_exec = nullptr;
  cudaGraph_t graph = nullptr;
  int *h_gdraw = nullptr;                   // pinned: draw
  unsigned long long *h_gx = nullptr;       // pinned: signal pointer
  unsigned long long *d_gx = nullptr;
  cudaEvent_t g_ev = nullptr;               // staging buffers consumed
  int plain_execs = 0, graph_kernels = 0;
  int *h_draw[kStageSlots] = {nullptr};
  cudaEvent_t ev[kStageSlots] = {nullptr};
  int next_slot = 0;
};

namespace {

struct V3Geom {
  int n, logn, k;
  int W, logW, B1, logB1, B2, logB2;
  int w1, w2;
  int fw_half1, fw_half2;
  long long nslots;
  int est_cap, ans_cap, tgt_cap, hash_size, log_hash;
  int flag_words;          // ceil(max bucket count / 32)
  int aux_ints;            // per-signal ints of the team's scratch: est_cap + flag_words + exchange + counters
};

// draw layout per signal: a, ai, b, shift, init_offset, init_G_offset
enum { D_A = 0, D_AI, D_B, D_SHIFT, D_OFF, D_GOFF, D_INTS = 8 };

__device__ __forceinline__ int g2_base(const V3Geom &g) { (void)g; return 0; }
__device__ __forceinline__ int g1_base(const V3Geom &g) { return 2 * g.B2; }
__device__ __forceinline__ int man_base(const V3Geom &g) { return 2 * g.B2 + 2 * g.B1; }

// ---- bucketisation kernels -------------------------------------------------

// x_man[shift][i] = x[(off + shift + i*sigma) mod n]   (computefourier-3.0.cc:111-121)
// `xi`, when set, is a device slot holding the signal pointer (CUDA-graph replay)
__device__ __forceinline__ const cplx *signal_base(const cplx *x, const unsigned long long *xi)
{
  return xi ? reinterpret_cast<const cplx *>(*xi) : x;
}

__global__ void v3_mansour_kernel(V3Geom g, const cplx *__restrict__ x_direct, const unsigned long long *xi,
                                  long long x_stride, const int *__restrict__ draw, cplx *samp)
{
  const cplx *__restrict__ x = signal_base(x_direct, xi);
  const int s = blockIdx.z, l = blockIdx.y;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)g.W) return;
  const int off = draw[s * D_INTS + D_OFF];
  const unsigned idx = ((unsigned)off + (unsigned)l + (i << (g.logn - g.logW))) & (unsigned)(g.n - 1);
  samp[(long long)s * g.nslots + man_base(g) + l * g.W + bitrev(i, g.logW)] =
      ldg_stream(x + (long long)s * x_stride + idx);
}

// contiguous window: S[l][b] = sum_c x[(G + cB + b + l) mod n] * taps[cB + b], c < floor(w/B)
// (computefourier-3.0.cc:155-202)
__global__ void v3_gauss_kernel(V3Geom g, const cplx *__restrict__ x_direct, const unsigned long long *xi,
                                long long x_stride, const int *__restrict__ draw, const cplx *__restrict__ taps,
                                cplx *samp)
{
  const cplx *__restrict__ x = signal_base(x_direct, xi);
  const int s = blockIdx.z, l = blockIdx.y;
  const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (unsigned)g.B1) return;
  const unsigned G = (unsigned)draw[s * D_INTS + D_GOFF];
  const int chunks = g.w1 / g.B1;
  const cplx *__restrict__ xs = x + (long long)s * x_stride;
  double ar = 0.0, ai = 0.0;
  for (int c = 0; c < chunks; c++) {
    const unsigned t = (unsigned)c * g.B1 + b;
    const cplx v = xs[(G + t + l) & (unsigned)(g.n - 1)];
    const cplx f = __ldg(&taps[t]);
    // (c*a - d*b, d*a + c*b) with c+di the tap   (:179-185)
    ar = __dadd_rn(ar, __dsub_rn(__dmul_rn(f.x, v.x), __dmul_rn(f.y, v.y)));
    ai = __dadd_rn(ai, __dadd_rn(__dmul_rn(f.y, v.x), __dmul_rn(f.x, v.y)));
  }
  samp[(long long)s * g.nslots + g1_base(g) + l * g.B1 + bitrev(b, g.logB1)] = make_double2(ar, ai);
}

__device__ __forceinline__ cplx cpow_int(cplx base, unsigned e)
{
  cplx r = make_double2(1.0, 0.0);
  while (e) {
    if (e & 1u) r = cmul_rn(r, base);
    base = cmul_rn(base, base);
    e >>= 1;
  }
  return r;
}

// permuted window: P[i] = x[(G + i*ai) mod n] * e^{2 pi i (G + i*ai) b / n} (:66-87, the reference
// runs the phase as a running product); S[l][bk] = sum_c P[cB + bk + l] * taps[cB + bk] (:245-287)
__global__ void v3_gauss_perm_kernel(V3Geom g, const cplx *__restrict__ x_direct, const unsigned long long *xi,
                                     long long x_stride, const int *__restrict__ draw,
                                     const cplx *__restrict__ taps, cplx *samp)
{
  const cplx *__restrict__ x = signal_base(x_direct, xi);
  const int s = blockIdx.z, l = blockIdx.y;
  const unsigned bk = blockIdx.x * blockDim.x + threadIdx.x;
  if (bk >= (unsigned)g.B2) return;
  const int *d = draw + s * D_INTS;
  const int G = d[D_GOFF], ai = d[D_AI], b = d[D_B];
  const double nn = (double)g.n;
  // same expression order as the reference so the (large) arguments round identically
  const double arg0 = 2 * M_PI * G * b / nn, arg1 = 2 * M_PI * ai * b / nn;
  double s0, c0, s1, c1;
  sincos(arg0, &s0, &c0);
  sincos(arg1, &s1, &c1);
  const cplx shift0 = make_double2(c0, s0), step = make_double2(c1, s1);
  const int chunks = g.w2 / g.B2;
  const unsigned i0 = bk + (unsigned)l;
  cplx phase = cmul_rn(shift0, cpow_int(step, i0));
  const cplx stepB = cpow_int(step, (unsigned)g.B2);
  const unsigned mask = (unsigned)(g.n - 1);
  unsigned idx = (unsigned)(((unsigned long long)(unsigned)(G % g.n) + (unsigned long long)i0 * (unsigned)ai) & mask);
  const unsigned idx_step = (unsigned)(((unsigned long long)g.B2 * (unsigned)ai) & mask);
  const cplx *__restrict__ xs = x + (long long)s * x_stride;
  double ar = 0.0, aim = 0.0;
  for (int c = 0; c < chunks; c++) {
    const cplx v = ldg_stream(xs + idx);
    const cplx P = cmul_rn(v, phase);
    const cplx f = __ldg(&taps[(unsigned)c * g.B2 + bk]);
    ar = __dadd_rn(ar, __dsub_rn(__dmul_rn(f.x, P.x), __dmul_rn(f.y, P.y)));
    aim = __dadd_rn(aim, __dadd_rn(__dmul_rn(f.y, P.x), __dmul_rn(f.x, P.y)));
    phase = cmul_rn(phase, stepB);
    idx = (idx + idx_step) & mask;
  }
  samp[(long long)s * g.nslots + g2_base(g) + l * g.B2 + bitrev(bk, g.logB2)] = make_double2(ar, aim);
}

// ---- the peeling loop ------------------------------------------------------
constexpr int kPeelThreads = 512;

struct PeelArgs {
  const int *draw;
  cplx *samp;
  int *head;
  int *est_key; cplx *est_val;
  int *t_slot, *t_next; cplx *t_delta;
  int *hkey, *hidx;
  int *ans_key; cplx *ans_val;
  int *count, *rounds;
  const cplx *fwin1, *fwin2;
  int *aux;             // [cap][aux_ints] team scratch
  long long *prof;      // per-signal cycle counters of the peeling phases (8 slots)
};

struct PeelCtx {
  V3Geom g;
  int a, ai, b, shift, off, goff;
  cplx *samp;
  int *head;
  int *est_key; cplx *est_val;
  int *t_slot, *t_next; cplx *t_delta;
  int *hkey, *hidx;
  int *ans_key; cplx *ans_val;
  const cplx *fwin1, *fwin2;
  int *scr;             // [est_cap] per-item scratch of ans_accumulate
  unsigned *flagw;      // [flag_words] decode flags
};

__device__ __forceinline__ cplx fwin_at(const cplx *__restrict__ fwin, int half, int n, int dist)
{
  // filterf[dist] with dist in [0, n): the window holds indices (-half .. +half) mod n
  const int sd = dist < n / 2 ? dist : dist - n;
  return __ldg(&fwin[half + sd]);
}

// C99 complex division as libgcc's __divdc3 evaluates it in the normal range
// (Smith's method, with its alternate order when the ratio underflows);
// reference: `median_value / filter_value`, computefourier-3.0.cc:619-620
__device__ __forceinline__ cplx cdiv_smith(cplx x, cplx y)
{
  const double a = x.x, b = x.y, c = y.x, d = y.y;
  const double RMIN = 2.2250738585072014e-308;
  double rx, ry;
  if (fabs(c) < fabs(d)) {
    const double ratio = __ddiv_rn(c, d);
    const double denom = __dadd_rn(__dmul_rn(c, ratio), d);
    if (fabs(ratio) > RMIN) {
      rx = __ddiv_rn(__dadd_rn(__dmul_rn(a, ratio), b), denom);
      ry = __ddiv_rn(__dsub_rn(__dmul_rn(b, ratio), a), denom);
    } else {
      rx = __ddiv_rn(__dadd_rn(__dmul_rn(c, __ddiv_rn(a, d)), b), denom);
      ry = __ddiv_rn(__dsub_rn(__dmul_rn(c, __ddiv_rn(b, d)), a), denom);
    }
  } else {
    const double ratio = __ddiv_rn(d, c);
    const double denom = __dadd_rn(__dmul_rn(d, ratio), c);
    if (fabs(ratio) > RMIN) {
      rx = __ddiv_rn(__dadd_rn(__dmul_rn(b, ratio), a), denom);
      ry = __ddiv_rn(__dsub_rn(b, __dmul_rn(a, ratio)), denom);
    } else {
      rx = __ddiv_rn(__dadd_rn(__dmul_rn(d, __ddiv_rn(b, c)), a), denom);
      ry = __ddiv_rn(__dsub_rn(b, __dmul_rn(d, __ddiv_rn(a, c))), denom);
    }
  }
  return make_double2(rx, ry);
}

// block-wide exclusive scan of one int per thread; `total` is the block sum
__device__ int block_scan_excl(int v, int &total, unsigned *warp_tot)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += u;
  }
  if (lane == 31) warp_tot[warp] = (unsigned)incl;
  __syncthreads();
  if (warp == 0) {
    const int t = lane < (int)(blockDim.x >> 5) ? (int)warp_tot[lane] : 0;      // CTAs of fewer than 32 warps
    int sc = t;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, sc, off);
      if (lane >= off) sc += u;
    }
    warp_tot[lane] = (unsigned)(sc - t);
    if (lane == 31) warp_tot[32] = (unsigned)sc;
  }
  __syncthreads();
  const int excl = (int)warp_tot[warp] + incl - v;
  total = (int)warp_tot[32];
  __syncthreads();
  return excl;
}

__device__ int block_sum(int v, unsigned *warp_tot)
{
  int total;
  block_scan_excl(v, total, warp_tot);
  return total;
}

// computefourier-3.0.cc:642-776, one bucket
__device__ __forceinline__ bool decode_one_mansour(const PeelCtx &c, int bk, cplx s0, cplx s1, int &key, cplx &val)
{
  const V3Geom &g = c.g;
  const double PI2 = 2 * M_PI, N_OVER_PI2 = (double)g.n / PI2, PI2_OVER_N = PI2 / (double)g.n;
  const unsigned FREQ_MASK = ((unsigned)(g.n - 1)) & ~((unsigned)g.W - 1u);
  const double NORM = 1. / (double)g.W, NORM2 = NORM * NORM;
  const double e0 = __dadd_rn(__dmul_rn(s0.x, s0.x), __dmul_rn(s0.y, s0.y));
  const double e1 = __dadd_rn(__dmul_rn(s1.x, s1.x), __dmul_rn(s1.y, s1.y));
  const double zero_check = __dadd_rn(__dmul_rn(e0, NORM2), __dmul_rn(e1, NORM2));
  if (!(zero_check > 1e-8)) return false;
  const double c0 = __dmul_rn(e0, NORM2), c1 = __dmul_rn(e1, NORM2);
  const double d0 = atan2(s0.y * NORM, s0.x * NORM), d1 = atan2(s1.y * NORM, s1.x * NORM);
  const double inv = 1. / c0;
  const double bb = c1 * inv - 1;
  const double error = bb * bb;
  if (!(error < g.n * 1e-10 && c0 > 0.01)) return false;
  const double slope = d1 - d0;
  const int freq1 = (int)llrint(slope * N_OVER_PI2);
  const int freq3 = (int)(((unsigned)freq1 & FREQ_MASK) | (unsigned)bk);
  const int freq_offset = (int)((unsigned)freq3 * (unsigned)c.off);     // 32-bit wrap, :743
  const double phase = d0 - PI2_OVER_N * freq_offset;
  const double mag = sqrt(c0);
  double sn, cs;
  sincos(phase, &sn, &cs);
  key = freq3;
  val = make_double2(mag * cs, mag * sn);
  return true;
}

// computefourier-3.0.cc:484-640, one bucket
__device__ __forceinline__ bool decode_one_gauss(const PeelCtx &c, int which, int bk, cplx s0, cplx s1, int &key,
                                                 cplx &val)
{
  const V3Geom &g = c.g;
  const int B = which == 1 ? g.B1 : g.B2;
  const int a = which == 1 ? 1 : c.a, b = which == 1 ? 0 : c.b;
  const cplx *fwin = which == 1 ? c.fwin1 : c.fwin2;
  const int half = which == 1 ? g.fw_half1 : g.fw_half2;
  const double PI2 = 2 * M_PI, N_OVER_PI2 = (double)g.n / PI2;
  const double PI2_A_OFFSET_OVER_N = PI2 * a * c.goff / (double)g.n;
  const double BUCKETS_OVER_N = (double)B / (double)g.n;
  const unsigned n1 = (unsigned)(g.n - 1), Bm = (unsigned)(B - 1);
  const unsigned N_OVER_BUCKETS = (unsigned)(g.n / B);
  double c0 = __dadd_rn(__dmul_rn(s0.x, s0.x), __dmul_rn(s0.y, s0.y));
  const double c1 = __dadd_rn(__dmul_rn(s1.x, s1.x), __dmul_rn(s1.y, s1.y));
  if (!(__dadd_rn(c0, c1) > 1e-8)) return false;
  const double d0 = atan2(s0.y, s0.x), d1 = atan2(s1.y, s1.x);
  const double error_b = c1 / c0 - 1;
  double error = error_b * error_b;
  error /= (double)g.n;
  if (!(error < 1e-12 && c0 > 0.01)) return false;
  const double slope = d1 - d0;
  int freq = (int)llrint(N_OVER_PI2 * slope) + g.n;
  freq = (int)((unsigned)freq & n1);
  const unsigned hashed_to = (unsigned)llrint(freq * BUCKETS_OVER_N) & Bm;
  if (hashed_to != (unsigned)bk) return false;
  const double phase = d0 - PI2_A_OFFSET_OVER_N * freq;
  c0 = sqrt(c0);
  double sn, cs;
  sincos(phase, &sn, &cs);
  cplx v = make_double2(c0 * cs, c0 * sn);
  const int dist = (int)((hashed_to * N_OVER_BUCKETS - (unsigned)freq + (unsigned)g.n) & n1);
  v = cdiv_smith(v, fwin_at(fwin, half, g.n, dist));
  const unsigned pf = (unsigned)(((unsigned long long)(unsigned)freq * (unsigned)a) & n1);   // timesmod
  key = (int)((pf - (unsigned)b + (unsigned)g.n) & n1);
  val = v;
  return true;
}

// ---- the team: the threads that peel ONE signal --------------------------------------
// One CTA for small plans; a thread-block cluster of up to 16 CTAs (16 SMs) for large ones.
// Everything the peeling loop does is either independent per bucket / per found coefficient
// (decode: two atan2, a sincos and a sqrt in double precision per occupied bucket; targets:
// six sincos per coefficient) or an ordered compaction, so it spreads over the team with one
// hardware cluster barrier where the single-CTA version had a __syncthreads(); one SM has
// 64 FP64 lanes, a 16-CTA team has 1024.  State shared by the team lives in global memory
// (L2); barrier.cluster's release/acquire at cluster scope orders it (and drops stale L1 lines).
constexpr int kMaxTeam = 16;
struct Team {
  int rank, size;          // this CTA's rank in the team, CTAs in the team
  int tid, nthreads;       // thread index in the team, threads in the team
  int *xch;                // [4][kMaxTeam][4] exchange slots (global)
  int *ctr;                // [4] global counters: 0 touched buckets, 1 segment allocator
  int seq;                 // exchanges done so far (uniform over the team)
  __device__ __forceinline__ void sync() const
  {
    if (size == 1) { __syncthreads(); return; }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  // every CTA contributes up to three ints; returns, per component, the sum over the CTAs
  // before this one (`before`) and over all (`total`).  One team barrier.
  __device__ void exchange3(int v0, int v1, int v2, int *before, int *total)
  {
    int *slot = xch + (seq & 3) * kMaxTeam * 4;
    seq++;
    if (threadIdx.x == 0) { slot[rank * 4 + 0] = v0; slot[rank * 4 + 1] = v1; slot[rank * 4 + 2] = v2; }
    sync();
    int b[3] = {0, 0, 0}, t[3] = {0, 0, 0};
    for (int q = 0; q < size; q++)
      for (int e = 0; e < 3; e++) {
        const int v = slot[q * 4 + e];
        if (q < rank) b[e] += v;
        t[e] += v;
      }
    for (int e = 0; e < 3; e++) { before[e] = b[e]; total[e] = t[e]; }
  }
};

// Decode every bucket of one filter (which: 0 aliasing, 1 first window, 2 permuted window)
// and list what was found in ascending bucket order, as the reference's loops do.
// Pass 1: all buckets in parallel over the team; a found (key, value) is parked at its bucket
// index and flagged in a bit map (global).  One exchange of per-CTA counts.  Pass 2: every CTA
// places the entries of its own bucket range behind those of the CTAs before it.
constexpr int kDecodeAhead = 4;     // loads of several chunks in flight
__device__ int decode_filter(const PeelCtx &c, Team &tm, int which, unsigned *warp_tot)
{
  const V3Geom &g = c.g;
  const int nb = which == 0 ? g.W : (which == 1 ? g.B1 : g.B2);
  const int tid = threadIdx.x, lane = tid & 31;
  int *park_key = c.t_slot;          // free between peel steps
  cplx *park_val = c.t_delta;
  unsigned *flagw = c.flagw;
  // contiguous bucket range of this CTA, a whole number of flag words; a filter of at most
  // 2048 buckets is CTA 0's alone (its placement then needs no counts from anybody else)
  const int nwords = (nb + 31) >> 5;
  const bool solo = nb <= 2 * kPeelThreads;
  const int wper = solo ? nwords : (nwords + tm.size - 1) / tm.size;
  const int w_lo = tm.rank * wper < nwords ? tm.rank * wper : nwords;
  const int w_hi = w_lo + wper < nwords ? w_lo + wper : nwords;
  const int b_lo = w_lo * 32, b_hi = w_hi * 32 < nb ? w_hi * 32 : nb;
  const cplx *p0 = c.samp + (which == 0 ? man_base(g) : (which == 1 ? g1_base(g) : g2_base(g)));
  const cplx *p1 = p0 + nb;
  int mine = 0;
  for (int cb0 = b_lo; cb0 < w_hi * 32; cb0 += kDecodeAhead * kPeelThreads) {
    cplx s0[kDecodeAhead], s1[kDecodeAhead];
#pragma unroll
    for (int u = 0; u < kDecodeAhead; u++) {
      const int bk = cb0 + u * kPeelThreads + tid;
      if (bk < b_hi) { s0[u] = p0[bk]; s1[u] = p1[bk]; }
    }
#pragma unroll
    for (int u = 0; u < kDecodeAhead; u++) {
      const int bk = cb0 + u * kPeelThreads + tid;
      if (cb0 + u * kPeelThreads >= w_hi * 32) break;
      int key = 0;
      cplx val = make_double2(0.0, 0.0);
      bool have = false;
      if (bk < b_hi)
        have = which == 0 ? decode_one_mansour(c, bk, s0[u], s1[u], key, val)
                          : decode_one_gauss(c, which, bk, s0[u], s1[u], key, val);
      if (have) { park_key[bk] = key; park_val[bk] = val; }
      const unsigned bal = __ballot_sync(0xffffffffu, have);
      if (lane == 0 && (bk >> 5) < w_hi) flagw[bk >> 5] = bal;
      mine += have;
    }
  }
  const int cta_found = block_sum(mine, warp_tot);       // (its barriers also publish the flag words inside the CTA)
  int before[3] = {0, 0, 0}, total[3];
  if (!solo) tm.exchange3(cta_found, 0, 0, before, total);
  // pass 2: ordered placement of this CTA's range, a tile of 2048 flag words at a time
  int pos0 = before[0];
  for (int wb = w_lo; wb < w_hi; wb += 2 * kPeelThreads) {
    const int w0 = wb + 2 * tid;
    const unsigned f0 = w0 < w_hi ? flagw[w0] : 0u, f1 = w0 + 1 < w_hi ? flagw[w0 + 1] : 0u;
    int tile_total;
    int pos = pos0 + block_scan_excl(__popc(f0) + __popc(f1), tile_total, warp_tot);
    unsigned bits = f0;
    int wbase = w0 * 32;
    for (int rep = 0; rep < 2; rep++) {
      while (bits) {
        const int bk = wbase + __ffs(bits) - 1;
        bits &= bits - 1;
        if (pos < g.est_cap) { c.est_key[pos] = park_key[bk]; c.est_val[pos] = park_val[bk]; }
        pos++;
      }
      bits = f1;
      wbase += 32;
    }
    pos0 += tile_total;
  }
  if (solo) tm.exchange3(cta_found, 0, 0, before, total);     // publishes the count and the list
  else tm.sync();
  return total[0] < g.est_cap ? total[0] : g.est_cap;
}

// the 6 deltas one coefficient leaves in a windowed filter's buckets (:354-463)
__device__ void gauss_targets(const PeelCtx &c, int which, int key, cplx value, int init_G_offset,
                              int *slots, cplx *deltas)
{
  const V3Geom &g = c.g;
  const int B = which == 1 ? g.B1 : g.B2;
  const int base = which == 1 ? g1_base(g) : g2_base(g);
  const cplx *fwin = which == 1 ? c.fwin1 : c.fwin2;
  const int half = which == 1 ? g.fw_half1 : g.fw_half2;
  const int n = g.n, n1 = n - 1;
  const double PI2_DIV_N = 2 * M_PI / (double)n;
  const unsigned n_over_B = (unsigned)(n / B);
  const int h = (int)(key * 1. / n_over_B + 0.5) % B;
  const int dist1 = (int)(((unsigned)h * n_over_B - (unsigned)key + (unsigned)n) & (unsigned)n1);
  const int dist2 = (int)(((unsigned)dist1 + (unsigned)n + n_over_B) & (unsigned)n1);
  const int dist3 = (int)(((unsigned)dist1 + (unsigned)n + (unsigned)n - n_over_B) & (unsigned)n1);
  const int key_offset = (int)((unsigned)key * (unsigned)init_G_offset) % n;            // :381
  const int key_offset2 = (int)((unsigned)key * (unsigned)(init_G_offset + 1)) % n;     // :382
  double a2, b2, a22, b22;
  sincos(PI2_DIV_N * key_offset, &b2, &a2);
  sincos(PI2_DIV_N * key_offset2, &b22, &a22);
  const cplx v1 = make_double2(value.x * a2 - value.y * b2, value.x * b2 + value.y * a2);
  const cplx v2 = make_double2(value.x * a22 - value.y * b22, value.x * b22 + value.y * a22);
  const cplx f1 = fwin_at(fwin, half, n, dist1), f2 = fwin_at(fwin, half, n, dist2),
             f3 = fwin_at(fwin, half, n, dist3);
  const int hp = (h + 1) % B, hm = (h + B - 1) % B;
  slots[0] = base + h;      deltas[0] = cmul_rn(v1, f1);
  slots[1] = base + hp;     deltas[1] = cmul_rn(v1, f2);
  slots[2] = base + hm;     deltas[2] = cmul_rn(v1, f3);
  slots[3] = base + B + h;  deltas[3] = cmul_rn(v2, f1);
  slots[4] = base + B + hp; deltas[4] = cmul_rn(v2, f2);
  slots[5] = base + B + hm; deltas[5] = cmul_rn(v2, f3);
}

// :300-351
__device__ void mansour_targets(const PeelCtx &c, int key, cplx value, int *slots, cplx *deltas)
{
  const V3Geom &g = c.g;
  const double PI2_OVER_N = 2 * M_PI / (double)g.n;
  const int h = key & (g.W - 1);
  double s0, c0, s1, c1;
  sincos(PI2_OVER_N * key * c.off, &s0, &c0);
  sincos(PI2_OVER_N * key * (c.off + 1), &s1, &c1);
  const double W = (double)g.W;
  slots[0] = man_base(g) + h;
  deltas[0] = make_double2(W * (value.x * c0 - value.y * s0), W * (value.x * s0 + value.y * c0));
  slots[1] = man_base(g) + g.W + h;
  deltas[1] = make_double2(W * (value.x * c1 - value.y * s1), W * (value.x * s1 + value.y * c1));
}

// Reserve `want` consecutive units of *counter for every lane of the (converged) calling warp with
// ONE atomic: returns this lane's first unit.
__device__ __forceinline__ int warp_reserve(int *counter, int want)
{
  const unsigned active = __activemask();
  const int lane = threadIdx.x & 31;
  int incl = want;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(active, incl, off);
    if (lane >= off) incl += v;
  }
  const int last = 31 - __clz(active);
  const int total = __shfl_sync(active, incl, last);
  int base = 0;
  if (lane == last && total > 0) base = atomicAdd(counter, total);
  base = __shfl_sync(active, base, last);
  return base + incl - want;
}

// Subtract, from every touched bucket, the deltas of items [0, F) in item order.
// keys/vals: the items (plain frequencies; the permuted filter sees (key*ai + shift) mod n,
// :465-482, :944); the flags say which filters to peel.
// Deterministic whatever the thread timing: targets are grouped by bucket (atomics decide
// only where a bucket's segment sits and the order INSIDE it), then one thread per touched
// bucket applies its segment in ascending target id, i.e. in item order.  Four team phases.
constexpr int kSmallTargets = 2048;     // peel_apply: up to this many targets are handled by CTA 0 in shared memory
constexpr int kSmallHash = 4096;        // open-addressing table over the touched buckets (load <= 0.5)
constexpr int kSmallHashShift = 20;     // 32 - log2(kSmallHash)
// dynamic shared memory of the peeling kernel (only CTA 0 of a team uses it): 96 KB, which leaves
// the SM ~130 KB of L1 (with 4096 targets / 192 KB the decode phases slowed down by 20 %)
struct PeelSmall {
  cplx delta[kSmallTargets];
  int slot[kSmallTargets];
  int order[kSmallTargets];             // target ids grouped by bucket
  int hkey[kSmallHash];                 // bucket (slot) owning a table entry, -1: free
  int hfill[kSmallHash];                // targets counted, then the running fill pointer of the entry's segment
  int hstart[kSmallHash];               // first position of the entry's segment in `order`
  int alloc;                            // segment allocator
  unsigned char ord[kPeelThreads / 32][64];   // large rounds: per warp, rank -> lane holding that delta
};

__device__ __forceinline__ int small_find(const PeelSmall &sm, int slot)
{
  for (unsigned h = ((unsigned)slot * 2654435761u) >> kSmallHashShift;; h = (h + 1) & (kSmallHash - 1))
    if (sm.hkey[h] == slot) return (int)h;
}

// the targets of item i, part 0 (permuted window, 6), 1 (first window, 6) or 2 (aliasing, 2),
// written at slots[0..) / deltas[0..) of that part; absent parts leave slot -1
__device__ __forceinline__ void item_targets(const PeelCtx &c, int part, int key, cplx v, bool enabled, int *slots,
                                             cplx *deltas)
{
  const V3Geom &g = c.g;
  const int cnt = part == 2 ? 2 : 6;
  for (int q = 0; q < cnt; q++) slots[q] = -1;
  if (!enabled) return;
  if (part == 0) {
    const int key2 = (int)(((((unsigned long long)(unsigned)key * (unsigned)c.ai) & (unsigned)(g.n - 1)) +
                            (unsigned)c.shift) % (unsigned)g.n);
    const int a_off = (int)((unsigned)c.a * (unsigned)c.goff);
    gauss_targets(c, 2, key2, v, a_off, slots, deltas);
  } else if (part == 1) {
    gauss_targets(c, 1, key, v, c.goff, slots, deltas);
  } else {
    mansour_targets(c, key, v, slots, deltas);
  }
}

__device__ void peel_apply(const PeelCtx &c, Team &tm, const int *keys, const cplx *vals, int F, bool do_g2,
                           bool do_g1, bool do_man, PeelSmall &sm)
{
  const V3Geom &g = c.g;
  if (F == 0) return;                             // uniform over the team
  if (F * 14 <= kSmallTargets) {
    // Up to 146 coefficients (every round but the first ones): CTA 0 alone, everything in shared
    // memory.  One thread per (item, filter) for the trigonometry; the targets are grouped by
    // bucket through a small hash table (shared-memory atomics decide only where a bucket's
    // segment sits and the order inside it); one thread per touched bucket then subtracts its
    // segment in ascending target id = item order.  No global atomics, one team barrier.
    if (tm.rank == 0) {
      const int T = F * 14;
      for (int h = threadIdx.x; h < kSmallHash; h += kPeelThreads) { sm.hkey[h] = -1; sm.hfill[h] = 0; }
      if (threadIdx.x == 0) sm.alloc = 0;
      for (int u = threadIdx.x; u < 3 * F; u += kPeelThreads) {
        const int i = u / 3, part = u - 3 * i;
        const bool on = part == 0 ? do_g2 : (part == 1 ? do_g1 : do_man);
        int sl[6];
        cplx dl[6];
        item_targets(c, part, keys[i], vals[i], on, sl, dl);
        const int base = i * 14 + part * 6;
        const int cntp = part == 2 ? 2 : 6;
        for (int q = 0; q < cntp; q++) { sm.slot[base + q] = sl[q]; if (sl[q] >= 0) sm.delta[base + q] = dl[q]; }
      }
      __syncthreads();
      // count the targets of every touched bucket
      for (int t = threadIdx.x; t < T; t += kPeelThreads) {
        const int slot = sm.slot[t];
        if (slot < 0) continue;
        for (unsigned h = ((unsigned)slot * 2654435761u) >> kSmallHashShift;; h = (h + 1) & (kSmallHash - 1)) {
          const int prev = atomicCAS(&sm.hkey[h], -1, slot);
          if (prev == -1 || prev == slot) { atomicAdd(&sm.hfill[h], 1); break; }
        }
      }
      __syncthreads();
      for (int h = threadIdx.x; h < kSmallHash; h += kPeelThreads) {
        const int len = sm.hfill[h];
        if (len > 0) {
          const int start = atomicAdd(&sm.alloc, len);
          sm.hstart[h] = start;
          sm.hfill[h] = start;
        }
      }
      __syncthreads();
      for (int t = threadIdx.x; t < T; t += kPeelThreads) {
        const int slot = sm.slot[t];
        if (slot >= 0) sm.order[atomicAdd(&sm.hfill[small_find(sm, slot)], 1)] = t;
      }
      __syncthreads();
      for (int h = threadIdx.x; h < kSmallHash; h += kPeelThreads) {
        const int slot = sm.hkey[h];
        if (slot < 0) continue;
        const int start = sm.hstart[h], len = sm.hfill[h] - start;
        cplx val = c.samp[slot];
        int last = -1;
        for (int rep = 0; rep < len; rep++) {          // ascending target id == item order
          int best = 0x7fffffff;
          for (int q = 0; q < len; q++) {
            const int id = sm.order[start + q];
            if (id > last && id < best) best = id;
          }
          val = csub_rn(val, sm.delta[best]);
          last = best;
        }
        c.samp[slot] = val;
      }
    }
    tm.sync();
    return;
  }
  int *cnt = c.head;                              // [nslots], zero between calls
  int *fill = c.t_next;                           // [nslots] running fill pointers
  int *touched = c.t_next + g.nslots;             // [nslots] buckets with at least one target
  int *seg = c.t_next + 2 * g.nslots;             // [F*14] target ids grouped by bucket
  int *ntouched = tm.ctr + 0, *seg_alloc = tm.ctr + 1;   // zero between calls
  // one thread per (item, filter): the trigonometry of the three filters runs side by side
  for (int u = tm.tid; u < 3 * F; u += tm.nthreads) {
    const int i = u / 3, part = u - 3 * i;
    const bool on = part == 0 ? do_g2 : (part == 1 ? do_g1 : do_man);
    const int cntp = part == 2 ? 2 : 6;
    int slots[6];
    cplx deltas[6];
    item_targets(c, part, keys[i], vals[i], on, slots, deltas);
    // all counters at once (one round trip), then one slot reservation for the buckets
    // this thread touched first
    int old[6], nfirst = 0;
    const int tbase = i * 14 + part * 6;
#pragma unroll
    for (int q = 0; q < 6; q++) {
      old[q] = 1;
      if (q < cntp) {
        c.t_slot[tbase + q] = slots[q];
        if (slots[q] >= 0) {
          c.t_delta[tbase + q] = deltas[q];
          old[q] = atomicAdd(&cnt[slots[q]], 1);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 6; q++) nfirst += old[q] == 0;
    int tpos = warp_reserve(ntouched, nfirst);        // one atomic per warp: same-address atomics serialise in L2
#pragma unroll
    for (int q = 0; q < 6; q++)
      if (old[q] == 0) touched[tpos++] = slots[q];
  }
  tm.sync();
  const int T = __ldcg(ntouched);
  // a segment per touched bucket; where it sits does not matter
  for (int j0 = 0; j0 < T; j0 += tm.nthreads) {      // whole warps stay together for the reservation
    const int j = j0 + tm.tid;
    const int sl = j < T ? touched[j] : -1;
    const int base = warp_reserve(seg_alloc, sl >= 0 ? cnt[sl] : 0);
    if (sl >= 0) fill[sl] = base;
  }
  tm.sync();
  for (int t = tm.tid; t < F * 14; t += tm.nthreads) {
    const int slot = c.t_slot[t];
    if (slot >= 0) seg[atomicAdd(&fill[slot], 1)] = t;
  }
  tm.sync();
  // One WARP per touched bucket.  Its lanes fetch the segment's (target id, delta) pairs (two per
  // lane: up to 64), rank the ids against each other with shuffles, leave "rank -> holder" in
  // shared memory, and the deltas are then subtracted in ascending id = item order, each fetched
  // from its holder by one shuffle -- no dependent global loads and no per-thread O(len^2) chain
  // (one thread per bucket spent ~15 us on the 22-long segments of the 512 permuted-window slots).
  {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *ord = sm.ord[wib];
    for (int j = tm.tid >> 5; j < T; j += tm.nthreads >> 5) {
      const int sl = touched[j];
      const int len = cnt[sl];
      const int *ids = seg + (fill[sl] - len);
      cplx val = c.samp[sl];
      if (len <= 64) {
        const int id0 = lane < len ? ids[lane] : 0x7fffffff;
        const int id1 = lane + 32 < len ? ids[lane + 32] : 0x7fffffff;
        const cplx zero = make_double2(0.0, 0.0);
        const cplx d0 = lane < len ? c.t_delta[id0] : zero;
        const cplx d1 = lane + 32 < len ? c.t_delta[id1] : zero;
        int r0 = 0, r1 = 0;
        for (int q = 0; q < 32; q++) {
          const int o0 = __shfl_sync(0xffffffffu, id0, q), o1 = __shfl_sync(0xffffffffu, id1, q);
          r0 += (o0 < id0) + (o1 < id0);
          r1 += (o0 < id1) + (o1 < id1);
        }
        if (lane < len) ord[r0] = (unsigned char)lane;
        if (lane + 32 < len) ord[r1] = (unsigned char)(lane + 32);
        __syncwarp();
        for (int r = 0; r < len; r++) {              // ascending target id == item order
          const int holder = ord[r];
          const cplx mine = holder >= 32 ? d1 : d0;
          const cplx d = make_double2(__shfl_sync(0xffffffffu, mine.x, holder & 31),
                                      __shfl_sync(0xffffffffu, mine.y, holder & 31));
          val = csub_rn(val, d);
        }
        __syncwarp();
      } else {
        int last = -1;
        for (int rep = 0; rep < len; rep++) {
          int best = 0x7fffffff;
          for (int q = 0; q < len; q++) {
            const int id = ids[q];
            if (id > last && id < best) best = id;
          }
          val = csub_rn(val, c.t_delta[best]);
          last = best;
        }
      }
      if (lane == 0) { c.samp[sl] = val; cnt[sl] = 0; }
    }
  }
  if (tm.tid == 0) { *ntouched = 0; *seg_alloc = 0; }
  tm.sync();
}

__device__ __forceinline__ unsigned hash_key(int key, int log_hash)
{
  return ((unsigned)key * 2654435761u) >> (32 - log_hash);
}

__device__ int hash_find(const PeelCtx &c, int key)
{
  const unsigned m = (unsigned)c.g.hash_size - 1u;
  for (unsigned h = hash_key(key, c.g.log_hash);; h = (h + 1) & m) {
    const int k = c.hkey[h];
    if (k == key) return c.hidx[h];
    if (k == -1) return -1;
  }
}

__device__ void hash_insert(const PeelCtx &c, int key, int idx)
{
  const unsigned m = (unsigned)c.g.hash_size - 1u;
  for (unsigned h = hash_key(key, c.g.log_hash);; h = (h + 1) & m) {
    const int prev = atomicCAS(&c.hkey[h], -1, key);
    if (prev == -1 || prev == key) { c.hidx[h] = idx; return; }
  }
}

// ans[key] (+)= val for the F decoded items, new keys appended in item order
// (:850, :909-913, :969-973, :1025-1029).  Every CTA takes a contiguous block of the items:
// pass 1 looks every key up and ranks the new ones inside the block, one exchange of the
// per-CTA counts, pass 2 appends / accumulates.  Keys of one decode pass are distinct.
__device__ int ans_accumulate(const PeelCtx &c, Team &tm, int F, int ans_count, bool assign, unsigned *warp_tot)
{
  if (F == 0) return ans_count;                     // uniform over the team
  if (F <= kPeelThreads) {
    // one item per thread of CTA 0; the others only learn the new count
    int newc = 0;
    if (tm.rank == 0) {
      const int i = threadIdx.x;
      const bool have = i < F;
      int idx = -1;
      if (have) idx = hash_find(c, c.est_key[i]);
      const bool is_new = have && idx < 0;
      const int excl = block_scan_excl(is_new ? 1 : 0, newc, warp_tot);      // its barriers separate lookups from inserts
      if (is_new) {
        idx = ans_count + excl;
        if (idx < c.g.ans_cap) {
          c.ans_key[idx] = c.est_key[i];
          c.ans_val[idx] = make_double2(0.0, 0.0);
          hash_insert(c, c.est_key[i], idx);
        }
      }
      if (have && idx >= 0 && idx < c.g.ans_cap) {
        const cplx v = c.est_val[i];
        c.ans_val[idx] = assign ? v : cadd_rn(c.ans_val[idx], v);
      }
    }
    int before[3], total[3];
    tm.exchange3(newc, 0, 0, before, total);
    ans_count += total[0];
    return ans_count > c.g.ans_cap ? c.g.ans_cap : ans_count;
  }
  const int per = (F + tm.size - 1) / tm.size;
  const int lo = tm.rank * per < F ? tm.rank * per : F;
  const int hi = lo + per < F ? lo + per : F;
  int *found_idx = c.scr;            // [est_cap]: >= 0 index in ans, < 0: -(rank among this CTA's new keys) - 1
  int newc = 0;
  for (int base = lo; base < hi; base += kPeelThreads) {
    const int i = base + threadIdx.x;
    const bool have = i < hi;
    int idx = -1;
    if (have) idx = hash_find(c, c.est_key[i]);
    const bool is_new = have && idx < 0;
    int total;
    const int excl = block_scan_excl(is_new ? 1 : 0, total, warp_tot);
    if (have) found_idx[i] = is_new ? -(newc + excl) - 1 : idx;
    newc += total;
  }
  int before[3], total[3];
  tm.exchange3(newc, 0, 0, before, total);
  for (int i = lo + (int)threadIdx.x; i < hi; i += kPeelThreads) {
    int idx = found_idx[i];
    if (idx < 0) {
      idx = ans_count + before[0] + (-idx - 1);
      if (idx < c.g.ans_cap) {
        c.ans_key[idx] = c.est_key[i];
        c.ans_val[idx] = make_double2(0.0, 0.0);
        hash_insert(c, c.est_key[i], idx);
      }
    }
    if (idx >= 0 && idx < c.g.ans_cap) {
      const cplx v = c.est_val[i];
      c.ans_val[idx] = assign ? v : cadd_rn(c.ans_val[idx], v);
    }
  }
  ans_count += total[0];
  if (ans_count > c.g.ans_cap) ans_count = c.g.ans_cap;
  tm.sync();
  return ans_count;
}

__global__ void __launch_bounds__(kPeelThreads)
v3_peel_kernel(V3Geom g, PeelArgs a, int team)
{
  __shared__ unsigned warp_tot[33];
  extern __shared__ __align__(16) unsigned char peel_dyn[];
  PeelSmall &small = *reinterpret_cast<PeelSmall *>(peel_dyn);
  const int s = blockIdx.x / team;
  Team tm;
  tm.size = team;
  tm.rank = blockIdx.x % team;
  tm.tid = tm.rank * kPeelThreads + threadIdx.x;
  tm.nthreads = team * kPeelThreads;
  tm.xch = a.aux + (long long)s * g.aux_ints + g.est_cap + g.flag_words;
  tm.ctr = tm.xch + 4 * kMaxTeam * 4;
  tm.seq = 0;
  PeelCtx c;
  c.g = g;
  const int *d = a.draw + s * D_INTS;
  c.a = d[D_A]; c.ai = d[D_AI]; c.b = d[D_B]; c.shift = d[D_SHIFT]; c.off = d[D_OFF]; c.goff = d[D_GOFF];
  c.samp = a.samp + (long long)s * g.nslots;
  c.head = a.head + (long long)s * g.nslots;
  c.est_key = a.est_key + (long long)s * g.est_cap;
  c.est_val = a.est_val + (long long)s * g.est_cap;
  c.t_slot = a.t_slot + (long long)s * g.tgt_cap * 14;
  c.t_next = a.t_next + (long long)s * (g.tgt_cap * 14 + 2 * g.nslots);
  c.t_delta = a.t_delta + (long long)s * g.tgt_cap * 14;
  c.hkey = a.hkey + (long long)s * g.hash_size;
  c.hidx = a.hidx + (long long)s * g.hash_size;
  c.ans_key = a.ans_key + (long long)s * g.ans_cap;
  c.ans_val = a.ans_val + (long long)s * g.ans_cap;
  c.fwin1 = a.fwin1; c.fwin2 = a.fwin2;
  c.scr = a.aux + (long long)s * g.aux_ints;
  c.flagw = reinterpret_cast<unsigned *>(c.scr + g.est_cap);

  for (long long i = tm.tid; i < g.nslots; i += tm.nthreads) c.head[i] = 0;
  for (int i = tm.tid; i < g.hash_size; i += tm.nthreads) c.hkey[i] = -1;
  if (tm.tid < 4) tm.ctr[tm.tid] = 0;
  tm.sync();

  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tmark = clock64();
#define PROF(slot) { const long long now_ = clock64(); prof[slot] += now_ - tmark; tmark = now_; }
  int ans_count = 0;
  // ---- aliasing filter: decode, record, clear the decoded buckets (:842-855) ----
  int F = decode_filter(c, tm, 0, warp_tot);
  PROF(0)
  ans_count = ans_accumulate(c, tm, F, ans_count, true, warp_tot);
  for (int i = tm.tid; i < F; i += tm.nthreads) {
    // MAN_SAMP[j + 2*(f % W)] = 0 for j = 0,1: in the interleaved layout that is
    // (bucket f % W, shift 0) and (bucket f % W, shift 1)
    const int h = c.est_key[i] & (g.W - 1);
    c.samp[man_base(g) + h] = make_double2(0.0, 0.0);
    c.samp[man_base(g) + g.W + h] = make_double2(0.0, 0.0);
  }
  tm.sync();
  if (F == g.k) {
    // :859-860: the reference returns here BEFORE copying its map to the output
    if (tm.tid == 0) { a.count[s] = 0; a.rounds[s] = 0; }
    return;
  }
  // ---- first window: peel what is known, decode, peel from window 1 + aliasing (:881-924) ----
  PROF(2)
  peel_apply(c, tm, c.ans_key, c.ans_val, ans_count, false, true, false, small);
  PROF(4)
  F = decode_filter(c, tm, 1, warp_tot);
  PROF(1)
  ans_count = ans_accumulate(c, tm, F, ans_count, false, warp_tot);
  PROF(2)
  peel_apply(c, tm, c.est_key, c.est_val, F, false, true, true, small);
  PROF(3)
  // ---- permuted window (:940-981) ----
  peel_apply(c, tm, c.ans_key, c.ans_val, ans_count, true, false, false, small);
  PROF(4)
  F = decode_filter(c, tm, 2, warp_tot);
  PROF(1)
  ans_count = ans_accumulate(c, tm, F, ans_count, false, warp_tot);
  PROF(2)
  peel_apply(c, tm, c.est_key, c.est_val, F, true, true, true, small);
  PROF(3)
  // ---- round robin until the occupied-bucket counts repeat (:991-1076) ----
  int prev_m = 0, prev_1 = 0, prev_2 = 0, rounds = 0;
  for (int nana = 0;; nana++) {
    if (nana % 3 == 0) { F = decode_filter(c, tm, 0, warp_tot); PROF(0) }
    else { F = decode_filter(c, tm, nana % 3 == 1 ? 1 : 2, warp_tot); PROF(1) }
    ans_count = ans_accumulate(c, tm, F, ans_count, false, warp_tot);
    PROF(2)
    peel_apply(c, tm, c.est_key, c.est_val, F, true, true, true, small);
    PROF(3)
    rounds = nana + 1;
    if (nana % 3 == 2) {
      // the reference indexes its interleaved arrays with a plain j < B (:1047-1057):
      // entry j is (bucket j/2, shift j%2)
      int l1 = 0, l2 = 0, lm = 0;
      for (int j = tm.tid; j < g.B1; j += tm.nthreads)
        l1 += cabs2_rn(c.samp[g1_base(g) + (j & 1) * g.B1 + (j >> 1)]) > 1e-6;
      for (int j = tm.tid; j < g.B2; j += tm.nthreads)
        l2 += cabs2_rn(c.samp[g2_base(g) + (j & 1) * g.B2 + (j >> 1)]) > 1e-6;
#pragma unroll 4
      for (int j = tm.tid; j < g.W; j += tm.nthreads) {
        const cplx v = c.samp[man_base(g) + j];
        const double r = v.x / (double)g.W, m = v.y / (double)g.W;
        lm += __dadd_rn(__dmul_rn(r, r), __dmul_rn(m, m)) > 1e-6;
      }
      const int b1 = block_sum(l1, warp_tot), b2 = block_sum(l2, warp_tot), bm = block_sum(lm, warp_tot);
      int before[3], total[3];
      tm.exchange3(b1, b2, bm, before, total);
      const int c1 = total[0], c2 = total[1], cm = total[2];
      PROF(5)
      if (prev_m == cm && prev_1 == c1 && prev_2 == c2) break;
      prev_m = cm; prev_1 = c1; prev_2 = c2;
      if (nana > 3000) break;      // safety net; the reference has none
    }
  }
  if (tm.tid == 0) {
    a.count[s] = ans_count;
    a.rounds[s] = rounds;
    if (a.prof)
      for (int q = 0; q < 8; q++) a.prof[s * 8 + q] = prof[q];
  }
#undef PROF
}

}  // namespace

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int floor_to_pow2_v3(double x)
{
  unsigned int ans;
  for (ans = 1; ans <= x; ans <<= 1) {}
  return (int)(ans / 2);
}

static V3Geom make_geom(const PlanImpl *p)
{
  const PlanV3 &v = *p->v3;
  V3Geom g;
  g.n = p->n; g.logn = p->logn; g.k = p->k;
  g.W = v.W_Man; g.logW = v.logW; g.B1 = v.B_g1; g.logB1 = v.logB1; g.B2 = v.B_g2; g.logB2 = v.logB2;
  g.w1 = v.filt[0].w; g.w2 = v.filt[1].w;
  g.fw_half1 = v.filt[0].fw_half; g.fw_half2 = v.filt[1].fw_half;
  g.nslots = v.nslots;
  g.est_cap = v.est_cap; g.ans_cap = v.ans_cap; g.tgt_cap = v.tgt_cap; g.hash_size = v.hash_size; g.log_hash = v.log_hash;
  g.flag_words = v.flag_words; g.aux_ints = v.aux_ints;
  return g;
}

static void v3_free_scratch(PlanV3 &v)
{
  // a captured graph holds the scratch pointers
  if (v.graph_exec) cudaGraphExecDestroy(v.graph_exec);
  if (v.graph) cudaGraphDestroy(v.graph);
  v.graph_exec = nullptr; v.graph = nullptr;
  cudaFree(v.d_samp); cudaFree(v.d_head); cudaFree(v.d_est_key); cudaFree(v.d_est_val);
  cudaFree(v.d_t_slot); cudaFree(v.d_t_next); cudaFree(v.d_t_delta); cudaFree(v.d_hkey);
  cudaFree(v.d_hidx); cudaFree(v.d_ans_key); cudaFree(v.d_ans_val); cudaFree(v.d_count);
  cudaFree(v.d_rounds); cudaFree(v.d_draw); cudaFree(v.d_prof); v.d_prof = nullptr;
  cudaFree(v.d_aux); v.d_aux = nullptr;
  for (int i = 0; i < kStageSlots; i++) {
    if (v.h_draw[i]) cudaFreeHost(v.h_draw[i]);
    v.h_draw[i] = nullptr;
  }
  v.d_samp = nullptr; v.d_head = nullptr; v.d_est_key = nullptr; v.d_est_val = nullptr;
  v.d_t_slot = nullptr; v.d_t_next = nullptr; v.d_t_delta = nullptr; v.d_hkey = nullptr;
  v.d_hidx = nullptr; v.d_ans_key = nullptr; v.d_ans_val = nullptr; v.d_count = nullptr;
  v.d_rounds = nullptr; v.d_draw = nullptr;
  v.cap = 0;
}

static int v3_ensure_capacity(PlanImpl *p, int nsig)
{
  PlanV3 &v = *p->v3;
  if (nsig <= v.cap) return 0;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  v3_free_scratch(v);
  const long long S = nsig;
  SFFTB_CUDA(cudaMalloc(&v.d_samp, sizeof(cplx) * S * v.nslots));
  SFFTB_CUDA(cudaMalloc(&v.d_head, sizeof(int) * S * v.nslots));
  SFFTB_CUDA(cudaMalloc(&v.d_est_key, sizeof(int) * S * v.est_cap));
  SFFTB_CUDA(cudaMalloc(&v.d_est_val, sizeof(cplx) * S * v.est_cap));
  SFFTB_CUDA(cudaMalloc(&v.d_t_slot, sizeof(int) * S * v.tgt_cap * 14));
  SFFTB_CUDA(cudaMalloc(&v.d_t_next, sizeof(int) * S * (v.tgt_cap * 14 + 2 * v.nslots)));
  SFFTB_CUDA(cudaMalloc(&v.d_t_delta, sizeof(cplx) * S * v.tgt_cap * 14));
  SFFTB_CUDA(cudaMalloc(&v.d_hkey, sizeof(int) * S * v.hash_size));
  SFFTB_CUDA(cudaMalloc(&v.d_hidx, sizeof(int) * S * v.hash_size));
  SFFTB_CUDA(cudaMalloc(&v.d_ans_key, sizeof(int) * S * v.ans_cap));
  SFFTB_CUDA(cudaMalloc(&v.d_ans_val, sizeof(cplx) * S * v.ans_cap));
  SFFTB_CUDA(cudaMalloc(&v.d_count, sizeof(int) * S));
  SFFTB_CUDA(cudaMalloc(&v.d_rounds, sizeof(int) * S));
  SFFTB_CUDA(cudaMalloc(&v.d_draw, sizeof(int) * S * D_INTS));
  SFFTB_CUDA(cudaMalloc(&v.d_prof, sizeof(long long) * S * 8));
  SFFTB_CUDA(cudaMalloc(&v.d_aux, sizeof(int) * S * v.aux_ints));
  for (int i = 0; i < kStageSlots; i++)
    SFFTB_CUDA(cudaHostAlloc(&v.h_draw[i], sizeof(int) * S * D_INTS, cudaHostAllocDefault));
  v.cap = nsig;
  return 0;
}

int v3_build(PlanImpl *p, int n_req, int k)
{
  // src/sfft.cc:506-579
  const unsigned n = (unsigned)floor_to_pow2_v3(n_req);
  if ((int)n != n_req || n < 16) {
    set_error("sfft_make_plan: n must be a power of two >= 16");
    return -1;
  }
  if (k < 8) { set_error("sfft_make_plan(v3): k must be >= 8"); return -1; }
  p->v3 = new PlanV3();
  PlanV3 &v = *p->v3;
  p->n = (int)n; p->logn = ilog2(n); p->k = k;
  v.W_Man = floor_to_pow2_v3(10.0 * (double)k);
  if ((unsigned)v.W_Man > n / 2) v.W_Man = (int)(n / 2);
  const double BB = (unsigned)(1.0 * (double)k);
  v.B_g1 = floor_to_pow2_v3(BB);
  const int b_g1 = (int)(1.00 * ((double)n / v.B_g1));
  const double BB2 = (unsigned)(0.25 * (double)k);
  v.B_g2 = floor_to_pow2_v3(BB2);
  const int b_g2 = (int)(1.00 * ((double)n / v.B_g2));
  if (v.B_g1 < 4 || v.B_g2 < 4 || (unsigned)v.B_g1 > n || (unsigned)v.W_Man < 2) {
    set_error("sfft_make_plan(v3): bucket counts out of range for this (n, k)");
    return -1;
  }
  v.logW = ilog2((unsigned)v.W_Man); v.logB1 = ilog2((unsigned)v.B_g1); v.logB2 = ilog2((unsigned)v.B_g2);
  // response entries read: +-n/2B around 0 and the same one bucket up/down (:375-379, :612-616)
  const int half1 = 3 * (int)(n / v.B_g1) / 2 + 2, half2 = 3 * (int)(n / v.B_g2) / 2 + 2;
  const int cap1 = (int)n / 2 - 1;
  FilterSpec specs[2] = {{0.5 / BB, 1.e-8, b_g1, half1 < cap1 ? half1 : cap1},
                         {0.5 / BB2, 1.e-8, b_g2, half2 < cap1 ? half2 : cap1}};
  DeviceFilter *outs[2] = {&v.filt[0], &v.filt[1]};
  if (build_filters(p->logn, 2, specs, outs, p->stream)) return -1;
  if (v.filt[0].w / v.B_g1 < 1 || v.filt[1].w / v.B_g2 < 1 || v.filt[1].w + 2 >= (int)n) {
    set_error("sfft_make_plan(v3): window shorter than one bucket sweep (reference asserts, cf3.cc:221)");
    return -1;
  }
  int twN = v.W_Man;
  if (v.B_g1 > twN) twN = v.B_g1;
  if (v.B_g2 > twN) twN = v.B_g2;
  v.log_twN = ilog2((unsigned)twN);
  std::vector<cplx> tw((size_t)(twN > 1 ? twN - 1 : 1));
  host_twiddle_levels(twN, tw.data());
  SFFTB_CUDA(cudaMalloc(&v.d_tw, sizeof(cplx) * tw.size()));
  SFFTB_CUDA(cudaMemcpyAsync(v.d_tw, tw.data(), sizeof(cplx) * tw.size(), cudaMemcpyHostToDevice, p->stream));
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  v.nslots = 2ll * v.B_g2 + 2ll * v.B_g1 + 2ll * v.W_Man;
  int cap = v.W_Man;
  if (v.B_g1 > cap) cap = v.B_g1;
  v.est_cap = cap;
  v.ans_cap = 8 * k + 2 * v.W_Man;
  v.tgt_cap = v.ans_cap > v.est_cap ? v.ans_cap : v.est_cap;
  v.hash_size = 1;
  v.log_hash = 0;
  while (v.hash_size < 4 * v.ans_cap) { v.hash_size <<= 1; v.log_hash++; }
  v.flag_words = (cap + 31) / 32;
  v.aux_ints = v.est_cap + v.flag_words + 4 * kMaxTeam * 4 + 8;
  // CTAs per signal in the peeling kernel: one SM per 1024 aliasing buckets, at most a
  // 16-CTA cluster (non-portable size; 8 if the device will not schedule 16)
  v.team = 1;
  while (v.team < kMaxTeam && v.team * kPeelThreads < v.W_Man) v.team <<= 1;
  if (const char *e = getenv("SFFTB_V3_TEAM")) {
    const int t = atoi(e);
    if (t == 1 || t == 2 || t == 4 || t == 8 || t == 16) v.team = t;
  }
  for (int i = 0; i < kStageSlots; i++) SFFTB_CUDA(cudaEventCreateWithFlags(&v.ev[i], cudaEventDisableTiming));
  return v3_ensure_capacity(p, 1);
}

void v3_free(PlanImpl *p)
{
  if (!p->v3) return;
  PlanV3 &v = *p->v3;
  if (v.graph_exec) cudaGraphExecDestroy(v.graph_exec);
  if (v.graph) cudaGraphDestroy(v.graph);
  if (v.h_gdraw) cudaFreeHost(v.h_gdraw);
  if (v.h_gx) cudaFreeHost(v.h_gx);
  cudaFree(v.d_gx);
  if (v.g_ev) cudaEventDestroy(v.g_ev);
  v3_free_scratch(v);
  free_filter(&v.filt[0]);
  free_filter(&v.filt[1]);
  cudaFree(v.d_tw);
  for (int i = 0; i < kStageSlots; i++)
    if (v.ev[i]) cudaEventDestroy(v.ev[i]);
  delete p->v3;
  p->v3 = nullptr;
}

// src/computefourier-3.0.cc:800-810: random() until odd, one more random(), two drand48()
int v3_draw(const PlanImpl *p, sfftb_draw *d)
{
  const int n = p->n;
  int a = 0;
  while (gcd_pub(a, n) != 1) a = (int)(random() % n);
  d->v3_a = a;
  d->v3_ai = mod_inverse_pub(a, n);
  d->v3_b = (int)(random() % n);
  d->v3_init_offset = (int)(unsigned)floor(drand48() * n);
  d->v3_init_G_offset = (int)(unsigned)floor(drand48() * n);
  return 0;
}

// one team (thread-block cluster of `team` CTAs) per signal
static int launch_peel(const V3Geom &g, const PeelArgs &a, int &team, bool &team_checked, int nsig, cudaStream_t st)
{
  SFFTB_ONCE_PER_DEVICE({
    SFFTB_CUDA(cudaFuncSetAttribute(v3_peel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PeelSmall)));
    SFFTB_CUDA(cudaFuncSetAttribute(v3_peel_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  });
  for (;;) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(nsig * team));
    cfg.blockDim = dim3(kPeelThreads);
    cfg.dynamicSmemBytes = sizeof(PeelSmall);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)team;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = team > 1 ? 1 : 0;
    if (team > 1 && !team_checked) {
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, v3_peel_kernel, &cfg) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        team >>= 1;            // this device will not co-schedule that many CTAs: halve the team
        continue;
      }
      team_checked = true;
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, v3_peel_kernel, g, a, team);
    if (e != cudaSuccess && team > 1) {
      cudaGetLastError();
      team >>= 1;
      continue;
    }
    if (e != cudaSuccess) {
      set_error(std::string("v3 peel kernel launch -> ") + cudaGetErrorString(e));
      return -1;
    }
    g_launches++;
    return 0;
  }
}

static void v3_fill_draws(const PlanImpl *p, int *h, int nsig, const sfftb_draw *draws)
{
  for (int s = 0; s < nsig; s++) {
    const sfftb_draw &d = draws[s];
    int *o = h + s * D_INTS;
    o[D_A] = d.v3_a; o[D_AI] = d.v3_ai; o[D_B] = d.v3_b;
    o[D_SHIFT] = ((int)((unsigned)d.v3_ai * (unsigned)d.v3_b) % p->n + p->n) % p->n;   // :807, 32-bit wrap
    o[D_OFF] = d.v3_init_offset; o[D_GOFF] = d.v3_init_G_offset;
    o[6] = 0; o[7] = 0;
  }
}

// everything after the draws are on the device: bucketise, bucket FFTs, peel
static int v3_stages(PlanImpl *p, const cplx *d_in, const unsigned long long *x_ind, long long stride, int nsig)
{
  PlanV3 &v = *p->v3;
  cudaStream_t st = p->stream;
  const V3Geom g = make_geom(p);
  {
    dim3 grid((unsigned)ceil_div(g.W, 256), 2, (unsigned)nsig);
    v3_mansour_kernel<<<grid, 256, 0, st>>>(g, d_in, x_ind, stride, v.d_draw, v.d_samp);
    SFFTB_LAUNCH_CHECK();
    dim3 grid1((unsigned)ceil_div(g.B1, 128), 2, (unsigned)nsig);
    v3_gauss_kernel<<<grid1, 128, 0, st>>>(g, d_in, x_ind, stride, v.d_draw, v.filt[0].time, v.d_samp);
    SFFTB_LAUNCH_CHECK();
    dim3 grid2((unsigned)ceil_div(g.B2, 64), 2, (unsigned)nsig);
    v3_gauss_perm_kernel<<<grid2, 64, 0, st>>>(g, d_in, x_ind, stride, v.d_draw, v.filt[1].time, v.d_samp);
    SFFTB_LAUNCH_CHECK();
  }
  timer_mark(p, "bucketise");
  if (fft_dit_inplace(v.d_samp + 0, g.logB2, 2, g.B2, nsig, g.nslots, v.d_tw, v.log_twN, -1, st)) return -1;
  if (fft_dit_inplace(v.d_samp + 2 * g.B2, g.logB1, 2, g.B1, nsig, g.nslots, v.d_tw, v.log_twN, -1, st)) return -1;
  if (fft_dit_inplace(v.d_samp + 2 * g.B2 + 2 * g.B1, g.logW, 2, g.W, nsig, g.nslots, v.d_tw, v.log_twN, -1, st)) return -1;
  timer_mark(p, "bucket_fft");

  PeelArgs a;
  a.draw = v.d_draw; a.samp = v.d_samp; a.head = v.d_head;
  a.est_key = v.d_est_key; a.est_val = v.d_est_val;
  a.t_slot = v.d_t_slot; a.t_next = v.d_t_next; a.t_delta = v.d_t_delta;
  a.hkey = v.d_hkey; a.hidx = v.d_hidx; a.ans_key = v.d_ans_key; a.ans_val = v.d_ans_val;
  a.count = v.d_count; a.rounds = v.d_rounds;
  a.fwin1 = v.filt[0].fwin; a.fwin2 = v.filt[1].fwin;
  a.prof = v.d_prof;
  a.aux = v.d_aux;
  if (launch_peel(g, a, v.team, v.team_checked, nsig, st)) return -1;
  timer_mark(p, "peel");
  p->last_nsig = nsig;
  return 0;
}

// Single-signal transform replayed from a CUDA graph: eight launches and a copy become one,
// which matters as soon as the host is busy (eight ranks of a box launching at once measured
// 5 ms per transform launch by launch against 0.6 ms of device work).
static int v3_exec_graph(PlanImpl *p, const cplx *d_in, const sfftb_draw *draw)
{
  PlanV3 &v = *p->v3;
  cudaStream_t st = p->stream;
  if (!v.graph_exec) {
    if (!v.h_gdraw) {
      SFFTB_CUDA(cudaHostAlloc(&v.h_gdraw, sizeof(int) * D_INTS, cudaHostAllocDefault));
      SFFTB_CUDA(cudaHostAlloc(&v.h_gx, sizeof(unsigned long long), cudaHostAllocDefault));
      SFFTB_CUDA(cudaMalloc(&v.d_gx, sizeof(unsigned long long)));
      SFFTB_CUDA(cudaEventCreateWithFlags(&v.g_ev, cudaEventDisableTiming));
    }
    SFFTB_CUDA(cudaStreamSynchronize(st));
    SFFTB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    if (cudaMemcpyAsync(v.d_draw, v.h_gdraw, sizeof(int) * D_INTS, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (cudaMemcpyAsync(v.d_gx, v.h_gx, sizeof(unsigned long long), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (cudaEventRecordWithFlags(v.g_ev, st, cudaEventRecordExternal) != cudaSuccess) rc = -1;
    if (!rc) rc = v3_stages(p, nullptr, v.d_gx, p->n, 1);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc || e != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      set_error("CUDA graph capture of the v3 transform failed");
      return -1;
    }
    v.graph = g;
    SFFTB_CUDA(cudaGraphInstantiate(&v.graph_exec, v.graph, 0));
  }
  SFFTB_CUDA(cudaEventSynchronize(v.g_ev));      // previous replay has consumed the staging buffers
  v3_fill_draws(p, v.h_gdraw, 1, draw);
  *v.h_gx = (unsigned long long)(uintptr_t)d_in;
  SFFTB_CUDA(cudaGraphLaunch(v.graph_exec, st));
  g_launches += v.graph_kernels;
  p->last_nsig = 1;
  return 0;
}

int v3_exec(PlanImpl *p, const cplx *d_in, long long stride, int nsig, const sfftb_draw *draws)
{
  PlanV3 &v = *p->v3;
  cudaStream_t st = p->stream;
  if (v3_ensure_capacity(p, nsig)) return -1;
  // the first transform runs launch by launch (one-time kernel attributes, the team size the
  // device accepts), then single-signal transforms are replayed from a captured graph
  static const bool graphs = getenv("SFFTB_NO_GRAPH") == nullptr;
  if (nsig == 1 && !p->timer.enabled && graphs && v.plain_execs >= 1) {
    if (v3_exec_graph(p, d_in, draws) == 0) return 0;
    v.plain_execs = -1000000;      // capture failed once: stay on the plain path
  }
  v.plain_execs++;
  timer_begin(p);
  const long long launches0 = g_launches;
  const int slot = v.next_slot;
  v.next_slot = (v.next_slot + 1) % kStageSlots;
  SFFTB_CUDA(cudaEventSynchronize(v.ev[slot]));
  int *h = v.h_draw[slot];
  v3_fill_draws(p, h, nsig, draws);
  SFFTB_CUDA(cudaMemcpyAsync(v.d_draw, h, sizeof(int) * (long long)nsig * D_INTS, cudaMemcpyHostToDevice, st));
  SFFTB_CUDA(cudaEventRecord(v.ev[slot], st));
  timer_mark(p, "stage_draws");
  if (v3_stages(p, d_in, nullptr, stride, nsig)) return -1;
  v.graph_kernels = (int)(g_launches - launches0);
  return 0;
}

int v3_info(const PlanImpl *p, sfftb_info *info)
{
  const PlanV3 &v = *p->v3;
  info->B_g1 = v.B_g1; info->w_g1 = v.filt[0].w; info->B_g2 = v.B_g2; info->w_g2 = v.filt[1].w;
  info->W_Man = v.W_Man;
  info->max_hits = v.ans_cap;
  info->gather_samples = 2ll * v.W_Man + (long long)(v.filt[0].w / v.B_g1) * v.B_g1 + 1 +
                         (long long)(v.filt[1].w / v.B_g2) * v.B_g2 + 1;
  info->gather_tap_bytes = 16ll * (v.filt[0].w + v.filt[1].w);
  return 0;
}

int v3_result(PlanImpl *p, const int **loc, const cplx **val, const int **count, long long *cap)
{
  PlanV3 &v = *p->v3;
  *loc = v.d_ans_key; *val = v.d_ans_val; *count = v.d_count; *cap = v.ans_cap;
  return 0;
}

int v3_filter_sizes(const PlanImpl *p, int which, int *w, int *fw_len)
{
  const PlanV3 &v = *p->v3;
  if (w) *w = v.filt[which].w;
  if (fw_len) *fw_len = 2 * v.filt[which].fw_half + 1;
  return 0;
}

DeviceFilter *v3_filter(PlanImpl *p, int which) { return &p->v3->filt[which]; }

long long v3_debug_fetch(PlanImpl *p, const char *what, void *dst, size_t capacity)
{
  PlanV3 &v = *p->v3;
  const std::string w(what);
  const void *src = nullptr;
  long long bytes = 0;
  if (w == "gauss_perm_samp") { src = v.d_samp; bytes = sizeof(cplx) * 2ll * v.B_g2; }
  else if (w == "gauss_samp") { src = v.d_samp + 2 * v.B_g2; bytes = sizeof(cplx) * 2ll * v.B_g1; }
  else if (w == "man_samp") { src = v.d_samp + 2 * v.B_g2 + 2 * v.B_g1; bytes = sizeof(cplx) * 2ll * v.W_Man; }
  else if (w == "rounds") { src = v.d_rounds; bytes = sizeof(int); }
  else if (w == "peel_cycles") { src = v.d_prof; bytes = sizeof(long long) * 8; }
  else if (w == "twiddle") { src = v.d_tw; bytes = sizeof(cplx) * ((1ll << v.log_twN) - 1); }
  else { set_error("sfftb_debug_fetch: unknown v3 array name"); return -1; }
  if (bytes > (long long)capacity) { set_error("sfftb_debug_fetch: destination too small"); return -1; }
  if (cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
    set_error("sfftb_debug_fetch: copy failed");
    return -1;
  }
  return bytes;
}

}  // namespace sfftb
