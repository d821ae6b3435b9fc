// v12_kernels.cu -- sm_100a kernels for sFFT v1/v2 (see v12_kernels.cuh).
#include <stdlib.h>
#include <string.h>

#include "v12_kernels.cuh"

namespace sfftb {

__device__ __forceinline__ long long loop_offset(const LoopGeom &g, int j)
{
  // cf12.cc:228-230
  return j < g.loops_loc
             ? ((long long)j << g.logB[0])
             : (((long long)g.loops_loc << g.logB[0]) + ((long long)(j - g.loops_loc) << g.logB[1]));
}

// ---------------------------------------------------------------------------
// K1  permuted windowed gather, folded into B buckets   (cf12.cc:222-261)
//
//   xs[j][b] = sum_{c} x[((c*B + b) * ai_j) mod n] * taps[c*B + b]
//
// One thread per (bucket, loop, signal).  The thread walks its bucket's taps in
// ascending i, which is the reference's accumulation order, so the folded sums
// are bit-identical; the ~w/B samples it needs are independent random 16-byte
// reads, issued U at a time (no_allocate: each 32-byte sector is touched once),
// while the taps stream coalesced and stay in L2 across loops.
// ---------------------------------------------------------------------------
constexpr int kGatherThreads = 256;

// One CTA per 256 buckets of one (loop, signal), CTAs handed out by the hardware scheduler.
// U = independent sample loads a thread issues back to back.  Measured (profiles/r02_gather_ab.md):
// the gather is bound by the random-request rate of the memory system; with every warp slot of
// the SM occupied, SHORT bursts per thread are faster when the signal lives in HBM -- U = 2
// beats U = 8 by 17-19 % at C2 and C4 -- while U = 8 wins when it is L2-resident (n <= 2^22:
// C1, C5).  Capping the resident CTAs, or a persistent statically strided form of the same
// loop, was slower at every shape; both were dropped.
// FILL64: ask L2 for 64-byte fills around a sample instead of the whole 128-byte line; the
// neighbours of a permuted sample are not wanted soon (halves the DRAM traffic, same or
// better time at every BASELINE shape).
template <bool FILL64, int U>
__global__ void __launch_bounds__(kGatherThreads)
gather_kernel(LoopGeom g, GatherArgs a)
{
  const int j = a.loop_begin + blockIdx.y;
  const int s = blockIdx.z;
  const bool est = j >= g.loops_loc;
  const int logB = est ? g.logB[1] : g.logB[0];
  const unsigned B = 1u << logB;
  const unsigned b = blockIdx.x * kGatherThreads + threadIdx.x;
  if (b >= B) return;
  const int w = est ? g.w[1] : g.w[0];
  const cplx *__restrict__ taps = est ? a.taps[1] : a.taps[0];
  const cplx *__restrict__ x =
      (a.x_indirect ? reinterpret_cast<const cplx *>(*a.x_indirect) : a.x) + (long long)s * a.x_stride;
  const unsigned ai = (unsigned)a.perm[(long long)s * perm_stride(g.loops) + g.loops + j];
  const unsigned mask = (unsigned)g.n_mask;
  unsigned idx = (b * ai) & mask;                    // n is a power of two <= 2^31
  const unsigned stepB = (B * ai) & mask;
  double acc_re = 0.0, acc_im = 0.0;
  for (unsigned i = b; i < (unsigned)w; i += U * B) {
    cplx xv[U], tv[U];
    unsigned id = idx;
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned ii = i + u * B;
      xv[u] = FILL64 ? ldg_stream64(x + id) : ldg_stream(x + id);
      tv[u] = __ldg(taps + (ii < (unsigned)w ? ii : 0u));
      id = (id + stepB) & mask;
    }
    idx = id;
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned ii = i + u * B;
      if (ii < (unsigned)w) {
        const cplx p = cmul_rn(xv[u], tv[u]);
        acc_re = __dadd_rn(acc_re, p.x);
        acc_im = __dadd_rn(acc_im, p.y);
      }
    }
  }
  cplx *xs = a.xs + (long long)s * g.x_samp_size + loop_offset(g, j);
  xs[bitrev(b, logB)] = make_double2(acc_re, acc_im);
}

int launch_gather(const LoopGeom &g, const GatherArgs &a, int nloops, int nsig, cudaStream_t st)
{
  if (nloops <= 0) return 0;
  const int maxlog = g.logB[0] > g.logB[1] ? g.logB[0] : g.logB[1];
  dim3 grid((unsigned)ceil_div(1ll << maxlog, kGatherThreads), (unsigned)nloops, (unsigned)nsig);
  int unroll = g.logn <= 22 ? 8 : 2;
  bool fill64 = true;
  size_t pad = 0;                                    // dynamic shared memory nobody uses: caps the CTAs per SM
  if (getenv("SFFTB_TUNE")) {
    // read at every launch so that one process can sweep them (tools/gather_sweep.py)
    if (const char *e = getenv("SFFTB_GATHER_UNROLL")) unroll = atoi(e);
    if (const char *e = getenv("SFFTB_GATHER_FILL")) fill64 = atoi(e) != 128;
    if (const char *e = getenv("SFFTB_GATHER_SMEM")) pad = (size_t)atoi(e);
    if (pad > 48 * 1024) {
      cudaFuncSetAttribute(gather_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(gather_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(gather_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
  }
  if (fill64) {
    if (unroll >= 8) gather_kernel<true, 8><<<grid, kGatherThreads, pad, st>>>(g, a);
    else if (unroll >= 4) gather_kernel<true, 4><<<grid, kGatherThreads, pad, st>>>(g, a);
    else gather_kernel<true, 2><<<grid, kGatherThreads, pad, st>>>(g, a);
  } else {
    if (unroll >= 8) gather_kernel<false, 8><<<grid, kGatherThreads, pad, st>>>(g, a);
    else gather_kernel<false, 4><<<grid, kGatherThreads, pad, st>>>(g, a);
  }
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// K3+K4  |.|^2 and top-num bucket selection   (cf12.cc:278-302, utils.cc:131-158)
//
// One CTA per (row, signal).  Squared magnitudes become order-preserving 64-bit
// keys held in shared memory (padded against bank conflicts); an MSD radix select
// (8 digits of 8 bits, run-length-aggregated shared atomics) finds the
// (num+1)-th largest key = the reference's cutoff.  Every thread owns a contiguous
// chunk of bucket indices, so ONE block-wide scan of (count > cutoff, count ==
// cutoff) places its selections: indices with key > cutoff and, if short, the first
// few with key == cutoff -- the reference's tie rule -- come out ascending, together
// with a B-bit membership bitmap for the voting stage.
// ---------------------------------------------------------------------------
constexpr int kSelectThreads = 1024;
constexpr int kSelectSmemKeys = 16384;   // rows up to this size keep their keys in shared memory
constexpr int kSelectCand = 1024;        // survivors of the radix passes finished from a compact list

__device__ __forceinline__ int key_slot(int i, bool padded) { return padded ? i + (i >> 4) : i; }

__global__ void __launch_bounds__(kSelectThreads)
select_kernel(SelectArgs a)
{
  extern __shared__ unsigned long long sel_smem[];
  __shared__ unsigned hist[256];
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long sh_prefix;
  __shared__ unsigned sh_K, sh_bin, sh_ncand;
  __shared__ unsigned long long cand[kSelectCand];

  const int row = a.row_begin + blockIdx.x * a.row_step;
  const int s = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = 1 << a.logB;
  const int words = B >= 32 ? B / 32 : 1;
  const bool in_smem = a.gkeys == nullptr;
  const cplx *__restrict__ src = a.xs + (long long)s * a.xs_stride + (long long)row * a.row_stride;
  // dynamic shared memory: [bitmap words (rounded to 8 B)] [padded keys, if they fit]
  unsigned *bm_s = reinterpret_cast<unsigned *>(sel_smem);
  unsigned long long *keys =
      in_smem ? sel_smem + (words + 1) / 2
              : a.gkeys + (long long)s * a.gk_sig_stride + (long long)row * B;

  for (int i = tid; i < B; i += kSelectThreads)
    keys[key_slot(i, in_smem)] = (unsigned long long)__double_as_longlong(cabs2_rn(src[i]));
  for (int i = tid; i < words; i += kSelectThreads) bm_s[i] = 0u;
  __syncthreads();

  // contiguous ownership: thread t holds indices [t*E, (t+1)*E)
  const int E = B >= kSelectThreads ? B / kSelectThreads : 1;
  const int lo = tid * E;
  const int mine = lo < B ? E : 0;

  unsigned long long prefix = 0;
  unsigned K = (unsigned)a.num + 1u;      // rank from the top of the wanted key
  bool compacted = false;                 // block-uniform
  for (int pass = 0; pass < 8; pass++) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    if (!compacted) {
      int run_d = -1;
      unsigned run_c = 0;
      for (int e = 0; e < mine; e++) {
        // the histogram does not care which keys a thread counts: shared-memory keys are
        // read chunk-wise (padded, conflict-free), global ones strided so warps coalesce
        const unsigned long long key = in_smem ? keys[key_slot(lo + e, true)] : keys[e * kSelectThreads + tid];
        const bool match = pass == 0 ? true : ((key >> (shift + 8)) == prefix);
        if (match) {
          const int d = (int)((key >> shift) & 255ull);
          if (d == run_d) {
            run_c++;
          } else {
            if (run_c) atomicAdd(&hist[run_d], run_c);
            run_d = d;
            run_c = 1;
          }
        }
      }
      if (run_c) atomicAdd(&hist[run_d], run_c);
    } else if (tid < (int)sh_ncand) {
      // few keys still share the prefix: they were copied out, one per thread
      const unsigned long long key = cand[tid];
      if ((key >> (shift + 8)) == prefix) atomicAdd(&hist[(int)((key >> shift) & 255ull)], 1u);
    }
    __syncthreads();
    if (warp == 0) {
      // lane l owns bins 255-8l .. 248-8l, i.e. lanes ascend as keys descend
      unsigned c8[8], tot = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        c8[q] = hist[255 - 8 * lane - q];
        tot += c8[q];
      }
      unsigned incl = tot;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      const unsigned excl = incl - tot;
      if (excl < K && K <= incl) {
        unsigned run = excl;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          if (K <= run + c8[q]) {
            sh_prefix = (prefix << 8) | (unsigned long long)(255 - 8 * lane - q);
            sh_K = K - run;
            sh_bin = c8[q];
            break;
          }
          run += c8[q];
        }
      }
    }
    __syncthreads();
    prefix = sh_prefix;
    K = sh_K;
    // once at most kSelectCand keys share the decided prefix, gather them and finish on those
    if (!compacted && pass < 7 && sh_bin <= (unsigned)kSelectCand) {
      if (tid == 0) sh_ncand = 0;
      __syncthreads();
      for (int e = 0; e < mine; e++) {
        const unsigned long long key = in_smem ? keys[key_slot(lo + e, true)] : keys[e * kSelectThreads + tid];
        if ((key >> shift) == prefix) cand[atomicAdd(&sh_ncand, 1u)] = key;
      }
      compacted = true;
      __syncthreads();
    }
  }
  const unsigned long long cutoff = prefix;
  const unsigned need = K - 1u;   // ties at the cutoff to admit, in index order

  // per-thread counts over the owned chunk, packed (gt << 32 | eq), then one block scan
  unsigned long long cnt = 0;
  for (int e = 0; e < mine; e++) {
    const unsigned long long key = keys[key_slot(lo + e, in_smem)];
    cnt += key > cutoff ? (1ull << 32) : (key == cutoff ? 1ull : 0ull);
  }
  unsigned long long incl = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const unsigned long long t = warp_tot[lane];
    unsigned long long sc = t;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, sc, off);
      if (lane >= off) sc += v;
    }
    warp_tot[lane] = sc - t;     // exclusive prefix of warp totals
  }
  __syncthreads();
  const unsigned long long before = warp_tot[warp] + incl - cnt;
  unsigned gt_b = (unsigned)(before >> 32), eq_b = (unsigned)(before & 0xffffffffu);

  int *J = a.J + (long long)s * a.J_sig_stride + (long long)row * a.num;
  for (int e = 0; e < mine; e++) {
    const int i = lo + e;
    const unsigned long long key = keys[key_slot(i, in_smem)];
    const bool f_gt = key > cutoff, f_eq = key == cutoff;
    if (f_gt || (f_eq && eq_b < need)) {
      J[gt_b + (eq_b < need ? eq_b : need)] = i;
      atomicOr(&bm_s[i >> 5], 1u << (i & 31));
    }
    gt_b += f_gt;
    eq_b += f_eq;
  }
  __syncthreads();
  unsigned *bm = a.bitmap + (long long)s * a.bm_sig_stride + (long long)row * words;
  for (int i = tid; i < words; i += kSelectThreads) bm[i] = bm_s[i];
}

// ---------------------------------------------------------------------------
// Rows above 16384 buckets (a v2 Comb spectrum of 2^17 points, location rows of 2^15 at
// n = 2^27): the same MSD radix select spread over B/1024 CTAs, one key per thread, as a
// chain of small kernels -- eight digit passes (block histogram in shared memory, flushed to
// a per-row global histogram; every CTA re-derives the prefix decided so far from the
// earlier passes' histograms, so no "decide" kernels are needed), a count pass and a
// placement pass.  One CTA doing all of it was 290 us at W = 2^17 (one SM's L2 bandwidth).
// Scratch per row: [8][256] u32 histogram, then B/1024 packed (greater << 32 | equal) counts.
// ---------------------------------------------------------------------------
constexpr int kBigThreads = 1024;
long long select_big_scratch(int B) { return 1024 + (B + kBigThreads - 1) / kBigThreads; }

struct BigRow {
  unsigned long long *keys;
  unsigned *hist;               // [8][256]
  unsigned long long *cnt;      // [G]
};
__device__ __forceinline__ BigRow big_row(const SelectArgs &a, int launched_row, int s)
{
  const int row = a.row_begin + launched_row * a.row_step;
  const long long B = 1ll << a.logB;
  unsigned long long *base = a.gkeys + (long long)s * a.gk_sig_stride;
  unsigned long long *scr = base + a.gk_scratch_off + (long long)row * (1024 + (B + kBigThreads - 1) / kBigThreads);
  BigRow r;
  r.keys = base + (long long)row * B;
  r.hist = reinterpret_cast<unsigned *>(scr);
  r.cnt = scr + 1024;
  return r;
}

// replay the decisions of digit passes [0, upto): the wanted key's prefix and its rank from
// the top among the keys sharing it.  The whole CTA first copies those passes' global
// histograms into shared memory (one round trip instead of `upto` dependent ones), then
// warp 0 scans them; result broadcast through shared memory (caller syncs afterwards).
__device__ __forceinline__ void big_decide(const unsigned *ghist, int upto, unsigned num,
                                           unsigned *hist, unsigned long long *sh_prefix, unsigned *sh_K)
{
  for (int t = threadIdx.x; t < upto * 256; t += blockDim.x) hist[t] = ghist[t];
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x & 31;
  unsigned long long prefix = 0;
  unsigned K = num + 1u;
  for (int q = 0; q < upto; q++) {
    // lane l owns bins 255-8l .. 248-8l, i.e. lanes ascend as keys descend
    unsigned c8[8], tot = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      c8[t] = hist[q * 256 + 255 - 8 * lane - t];
      tot += c8[t];
    }
    unsigned incl = tot;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    const unsigned excl = incl - tot;
    unsigned digit = 0, newK = 0;
    const bool owner = excl < K && K <= incl;
    if (owner) {
      unsigned run = excl;
#pragma unroll
      for (int t = 0; t < 8; t++) {
        if (newK == 0 && K <= run + c8[t]) { digit = 255u - 8u * lane - t; newK = K - run; }
        run += c8[t];
      }
    }
    const unsigned who = __ffs(__ballot_sync(0xffffffffu, owner)) - 1;
    digit = __shfl_sync(0xffffffffu, digit, who);
    K = __shfl_sync(0xffffffffu, newK, who);
    prefix = (prefix << 8) | digit;
  }
  if (lane == 0) { *sh_prefix = prefix; *sh_K = K; }
}

__global__ void __launch_bounds__(kBigThreads)
select_big_pass_kernel(SelectArgs a, int pass)
{
  __shared__ unsigned h[256];
  __shared__ unsigned prev[8 * 256];
  __shared__ unsigned long long sh_prefix;
  __shared__ unsigned sh_K;
  const BigRow r = big_row(a, blockIdx.y, blockIdx.z);
  const int i = blockIdx.x * kBigThreads + threadIdx.x;
  const int tid = threadIdx.x;
  if (tid < 256) h[tid] = 0;
  unsigned long long key;
  if (pass == 0) {
    const int row = a.row_begin + blockIdx.y * a.row_step;
    const cplx *__restrict__ src = a.xs + (long long)blockIdx.z * a.xs_stride + (long long)row * a.row_stride;
    key = (unsigned long long)__double_as_longlong(cabs2_rn(src[i]));
    r.keys[i] = key;
  } else {
    key = r.keys[i];
    big_decide(r.hist, pass, (unsigned)a.num, prev, &sh_prefix, &sh_K);
  }
  __syncthreads();
  const int shift = 56 - 8 * pass;
  if (pass == 0 || (key >> (shift + 8)) == sh_prefix) {
    // aggregate equal digits inside the warp before touching shared memory
    const unsigned d = (unsigned)((key >> shift) & 255ull);
    const unsigned peers = __match_any_sync(__activemask(), d);
    if ((int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&h[d], (unsigned)__popc(peers));
  }
  __syncthreads();
  if (tid < 256 && h[tid]) atomicAdd(&r.hist[pass * 256 + tid], h[tid]);
}

__global__ void __launch_bounds__(kBigThreads)
select_big_count_kernel(SelectArgs a)
{
  __shared__ unsigned prev[8 * 256];
  __shared__ unsigned long long sh_prefix;
  __shared__ unsigned sh_K;
  __shared__ unsigned long long wsum[32];
  const BigRow r = big_row(a, blockIdx.y, blockIdx.z);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long key = r.keys[blockIdx.x * kBigThreads + tid];
  big_decide(r.hist, 8, (unsigned)a.num, prev, &sh_prefix, &sh_K);
  __syncthreads();
  const unsigned long long cutoff = sh_prefix;
  unsigned long long c = key > cutoff ? (1ull << 32) : (key == cutoff ? 1ull : 0ull);
#pragma unroll
  for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if (lane == 0) wsum[warp] = c;
  __syncthreads();
  if (warp == 0) {
    unsigned long long t = wsum[lane];
#pragma unroll
    for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (lane == 0) r.cnt[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kBigThreads)
select_big_place_kernel(SelectArgs a)
{
  __shared__ unsigned prev[8 * 256];
  __shared__ unsigned long long sh_prefix;
  __shared__ unsigned sh_K;
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long sh_base;
  const BigRow r = big_row(a, blockIdx.y, blockIdx.z);
  const int row = a.row_begin + blockIdx.y * a.row_step;
  const int s = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x * kBigThreads + tid;
  const unsigned long long key = r.keys[i];
  // counts of the CTAs before this one
  if (warp == 1) {
    unsigned long long t = 0;
    for (int c = lane; c < (int)blockIdx.x; c += 32) t += r.cnt[c];
#pragma unroll
    for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (lane == 0) sh_base = t;
  }
  big_decide(r.hist, 8, (unsigned)a.num, prev, &sh_prefix, &sh_K);
  __syncthreads();
  const unsigned long long cutoff = sh_prefix;
  const unsigned need = sh_K - 1u;          // ties at the cutoff to admit, in index order
  const bool f_gt = key > cutoff, f_eq = key == cutoff;
  const unsigned long long mine = f_gt ? (1ull << 32) : (f_eq ? 1ull : 0ull);
  unsigned long long incl = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const unsigned long long t = wsum[lane];
    unsigned long long sc = t;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, sc, off);
      if (lane >= off) sc += v;
    }
    wsum[lane] = sc - t;
  }
  __syncthreads();
  const unsigned long long before = sh_base + wsum[warp] + incl - mine;
  const unsigned gt_b = (unsigned)(before >> 32), eq_b = (unsigned)(before & 0xffffffffu);
  const bool take = f_gt || (f_eq && eq_b < need);
  if (take) {
    int *J = a.J + (long long)s * a.J_sig_stride + (long long)row * a.num;
    J[gt_b + (eq_b < need ? eq_b : need)] = i;
  }
  // a warp covers 32 consecutive buckets = one bitmap word
  const unsigned word = __ballot_sync(0xffffffffu, take);
  if (lane == 0) {
    const int words = 1 << (a.logB - 5);
    a.bitmap[(long long)s * a.bm_sig_stride + (long long)row * words + (i >> 5)] = word;
  }
}

__global__ void select_big_zero_kernel(SelectArgs a)
{
  const BigRow r = big_row(a, blockIdx.x, blockIdx.y);
  for (int t = threadIdx.x; t < 8 * 256; t += blockDim.x) r.hist[t] = 0u;
}

static int launch_select_big(const SelectArgs &a, int nrows, int nsig, cudaStream_t st)
{
  const int B = 1 << a.logB;
  const dim3 grid((unsigned)(B / kBigThreads), (unsigned)nrows, (unsigned)nsig);
  select_big_zero_kernel<<<dim3((unsigned)nrows, (unsigned)nsig), 256, 0, st>>>(a);
  SFFTB_LAUNCH_CHECK();
  for (int pass = 0; pass < 8; pass++) {
    select_big_pass_kernel<<<grid, kBigThreads, 0, st>>>(a, pass);
    SFFTB_LAUNCH_CHECK();
  }
  select_big_count_kernel<<<grid, kBigThreads, 0, st>>>(a);
  SFFTB_LAUNCH_CHECK();
  select_big_place_kernel<<<grid, kBigThreads, 0, st>>>(a);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// Rows of 2^15 .. 2^18 buckets in ONE kernel: a thread-block cluster per row, every CTA
// keeping 16384 keys of the row in its own shared memory.  The MSD radix select runs as in
// select_kernel; the only cross-CTA steps go through distributed shared memory -- each pass
// adds the CTAs' digit histograms into CTA 0's (red.shared::cluster), one cluster barrier,
// and everybody reads the merged histogram back; once at most kSelectCand keys share the
// decided prefix they are gathered into CTA 0, which finishes alone and publishes the
// cutoff; the ordered placement needs the (greater, equal) counts of the lower-ranked CTAs.
// ~8 cluster barriers instead of a chain of 11 kernels (60 us -> see profiles/r02_*).
// ---------------------------------------------------------------------------
constexpr int kClusterKeys = 16384;     // keys per CTA
constexpr int kMaxSelectCluster = 16;

__device__ __forceinline__ unsigned map_to_rank(const void *smem_ptr, unsigned rank)
{
  unsigned local = (unsigned)__cvta_generic_to_shared(smem_ptr), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
  return remote;
}
__device__ __forceinline__ void dsmem_red_add(unsigned addr, unsigned v)
{
  asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned dsmem_atom_add(unsigned addr, unsigned v)
{
  unsigned old;
  asm volatile("atom.relaxed.cluster.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned dsmem_ld_u32(unsigned addr)
{
  unsigned v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long dsmem_ld_u64(unsigned addr)
{
  unsigned long long v;
  asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void dsmem_st_u64(unsigned addr, unsigned long long v)
{
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void cluster_barrier()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kSelectThreads)
select_cluster_kernel(SelectArgs a, int csize)
{
  extern __shared__ unsigned long long selc_keys[];        // [kClusterKeys, padded]
  __shared__ unsigned hist[256];
  __shared__ unsigned merged[3][256];                       // CTA 0's copies are the cluster's (rotating)
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long sh_prefix;
  __shared__ unsigned sh_K, sh_bin, sh_ncand;
  __shared__ unsigned long long cand[kSelectCand];          // CTA 0's
  __shared__ unsigned long long cta_cnt[kMaxSelectCluster]; // CTA 0's: packed (greater << 32 | equal) per CTA
  __shared__ unsigned long long fin[2];                     // CTA 0's: cutoff, ties to admit

  const unsigned rank = blockIdx.x % (unsigned)csize;
  const int row = a.row_begin + (int)(blockIdx.x / (unsigned)csize) * a.row_step;
  const int s = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = 1 << a.logB;
  const cplx *__restrict__ src =
      a.xs + (long long)s * a.xs_stride + (long long)row * a.row_stride + (long long)rank * kClusterKeys;
  unsigned long long *keys = selc_keys;
  for (int i = tid; i < kClusterKeys; i += kSelectThreads)
    keys[key_slot(i, true)] = (unsigned long long)__double_as_longlong(cabs2_rn(src[i]));
  if (tid < 256) { merged[0][tid] = 0u; merged[1][tid] = 0u; merged[2][tid] = 0u; }
  if (tid == 0) sh_ncand = 0u;
  cluster_barrier();

  const unsigned merged0 = map_to_rank(&merged[0][0], 0);
  const int E = kClusterKeys / kSelectThreads;               // contiguous ownership: [tid*E, (tid+1)*E)
  const int lo = tid * E;
  unsigned long long prefix = 0;
  unsigned K = (unsigned)a.num + 1u;
  bool solo = false;               // candidates gathered in CTA 0, which finishes alone
  int pass = 0;
  for (; pass < 8 && !solo; pass++) {
    const int shift = 56 - 8 * pass;
    const unsigned mcur = merged0 + (unsigned)(pass % 3) * 256u * 4u;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    {
      int run_d = -1;
      unsigned run_c = 0;
      for (int e = 0; e < E; e++) {
        const unsigned long long key = keys[key_slot(lo + e, true)];
        if (pass == 0 || (key >> (shift + 8)) == prefix) {
          const int d = (int)((key >> shift) & 255ull);
          if (d == run_d) {
            run_c++;
          } else {
            if (run_c) atomicAdd(&hist[run_d], run_c);
            run_d = d;
            run_c = 1;
          }
        }
      }
      if (run_c) atomicAdd(&hist[run_d], run_c);
    }
    __syncthreads();
    if (tid < 256) {
      if (hist[tid]) dsmem_red_add(mcur + (unsigned)tid * 4u, hist[tid]);
      // next pass's accumulator: last read in pass-2, i.e. before every CTA's previous barrier
      if (rank == 0) merged[(pass + 1) % 3][tid] = 0u;
    }
    cluster_barrier();
    if (warp == 0) {
      // lane l owns bins 255-8l .. 248-8l, i.e. lanes ascend as keys descend
      unsigned c8[8], tot = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        c8[q] = dsmem_ld_u32(mcur + (unsigned)(255 - 8 * lane - q) * 4u);
        tot += c8[q];
      }
      unsigned incl = tot;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      const unsigned excl = incl - tot;
      if (excl < K && K <= incl) {
        unsigned run = excl;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          if (K <= run + c8[q]) {
            sh_prefix = (prefix << 8) | (unsigned long long)(255 - 8 * lane - q);
            sh_K = K - run;
            sh_bin = c8[q];
            break;
          }
          run += c8[q];
        }
      }
    }
    __syncthreads();
    prefix = sh_prefix;
    K = sh_K;
    if (pass < 7 && sh_bin <= (unsigned)kSelectCand) {
      // few keys still share the prefix: gather them in CTA 0
      const unsigned ncand0 = map_to_rank(&sh_ncand, 0), cand0 = map_to_rank(&cand[0], 0);
      for (int e = 0; e < E; e++) {
        const unsigned long long key = keys[key_slot(lo + e, true)];
        if ((key >> shift) == prefix) dsmem_st_u64(cand0 + dsmem_atom_add(ncand0, 1u) * 8u, key);
      }
      solo = true;
    }
  }
  if (solo) {
    cluster_barrier();                  // the candidates have landed in CTA 0
    if (rank == 0) {
      for (; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        if (tid < (int)sh_ncand) {
          const unsigned long long key = cand[tid];
          if ((key >> (shift + 8)) == prefix) atomicAdd(&hist[(int)((key >> shift) & 255ull)], 1u);
        }
        __syncthreads();
        if (warp == 0) {
          unsigned c8[8], tot = 0;
#pragma unroll
          for (int q = 0; q < 8; q++) {
            c8[q] = hist[255 - 8 * lane - q];
            tot += c8[q];
          }
          unsigned incl = tot;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
          }
          const unsigned excl = incl - tot;
          if (excl < K && K <= incl) {
            unsigned run = excl;
#pragma unroll
            for (int q = 0; q < 8; q++) {
              if (K <= run + c8[q]) {
                sh_prefix = (prefix << 8) | (unsigned long long)(255 - 8 * lane - q);
                sh_K = K - run;
                break;
              }
              run += c8[q];
            }
          }
        }
        __syncthreads();
        prefix = sh_prefix;
        K = sh_K;
      }
    }
  }
  if (rank == 0 && tid == 0) { fin[0] = prefix; fin[1] = (unsigned long long)(K - 1u); }
  cluster_barrier();
  const unsigned fin0 = map_to_rank(&fin[0], 0);
  const unsigned long long cutoff = dsmem_ld_u64(fin0);
  const unsigned need = (unsigned)dsmem_ld_u64(fin0 + 8u);   // ties at the cutoff to admit, in index order

  // per-thread counts over the owned chunk, packed (gt << 32 | eq); block scan; CTA totals to CTA 0
  unsigned long long cnt = 0;
  for (int e = 0; e < E; e++) {
    const unsigned long long key = keys[key_slot(lo + e, true)];
    cnt += key > cutoff ? (1ull << 32) : (key == cutoff ? 1ull : 0ull);
  }
  unsigned long long incl = cnt;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const unsigned long long t = warp_tot[lane];
    unsigned long long sc = t;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, sc, off);
      if (lane >= off) sc += v;
    }
    warp_tot[lane] = sc - t;     // exclusive prefix of warp totals
    if (lane == 31) dsmem_st_u64(map_to_rank(&cta_cnt[0], 0) + rank * 8u, sc);
  }
  cluster_barrier();
  unsigned long long base = 0;
  {
    const unsigned cnt0 = map_to_rank(&cta_cnt[0], 0);
    for (unsigned q = 0; q < rank; q++) base += dsmem_ld_u64(cnt0 + q * 8u);
  }
  const unsigned long long before = base + warp_tot[warp] + incl - cnt;
  unsigned gt_b = (unsigned)(before >> 32), eq_b = (unsigned)(before & 0xffffffffu);
  const int words = B / 32;
  int *J = a.J + (long long)s * a.J_sig_stride + (long long)row * a.num;
  unsigned *bm = a.bitmap + (long long)s * a.bm_sig_stride + (long long)row * words + rank * (kClusterKeys / 32);
  // E = 16 consecutive buckets per thread: two threads share a bitmap word
  unsigned bits = 0;
  for (int e = 0; e < E; e++) {
    const int i = lo + e;
    const unsigned long long key = keys[key_slot(i, true)];
    const bool f_gt = key > cutoff, f_eq = key == cutoff;
    if (f_gt || (f_eq && eq_b < need)) {
      J[gt_b + (eq_b < need ? eq_b : need)] = (int)(rank * kClusterKeys) + i;
      bits |= 1u << (i & 31);
    }
    gt_b += f_gt;
    eq_b += f_eq;
  }
  const unsigned other = __shfl_xor_sync(0xffffffffu, bits, 1);
  if ((tid & 1) == 0) bm[lo >> 5] = bits | other;
  // no CTA may exit while another still reads its shared memory
  cluster_barrier();
}

static size_t select_cluster_smem_bytes() { return (size_t)(kClusterKeys + (kClusterKeys >> 4) + 1) * 8; }

// one cluster of B/16384 CTAs per row; false if this device will not take the launch
static bool launch_select_cluster(const SelectArgs &a, int nrows, int nsig, cudaStream_t st)
{
  const int B = 1 << a.logB;
  const int csize = B / kClusterKeys;
  if (csize < 2 || csize > kMaxSelectCluster) return false;
  static int usable = -1;     // -1 unknown, 0 no, 1 yes
  if (usable == 0) return false;
  const size_t smem = select_cluster_smem_bytes();
  if (usable < 0) {
    if (cudaFuncSetAttribute(select_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(select_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      usable = 0;
      return false;
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)(nrows * csize), (unsigned)nsig);
  cfg.blockDim = dim3(kSelectThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (usable < 0) {
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, select_cluster_kernel, &cfg) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      usable = 0;
      return false;
    }
    usable = 1;
  }
  if (cudaLaunchKernelEx(&cfg, select_cluster_kernel, a, csize) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  g_launches++;
  return true;
}

static size_t select_smem_bytes(int B, bool keys_in_smem)
{
  const int words = B >= 32 ? B / 32 : 1;
  size_t bytes = (size_t)((words + 1) / 2) * 8;
  if (keys_in_smem) bytes += (size_t)(B + (B >> 4) + 1) * 8;
  return bytes;
}

int launch_select(const SelectArgs &a, int nrows, int nsig, cudaStream_t st)
{
  if (nrows <= 0) return 0;
  const int B = 1 << a.logB;
  if (B > kSelectSmemKeys) {
    if (!a.gkeys) {
      set_error("launch_select: rows above 16384 buckets need the global key scratch");
      return -1;
    }
    if (!getenv("SFFTB_NO_SELECT_CLUSTER") && launch_select_cluster(a, nrows, nsig, st)) return 0;
    return launch_select_big(a, nrows, nsig, st);
  }
  SFFTB_ONCE_PER_DEVICE(SFFTB_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                        (int)select_smem_bytes(kSelectSmemKeys, true))));
  dim3 grid((unsigned)nrows, (unsigned)nsig);
  select_kernel<<<grid, kSelectThreads, select_smem_bytes(B, a.gkeys == nullptr), st>>>(a);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// K5  reverse-hash voting without a dense score array   (cf12.cc:92-116, :126-184)
//
// A frequency loc is voted by location loop j iff the bucket its permuted image
// ai_j*loc falls into (shifted by half a bucket) is in J_j -- a bit test on a
// B-bit map.  So score(loc) is a sum of loops_loc bit tests, and the hit list is
// {loc : score >= threshold}.  Candidates are generated from the selected
// buckets of the first (loops_loc - threshold + 1) loops (a hit must receive its
// first vote there); a candidate is emitted by the loop that votes for it first,
// which removes duplicates without any table.  v2 adds one more bit test:
// loc mod W_Comb must be a Comb-approved residue.
// ---------------------------------------------------------------------------
// One CTA per selected bucket (signal s, loop j, entry Ji of J_j): its threads walk the n/B
// permuted positions of that bucket (cf12.cc:102-115).  The B-bit maps of all location loops
// sit in shared memory (loops_loc * B/8 bytes: 16 KiB at n = 2^27), the per-loop multipliers in
// registers, and nothing in the loop divides: 64-bit divisions to recover (loop, entry) from
// a flat index were most of the old kernel's instructions.
constexpr int kVoteThreads = 256;
constexpr int kVoteMaxLoops = 8;        // location loops kept in registers; more -> generic path
constexpr int kVoteMaxEntries = 16;     // selected buckets one CTA may own
constexpr int kVoteHitCap = 2048;       // hits a CTA collects in shared memory before it falls back to global atomics

// AGGREGATE: collect the CTA's hits in shared memory and append them with one global atomic
// (batches: thousands of atomics on one signal's counter serialise in L2); a single signal's
// short CTAs are better off with direct atomics.
// NLOC = loops_loc at compile time: the bit tests of the other location loops unroll exactly
// (with a run-time count the eight-way unrolled loop issued every predicated-off test too:
// ~170 instructions per candidate, the kernel was 83 % issue-bound).
template <bool SMEM_MAPS, bool AGGREGATE, int NLOC>
__global__ void __launch_bounds__(kVoteThreads)
vote_kernel(LoopGeom g, VoteArgs a, int first_loops, int entries_per_cta)
{
  extern __shared__ unsigned vote_bm[];          // [loops_loc][words] when SMEM_MAPS
  // hits of this CTA, appended to the global list with ONE atomic: thousands of atomics on one
  // signal's counter serialise in L2 (~3 ns each), which was most of a batch's voting time
  __shared__ int s_hits[AGGREGATE ? kVoteHitCap : 1];
  __shared__ int s_nhit, s_base;
  __shared__ unsigned s_low[kVoteMaxEntries];    // first permuted position of each entry's bucket
  __shared__ unsigned s_aj[kVoteMaxEntries];     // a_j of the entry's loop
  const int s = blockIdx.y;
  const int logB = g.logB[0];
  const int logseg = g.logn - logB;
  const unsigned seg = 1u << logseg, half = seg >> 1;
  const unsigned mask = (unsigned)g.n_mask, Bm = (1u << logB) - 1u;
  const int words = logB >= 5 ? (1 << (logB - 5)) : 1;
  constexpr int L = NLOC;
  const unsigned *gbm = a.bitmap + (long long)s * a.bm_sig_stride;
  if (SMEM_MAPS)
    for (int i = threadIdx.x; i < L * words; i += kVoteThreads) vote_bm[i] = gbm[i];
  const int total_entries = first_loops * a.num;
  const int e0 = blockIdx.x * entries_per_cta;
  const int e1 = e0 + entries_per_cta < total_entries ? e0 + entries_per_cta : total_entries;
  if ((int)threadIdx.x < e1 - e0) {
    // J is [loops_loc][num]; entry e of the flattened list is J[j][e - j*num], i.e. J[e]
    const int e = e0 + (int)threadIdx.x;
    const unsigned Jv = (unsigned)a.J[(long long)s * a.J_sig_stride + e];
    // cf12.cc:102: low = ceil((J - 0.5) * n/B) mod n  == J*seg - seg/2 (exact for seg >= 2)
    s_low[threadIdx.x] = ((Jv << logseg) - half) & mask;
    s_aj[threadIdx.x] = (unsigned)a.perm[(long long)s * perm_stride(g.loops) + e / a.num];
  }
  if (threadIdx.x == 0) s_nhit = 0;
  __syncthreads();
  const unsigned *bm = SMEM_MAPS ? vote_bm : gbm;
  unsigned ai[L];
#pragma unroll
  for (int q = 0; q < L; q++) ai[q] = (unsigned)a.perm[(long long)s * perm_stride(g.loops) + g.loops + q];
  const unsigned *cb = a.comb_bitmap ? a.comb_bitmap + (long long)s * a.comb_sig_stride : nullptr;

  // this CTA's (loop, entry) pairs, flattened with their n/B positions: every thread walks
  // candidates of several entries, whose constants were fetched side by side above
  const unsigned ncand = (unsigned)(e1 - e0) << logseg;
  int j = e0 / a.num;                              // loops of consecutive entries only ever increase
  for (unsigned c = threadIdx.x; c < ncand; c += kVoteThreads) {
    const int le = (int)(c >> logseg);
    const unsigned t = c & (seg - 1u);
    while (e0 + le >= (j + 1) * a.num) j++;
    const unsigned p = (s_low[le] + t) & mask;
    const unsigned loc = (s_aj[le] * p) & mask;                   // n is a power of two <= 2^31
    const unsigned rres = loc & (unsigned)a.W_mask;              // v2: loc mod W_Comb must be approved
    if (cb && !((__ldg(&cb[rres >> 5]) >> (rres & 31u)) & 1u)) continue;
    // emitted by the first loop that votes for it; score = votes over all location loops
    bool earlier = false;
    int score = 1;
#pragma unroll
    for (int q = 0; q < L; q++) {
      if (q != j) {
        const unsigned Jb = ((((ai[q] * loc) & mask) + half) >> logseg) & Bm;
        const bool v = (bm[q * words + (Jb >> 5)] >> (Jb & 31u)) & 1u;
        if (q < j) earlier |= v;
        else score += v ? 1 : 0;
      }
    }
    if (!earlier && score >= a.thresh) {
      const int slot = AGGREGATE ? atomicAdd(&s_nhit, 1) : kVoteHitCap;
      if (slot < kVoteHitCap) {
        s_hits[slot] = (int)loc;
      } else {
        const int pos = atomicAdd(&a.count[s], 1);
        if (pos < a.hits_cap) a.hits[(long long)s * a.hits_cap + pos] = (int)loc;
      }
    }
  }
  if (!AGGREGATE) return;
  __syncthreads();
  const int mine = s_nhit < kVoteHitCap ? s_nhit : kVoteHitCap;
  if (threadIdx.x == 0 && mine > 0) s_base = atomicAdd(&a.count[s], mine);
  __syncthreads();
  for (int i = threadIdx.x; i < mine; i += kVoteThreads) {
    const int pos = s_base + i;
    if (pos < a.hits_cap) a.hits[(long long)s * a.hits_cap + pos] = s_hits[i];
  }
}

// more than kVoteMaxLoops location loops (no table of the reference has that many)
__global__ void __launch_bounds__(kVoteThreads)
vote_generic_kernel(LoopGeom g, VoteArgs a, int first_loops)
{
  const int s = blockIdx.y;
  const int logB = g.logB[0];
  const int logseg = g.logn - logB;
  const unsigned seg = 1u << logseg, half = seg >> 1;
  const unsigned mask = (unsigned)g.n_mask, Bm = (1u << logB) - 1u;
  const int words = logB >= 5 ? (1 << (logB - 5)) : 1;
  const unsigned *bm = a.bitmap + (long long)s * a.bm_sig_stride;
  const int *perm = a.perm + (long long)s * perm_stride(g.loops);
  const unsigned *cb = a.comb_bitmap ? a.comb_bitmap + (long long)s * a.comb_sig_stride : nullptr;
  for (int e = blockIdx.x; e < first_loops * a.num; e += gridDim.x) {
    const int j = e / a.num, Ji = e - j * a.num;
    const unsigned Jv = (unsigned)a.J[(long long)s * a.J_sig_stride + (long long)j * a.num + Ji];
    const unsigned low = ((Jv << logseg) - half) & mask;
    const unsigned aj = (unsigned)perm[j];
    for (unsigned t = threadIdx.x; t < seg; t += kVoteThreads) {
      const unsigned loc = (aj * ((low + t) & mask)) & mask;
      const unsigned rres = loc & (unsigned)a.W_mask;
      if (cb && !((__ldg(&cb[rres >> 5]) >> (rres & 31u)) & 1u)) continue;
      bool earlier = false;
      int score = 1;
      for (int q = 0; q < g.loops_loc; q++) {
        if (q == j) continue;
        const unsigned Jb = (((((unsigned)perm[g.loops + q] * loc) & mask) + half) >> logseg) & Bm;
        const bool v = (__ldg(&bm[q * words + (Jb >> 5)]) >> (Jb & 31u)) & 1u;
        if (q < j) earlier |= v;
        else score += v ? 1 : 0;
      }
      if (!earlier && score >= a.thresh) {
        const int pos = atomicAdd(&a.count[s], 1);
        if (pos < a.hits_cap) a.hits[(long long)s * a.hits_cap + pos] = (int)loc;
      }
    }
  }
}

int launch_vote(const LoopGeom &g, const VoteArgs &a, int nsig, cudaStream_t st)
{
  const int first_loops = g.loops_loc - a.thresh + 1;
  if (first_loops <= 0) return 0;
  const int words = g.logB[0] >= 5 ? (1 << (g.logB[0] - 5)) : 1;
  const long long entries = (long long)first_loops * a.num;
  if (g.loops_loc > kVoteMaxLoops) {
    long long blocks = entries;
    const long long share = 148ll * 64 / (nsig < 64 ? nsig : 64);
    const long long cap = share > 148 ? share : 148;
    if (blocks > cap) blocks = cap;
    vote_generic_kernel<<<dim3((unsigned)blocks, (unsigned)nsig), kVoteThreads, 0, st>>>(g, a, first_loops);
  } else {
    // voting is latency-bound per CTA: as many CTAs as there are entries, until the grid is a
    // few waves deep; beyond that (batches) a CTA takes up to 8192 candidates' worth of entries
    const int logseg = g.logn - g.logB[0];
    long long per_cta = entries * nsig / 600;
    long long most = logseg >= 12 ? 1 : (8192 >> logseg);
    if (most > kVoteMaxEntries) most = kVoteMaxEntries;
    if (per_cta > most) per_cta = most;
    if (per_cta < 1) per_cta = 1;
    const long long blocks = (entries + per_cta - 1) / per_cta;
    const dim3 grid((unsigned)blocks, (unsigned)nsig);
    const size_t smem = sizeof(unsigned) * (size_t)g.loops_loc * words;
    const bool maps = smem <= 32 * 1024, agg = nsig > 1;
    const size_t dyn = maps ? smem : 0;
#define SFFTB_VOTE_CASE(N)                                                                                        \
  case N:                                                                                                         \
    if (maps && agg) vote_kernel<true, true, N><<<grid, kVoteThreads, dyn, st>>>(g, a, first_loops, (int)per_cta);  \
    else if (maps) vote_kernel<true, false, N><<<grid, kVoteThreads, dyn, st>>>(g, a, first_loops, (int)per_cta);   \
    else if (agg) vote_kernel<false, true, N><<<grid, kVoteThreads, dyn, st>>>(g, a, first_loops, (int)per_cta);    \
    else vote_kernel<false, false, N><<<grid, kVoteThreads, dyn, st>>>(g, a, first_loops, (int)per_cta);            \
    break;
    switch (g.loops_loc) {
      SFFTB_VOTE_CASE(1) SFFTB_VOTE_CASE(2) SFFTB_VOTE_CASE(3) SFFTB_VOTE_CASE(4)
      SFFTB_VOTE_CASE(5) SFFTB_VOTE_CASE(6) SFFTB_VOTE_CASE(7) SFFTB_VOTE_CASE(8)
    }
#undef SFFTB_VOTE_CASE
  }
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// K7  per-hit estimation: median over loops of bucket / filter response
// (cf12.cc:341-419).  Arithmetic follows the reference instruction by instruction,
// including the sign it actually applies to the imaginary part (:388-392 form
// (a*d - b*c), not (b*c - a*d)).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double select_rank(const double *v, int cnt, int want)
{
  // value with `want` elements before it in ascending order (stable on ties)
  for (int i = 0; i < cnt; i++) {
    const double vi = v[i];
    int rank = 0;
    for (int q = 0; q < cnt; q++) rank += (v[q] < vi) || (v[q] == vi && q < i);
    if (rank == want) return vi;
  }
  return v[0];
}

// median_select networks (generated): MedianNet<L>::run(v) = v[(L-1)/2] of sorted v
#include "median_networks.inc"

// one (hit, loop) term: bucket / filter response, as the reference executes it
__device__ __forceinline__ void estimate_term(const LoopGeom &g, const EstimateArgs &a,
                                              const cplx *__restrict__ xs, unsigned ai, unsigned loc,
                                              int j, double &out_re, double &out_im)
{
  const bool est = j >= g.loops_loc;
  const int logB = est ? g.logB[1] : g.logB[0];
  const int logseg = g.logn - logB;
  const int seg = 1 << logseg;
  const unsigned pos = (unsigned)(((unsigned long long)ai * loc) & (unsigned)g.n_mask);   // cf12.cc:370
  unsigned bucket = pos >> logseg;
  int dist = (int)(pos & (unsigned)(seg - 1));
  if (dist > seg / 2) {                                                                   // :373-377
    bucket = (bucket + 1) & ((1u << logB) - 1u);
    dist -= seg;
  }
  const cplx sv = xs[loop_offset(g, j) + bucket];
  const cplx *__restrict__ fw = est ? a.fwin[1] : a.fwin[0];
  const cplx f = __ldg(&fw[(est ? a.fw_half[1] : a.fw_half[0]) - dist]);   // freq[(n - dist) % n], :378
  const double ac = __dmul_rn(sv.x, f.x), bd = __dmul_rn(sv.y, f.y);
  const double ad = __dmul_rn(sv.x, f.y), bc = __dmul_rn(sv.y, f.x);
  const double den = __dadd_rn(__dmul_rn(f.x, f.x), __dmul_rn(f.y, f.y));
  out_re = __ddiv_rn(__dadd_rn(ac, bd), den);
  out_im = __ddiv_rn(__dsub_rn(ad, bc), den);              // :390-392: (a*d) + (-(b*c))
}

// a / b with y = RN(1/b) precomputed: two Markstein correction steps give the
// correctly rounded quotient (== __ddiv_rn) whenever nothing under/overflows; outside
// a wide safe exponent band, and for a == 0 (sign of zero), take the real division.
// Checked against __ddiv_rn by sfftb_debug_div_check (tests/test_gpu_stages.py).
__device__ __forceinline__ double div_by_rcp_rn(double a, double b, double y)
{
  const unsigned ea = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
  const unsigned eb = ((unsigned)__double2hiint(b) >> 20) & 0x7ffu;
  if (ea - 640u < 768u && eb - 640u < 768u) {      // 2^-383 <= |a|,|b| < 2^385
    double q = __dmul_rn(a, y);
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    r = __fma_rn(-b, q, a);
    q = __fma_rn(r, y, q);
    return q;
  }
  return __ddiv_rn(a, b);
}

// (den, RN(1/den)) for every entry of a filter-response window
__global__ void filter_den_kernel(const cplx *__restrict__ fwin, int len, double2 *fdr)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const cplx f = fwin[i];
  const double den = __dadd_rn(__dmul_rn(f.x, f.x), __dmul_rn(f.y, f.y));   // cf12.cc:394-396
  fdr[i] = make_double2(den, __drcp_rn(den));
}

int launch_filter_den(const cplx *fwin, int len, double2 *fdr, cudaStream_t st)
{
  filter_den_kernel<<<ceil_div(len, 256), 256, 0, st>>>(fwin, len, fdr);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

__global__ void div_check_kernel(unsigned long long seed, long long count, unsigned long long *mismatch)
{
  unsigned long long bad = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    // splitmix64 -> two doubles with random mantissas, exponents spread over +-40 binades
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    unsigned long long r[2];
    for (int q = 0; q < 2; q++) {
      z += 0x9E3779B97F4A7C15ull;
      unsigned long long t = z;
      t = (t ^ (t >> 30)) * 0xBF58476D1CE4E5B9ull;
      t = (t ^ (t >> 27)) * 0x94D049BB133111EBull;
      r[q] = t ^ (t >> 31);
    }
    const int ex_a = (int)((r[0] >> 52) % 81) - 40, ex_b = (int)((r[1] >> 52) % 81) - 40;
    unsigned long long ma = r[0] & 0xFFFFFFFFFFFFFull, mb = r[1] & 0xFFFFFFFFFFFFFull;
    if ((i & 15) == 0) mb = (i & 16) ? 0xFFFFFFFFFFFFFull : 0ull;       // all-ones / power-of-two divisors
    if ((i & 31) == 1) ma = mb;                                          // equal mantissas
    double a = __longlong_as_double((long long)(((unsigned long long)(1023 + ex_a) << 52) | ma));
    const double b = __longlong_as_double((long long)(((unsigned long long)(1023 + ex_b) << 52) | mb));
    if (r[0] >> 63) a = -a;
    const double y = __drcp_rn(b);
    const double q1 = div_by_rcp_rn(a, b, y), q2 = __ddiv_rn(a, b);
    bad += __double_as_longlong(q1) != __double_as_longlong(q2);
  }
  if (bad) atomicAdd(mismatch, bad);
}

long long run_div_check(unsigned long long seed, long long count)
{
  unsigned long long *d = nullptr, h = 0;
  if (cudaMalloc(&d, 8) != cudaSuccess) return -1;
  cudaMemset(d, 0, 8);
  div_check_kernel<<<148 * 8, 256>>>(seed, count, d);
  g_launches++;
  if (cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(d); return -1; }
  cudaFree(d);
  return (long long)h;
}

// Estimation kernel, L loops known at compile time.  Two adjacent lanes share one
// hit: the even lane carries the L real parts, the odd lane the L imaginary parts;
// each keeps its L quotients in registers and runs one median network.  Both lanes
// read the same bucket / filter entries (one memory transaction per pair).
template <int L>
__global__ void __launch_bounds__(256, 2)
estimate_pair_kernel(LoopGeom g, EstimateArgs a, long long max_per_sig)
{
  const int s = blockIdx.y;
  long long total;
  int nc = 1;
  if (a.approved) {
    nc = a.num_comb[s];
    total = (long long)nc * a.n_over_W;
    if (nc < 1) nc = 1;
  } else {
    total = a.count[s];
  }
  if (total > max_per_sig) total = max_per_sig;
  // multi-GPU slice of the list: [lo, lo + total)
  long long lo = 0;
  if (a.slice_world > 1) {
    lo = total * a.slice_rank / a.slice_world;
    total = total * (a.slice_rank + 1) / a.slice_world - lo;
  }
  if (a.slice_count && blockIdx.x == 0 && threadIdx.x == 0) a.slice_count[s] = (int)total;
  const int *__restrict__ perm = a.perm + (long long)s * perm_stride(g.loops) + g.loops;   // ai[]
  const cplx *__restrict__ xs = a.xs + (long long)s * a.xs_stride;
  const bool imag = threadIdx.x & 1;
  const int pair = threadIdx.x >> 1;
  const unsigned mask = (unsigned)g.n_mask;

  for (long long base = (long long)blockIdx.x * 128; base < total; base += (long long)gridDim.x * 128) {
    const long long h = base + pair;
    const bool active = h < total;
    const long long hc = lo + (active ? h : total - 1);
    unsigned loc;
    if (a.approved) {
      const long long jj = hc / nc;
      const int i = (int)(hc - jj * nc);
      loc = (unsigned)(jj * a.W + __ldg(&a.approved[(long long)s * a.approved_stride + i]));   // cf12.cc:508-511
    } else {
      loc = (unsigned)a.hits[(long long)s * a.hits_cap + hc];
    }
    double v[L];
#pragma unroll
    for (int j = 0; j < L; j++) {
      const bool est = j >= g.loops_loc;
      const int logB = est ? g.logB[1] : g.logB[0];
      const int logseg = g.logn - logB;
      const int seg = 1 << logseg;
      const unsigned pos = (unsigned)(((unsigned long long)(unsigned)__ldg(&perm[j]) * loc) & mask);   // cf12.cc:370
      unsigned bucket = pos >> logseg;
      int dist = (int)(pos & (unsigned)(seg - 1));
      if (dist > seg / 2) {                                                                 // :373-377
        bucket = (bucket + 1) & ((1u << logB) - 1u);
        dist -= seg;
      }
      const cplx sv = xs[loop_offset(g, j) + bucket];
      const int fi = (est ? a.fw_half[1] : a.fw_half[0]) - dist;                           // freq[(n - dist) % n]
      const cplx f = __ldg(&(est ? a.fwin[1] : a.fwin[0])[fi]);
      const double2 dr = __ldg(&(est ? a.fdr[1] : a.fdr[0])[fi]);
      // even lane: (a*c + b*d) / den      odd lane: (a*d - b*c) / den   (:388-398)
      const double t1 = __dmul_rn(sv.x, imag ? f.y : f.x);
      const double t2 = __dmul_rn(sv.y, imag ? f.x : f.y);
      const double num = __dadd_rn(t1, imag ? -t2 : t2);
      v[j] = div_by_rcp_rn(num, dr.x, dr.y);
    }
    const double mine = MedianNet<L>::run(v);                                               // :406-412
    const double other = __shfl_xor_sync(0xffffffffu, mine, 1);
    if (active && !imag) {
      a.out_loc[(long long)s * a.out_cap + h] = (int)loc;
      a.out_val[(long long)s * a.out_cap + h] = make_double2(mine, other);
    }
  }
}

// generic fallback: loop count at run time, values in local memory, plain division
__global__ void __launch_bounds__(128)
estimate_generic_kernel(LoopGeom g, EstimateArgs a, long long max_per_sig)
{
  const int s = blockIdx.y;
  long long total;
  int nc = 0;
  if (a.approved) {
    nc = a.num_comb[s];
    total = (long long)nc * a.n_over_W;
  } else {
    total = a.count[s];
  }
  if (total > max_per_sig) total = max_per_sig;
  long long lo = 0;
  if (a.slice_world > 1) {
    lo = total * a.slice_rank / a.slice_world;
    total = total * (a.slice_rank + 1) / a.slice_world - lo;
  }
  if (a.slice_count && blockIdx.x == 0 && threadIdx.x == 0) a.slice_count[s] = (int)total;
  const int loops = g.loops;
  const int mid = (loops - 1) / 2;     // cf12.cc:406
  const int *__restrict__ perm = a.perm + (long long)s * perm_stride(g.loops) + g.loops;   // ai[]
  const cplx *__restrict__ xs = a.xs + (long long)s * a.xs_stride;
  for (long long h = blockIdx.x * (long long)blockDim.x + threadIdx.x; h < total;
       h += (long long)gridDim.x * blockDim.x) {
    unsigned loc;
    const long long hc = lo + h;
    if (a.approved) {
      const long long jj = hc / nc;
      const int i = (int)(hc - jj * nc);
      loc = (unsigned)(jj * a.W + __ldg(&a.approved[(long long)s * a.approved_stride + i]));
    } else {
      loc = (unsigned)a.hits[(long long)s * a.hits_cap + hc];
    }
    double vr[kMaxLoops], vi[kMaxLoops];
    for (int j = 0; j < loops; j++)
      estimate_term(g, a, xs, (unsigned)perm[j], loc, j, vr[j], vi[j]);
    a.out_loc[(long long)s * a.out_cap + h] = (int)loc;
    a.out_val[(long long)s * a.out_cap + h] =
        make_double2(select_rank(vr, loops, mid), select_rank(vi, loops, mid));
  }
}

int launch_estimate(const LoopGeom &g, const EstimateArgs &a, int nsig, long long max_per_sig,
                    cudaStream_t st)
{
  // the hit count is only known on the device: size the grid for a full machine and let
  // the kernel stride; in a batch every signal gets its share of that machine
  long long blocks = (max_per_sig + 127) / 128;
  long long cap = 148ll * 16 / nsig;
  if (cap < 8) cap = 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)nsig);
  switch (g.loops) {
#define SFFTB_EST_CASE(N) case N: estimate_pair_kernel<N><<<grid, 256, 0, st>>>(g, a, max_per_sig); break;
    SFFTB_EST_CASE(2) SFFTB_EST_CASE(3) SFFTB_EST_CASE(4) SFFTB_EST_CASE(5) SFFTB_EST_CASE(6)
    SFFTB_EST_CASE(7) SFFTB_EST_CASE(8) SFFTB_EST_CASE(9) SFFTB_EST_CASE(10)
    // every total loop count of the reference's tables (parameters.cc) and defaults
    SFFTB_EST_CASE(11) SFFTB_EST_CASE(12) SFFTB_EST_CASE(13) SFFTB_EST_CASE(14)
    SFFTB_EST_CASE(15) SFFTB_EST_CASE(16) SFFTB_EST_CASE(17) SFFTB_EST_CASE(18)
    SFFTB_EST_CASE(19) SFFTB_EST_CASE(20) SFFTB_EST_CASE(21) SFFTB_EST_CASE(22)
    SFFTB_EST_CASE(23) SFFTB_EST_CASE(24) SFFTB_EST_CASE(25) SFFTB_EST_CASE(26)
    SFFTB_EST_CASE(27) SFFTB_EST_CASE(28) SFFTB_EST_CASE(29) SFFTB_EST_CASE(30)
    SFFTB_EST_CASE(31) SFFTB_EST_CASE(32)
#undef SFFTB_EST_CASE
    default: estimate_generic_kernel<<<grid, 128, 0, st>>>(g, a, max_per_sig); break;
  }
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// v2 structured estimation (see v12_kernels.cuh)
// ---------------------------------------------------------------------------
// hits per tile = threads per CTA: one CTA per SM, 8 KB runs.  256-hit tiles with two CTAs per SM
// are slower (0.965 vs 0.955 ms at C2), also when the second CTA starts half a tile period late
// so that the two sit in opposite phases (profiles/r02_v2_tile256_skew.jsonl).
constexpr int kV2LogTile = 9;
constexpr int kV2MaxSmem = 227 * 1024 - 8192;   // dynamic shared memory a CTA may ask for (static: parameters)

__device__ __forceinline__ void v2_slice(const V2StructArgs &a, long long total, long long &lo, long long &hi)
{
  lo = 0; hi = total;
  if (a.slice_world > 1) {
    lo = total * a.slice_rank / a.slice_world;
    hi = total * (a.slice_rank + 1) / a.slice_world;
  }
}

// class-major copy of every bucket spectrum: xt[(b mod 2^t) * T + (b >> t)] = xs[b],
// T = 2^kV2LogTile, t = logB - kV2LogTile.  Buckets that agree modulo 2^t become one
// contiguous run of T elements.  One CTA writes one run and records whether every
// component of it lies in the exponent band where the fast division of the estimation
// kernel is exact (nonzero, 2^-160 <= |x| < 2^190); it also resets the tile counter.
__global__ void __launch_bounds__(1 << kV2LogTile)
v2_regroup_kernel(LoopGeom g, const cplx *__restrict__ xs, cplx *__restrict__ xt,
                  unsigned char *__restrict__ run_unsafe, unsigned *__restrict__ tile_counter)
{
  constexpr int logT = kV2LogTile;
  const int j = blockIdx.y;
  const long long sig = (long long)blockIdx.z * g.x_samp_size;
  const int logB = j >= g.loops_loc ? g.logB[1] : g.logB[0];
  const int t = logB - logT;
  if (blockIdx.x == 0 && j == 0 && threadIdx.x == 0) tile_counter[blockIdx.z] = 0u;
  const unsigned o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (1u << logB)) return;
  const unsigned b = (o >> logT) | ((o & ((1u << logT) - 1u)) << t);
  const long long off = sig + loop_offset(g, j);
  const cplx v = xs[off + b];
  xt[off + o] = v;
  const unsigned ex = ((unsigned)__double2hiint(v.x) >> 20) & 0x7ffu;
  const unsigned ey = ((unsigned)__double2hiint(v.y) >> 20) & 0x7ffu;
  const int bad = (ex - (1023u - 160u) >= 350u) || (ey - (1023u - 160u) >= 350u);
  const int any = __syncthreads_or(bad);
  if (threadIdx.x == 0) run_unsafe[(off + o) >> logT] = (unsigned char)any;
}

// per-(tile, loop) constants, split so that every access is one aligned vector load
struct V2TileParams {
  cplx f[32];          // filter response at this residue's in-bucket offset
  double2 dr[32];      // (|f|^2, RN(1/|f|^2))
  uint2 pm[32];        // byte offset of hit u's element in its run: (pm.x + pm.y*u) mod 16T
  unsigned r;          // the tile's residue
  int unsafe;          // this tile must use real divisions
  unsigned chunk;      // first chunk claimed by the CTA
  long long tile;      // the tile these parameters belong to; -1: no more tiles
};

// a / b with y = RN(1/b): two Markstein corrections, as div_by_rcp_rn without the guard.
// Exact (== __ddiv_rn) when 2^-383 <= |a|,|b| < 2^385 or a == +0.  The estimation kernel
// guarantees that per tile instead of per division: numerators are sums of two products
// of a spectrum component and a filter component; with every spectrum component of the
// tile's runs in [2^-160, 2^190) (v2_regroup_kernel) and every filter component zero or in
// that band (not both zero), a numerator is +0 (exact cancellation) or in [2^-373, 2^381).
__device__ __forceinline__ double div_fast(double a, double b, double y)
{
  double q = __dmul_rn(a, y);
  double r = __fma_rn(-b, q, a);
  q = __fma_rn(r, y, q);
  r = __fma_rn(-b, q, a);
  return __fma_rn(r, y, q);
}

__device__ __forceinline__ cplx lds_cplx(unsigned addr)
{
  cplx r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
  unsigned done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// one contiguous global -> shared copy by the TMA engine, completion counted on `bar`
__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ bool v2_band_or_zero(double x)
{
  const unsigned h = (unsigned)__double2hiint(x) & 0x7fffffffu;
  return (h | (unsigned)__double2loint(x)) == 0u || ((h >> 20) - (1023u - 160u)) < 350u;
}

constexpr int kV2Chunk = 4;           // tiles claimed per atomic

// v[(L-1)/2] of the ascending order of v, selected on 32-bit keys.  The high word of a double,
// read as a float, orders like the double itself (sign, then exponent and leading mantissa
// bits) as long as it is neither a float NaN/Inf/subnormal pattern nor -0 -- true for every
// quotient of a tile whose inputs are in the proven band (div_fast above: |q| in
// [2^-758, 2^766) or +0).  The network then costs one FMNMX per min / max instead of one
// 64-bit compare and two selects per live word.  The key of the median is its high word; the
// low word is picked up from the one quotient carrying that high word.  When several do (they
// agree to 2^-20: the loops' estimates of a real coefficient), or when the tile is outside the
// band, the exact 64-bit network decides -- v is untouched until then.
template <int L>
__device__ __forceinline__ double v2_median(double (&v)[L], int exact)
{
  float key[L];
#pragma unroll
  for (int j = 0; j < L; j++) key[j] = __int_as_float(__double2hiint(v[j]));
  const int mh = __float_as_int(MedianNet<L>::run_hi(key));
  unsigned lo = 0u, cnt = 0u;
#pragma unroll
  for (int j = 0; j < L; j++)
    if (__double2hiint(v[j]) == mh) { lo = (unsigned)__double2loint(v[j]); cnt++; }
  if (exact | (cnt != 1u)) return MedianNet<L>::run(v);
  return __hiloint2double(mh, (int)lo);
}

__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Tile = (residue i, jj mod 2^s == c): T = (n/W)/2^s hits, one per thread.  In loop j those
// hits read T buckets that agree modulo q*2^s -- one contiguous 16*T-byte run of xt -- so
// L bulk copies bring the tile's inputs into shared memory; each thread picks its L values,
// divides by the filter response, keeps the 2L quotients in registers and takes the two
// medians.  No intermediate leaves the SM.
// CTAs are persistent and claim chunks of consecutive tiles from a counter.  Two mbarriers
// pipeline them: `full` (the L copies of a tile have landed, parameters are written) and
// `empty` (every warp has read its inputs).  The L parameter threads wait on `empty`, write
// the next tile's parameters and issue its copies; everybody else goes straight from the
// divisions to the medians, so the copies fly under the medians and warps of one CTA
// drift apart by up to a phase -- FP64-heavy divisions and ALU-heavy medians overlap.
template <int L>
__global__ void __launch_bounds__(1 << kV2LogTile, 512 >> kV2LogTile)
v2_fused_kernel(LoopGeom g, V2StructArgs a)
{
  constexpr int logT = kV2LogTile, T = 1 << logT;
  constexpr unsigned kRunBytes = T * sizeof(cplx);
  extern __shared__ __align__(128) cplx v2_stage[];              // [L][T], then the run flags
  __shared__ V2TileParams prm;
  __shared__ __align__(8) unsigned long long bars[2];            // full, empty
  const int sig = blockIdx.y;
  const int logNW = g.logn - a.logW;
  const int sbits = logNW - logT;
  const int nc = a.num_comb[sig];
  long long lo, hi;
  v2_slice(a, (long long)nc << sbits, lo, hi);
  if (a.slice_count && blockIdx.x == 0 && threadIdx.x == 0) a.slice_count[sig] = (int)((hi - lo) << logT);
  const unsigned u = threadIdx.x;
  const unsigned stage0 = (unsigned)__cvta_generic_to_shared(v2_stage);
  const unsigned full = (unsigned)__cvta_generic_to_shared(&bars[0]);
  const unsigned empty = (unsigned)__cvta_generic_to_shared(&bars[1]);
  const cplx *__restrict__ xt = a.xt + (long long)sig * g.x_samp_size;
  unsigned char *s_unsafe = reinterpret_cast<unsigned char *>(v2_stage + L * T);
  unsigned *ctr = a.tile_counter + sig;

  {
    const int nruns = (int)(g.x_samp_size >> logT);
    const unsigned char *src = a.run_unsafe + (((long long)sig * g.x_samp_size) >> logT);
    for (int t = (int)u; t < nruns; t += T) s_unsafe[t] = src[t];
  }
  if (u == 0) {
    mbar_init(full, L);
    mbar_init(empty, T / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prm.chunk = atomicAdd(ctr, 1u);
  }
  __syncthreads();

  // ---- parameter threads: loop j = u ----
  const bool is_param = u < (unsigned)L;
  const unsigned param_mask = L == 32 ? 0xffffffffu : ((1u << L) - 1u);
  const int pj = is_param ? (int)u : 0;
  const bool p_est = pj >= g.loops_loc;
  const int p_logB = p_est ? g.logB[1] : g.logB[0];
  const int p_logseg = g.logn - p_logB;
  const int p_logq = a.logW - p_logseg;
  const unsigned p_ai = (unsigned)a.perm[(long long)sig * perm_stride(g.loops) + g.loops + pj];
  const unsigned p_m = p_ai & (unsigned)((1 << logNW) - 1);      // ai*W mod n = W * (ai mod n/W)
  const unsigned p_off = (unsigned)loop_offset(g, pj);
  int p_i = -1;
  unsigned p_bucket = 0;
  bool p_fbad = false;
  long long p_tile = lo + (long long)prm.chunk * kV2Chunk;      // next tile to issue
  long long p_chunk_end = p_tile + kV2Chunk;
  unsigned p_pending = 0;
  // writes the parameters of tile `p_tile` (or the end marker) and releases `full`
  auto issue_tile = [&]() {
    if (p_tile >= hi) {
      if (pj == 0) prm.tile = -1;
      mbar_arrive(full);
      return;
    }
    const long long tile = p_tile;
    if (pj == 0 && tile + kV2Chunk == p_chunk_end) p_pending = atomicAdd(ctr, 1u);   // first of its chunk
    const int i = (int)(tile >> sbits);
    const unsigned c = (unsigned)(tile & ((1ll << sbits) - 1));
    if (i != p_i) {
      p_i = i;
      const unsigned r = (unsigned)__ldg(&a.approved[(long long)sig * a.approved_stride + i]);
      const int seg = 1 << p_logseg;
      const unsigned pos = (unsigned)(((unsigned long long)p_ai * r) & (unsigned)g.n_mask);   // jj = 0
      unsigned bucket = pos >> p_logseg;
      int dist = (int)(pos & (unsigned)(seg - 1));
      if (dist > seg / 2) {                                                                 // cf12.cc:373-377
        bucket = (bucket + 1) & ((1u << p_logB) - 1u);
        dist -= seg;
      }
      p_bucket = bucket;
      const int half = p_est ? a.fw_half[1] : a.fw_half[0];
      const double2 dr = __ldg(&(p_est ? a.fdr[1] : a.fdr[0])[half - dist]);
      const cplx f = __ldg(&(p_est ? a.fwin[1] : a.fwin[0])[half - dist]);
      prm.f[pj] = f;
      prm.dr[pj] = dr;
      if (pj == 0) prm.r = r;
      const unsigned eb = ((unsigned)__double2hiint(dr.x) >> 20) & 0x7ffu;
      p_fbad = (eb - 640u >= 768u) || !v2_band_or_zero(f.x) || !v2_band_or_zero(f.y);
    }
    const unsigned cls = p_bucket & ((1u << p_logq) - 1u);
    const unsigned P = ((p_bucket >> p_logq) + p_m * c) & (unsigned)((1 << logNW) - 1);
    const unsigned beta = cls | ((P & ((1u << sbits) - 1u)) << p_logq);
    const unsigned run0 = p_off + (beta << logT);
    prm.pm[pj] = make_uint2(((P >> sbits) & (unsigned)(T - 1)) << 4, (p_m & (unsigned)(T - 1)) << 4);
    const int bad = __any_sync(param_mask, p_fbad || s_unsafe[run0 >> logT]);
    if (pj == 0) { prm.unsafe = bad; prm.tile = tile; }
    // the CTA's reads of the previous tile (generic proxy) precede this copy (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(full, kRunBytes);
    bulk_copy_g2s(stage0 + (unsigned)pj * kRunBytes, xt + run0, kRunBytes, full);
    // advance: within the chunk, or to the chunk claimed while this one was running
    p_tile = tile + 1;
    if (p_tile == p_chunk_end) {
      p_tile = lo + (long long)__shfl_sync(param_mask, p_pending, 0) * kV2Chunk;
      p_chunk_end = p_tile + kV2Chunk;
    }
  };
  if (is_param) issue_tile();

  unsigned parity = 0;
  while (true) {
    mbar_wait(full, parity);
    const long long tile = prm.tile;
    if (tile < 0) break;
    const unsigned r = prm.r;
    const int unsafe = prm.unsafe;
    double vr[L], vi[L];
#pragma unroll
    for (int j = 0; j < L; j++) {
      const uint2 pm = prm.pm[j];
      const cplx f = prm.f[j];
      const double2 dr = prm.dr[j];
      const cplx sv = lds_cplx(stage0 + (unsigned)j * kRunBytes + ((pm.x + pm.y * u) & (kRunBytes - 16u)));
      const double ac = __dmul_rn(sv.x, f.x), bd = __dmul_rn(sv.y, f.y);
      const double ad = __dmul_rn(sv.x, f.y), bc = __dmul_rn(sv.y, f.x);
      vr[j] = div_fast(__dadd_rn(ac, bd), dr.x, dr.y);                             // :388-398
      vi[j] = div_fast(__dsub_rn(ad, bc), dr.x, dr.y);
    }
    if (unsafe) {
      // some input of this tile is zero or outside the band where div_fast is proven
#pragma unroll 1
      for (int j = 0; j < L; j++) {
        const uint2 pm = prm.pm[j];
        const cplx f = prm.f[j];
        const double den = prm.dr[j].x;
        const cplx sv = v2_stage[j * T + (((pm.x + pm.y * u) >> 4) & (unsigned)(T - 1))];
        const double ac = __dmul_rn(sv.x, f.x), bd = __dmul_rn(sv.y, f.y);
        const double ad = __dmul_rn(sv.x, f.y), bc = __dmul_rn(sv.y, f.x);
        const double qr = __ddiv_rn(__dadd_rn(ac, bd), den), qi = __ddiv_rn(__dsub_rn(ad, bc), den);
        // the register arrays cannot be indexed dynamically: a predicated sweep puts the
        // quotient in place
#pragma unroll
        for (int t = 0; t < L; t++)
          if (t == j) { vr[t] = qr; vi[t] = qi; }
      }
    }
    // this warp is done with the tile's inputs and parameters
    __syncwarp();
    if ((u & 31u) == 0) mbar_arrive(empty);
    if (is_param) {
      mbar_wait(empty, parity);
      issue_tile();
    }
    parity ^= 1u;

    const double re = v2_median<L>(vr, unsafe);
    const double im = v2_median<L>(vi, unsafe);
    const unsigned c = (unsigned)(tile & ((1ll << sbits) - 1));
    const unsigned jj = c + (u << sbits);
    const long long o = (long long)sig * a.out_cap + ((tile - lo) << logT) + u;
    a.out_loc[o] = (int)((jj << a.logW) + r);                                      // cf12.cc:508-511
    a.out_val[o] = make_double2(re, im);
  }
}

bool v2_struct_supported(const LoopGeom &g, int logW)
{
  for (int grp = 0; grp < 2; grp++) {
    const int logseg = g.logn - g.logB[grp];
    if (logW < logseg) return false;                 // W must be a multiple of the bucket width
  }
  const int logNW = g.logn - logW;
  const long long flag_bytes = ((g.x_samp_size >> kV2LogTile) + 15) & ~15ll;
  return logNW >= kV2LogTile && g.loops >= 2 && g.loops <= 32 &&
         ((long long)sizeof(cplx) * g.loops << kV2LogTile) + flag_bytes <= kV2MaxSmem &&
         g.logB[0] >= kV2LogTile && g.logB[1] >= kV2LogTile;
}

int v2_struct_log_tile(const LoopGeom &g, int logW)
{
  const int logNW = g.logn - logW;
  return logNW < kV2LogTile ? logNW : kV2LogTile;
}

int launch_v2_struct(const LoopGeom &g, const V2StructArgs &a, int max_comb, int nsig, cudaStream_t st)
{
  const int logNW = g.logn - a.logW;
  constexpr int logT = kV2LogTile, T = 1 << logT;
  const int maxlog = g.logB[0] > g.logB[1] ? g.logB[0] : g.logB[1];
  dim3 rgrid(1u << (maxlog - logT), (unsigned)g.loops, (unsigned)nsig);
  v2_regroup_kernel<<<rgrid, T, 0, st>>>(g, a.xs, a.xt, a.run_unsafe, a.tile_counter);
  SFFTB_LAUNCH_CHECK();
  const size_t flag_bytes = ((size_t)(g.x_samp_size >> logT) + 15) & ~(size_t)15;
  const size_t smem = sizeof(cplx) * (size_t)g.loops * T + flag_bytes;
  // persistent CTAs, as many as are resident at once (512 threads per SM)
  long long ctas = 148ll * (512 >> kV2LogTile) / nsig;
  const long long chunks = (((long long)max_comb << (logNW - logT)) + kV2Chunk - 1) / kV2Chunk;
  if (ctas < 1) ctas = 1;
  if (ctas > chunks) ctas = chunks;
  dim3 grid((unsigned)ctas, (unsigned)nsig);
  switch (g.loops) {
#define SFFTB_V2F_CASE(N)                                                                         \
  case N: {                                                                                       \
    SFFTB_ONCE_PER_DEVICE(SFFTB_CUDA(cudaFuncSetAttribute(                                        \
        v2_fused_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, kV2MaxSmem)));           \
    v2_fused_kernel<N><<<grid, T, smem, st>>>(g, a);                                              \
  } break;
    SFFTB_V2F_CASE(2) SFFTB_V2F_CASE(3) SFFTB_V2F_CASE(4) SFFTB_V2F_CASE(5) SFFTB_V2F_CASE(6)
    SFFTB_V2F_CASE(7) SFFTB_V2F_CASE(8) SFFTB_V2F_CASE(9) SFFTB_V2F_CASE(10) SFFTB_V2F_CASE(11)
    SFFTB_V2F_CASE(12) SFFTB_V2F_CASE(13) SFFTB_V2F_CASE(14) SFFTB_V2F_CASE(15) SFFTB_V2F_CASE(16)
    SFFTB_V2F_CASE(17) SFFTB_V2F_CASE(18) SFFTB_V2F_CASE(19) SFFTB_V2F_CASE(20) SFFTB_V2F_CASE(21)
    SFFTB_V2F_CASE(22) SFFTB_V2F_CASE(23) SFFTB_V2F_CASE(24) SFFTB_V2F_CASE(25) SFFTB_V2F_CASE(26)
    SFFTB_V2F_CASE(27) SFFTB_V2F_CASE(28) SFFTB_V2F_CASE(29) SFFTB_V2F_CASE(30) SFFTB_V2F_CASE(31)
    SFFTB_V2F_CASE(32)
#undef SFFTB_V2F_CASE
    default: set_error("launch_v2_struct: unsupported loop count"); return -1;
  }
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// K6  Comb pre-filter (v2)   (cf12.cc:49-82, :483-512)
// ---------------------------------------------------------------------------
__global__ void comb_sample_kernel(const cplx *__restrict__ x_direct, const unsigned long long *x_indirect,
                                   long long x_stride, const int *__restrict__ comb_off, int comb_loops,
                                   int logW, int logn, cplx *cxs, long long cxs_stride)
{
  const cplx *__restrict__ x = x_indirect ? reinterpret_cast<const cplx *>(*x_indirect) : x_direct;
  const int c = blockIdx.y, s = blockIdx.z;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << logW)) return;
  const int logsig = logn - logW;
  const unsigned off = (unsigned)comb_off[s * comb_loops + c];
  const cplx v = ldg_stream(x + (long long)s * x_stride + off + ((unsigned long long)i << logsig));
  cxs[(long long)s * cxs_stride + ((long long)c << logW) + bitrev(i, logW)] = v;
}

int launch_comb_sample(const cplx *x, const unsigned long long *x_indirect, long long x_stride,
                       const int *comb_off, int comb_loops, int logW, int logn, cplx *cxs,
                       long long cxs_stride, int nsig, cudaStream_t st)
{
  dim3 grid((unsigned)ceil_div(1ll << logW, 256), (unsigned)comb_loops, (unsigned)nsig);
  comb_sample_kernel<<<grid, 256, 0, st>>>(x, x_indirect, x_stride, comb_off, comb_loops, logW, logn, cxs,
                                           cxs_stride);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(1024)
comb_merge_kernel(const unsigned *__restrict__ loop_bitmaps, int comb_loops, int W, int n_over_W,
                  unsigned *approved_bitmap, int *approved, int *num_comb, int *count)
{
  __shared__ unsigned warp_tot[32];
  const int s = blockIdx.x;
  const int words = W >= 32 ? W / 32 : 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned *lb = loop_bitmaps + (long long)s * comb_loops * words;
  unsigned *ab = approved_bitmap + (long long)s * words;
  int *out = approved + (long long)s * W;
  unsigned base = 0;
  for (int w0 = 0; w0 < words; w0 += 1024) {
    const int wi = w0 + tid;
    unsigned bits = 0;
    if (wi < words)
      for (int c = 0; c < comb_loops; c++) bits |= lb[c * words + wi];
    if (wi < words) ab[wi] = bits;
    const unsigned cnt = __popc(bits);
    unsigned incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned woff = 0, tot = 0;
    for (int q = 0; q < 32; q++) {
      if (q < warp) woff += warp_tot[q];
      tot += warp_tot[q];
    }
    unsigned pos = base + woff + incl - cnt;
    unsigned bb = bits;
    while (bb) {
      const int bit = __ffs(bb) - 1;
      out[pos++] = wi * 32 + bit;
      bb &= bb - 1;
    }
    base += tot;
    __syncthreads();
  }
  if (tid == 0) {
    num_comb[s] = (int)base;
    count[s] = (int)((long long)base * n_over_W);
  }
}

int launch_comb_merge(const unsigned *loop_bitmaps, int comb_loops, int W, int n_over_W,
                      unsigned *approved_bitmap, int *approved, int *num_comb, int *count,
                      int nsig, cudaStream_t st)
{
  comb_merge_kernel<<<nsig, 1024, 0, st>>>(loop_bitmaps, comb_loops, W, n_over_W,
                                           approved_bitmap, approved, num_comb, count);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// legacy dense output
// ---------------------------------------------------------------------------
__global__ void scatter_kernel(const int *__restrict__ loc, const cplx *__restrict__ val,
                               const int *__restrict__ count, cplx *out)
{
  const long long total = *count;
  for (long long h = blockIdx.x * (long long)blockDim.x + threadIdx.x; h < total;
       h += (long long)gridDim.x * blockDim.x)
    out[loc[h]] = val[h];
}

int launch_scatter(const int *loc, const cplx *val, const int *count, cplx *out, cudaStream_t st)
{
  scatter_kernel<<<148 * 8, 256, 0, st>>>(loc, val, count, out);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sfftb
