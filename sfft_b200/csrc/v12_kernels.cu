// v12_kernels.cu -- sm_100a kernels for sFFT v1/v2 (see v12_kernels.cuh).
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
constexpr int kGatherUnroll = 8;

__global__ void __launch_bounds__(kGatherThreads)
gather_kernel(LoopGeom g, GatherArgs a)
{
  const int j = a.loop_begin + blockIdx.y * a.loop_step;
  const int s = blockIdx.z;
  const bool est = j >= g.loops_loc;
  const int logB = est ? g.logB[1] : g.logB[0];
  const unsigned B = 1u << logB;
  const unsigned b = blockIdx.x * kGatherThreads + threadIdx.x;
  if (b >= B) return;

  const int w = est ? g.w[1] : g.w[0];
  const cplx *__restrict__ taps = est ? a.taps[1] : a.taps[0];
  const cplx *__restrict__ x = a.x + (long long)s * a.x_stride;
  const unsigned ai = (unsigned)a.perm[(long long)s * perm_stride(g.loops) + g.loops + j];
  const unsigned mask = (unsigned)g.n_mask;

  unsigned idx = (unsigned)(((unsigned long long)b * ai) & mask);
  const unsigned stepB = (unsigned)(((unsigned long long)B * ai) & mask);

  double acc_re = 0.0, acc_im = 0.0;
  for (unsigned i = b; i < (unsigned)w; i += kGatherUnroll * B) {
    cplx xv[kGatherUnroll], tv[kGatherUnroll];
    unsigned id = idx;
#pragma unroll
    for (int u = 0; u < kGatherUnroll; u++) {
      const unsigned ii = i + u * B;
      xv[u] = ldg_stream(x + id);
      tv[u] = __ldg(taps + (ii < (unsigned)w ? ii : 0u));
      id = (id + stepB) & mask;
    }
    idx = id;
#pragma unroll
    for (int u = 0; u < kGatherUnroll; u++) {
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
  gather_kernel<<<grid, kGatherThreads, 0, st>>>(g, a);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// K3+K4  |.|^2 and top-num bucket selection   (cf12.cc:278-302, utils.cc:131-158)
//
// One CTA per (row, signal).  Squared magnitudes become order-preserving 64-bit
// keys; an MSD radix select (8 digits of 8 bits, run-length-aggregated shared
// atomics) finds the (num+1)-th largest key = the reference's cutoff; then an
// ordered compaction emits indices with key > cutoff and, if short, the first
// few with key == cutoff -- the reference's tie rule -- ascending, together with a
// B-bit membership bitmap for the voting stage.
// ---------------------------------------------------------------------------
constexpr int kSelectThreads = 1024;
constexpr int kSelectSmemKeys = 16384;   // 128 KB of keys in shared memory

__global__ void __launch_bounds__(kSelectThreads)
select_kernel(SelectArgs a)
{
  extern __shared__ unsigned long long skeys[];
  __shared__ unsigned hist[256];
  __shared__ unsigned warp_gt[32], warp_eq[32];
  __shared__ unsigned long long sh_prefix;
  __shared__ unsigned sh_K;

  const int row = a.row_begin + blockIdx.x * a.row_step;
  const int s = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = 1 << a.logB;
  const cplx *__restrict__ src = a.xs + (long long)s * a.xs_stride + (long long)row * a.row_stride;
  unsigned long long *keys =
      a.gkeys ? a.gkeys + (long long)s * a.gk_sig_stride + (long long)row * B : skeys;

  for (int i = tid; i < B; i += kSelectThreads)
    keys[i] = (unsigned long long)__double_as_longlong(cabs2_rn(src[i]));
  __syncthreads();

  unsigned long long prefix = 0;
  unsigned K = (unsigned)a.num + 1u;      // rank from the top of the wanted key
  for (int pass = 0; pass < 8; pass++) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    int run_d = -1;
    unsigned run_c = 0;
    for (int i = tid; i < B; i += kSelectThreads) {
      const unsigned long long key = keys[i];
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
  const unsigned long long cutoff = prefix;
  const unsigned need = K - 1u;   // ties at the cutoff to admit, in index order

  int *J = a.J + (long long)s * a.J_sig_stride + (long long)row * a.num;
  const int words = B >= 32 ? B / 32 : 1;
  unsigned *bm = a.bitmap + (long long)s * a.bm_sig_stride + (long long)row * words;
  unsigned base_gt = 0, base_eq = 0;
  for (int base = 0; base < B; base += kSelectThreads) {
    const int i = base + tid;
    const bool valid = i < B;
    const unsigned long long key = valid ? keys[i] : 0ull;
    const bool f_gt = valid && key > cutoff;
    const bool f_eq = valid && key == cutoff;
    const unsigned bal_gt = __ballot_sync(0xffffffffu, f_gt);
    const unsigned bal_eq = __ballot_sync(0xffffffffu, f_eq);
    if (lane == 0) {
      warp_gt[warp] = __popc(bal_gt);
      warp_eq[warp] = __popc(bal_eq);
    }
    __syncthreads();
    unsigned off_gt = 0, off_eq = 0, tot_gt = 0, tot_eq = 0;
    for (int wv = 0; wv < kSelectThreads / 32; wv++) {
      const unsigned g_ = warp_gt[wv], e_ = warp_eq[wv];
      if (wv < warp) { off_gt += g_; off_eq += e_; }
      tot_gt += g_;
      tot_eq += e_;
    }
    const unsigned lt = (1u << lane) - 1u;
    const unsigned gt_before = base_gt + off_gt + __popc(bal_gt & lt);
    const unsigned eq_before = base_eq + off_eq + __popc(bal_eq & lt);
    const bool sel = f_gt || (f_eq && eq_before < need);
    if (sel) J[gt_before + (eq_before < need ? eq_before : need)] = i;
    const unsigned bal_sel = __ballot_sync(0xffffffffu, sel);
    if (lane == 0 && base + 32 * warp < B)
      bm[(base >> 5) + warp] = bal_sel;
    base_gt += tot_gt;
    base_eq += tot_eq;
    __syncthreads();
  }
}

int launch_select(const SelectArgs &a, int nrows, int nsig, cudaStream_t st)
{
  if (nrows <= 0) return 0;
  const int B = 1 << a.logB;
  size_t smem = a.gkeys ? 0 : (size_t)B * sizeof(unsigned long long);
  static bool attr_set = false;
  if (!attr_set) {
    SFFTB_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSelectSmemKeys * (int)sizeof(unsigned long long)));
    attr_set = true;
  }
  dim3 grid((unsigned)nrows, (unsigned)nsig);
  select_kernel<<<grid, kSelectThreads, smem, st>>>(a);
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
__device__ __forceinline__ bool voted_by(const LoopGeom &g, const VoteArgs &a, int s, int j,
                                         unsigned loc)
{
  const int logB = g.logB[0];
  const int logseg = g.logn - logB;
  const unsigned ai = (unsigned)a.perm[(long long)s * perm_stride(g.loops) + g.loops + j];
  const unsigned p = (unsigned)(((unsigned long long)ai * loc) & (unsigned)g.n_mask);
  const unsigned half = (1u << logseg) >> 1;
  const unsigned Jb = (((p + half) & (unsigned)g.n_mask) >> logseg) & ((1u << logB) - 1u);
  const int words = logB >= 5 ? (1 << (logB - 5)) : 1;
  const unsigned *bm = a.bitmap + (long long)s * a.bm_sig_stride + (long long)j * words;
  return (__ldg(&bm[Jb >> 5]) >> (Jb & 31)) & 1u;
}

__global__ void __launch_bounds__(256)
vote_kernel(LoopGeom g, VoteArgs a, long long total_per_sig)
{
  const int s = blockIdx.y;
  const int logB = g.logB[0];
  const int logseg = g.logn - logB;
  const unsigned seg = 1u << logseg;
  const unsigned mask = (unsigned)g.n_mask;
  const long long per_loop = (long long)a.num << logseg;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total_per_sig;
       q += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(q / per_loop);
    const long long r = q - (long long)j * per_loop;
    const int Ji = (int)(r >> logseg);
    const unsigned t = (unsigned)(r & (seg - 1));
    const unsigned Jv = (unsigned)a.J[(long long)s * a.J_sig_stride + (long long)j * a.num + Ji];
    // cf12.cc:102: low = ceil((J - 0.5) * n/B) mod n  == J*seg - seg/2 (exact for seg >= 2)
    const unsigned low = ((Jv << logseg) - (seg >> 1)) & mask;
    const unsigned p = (low + t) & mask;
    const unsigned aj = (unsigned)a.perm[(long long)s * perm_stride(g.loops) + j];
    const unsigned loc = (unsigned)(((unsigned long long)aj * p) & mask);
    if (a.comb_bitmap) {
      const unsigned rres = loc & (unsigned)a.W_mask;
      const unsigned *cb = a.comb_bitmap + (long long)s * a.comb_sig_stride;
      if (!((__ldg(&cb[rres >> 5]) >> (rres & 31)) & 1u)) continue;
    }
    bool earlier = false;
    for (int jj = 0; jj < j; jj++)
      if (voted_by(g, a, s, jj, loc)) { earlier = true; break; }
    if (earlier) continue;
    int score = 1;
    for (int jj = j + 1; jj < g.loops_loc; jj++) score += voted_by(g, a, s, jj, loc) ? 1 : 0;
    if (score >= a.thresh) {
      const int pos = atomicAdd(&a.count[s], 1);
      if (pos < a.hits_cap) a.hits[(long long)s * a.hits_cap + pos] = (int)loc;
    }
  }
}

int launch_vote(const LoopGeom &g, const VoteArgs &a, int nsig, cudaStream_t st)
{
  const int first_loops = g.loops_loc - a.thresh + 1;
  if (first_loops <= 0) return 0;
  const int logseg = g.logn - g.logB[0];
  const long long total = (long long)first_loops * ((long long)a.num << logseg);
  long long blocks = (total + 255) / 256;
  const long long cap = 148ll * 32;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, (unsigned)nsig);
  vote_kernel<<<grid, 256, 0, st>>>(g, a, total);
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

// L > 0: loop count known at compile time, everything in registers.
// L == 0: generic fallback (loop count at run time, values in local memory).
template <int L>
__global__ void __launch_bounds__(128)
estimate_kernel(LoopGeom g, EstimateArgs a, long long max_per_sig)
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
  const int loops = L > 0 ? L : g.loops;
  const int mid = (loops - 1) / 2;     // cf12.cc:406
  const int *__restrict__ perm = a.perm + (long long)s * perm_stride(g.loops) + g.loops;   // ai[]
  const cplx *__restrict__ xs = a.xs + (long long)s * a.xs_stride;

  for (long long h = blockIdx.x * (long long)blockDim.x + threadIdx.x; h < total;
       h += (long long)gridDim.x * blockDim.x) {
    unsigned loc;
    if (a.approved) {
      const long long jj = h / nc;
      const int i = (int)(h - jj * nc);
      loc = (unsigned)(jj * a.W + __ldg(&a.approved[(long long)s * a.approved_stride + i]));   // cf12.cc:508-511
    } else {
      loc = (unsigned)a.hits[(long long)s * a.hits_cap + h];
    }
    double re, im;
    if (L > 0) {
      double vr[L > 0 ? L : 2], vi[L > 0 ? L : 2];
#pragma unroll
      for (int j = 0; j < (L > 0 ? L : 2); j++)
        estimate_term(g, a, xs, (unsigned)__ldg(&perm[j]), loc, j, vr[j], vi[j]);
      re = MedianNet<(L > 0 ? L : 2)>::run(vr);
      im = MedianNet<(L > 0 ? L : 2)>::run(vi);
    } else {
      double vr[kMaxLoops], vi[kMaxLoops];
      for (int j = 0; j < loops; j++)
        estimate_term(g, a, xs, (unsigned)perm[j], loc, j, vr[j], vi[j]);
      re = select_rank(vr, loops, mid);
      im = select_rank(vi, loops, mid);
    }
    a.out_loc[(long long)s * a.out_cap + h] = (int)loc;
    a.out_val[(long long)s * a.out_cap + h] = make_double2(re, im);
  }
}

template <int L>
static void launch_estimate_t(const LoopGeom &g, const EstimateArgs &a, dim3 grid, long long max_per_sig,
                              cudaStream_t st)
{
  estimate_kernel<L><<<grid, 128, 0, st>>>(g, a, max_per_sig);
}

int launch_estimate(const LoopGeom &g, const EstimateArgs &a, int nsig, long long max_per_sig,
                    cudaStream_t st)
{
  long long blocks = (max_per_sig + 127) / 128;
  const long long cap = 148ll * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)nsig);
  switch (g.loops) {
#define SFFTB_EST_CASE(N) case N: launch_estimate_t<N>(g, a, grid, max_per_sig, st); break;
    // every total loop count of the reference's tables (parameters.cc) and defaults
    SFFTB_EST_CASE(2) SFFTB_EST_CASE(3) SFFTB_EST_CASE(4) SFFTB_EST_CASE(5) SFFTB_EST_CASE(6)
    SFFTB_EST_CASE(7) SFFTB_EST_CASE(8) SFFTB_EST_CASE(9) SFFTB_EST_CASE(10)
    SFFTB_EST_CASE(11) SFFTB_EST_CASE(12) SFFTB_EST_CASE(13) SFFTB_EST_CASE(14)
    SFFTB_EST_CASE(15) SFFTB_EST_CASE(16) SFFTB_EST_CASE(17) SFFTB_EST_CASE(18)
    SFFTB_EST_CASE(19) SFFTB_EST_CASE(20) SFFTB_EST_CASE(21) SFFTB_EST_CASE(22)
    SFFTB_EST_CASE(23) SFFTB_EST_CASE(24)
#undef SFFTB_EST_CASE
    default: launch_estimate_t<0>(g, a, grid, max_per_sig, st); break;
  }
  SFFTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------
// K6  Comb pre-filter (v2)   (cf12.cc:49-82, :483-512)
// ---------------------------------------------------------------------------
__global__ void comb_sample_kernel(const cplx *__restrict__ x, long long x_stride,
                                   const int *__restrict__ comb_off, int comb_loops, int logW,
                                   int logn, cplx *cxs, long long cxs_stride)
{
  const int c = blockIdx.y, s = blockIdx.z;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1u << logW)) return;
  const int logsig = logn - logW;
  const unsigned off = (unsigned)comb_off[s * comb_loops + c];
  const cplx v = ldg_stream(x + (long long)s * x_stride + off + ((unsigned long long)i << logsig));
  cxs[(long long)s * cxs_stride + ((long long)c << logW) + bitrev(i, logW)] = v;
}

int launch_comb_sample(const cplx *x, long long x_stride, const int *comb_off, int comb_loops,
                       int logW, int logn, cplx *cxs, long long cxs_stride, int nsig,
                       cudaStream_t st)
{
  dim3 grid((unsigned)ceil_div(1ll << logW, 256), (unsigned)comb_loops, (unsigned)nsig);
  comb_sample_kernel<<<grid, 256, 0, st>>>(x, x_stride, comb_off, comb_loops, logW, logn, cxs,
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
