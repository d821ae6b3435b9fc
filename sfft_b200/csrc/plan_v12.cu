// plan_v12.cu -- plan derivation and transform driver for sFFT v1 / v2.
//
// Host-side mirror of sfft_v1v2_make_plan (src/sfft.cc:298-392) and of outer_loop's
// control flow (src/computefourier-1.0-2.0.cc:438-541); all data-path work is in
// the kernels of v12_kernels.cu / fft.cu.
#include <math.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#include "fft.cuh"
#include "plan.cuh"

namespace sfftb {

namespace {

struct ParamRow {
  int by_k, with_comb, key;
  double Bcst_loc, Bcst_est, Comb_cst;
  int loc_loops, est_loops, threshold_loops, comb_loops;
  double tolerance_loc, tolerance_est;
};

const ParamRow kParamRows[] = {
#include "param_table.inc"
};

// src/utils.cc:243-248
int floor_to_pow2(double x)
{
  unsigned int ans;
  for (ans = 1; ans <= x; ans <<= 1) {}
  return (int)(ans / 2);
}

// src/utils.cc:41-46
int gcd_ref(int a, int b)
{
  while (a % b != 0) {
    const int r = a % b;
    a = b;
    b = r;
  }
  return b;
}

// src/utils.cc:85-100
int mod_inverse(int a, int n)
{
  int i = n, v = 0, d = 1;
  while (a > 0) {
    const int t = i / a, x = a;
    a = i % x;
    i = x;
    const int nd = v - t * d;
    v = d;
    d = nd;
  }
  v %= n;
  if (v < 0) v = (v + n) % n;
  return v;
}

}  // namespace

int mod_inverse_pub(int a, int n) { return mod_inverse(a, n); }
int gcd_pub(int a, int b) { return gcd_ref(a, b); }

int v12_derive(PlanImpl *p, int n_req, int k, int with_comb, int tuned_by_k)
{
  PlanV12 &v = p->v12;
  // defaults, src/sfft.cc:306-314
  double Bcst_loc = 1, Bcst_est = 1, Comb_cst = 2;
  int loc_loops = 4, est_loops = 16, threshold_loops = 3, comb_loops = 1;
  double tol_loc = 1.e-8, tol_est = 1.e-8;
  // src/sfft.cc:316-327: for k > 50 the by-K table is searched with n as the key
  // -- the by-K table is keyed by k (parameters.cc:282-513, tuned at n = 2^22), so that
  // lookup never matches and every k > 50 plan runs on the defaults above.  The reference's
  // behaviour is the default here too; SFFTB_PLAN_TUNED_BY_K (include/sfft.h) opts in to the
  // lookup by k the table was written for.
  const int by_k = (unsigned)k > 50 ? 1 : 0;
  const int key = by_k && tuned_by_k ? k : n_req;
  for (const ParamRow &r : kParamRows) {
    if (r.by_k == by_k && r.with_comb == with_comb && r.key == key) {
      Bcst_loc = r.Bcst_loc; Bcst_est = r.Bcst_est; Comb_cst = r.Comb_cst;
      loc_loops = r.loc_loops; est_loops = r.est_loops;
      threshold_loops = r.threshold_loops; comb_loops = r.comb_loops;
      tol_loc = r.tolerance_loc; tol_est = r.tolerance_est;
      break;
    }
  }
  const unsigned n = (unsigned)floor_to_pow2(n_req);
  if ((int)n != n_req || n < 4) {
    set_error("sfft_make_plan: n must be a power of two >= 4 (the reference's transform is "
              "undefined otherwise: it plans for floor_to_pow2(n) but executes with n)");
    return -1;
  }
  if (k < 1) { set_error("sfft_make_plan: k must be >= 1"); return -1; }

  // src/sfft.cc:332-349
  const double BB_loc = (unsigned)(Bcst_loc * sqrt((double)(int)n * (unsigned)k / (log2((double)n))));
  const double BB_est = (unsigned)(Bcst_est * sqrt((double)(int)n * (unsigned)k / (log2((double)n))));
  if (BB_loc < 1 || BB_est < 1) { set_error("sfft_make_plan: bucket count underflows"); return -1; }
  v.with_comb = with_comb;
  v.lobe_loc = 0.5 / BB_loc;
  v.lobe_est = 0.5 / BB_est;
  v.b_loc = (int)(1.2 * 1.1 * ((double)n / BB_loc));
  v.b_est = (int)(1.4 * 1.1 * ((double)n / BB_est));
  v.B_loc = floor_to_pow2(BB_loc);
  v.B_thresh = 2 * k;
  v.B_est = floor_to_pow2(BB_est);
  v.W_Comb = floor_to_pow2(Comb_cst * n / v.B_loc);
  v.Comb_loops = comb_loops;
  v.loops_loc = loc_loops;
  v.loops_thresh = threshold_loops;
  v.loops_est = est_loops;
  v.tol_loc = tol_loc;
  v.tol_est = tol_est;
  v.x_samp_size = (long long)v.loops_loc * v.B_loc + (long long)v.loops_est * v.B_est;

  const int loops = v.loops_loc + v.loops_est;
  if (loops > SFFTB_MAX_LOOPS || v.Comb_loops > SFFTB_MAX_COMB_LOOPS) {
    set_error("sfft_make_plan: loop count exceeds SFFTB_MAX_LOOPS");
    return -1;
  }
  if ((unsigned)v.B_loc > n || (unsigned)v.B_est > n) {
    set_error("sfft_make_plan: more buckets than samples (reference asserts n % B == 0, cf12.cc:215)");
    return -1;
  }
  // find_largest_indices asserts n >= num+1 (src/utils.cc:134) for every loop
  // (cf12.cc:301) and for the Comb spectrum (cf12.cc:78)
  if (v.B_loc < v.B_thresh + 1 || v.B_est < v.B_thresh + 1) {
    set_error("sfft_make_plan: 2k+1 exceeds the bucket count (reference asserts, utils.cc:134)");
    return -1;
  }
  if (with_comb && (v.W_Comb < v.B_thresh + 1 || (unsigned)v.W_Comb > n)) {
    set_error("sfft_make_plan: 2k+1 exceeds W_Comb (reference asserts, utils.cc:134 via cf12.cc:78)");
    return -1;
  }

  p->n = (int)n;
  p->logn = ilog2(n);
  p->k = k;
  LoopGeom &g = v.geom;
  g.n_mask = (int)n - 1;
  g.logn = p->logn;
  g.loops = loops;
  g.loops_loc = v.loops_loc;
  g.logB[0] = ilog2((unsigned)v.B_loc);
  g.logB[1] = ilog2((unsigned)v.B_est);
  g.x_samp_size = v.x_samp_size;
  return 0;
}

int v12_build(PlanImpl *p)
{
  PlanV12 &v = p->v12;
  cudaStream_t st = p->stream;
  const int n = p->n;
  // filters (src/sfft.cc:355-364): only n/B+1 response entries are ever read
  // (cf12.cc:371-385), so keep the window [-n/2B, +n/2B]
  const int half_loc = (n / v.B_loc) / 2, half_est = (n / v.B_est) / 2;
  FilterSpec specs[2] = {{v.lobe_loc, v.tol_loc, v.b_loc, half_loc}, {v.lobe_est, v.tol_est, v.b_est, half_est}};
  DeviceFilter *outs[2] = {&v.filt[0], &v.filt[1]};
  if (build_filters(p->logn, 2, specs, outs, st)) return -1;
  v.geom.w[0] = v.filt[0].w;
  v.geom.w[1] = v.filt[1].w;

  // twiddle table for the largest bucket FFT of this plan
  int twN = v.B_loc > v.B_est ? v.B_loc : v.B_est;
  if (v.with_comb && v.W_Comb > twN) twN = v.W_Comb;
  v.log_twN = ilog2((unsigned)twN);
  std::vector<cplx> tw((size_t)(twN > 1 ? twN - 1 : 1));
  host_twiddle_levels(twN, tw.data());
  SFFTB_CUDA(cudaMalloc(&v.d_tw, sizeof(cplx) * tw.size()));
  SFFTB_CUDA(cudaMemcpyAsync(v.d_tw, tw.data(), sizeof(cplx) * tw.size(), cudaMemcpyHostToDevice, st));
  SFFTB_CUDA(cudaStreamSynchronize(st));

  // result-list capacities
  const long long seg = n / v.B_loc;
  const int first_loops = v.loops_loc - v.loops_thresh + 1;
  long long voted = first_loops > 0 ? (long long)first_loops * v.B_thresh * seg : 0;
  if (voted > n) voted = n;
  if (voted < 1) voted = 1;
  v.max_voted = voted;
  if (v.with_comb) {
    long long nc = (long long)v.Comb_loops * v.B_thresh;
    if (nc > v.W_Comb) nc = v.W_Comb;
    v.max_hits = nc * (n / v.W_Comb);       // cf12.cc:505-512
  } else {
    v.max_hits = voted;
  }
  v.ints_per_sig = 2 * v.geom.loops + v.Comb_loops;
  for (int i = 0; i < kStageSlots; i++) SFFTB_CUDA(cudaEventCreateWithFlags(&v.stage_ev[i], cudaEventDisableTiming));
  if (v.B_loc != v.B_est || v.with_comb) {
    SFFTB_CUDA(cudaStreamCreateWithFlags(&v.side_stream, cudaStreamNonBlocking));
    SFFTB_CUDA(cudaEventCreateWithFlags(&v.side_fork, cudaEventDisableTiming));
    SFFTB_CUDA(cudaEventCreateWithFlags(&v.side_join, cudaEventDisableTiming));
  }
  return v12_ensure_capacity(p, 1);
}

static void v12_free_scratch(PlanV12 &v)
{
  if (v.graph_exec) cudaGraphExecDestroy(v.graph_exec);
  if (v.graph) cudaGraphDestroy(v.graph);
  v.graph_exec = nullptr; v.graph = nullptr;

  cudaFree(v.d_xs); cudaFree(v.d_J); cudaFree(v.d_bitmap); cudaFree(v.d_gkeys);
  cudaFree(v.d_voted); cudaFree(v.d_voted_count); cudaFree(v.d_hit_loc); cudaFree(v.d_hit_val);
  cudaFree(v.d_count); cudaFree(v.d_comb_xs); cudaFree(v.d_comb_J); cudaFree(v.d_comb_bm);
  cudaFree(v.d_appr_bm); cudaFree(v.d_approved); cudaFree(v.d_num_comb); cudaFree(v.d_stage);
  cudaFree(v.d_xt); v.d_xt = nullptr;
  cudaFree(v.d_run_unsafe); v.d_run_unsafe = nullptr;
  cudaFree(v.d_tile_counter); v.d_tile_counter = nullptr;
  for (int i = 0; i < kStageSlots; i++) {
    if (v.h_stage[i]) cudaFreeHost(v.h_stage[i]);
    v.h_stage[i] = nullptr;
  }
  if (v.h_counts) cudaFreeHost(v.h_counts);
  v.h_counts = nullptr;
  v.d_xs = nullptr; v.d_J = nullptr; v.d_bitmap = nullptr; v.d_gkeys = nullptr;
  v.d_voted = nullptr; v.d_voted_count = nullptr; v.d_hit_loc = nullptr; v.d_hit_val = nullptr;
  v.d_count = nullptr; v.d_comb_xs = nullptr; v.d_comb_J = nullptr; v.d_comb_bm = nullptr;
  v.d_appr_bm = nullptr; v.d_approved = nullptr; v.d_num_comb = nullptr; v.d_stage = nullptr;
  v.cap = 0;
}

int v12_ensure_capacity(PlanImpl *p, int nsig)
{
  PlanV12 &v = p->v12;
  if (nsig <= v.cap) return 0;
  if (v.shard.attached || v.shard.d_flags) {
    // peers hold mappings of d_xs (sfftb_shard_export / _attach): it must not move
    set_error("this plan's buffers are exported to peer GPUs (sharded transform): batches larger than the "
              "capacity at export time need sfftb_shard_detach first");
    return -1;
  }
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  v12_free_scratch(v);
  const long long S = nsig;
  const int num = v.B_thresh;
  const int words_loc = v.B_loc >= 32 ? v.B_loc / 32 : 1;
  SFFTB_CUDA(cudaMalloc(&v.d_xs, sizeof(cplx) * S * v.x_samp_size));
  SFFTB_CUDA(cudaMalloc(&v.d_J, sizeof(int) * S * v.loops_loc * num));
  SFFTB_CUDA(cudaMalloc(&v.d_bitmap, sizeof(unsigned) * S * v.loops_loc * words_loc));
  SFFTB_CUDA(cudaMalloc(&v.d_voted, sizeof(int) * S * v.max_voted));
  SFFTB_CUDA(cudaMalloc(&v.d_voted_count, sizeof(int) * S));
  SFFTB_CUDA(cudaMalloc(&v.d_hit_loc, sizeof(int) * S * v.max_hits));
  SFFTB_CUDA(cudaMalloc(&v.d_hit_val, sizeof(cplx) * S * v.max_hits));
  SFFTB_CUDA(cudaMalloc(&v.d_count, sizeof(int) * S));
  SFFTB_CUDA(cudaMalloc(&v.d_stage, sizeof(int) * S * v.ints_per_sig));
  long long gk = 0;
  if (v.B_loc > 16384) gk = (long long)v.loops_loc * select_gkeys_per_row(v.B_loc);
  if (v.with_comb) {
    const int W = v.W_Comb, words = W >= 32 ? W / 32 : 1;
    SFFTB_CUDA(cudaMalloc(&v.d_comb_xs, sizeof(cplx) * S * v.Comb_loops * W));
    SFFTB_CUDA(cudaMalloc(&v.d_comb_J, sizeof(int) * S * v.Comb_loops * num));
    SFFTB_CUDA(cudaMalloc(&v.d_comb_bm, sizeof(unsigned) * S * v.Comb_loops * words));
    SFFTB_CUDA(cudaMalloc(&v.d_appr_bm, sizeof(unsigned) * S * words));
    SFFTB_CUDA(cudaMalloc(&v.d_approved, sizeof(int) * S * W));
    SFFTB_CUDA(cudaMalloc(&v.d_num_comb, sizeof(int) * S));
    if (W > 16384 && (long long)v.Comb_loops * select_gkeys_per_row(W) > gk)
      gk = (long long)v.Comb_loops * select_gkeys_per_row(W);
    v.max_comb = (int)((long long)v.Comb_loops * num < W ? (long long)v.Comb_loops * num : W);
    if (v2_struct_supported(v.geom, ilog2((unsigned)W)) && !getenv("SFFTB_NO_V2_STRUCT")) {
      SFFTB_CUDA(cudaMalloc(&v.d_xt, sizeof(cplx) * S * v.x_samp_size));
      SFFTB_CUDA(cudaMalloc(&v.d_run_unsafe, (size_t)(S * v.x_samp_size) >> v2_struct_log_tile(v.geom, ilog2((unsigned)W))));
      SFFTB_CUDA(cudaMalloc(&v.d_tile_counter, sizeof(unsigned) * S));
    }
  }
  v.gkeys_per_sig = gk;
  if (gk) SFFTB_CUDA(cudaMalloc(&v.d_gkeys, sizeof(unsigned long long) * S * gk));
  for (int i = 0; i < kStageSlots; i++)
    SFFTB_CUDA(cudaHostAlloc(&v.h_stage[i], sizeof(int) * S * v.ints_per_sig, cudaHostAllocDefault));
  SFFTB_CUDA(cudaHostAlloc(&v.h_counts, sizeof(long long) * S, cudaHostAllocDefault));
  v.cap = nsig;
  return 0;
}

void v12_shard_release(PlanImpl *p)
{
  ShardPeers &sh = p->v12.shard;
  if (sh.graph_exec) cudaGraphExecDestroy(sh.graph_exec);
  if (sh.graph) cudaGraphDestroy(sh.graph);
  sh.graph_exec = nullptr; sh.graph = nullptr;
  for (int i = 0; i < sh.n_opened; i++) cudaIpcCloseMemHandle(sh.opened[i]);
  sh.n_opened = 0;
  cudaFree(sh.d_flags);
  sh.d_flags = nullptr;
  sh.attached = false;
  sh.world = 1; sh.rank = 0;
}

void v12_free(PlanImpl *p)
{
  PlanV12 &v = p->v12;
  v12_shard_release(p);
  v12_free_scratch(v);
  free_filter(&v.filt[0]);
  free_filter(&v.filt[1]);
  cudaFree(v.d_tw);
  v.d_tw = nullptr;
  if (v.h_gstage) cudaFreeHost(v.h_gstage);
  if (v.h_gx) cudaFreeHost(v.h_gx);
  cudaFree(v.d_gx);
  if (v.g_ev) cudaEventDestroy(v.g_ev);
  v.h_gstage = nullptr; v.h_gx = nullptr; v.d_gx = nullptr; v.g_ev = nullptr;
  for (int i = 0; i < kStageSlots; i++) {
    if (v.stage_ev[i]) cudaEventDestroy(v.stage_ev[i]);
    v.stage_ev[i] = nullptr;
  }
  if (v.side_stream) cudaStreamDestroy(v.side_stream);
  if (v.side_fork) cudaEventDestroy(v.side_fork);
  if (v.side_join) cudaEventDestroy(v.side_join);
  v.side_stream = nullptr; v.side_fork = nullptr; v.side_join = nullptr;
}

// One transform's worth of libc randomness, in the reference's order:
// `loops` rejection loops of random() % n until odd (cf12.cc:465-474), then one
// drand48() per Comb loop (cf12.cc:62).
int v12_draw(const PlanImpl *p, sfftb_draw *d)
{
  const PlanV12 &v = p->v12;
  const int n = p->n;
  d->loops = v.geom.loops;
  for (int i = 0; i < v.geom.loops; i++) {
    int a = 0;
    while (gcd_ref(a, n) != 1) a = (int)(random() % n);
    d->a[i] = a;
    d->ai[i] = mod_inverse(a, n);
  }
  if (v.with_comb) {
    const int sigma = n / v.W_Comb;
    for (int c = 0; c < v.Comb_loops; c++) d->comb_offset[c] = (int)(unsigned)floor(drand48() * sigma);
  }
  return 0;
}

// ---- transform stages (shared by the single-GPU driver and the sharded one) ----

static int v12_stage_draws(PlanImpl *p, int nsig, const sfftb_draw *draws)
{
  PlanV12 &v = p->v12;
  cudaStream_t st = p->stream;
  const int loops = v.geom.loops;
  const int slot = v.stage_next;
  v.stage_next = (v.stage_next + 1) % kStageSlots;
  SFFTB_CUDA(cudaEventSynchronize(v.stage_ev[slot]));
  int *hs = v.h_stage[slot];
  int *h_perm = hs;
  int *h_coff = hs + (long long)nsig * 2 * loops;
  for (int s = 0; s < nsig; s++) {
    const sfftb_draw &d = draws[s];
    memcpy(h_perm + (long long)s * 2 * loops, d.a, sizeof(int) * loops);
    memcpy(h_perm + (long long)s * 2 * loops + loops, d.ai, sizeof(int) * loops);
    for (int c = 0; c < v.Comb_loops; c++) h_coff[(long long)s * v.Comb_loops + c] = d.comb_offset[c];
  }
  SFFTB_CUDA(cudaMemcpyAsync(v.d_stage, hs, sizeof(int) * (long long)nsig * v.ints_per_sig,
                             cudaMemcpyHostToDevice, st));
  SFFTB_CUDA(cudaEventRecord(v.stage_ev[slot], st));
  SFFTB_CUDA(cudaMemsetAsync(v.d_voted_count, 0, sizeof(int) * nsig, st));
  v.cur_nsig = nsig;
  timer_mark(p, "stage_draws");
  return 0;
}

// Comb pre-filter (v2)  cf12.cc:483-512
static int v12_stage_comb(PlanImpl *p, const cplx *d_in, const unsigned long long *x_ind, long long stride, int nsig)
{
  PlanV12 &v = p->v12;
  if (!v.with_comb) return 0;
  // The Comb filter and the gather + bucket FFTs are independent until estimation: the Comb's four
  // small kernels (~38 us of latency) run on the side stream, under the gather (a parallel branch
  // of the captured graph); v12_join_side brings them back before estimation.
  cudaStream_t st = p->stream;
  const bool forked = v.side_stream && !p->timer.enabled;
  if (forked) {
    SFFTB_CUDA(cudaEventRecord(v.side_fork, p->stream));        // after the draws have been queued
    SFFTB_CUDA(cudaStreamWaitEvent(v.side_stream, v.side_fork, 0));
    st = v.side_stream;
  }
  const int num = v.B_thresh, loops = v.geom.loops;
  const int *d_coff = v.d_stage + (long long)nsig * 2 * loops;
  const int W = v.W_Comb, logW = ilog2((unsigned)W), words = W >= 32 ? W / 32 : 1;
  if (launch_comb_sample(d_in, x_ind, stride, d_coff, v.Comb_loops, logW, p->logn, v.d_comb_xs,
                         (long long)v.Comb_loops * W, nsig, st)) return -1;
  if (fft_dit_inplace(v.d_comb_xs, logW, v.Comb_loops, W, nsig, (long long)v.Comb_loops * W,
                      v.d_tw, v.log_twN, -1, st)) return -1;
  SelectArgs sa;
  sa.xs = v.d_comb_xs; sa.xs_stride = (long long)v.Comb_loops * W; sa.row_stride = W;
  sa.logB = logW; sa.num = num;
  sa.J = v.d_comb_J; sa.J_sig_stride = (long long)v.Comb_loops * num;
  sa.bitmap = v.d_comb_bm; sa.bm_sig_stride = (long long)v.Comb_loops * words;
  sa.gkeys = W > 16384 ? v.d_gkeys : nullptr; sa.gk_sig_stride = v.gkeys_per_sig;
  sa.gk_scratch_off = (long long)v.Comb_loops * W;
  sa.row_begin = 0; sa.row_step = 1;
  if (launch_select(sa, v.Comb_loops, nsig, st)) return -1;
  if (launch_comb_merge(v.d_comb_bm, v.Comb_loops, W, p->n / W, v.d_appr_bm, v.d_approved,
                        v.d_num_comb, v.d_count, nsig, st)) return -1;
  if (forked) {
    SFFTB_CUDA(cudaEventRecord(v.side_join, st));
    v.side_pending = true;
  }
  timer_mark(p, "comb");
  return 0;
}

// permuted windowed gather (cf12.cc:222-261) + bucket FFTs (cf12.cc:270-275) of loops [lb, le)
static int v12_stage_bucketize(PlanImpl *p, const cplx *d_in, const unsigned long long *x_ind, long long stride,
                               int nsig, int lb, int le)
{
  PlanV12 &v = p->v12;
  cudaStream_t st = p->stream;
  const LoopGeom &g = v.geom;
  if (le <= lb) return 0;
  GatherArgs ga;
  ga.x = d_in; ga.x_indirect = x_ind; ga.x_stride = stride;
  ga.taps[0] = v.filt[0].time; ga.taps[1] = v.filt[1].time;
  ga.perm = v.d_stage; ga.xs = v.d_xs;
  ga.loop_begin = lb; ga.loop_step = 1;
  if (launch_gather(g, ga, le - lb, nsig, st)) return -1;
  timer_mark(p, "gather");

  const int loc_b = lb < v.loops_loc ? lb : v.loops_loc, loc_e = le < v.loops_loc ? le : v.loops_loc;
  const int est_b = (lb > v.loops_loc ? lb : v.loops_loc) - v.loops_loc;
  const int est_e = (le > v.loops_loc ? le : v.loops_loc) - v.loops_loc;
  if (v.B_loc == v.B_est) {
    if (fft_dit_inplace(v.d_xs + (long long)lb * v.B_loc, g.logB[0], le - lb, v.B_loc, nsig,
                        v.x_samp_size, v.d_tw, v.log_twN, -1, st)) return -1;
  } else {
    // the estimation rows are not needed before the estimation stage: fork their FFT onto the
    // side stream (a parallel branch of the captured graph) and join in v12_stage_finish
    cudaStream_t est_st = st;
    if (est_e > est_b && loc_e > loc_b && v.side_stream && !p->timer.enabled) {
      SFFTB_CUDA(cudaEventRecord(v.side_fork, st));
      SFFTB_CUDA(cudaStreamWaitEvent(v.side_stream, v.side_fork, 0));
      est_st = v.side_stream;
    }
    if (est_e > est_b &&
        fft_dit_inplace(v.d_xs + (long long)v.loops_loc * v.B_loc + (long long)est_b * v.B_est, g.logB[1],
                        est_e - est_b, v.B_est, nsig, v.x_samp_size, v.d_tw, v.log_twN, -1, est_st)) return -1;
    if (est_st != st) {
      SFFTB_CUDA(cudaEventRecord(v.side_join, est_st));
      v.side_pending = true;
    }
    if (loc_e > loc_b &&
        fft_dit_inplace(v.d_xs + (long long)loc_b * v.B_loc, g.logB[0], loc_e - loc_b, v.B_loc, nsig,
                        v.x_samp_size, v.d_tw, v.log_twN, -1, st)) return -1;
  }
  timer_mark(p, "bucket_fft");
  return 0;
}

// make the main stream wait for the estimation rows' FFT if it was forked onto the side stream
static int v12_join_side(PlanImpl *p)
{
  PlanV12 &v = p->v12;
  if (v.side_pending) {
    SFFTB_CUDA(cudaStreamWaitEvent(p->stream, v.side_join, 0));
    v.side_pending = false;
  }
  return 0;
}

// |.|^2 + top-2k per location loop (cf12.cc:278-302) and voting (:304-323)
static int v12_stage_locate(PlanImpl *p, int nsig)
{
  PlanV12 &v = p->v12;
  cudaStream_t st = p->stream;
  const LoopGeom &g = v.geom;
  const int num = v.B_thresh;
  const int *d_perm = v.d_stage;
  const int words_loc = v.B_loc >= 32 ? v.B_loc / 32 : 1;
  SelectArgs sa;
  sa.xs = v.d_xs; sa.xs_stride = v.x_samp_size; sa.row_stride = v.B_loc;
  sa.logB = g.logB[0]; sa.num = num;
  sa.J = v.d_J; sa.J_sig_stride = (long long)v.loops_loc * num;
  sa.bitmap = v.d_bitmap; sa.bm_sig_stride = (long long)v.loops_loc * words_loc;
  sa.gkeys = v.B_loc > 16384 ? v.d_gkeys : nullptr; sa.gk_sig_stride = v.gkeys_per_sig;
  sa.gk_scratch_off = (long long)v.loops_loc * v.B_loc;
  sa.row_begin = 0; sa.row_step = 1;
  if (launch_select(sa, v.loops_loc, nsig, st)) return -1;
  timer_mark(p, "select");

  VoteArgs va;
  va.perm = d_perm;
  va.J = v.d_J; va.J_sig_stride = sa.J_sig_stride;
  va.bitmap = v.d_bitmap; va.bm_sig_stride = sa.bm_sig_stride;
  va.comb_bitmap = v.with_comb ? v.d_appr_bm : nullptr;
  va.comb_sig_stride = v.with_comb ? (v.W_Comb >= 32 ? v.W_Comb / 32 : 1) : 0;
  va.W_mask = v.W_Comb - 1;
  va.hits = v.d_voted; va.hits_cap = v.max_voted; va.count = v.d_voted_count;
  va.num = num; va.thresh = v.loops_thresh;
  if (launch_vote(g, va, nsig, st)) return -1;
  timer_mark(p, "vote");
  v.locate_skipped = false;
  return 0;
}

// v2 only: the selection/voting of the last transform was skipped (see v12_stage_finish);
// run it now on the bucket spectra still in place (debug hooks that read J / the voted set)
int v12_locate_on_demand(PlanImpl *p)
{
  PlanV12 &v = p->v12;
  if (!v.locate_skipped) return 0;
  const int nsig = v.cur_nsig > 0 ? v.cur_nsig : 1;
  SFFTB_CUDA(cudaMemsetAsync(v.d_voted_count, 0, sizeof(int) * nsig, p->stream));
  if (v12_stage_locate(p, nsig)) return -1;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  return 0;
}

// location (selection + voting), then estimation (cf12.cc:341-419)
static int v12_stage_finish(PlanImpl *p, int nsig, int slice_rank, int slice_world)
{
  PlanV12 &v = p->v12;
  cudaStream_t st = p->stream;
  const LoopGeom &g = v.geom;
  const int *d_perm = v.d_stage;
  // v2 estimates the pre-filled list {jj*W + r : r approved} (cf12.cc:505-512); every voted
  // location has an approved residue (cf12.cc:126-184), so it is in that list already and the
  // reference's voting only appends duplicates that `ans[loc] = value` overwrites with the
  // same bits.  The result is identical without it: skip selection and voting (7 % of a C2
  // transform) unless a debug hook asks for their outputs.
  if (v.with_comb) {
    v.locate_skipped = true;
  } else if (v12_stage_locate(p, nsig)) {
    return -1;
  }
  if (v12_join_side(p)) return -1;       // the estimation rows' FFT ran beside selection and voting

  EstimateArgs ea;
  ea.perm = d_perm;
  ea.xs = v.d_xs; ea.xs_stride = v.x_samp_size;
  ea.fwin[0] = v.filt[0].fwin; ea.fwin[1] = v.filt[1].fwin;
  ea.fw_half[0] = v.filt[0].fw_half; ea.fw_half[1] = v.filt[1].fw_half;
  ea.fdr[0] = v.filt[0].fdr; ea.fdr[1] = v.filt[1].fdr;
  ea.hits = v.d_voted; ea.hits_cap = v.max_voted;
  ea.count = v.d_voted_count;
  ea.approved = v.with_comb ? v.d_approved : nullptr;
  ea.approved_stride = v.W_Comb;
  ea.num_comb = v.d_num_comb;
  ea.W = v.W_Comb; ea.n_over_W = v.with_comb ? p->n / v.W_Comb : 0;
  ea.out_loc = v.d_hit_loc; ea.out_val = v.d_hit_val; ea.out_cap = v.max_hits;
  // only the v2 pre-filled list is worth slicing across GPUs; v1's hit order is not
  // identical across ranks (atomic append), so v1 estimates every hit on every rank
  const bool slice = slice_world > 1 && v.with_comb;
  ea.slice_rank = slice ? slice_rank : 0;
  ea.slice_world = slice ? slice_world : 1;
  ea.slice_count = slice ? v.d_count : nullptr;
  if (v.with_comb && v.d_xt) {
    V2StructArgs sa2;
    sa2.perm = d_perm; sa2.xs = v.d_xs; sa2.xt = v.d_xt;
    sa2.run_unsafe = v.d_run_unsafe; sa2.tile_counter = v.d_tile_counter;
    sa2.fwin[0] = ea.fwin[0]; sa2.fwin[1] = ea.fwin[1];
    sa2.fw_half[0] = ea.fw_half[0]; sa2.fw_half[1] = ea.fw_half[1];
    sa2.fdr[0] = ea.fdr[0]; sa2.fdr[1] = ea.fdr[1];
    sa2.approved = v.d_approved; sa2.approved_stride = v.W_Comb; sa2.num_comb = v.d_num_comb;
    sa2.logW = ilog2((unsigned)v.W_Comb);
    sa2.logT = v2_struct_log_tile(g, sa2.logW);
    sa2.out_loc = v.d_hit_loc; sa2.out_val = v.d_hit_val; sa2.out_cap = v.max_hits;
    sa2.slice_rank = ea.slice_rank; sa2.slice_world = ea.slice_world; sa2.slice_count = ea.slice_count;
    if (launch_v2_struct(g, sa2, v.max_comb, nsig, st)) return -1;
  } else {
    if (launch_estimate(g, ea, nsig, v.max_hits, st)) return -1;
  }
  timer_mark(p, "estimate");
  p->last_nsig = nsig;
  return 0;
}

static void fill_stage(const PlanV12 &v, int *hs, int nsig, const sfftb_draw *draws)
{
  const int loops = v.geom.loops;
  int *h_coff = hs + (long long)nsig * 2 * loops;
  for (int s = 0; s < nsig; s++) {
    const sfftb_draw &d = draws[s];
    memcpy(hs + (long long)s * 2 * loops, d.a, sizeof(int) * loops);
    memcpy(hs + (long long)s * 2 * loops + loops, d.ai, sizeof(int) * loops);
    for (int c = 0; c < v.Comb_loops; c++) h_coff[(long long)s * v.Comb_loops + c] = d.comb_offset[c];
  }
}

// Single-signal transform replayed from a CUDA graph: ~10 kernels + 2 small copies become
// one launch, which is what a 2-60 us transform needs (SURVEY 7 "Latency").
static int v12_exec_graph(PlanImpl *p, const cplx *d_in, const sfftb_draw *draw)
{
  PlanV12 &v = p->v12;
  cudaStream_t st = p->stream;
  if (!v.graph_exec) {
    if (!v.h_gstage) {
      SFFTB_CUDA(cudaHostAlloc(&v.h_gstage, sizeof(int) * v.ints_per_sig, cudaHostAllocDefault));
      SFFTB_CUDA(cudaHostAlloc(&v.h_gx, sizeof(unsigned long long), cudaHostAllocDefault));
      SFFTB_CUDA(cudaMalloc(&v.d_gx, sizeof(unsigned long long)));
      SFFTB_CUDA(cudaEventCreateWithFlags(&v.g_ev, cudaEventDisableTiming));
    }
    SFFTB_CUDA(cudaStreamSynchronize(st));
    SFFTB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    if (cudaMemcpyAsync(v.d_stage, v.h_gstage, sizeof(int) * v.ints_per_sig, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (cudaMemcpyAsync(v.d_gx, v.h_gx, sizeof(unsigned long long), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (cudaEventRecordWithFlags(v.g_ev, st, cudaEventRecordExternal) != cudaSuccess) rc = -1;
    if (cudaMemsetAsync(v.d_voted_count, 0, sizeof(int), st) != cudaSuccess) rc = -1;
    if (!rc) rc = v12_stage_comb(p, nullptr, v.d_gx, p->n, 1);
    if (!rc) rc = v12_stage_bucketize(p, nullptr, v.d_gx, p->n, 1, 0, v.geom.loops);
    if (!rc) rc = v12_stage_finish(p, 1, 0, 1);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc || e != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      set_error("CUDA graph capture of the transform failed");
      return -1;
    }
    v.graph = g;
    SFFTB_CUDA(cudaGraphInstantiate(&v.graph_exec, v.graph, 0));
  }
  SFFTB_CUDA(cudaEventSynchronize(v.g_ev));      // previous replay has consumed the staging buffers
  fill_stage(v, v.h_gstage, 1, draw);
  *v.h_gx = (unsigned long long)(uintptr_t)d_in;
  SFFTB_CUDA(cudaGraphLaunch(v.graph_exec, st));
  g_launches += v.graph_kernels;
  p->last_nsig = 1;
  return 0;
}

static bool graphs_enabled()
{
  static int on = -1;
  if (on < 0) on = getenv("SFFTB_NO_GRAPH") ? 0 : 1;
  return on == 1;
}

int v12_exec(PlanImpl *p, const cplx *d_in, long long stride, int nsig, const sfftb_draw *draws)
{
  PlanV12 &v = p->v12;
  if (v12_ensure_capacity(p, nsig)) return -1;
  // the first transforms run launch by launch (lazy one-time kernel attributes are set there),
  // then single-signal transforms are replayed from a captured graph
  if (nsig == 1 && !p->timer.enabled && graphs_enabled() && v.plain_execs >= 1) {
    if (v12_exec_graph(p, d_in, draws) == 0) return 0;
    v.plain_execs = -1000000;      // capture failed once: stay on the plain path
  }
  v.plain_execs++;
  timer_begin(p);
  const long long launches0 = g_launches;
  if (v12_stage_draws(p, nsig, draws)) return -1;
  if (v12_stage_comb(p, d_in, nullptr, stride, nsig)) return -1;
  if (v12_stage_bucketize(p, d_in, nullptr, stride, nsig, 0, v.geom.loops)) return -1;
  if (v12_stage_finish(p, nsig, 0, 1)) return -1;
  v.graph_kernels = (int)(g_launches - launches0);
  return 0;
}

// ---- multi-GPU sharding of one transform (include/sfft.h) ----

void v12_shard_loops(const PlanImpl *p, int rank, int world, int *begin, int *end)
{
  const int loops = p->v12.geom.loops;
  *begin = (int)((long long)loops * rank / world);
  *end = (int)((long long)loops * (rank + 1) / world);
}

int v12_shard_bucketize(PlanImpl *p, const cplx *d_in, const sfftb_draw *draw, int rank, int world)
{
  PlanV12 &v = p->v12;
  if (v12_ensure_capacity(p, 1)) return -1;
  timer_begin(p);
  if (v12_stage_draws(p, 1, draw)) return -1;
  if (v12_stage_comb(p, d_in, nullptr, p->n, 1)) return -1;       // tiny; replicated on every rank
  // rows this rank does not own must be exactly zero for the sum over ranks
  SFFTB_CUDA(cudaMemsetAsync(v.d_xs, 0, sizeof(cplx) * v.x_samp_size, p->stream));
  int lb, le;
  v12_shard_loops(p, rank, world, &lb, &le);
  if (v12_stage_bucketize(p, d_in, nullptr, p->n, 1, lb, le)) return -1;
  return v12_join_side(p);               // the caller's collective follows on the plan's stream
}

int v12_shard_finish(PlanImpl *p, int rank, int world) { return v12_stage_finish(p, 1, rank, world); }

// ---- peer exchange (shard.cu): one sharded transform, every step on the plan's stream ----
static int v12_shard_body(PlanImpl *p, const cplx *d_in, const unsigned long long *x_ind)
{
  PlanV12 &v = p->v12;
  const ShardPeers &sh = v.shard;
  int lb, le;
  v12_shard_loops(p, sh.rank, sh.world, &lb, &le);
  if (v12_stage_comb(p, d_in, x_ind, p->n, 1)) return -1;              // tiny; replicated on every rank
  if (v12_stage_bucketize(p, d_in, x_ind, p->n, 1, lb, le)) return -1;
  if (v12_join_side(p)) return -1;       // every owned row must be complete before it is stored into the peers
  // rows of loops [lb, le) are one contiguous block of the spectra buffer (cf12.cc:228-230)
  const long long off = lb < v.loops_loc ? (long long)lb * v.B_loc
                                         : (long long)v.loops_loc * v.B_loc + (long long)(lb - v.loops_loc) * v.B_est;
  const long long end = le < v.loops_loc ? (long long)le * v.B_loc
                                         : (long long)v.loops_loc * v.B_loc + (long long)(le - v.loops_loc) * v.B_est;
  if (launch_shard_push(sh, off, end - off, p->stream)) return -1;
  if (launch_shard_wait_ready(sh, p->stream)) return -1;
  timer_mark(p, "exchange");
  if (v12_stage_finish(p, 1, sh.rank, sh.world)) return -1;
  if (launch_shard_done(sh, p->stream)) return -1;
  return 0;
}

int v12_shard_exec(PlanImpl *p, const cplx *d_in, const sfftb_draw *draw)
{
  PlanV12 &v = p->v12;
  ShardPeers &sh = v.shard;
  cudaStream_t st = p->stream;
  if (sh.plain_execs < 1 || p->timer.enabled || !graphs_enabled()) {
    // first transform launch by launch (lazy one-time kernel attributes are set there)
    timer_begin(p);
    const long long launches0 = g_launches;
    if (v12_stage_draws(p, 1, draw)) return -1;
    if (v12_shard_body(p, d_in, nullptr)) return -1;
    sh.graph_kernels = (int)(g_launches - launches0);
    sh.plain_execs++;
    return 0;
  }
  if (!sh.graph_exec) {
    if (!v.h_gstage) {
      SFFTB_CUDA(cudaHostAlloc(&v.h_gstage, sizeof(int) * v.ints_per_sig, cudaHostAllocDefault));
      SFFTB_CUDA(cudaHostAlloc(&v.h_gx, sizeof(unsigned long long), cudaHostAllocDefault));
      SFFTB_CUDA(cudaMalloc(&v.d_gx, sizeof(unsigned long long)));
      SFFTB_CUDA(cudaEventCreateWithFlags(&v.g_ev, cudaEventDisableTiming));
    }
    SFFTB_CUDA(cudaStreamSynchronize(st));
    SFFTB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    if (cudaMemcpyAsync(v.d_stage, v.h_gstage, sizeof(int) * v.ints_per_sig, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (cudaMemcpyAsync(v.d_gx, v.h_gx, sizeof(unsigned long long), cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (cudaEventRecordWithFlags(v.g_ev, st, cudaEventRecordExternal) != cudaSuccess) rc = -1;
    if (cudaMemsetAsync(v.d_voted_count, 0, sizeof(int), st) != cudaSuccess) rc = -1;
    if (!rc) rc = v12_shard_body(p, nullptr, v.d_gx);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc || e != cudaSuccess || !g) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      set_error("CUDA graph capture of the sharded transform failed");
      return -1;
    }
    sh.graph = g;
    SFFTB_CUDA(cudaGraphInstantiate(&sh.graph_exec, sh.graph, 0));
  }
  SFFTB_CUDA(cudaEventSynchronize(v.g_ev));      // previous replay has consumed the staging buffers
  fill_stage(v, v.h_gstage, 1, draw);
  *v.h_gx = (unsigned long long)(uintptr_t)d_in;
  SFFTB_CUDA(cudaGraphLaunch(sh.graph_exec, st));
  g_launches += sh.graph_kernels;
  p->last_nsig = 1;
  v.cur_nsig = 1;
  return 0;
}

// Which part of the single-GPU result list rank `rank` of `world` produces, as (offset, count)
// in that list's order.  v1: everything (estimation is replicated).  v2: a block of the
// pre-filled list -- whole tiles when the structured estimation kernel is in use, else hits.
// Needs the Comb result of the last transform (synchronises the stream).
int v12_shard_slice(PlanImpl *p, int rank, int world, long long *offset, long long *count)
{
  PlanV12 &v = p->v12;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  if (!v.with_comb) {
    int c = 0;
    SFFTB_CUDA(cudaMemcpy(&c, v.d_voted_count, sizeof(int), cudaMemcpyDeviceToHost));
    *offset = 0;
    *count = c > v.max_voted ? v.max_voted : c;
    return 0;
  }
  int nc = 0;
  SFFTB_CUDA(cudaMemcpy(&nc, v.d_num_comb, sizeof(int), cudaMemcpyDeviceToHost));
  const int logNW = p->logn - ilog2((unsigned)v.W_Comb);
  if (v.d_xt) {
    const int logT = v2_struct_log_tile(v.geom, ilog2((unsigned)v.W_Comb));
    const long long tiles = (long long)nc << (logNW - logT);
    const long long lo = tiles * rank / world, hi = tiles * (rank + 1) / world;
    *offset = lo << logT;
    *count = (hi - lo) << logT;
  } else {
    const long long total = (long long)nc << logNW;
    const long long lo = total * rank / world, hi = total * (rank + 1) / world;
    *offset = lo;
    *count = hi - lo;
  }
  return 0;
}

}  // namespace sfftb
