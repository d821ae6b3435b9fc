// plan.cuh -- internal plan object behind the public sfft_plan (include/sfft.h).
//
// Mirrors the reference's plan data (struct sfft_v1v2_data / sfft_v3_data,
// src/sfft.h:78-154) but everything a transform touches lives in HBM, and mode
// (v1 vs v2) is per-plan state instead of the reference's process globals
// (src/common.cc:22-23).
#pragma once

#include <vector>

#include "../../include/sfft.h"
#include "common.cuh"
#include "plan_builder.cuh"
#include "v12_kernels.cuh"

namespace sfftb {

constexpr int kStageSlots = 8;     // ring of pinned staging buffers for per-transform draws
constexpr int kMaxStages = 16;

struct StageTimer {
  bool enabled = false;
  int count = 0;
  const char *names[kMaxStages];
  cudaEvent_t ev[kMaxStages + 1];
  bool created = false;
};

// NVLink peer exchange of the sharded transform (shard.cu): every rank's bucket-spectra
// buffer and a small flag block are mapped into every other rank through CUDA IPC; ranks
// store their rows straight into their peers' buffers and signal with flags.
constexpr int kMaxPeers = 16;
// layout of the flag block (unsigned words)
enum { SH_READY = 0, SH_DONE = kMaxPeers, SH_EPOCH = 2 * kMaxPeers, SH_ERR = 2 * kMaxPeers + 1,
       SH_CTR = 2 * kMaxPeers + 8, SH_WORDS = 3 * kMaxPeers + 8 };
struct ShardPeers {
  bool attached = false;
  int rank = 0, world = 1;
  unsigned *d_flags = nullptr;                 // local flag block, IPC-exported
  cplx *peer_xs[kMaxPeers] = {nullptr};        // [rank] = local d_xs
  unsigned *peer_flags[kMaxPeers] = {nullptr}; // [rank] = local d_flags
  void *opened[2 * kMaxPeers] = {nullptr};     // mappings to close on detach
  int n_opened = 0;
  cudaGraphExec_t graph_exec = nullptr;
  cudaGraph_t graph = nullptr;
  int graph_kernels = 0;
  int plain_execs = 0;
};

struct PlanV12 {
  // ---- derived parameters (src/sfft.cc:298-353) ----
  int with_comb = 0;
  int B_loc = 0, B_est = 0, B_thresh = 0, W_Comb = 0, Comb_loops = 0;
  int loops_loc = 0, loops_thresh = 0, loops_est = 0;
  int b_loc = 0, b_est = 0;
  double tol_loc = 0, tol_est = 0, lobe_loc = 0, lobe_est = 0;
  long long x_samp_size = 0;
  LoopGeom geom;
  DeviceFilter filt[2];            // [0] location, [1] estimation
  cplx *d_tw = nullptr;            // twiddle table for the bucket FFTs
  int log_twN = 0;

  // ---- scratch, sized for `cap` signals ----
  int cap = 0;
  long long max_hits = 0, max_voted = 0;
  cplx *d_xs = nullptr;            // [cap][x_samp_size]
  int *d_J = nullptr;              // [cap][loops_loc][num]
  unsigned *d_bitmap = nullptr;    // [cap][loops_loc][B_loc/32]
  unsigned long long *d_gkeys = nullptr;
  long long gkeys_per_sig = 0;
  int *d_voted = nullptr;          // [cap][max_voted]
  int *d_voted_count = nullptr;    // [cap]
  int *d_hit_loc = nullptr;        // [cap][max_hits]
  cplx *d_hit_val = nullptr;       // [cap][max_hits]
  int *d_count = nullptr;          // [cap]   (v2: size of the pre-filled list)
  // v2
  cplx *d_comb_xs = nullptr;       // [cap][Comb_loops][W]
  int *d_comb_J = nullptr;         // [cap][Comb_loops][num]
  unsigned *d_comb_bm = nullptr;   // [cap][Comb_loops][W/32]
  unsigned *d_appr_bm = nullptr;   // [cap][W/32]
  int *d_approved = nullptr;       // [cap][W]
  int *d_num_comb = nullptr;       // [cap]
  cplx *d_xt = nullptr;            // v2 structured estimation: class-major copy of d_xs
  unsigned char *d_run_unsafe = nullptr;   // [cap][x_samp_size / tile]
  unsigned *d_tile_counter = nullptr;      // [cap]
  int max_comb = 0;
  // per-transform draws: a[loops], ai[loops] per signal, then comb offsets
  int *d_stage = nullptr;          // [cap * ints_per_sig]
  int ints_per_sig = 0;
  int *h_stage[kStageSlots] = {nullptr};
  cudaEvent_t stage_ev[kStageSlots] = {nullptr};
  int stage_next = 0;
  int cur_nsig = 0;
  bool locate_skipped = false;             // v2: J / bitmap / voted list not computed for the last transform
  // CUDA-graph replay of the single-signal transform: the kernel sequence is captured once;
  // per call only the staged draw and the signal pointer (read indirectly) change
  cudaGraphExec_t graph_exec = nullptr;
  cudaGraph_t graph = nullptr;
  int *h_gstage = nullptr;                 // pinned: draws
  unsigned long long *h_gx = nullptr;      // pinned: signal pointer
  unsigned long long *d_gx = nullptr;
  cudaEvent_t g_ev = nullptr;              // staging buffers consumed
  int plain_execs = 0;
  int graph_kernels = 0;                   // kernels in one transform (for the launch counter)
  long long *h_counts = nullptr;   // pinned
  // plans whose location and estimation rows differ in size: the estimation rows' FFT runs on a
  // side stream, under the location rows' FFT + selection + voting (joined before estimation)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t side_fork = nullptr, side_join = nullptr;
  bool side_pending = false;
  ShardPeers shard;
};

struct PlanV3;   // v3.cu

struct PlanImpl {
  int version = 1;     // 1, 2, 3
  int n = 0, logn = 0, k = 0;
  int device = 0;
  int flags = 0;       // the fftw_optimization word the plan was made with
  cudaStream_t own_stream = nullptr, stream = nullptr;
  PlanV12 v12;
  PlanV3 *v3 = nullptr;
  // legacy host-pointer path
  cplx *d_in = nullptr;   long long d_in_elems = 0;
  cplx *d_out = nullptr;
  // legacy path, sparse results: zeros streamed to the caller's `out` while the input
  // streams in, then the few coefficients scattered by the host
  cudaStream_t zero_stream = nullptr;
  cplx *d_zero = nullptr;  long long zero_elems = 0;
  int *h_loc = nullptr;    cplx *h_val = nullptr;     // pinned, kHostScatterCap entries
  int last_nsig = 0;
  StageTimer timer;
};

// plan_v12.cu
int v12_derive(PlanImpl *p, int n, int k, int with_comb, int tuned_by_k);
int v12_build(PlanImpl *p);
int v12_ensure_capacity(PlanImpl *p, int nsig);
void v12_free(PlanImpl *p);
int v12_draw(const PlanImpl *p, sfftb_draw *d);
int v12_exec(PlanImpl *p, const cplx *d_in, long long stride, int nsig, const sfftb_draw *draws);
void v12_shard_loops(const PlanImpl *p, int rank, int world, int *begin, int *end);
int v12_shard_bucketize(PlanImpl *p, const cplx *d_in, const sfftb_draw *draw, int rank, int world);
int v12_shard_finish(PlanImpl *p, int rank, int world);
int v12_shard_exec(PlanImpl *p, const cplx *d_in, const sfftb_draw *draw);
int v12_shard_slice(PlanImpl *p, int rank, int world, long long *offset, long long *count);
void v12_shard_release(PlanImpl *p);
// shard.cu: the peer-exchange kernels
int launch_shard_push(const ShardPeers &sh, long long elem_off, long long elem_cnt, cudaStream_t st);
int launch_shard_wait_ready(const ShardPeers &sh, cudaStream_t st);
int launch_shard_done(const ShardPeers &sh, cudaStream_t st);
int v12_locate_on_demand(PlanImpl *p);

// timing helpers
void timer_begin(PlanImpl *p);
void timer_mark(PlanImpl *p, const char *name);

// v3.cu
int v3_build(PlanImpl *p, int n, int k);
void v3_free(PlanImpl *p);
int v3_draw(const PlanImpl *p, sfftb_draw *d);
int v3_exec(PlanImpl *p, const cplx *d_in, long long stride, int nsig, const sfftb_draw *draws);
int v3_info(const PlanImpl *p, sfftb_info *info);
int v3_result(PlanImpl *p, const int **loc, const cplx **val, const int **count, long long *cap);
int v3_filter_sizes(const PlanImpl *p, int which, int *w, int *fw_len);
DeviceFilter *v3_filter(PlanImpl *p, int which);
long long v3_debug_fetch(PlanImpl *p, const char *what, void *dst, size_t capacity);

}  // namespace sfftb
