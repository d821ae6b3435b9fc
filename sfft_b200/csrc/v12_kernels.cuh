// v12_kernels.cuh -- device stages of the sFFT v1/v2 transform.
//
// Reference path (single-threaded CPU): outer_loop, src/computefourier-1.0-2.0.cc:438-541.
// Here every stage is a grid over (bucket | candidate | hit) x loop x signal, so the
// same kernels serve sfft_exec (one signal) and sfft_exec_many (a batch).
#pragma once

#include "common.cuh"

namespace sfftb {

constexpr int kMaxLoops = 64;   // == SFFTB_MAX_LOOPS in include/sfft.h

// Layout of the per-signal bucket array x_samp (reference cf12.cc:228-230):
//   loops_loc rows of B_loc, then loops_est rows of B_est.
struct LoopGeom {
  int n_mask;          // n - 1
  int logn;
  int loops, loops_loc;
  int logB[2];         // [0] location loops, [1] estimation loops
  int w[2];            // taps per filter
  long long x_samp_size;
};

// permutation table per signal: a[0..loops) then ai[0..loops)  (cf12.cc:465-474)
__host__ __device__ __forceinline__ long long perm_stride(int loops) { return 2ll * loops; }

struct GatherArgs {
  const cplx *x;            // signals
  const unsigned long long *x_indirect;   // if set: device slot holding the signal pointer (CUDA-graph replay)
  long long x_stride;       // elements between signals
  const cplx *taps[2];
  const int *perm;
  cplx *xs;                 // [S][x_samp_size], written in bit-reversed bucket order
  int loop_begin, loop_step;   // loops handled: [loop_begin, loop_begin + nloops) (sharding); loop_step unused
};

struct SelectArgs {
  const cplx *xs;  long long xs_stride;   // bucket spectra (natural order)
  long long row_stride;                   // elements between consecutive rows
  int logB, num;
  int *J;          long long J_sig_stride;      // [S][rows][num]
  unsigned *bitmap; long long bm_sig_stride;    // [S][rows][max(1,B/32)]
  // rows above 16384 buckets: global scratch per signal = [rows_total * B keys]
  // [rows_total * select_big_scratch(B) words of histogram/count scratch at gk_scratch_off]
  unsigned long long *gkeys; long long gk_sig_stride; long long gk_scratch_off;
  int row_begin, row_step;
};
// 64-bit words of global scratch one row of B > 16384 buckets needs: keys, then work area
long long select_big_scratch(int B);
inline long long select_gkeys_per_row(int B) { return (long long)B + select_big_scratch(B); }

struct VoteArgs {
  const int *perm;
  const int *J;         long long J_sig_stride;
  const unsigned *bitmap; long long bm_sig_stride;
  const unsigned *comb_bitmap; long long comb_sig_stride; int W_mask;   // v2 only (else null)
  int *hits;            long long hits_cap;
  int *count;
  int num, thresh;
};

struct EstimateArgs {
  const int *perm;
  const cplx *xs;       long long xs_stride;
  const cplx *fwin[2];  int fw_half[2];
  const double2 *fdr[2];                 // (|f|^2, RN(1/|f|^2)) per window entry
  // v1: explicit hit list;  v2: implicit prefill list (jj*W + approved[i])
  const int *hits;      long long hits_cap;
  const int *count;
  const int *approved;  long long approved_stride; const int *num_comb; int W, n_over_W;
  int *out_loc;         cplx *out_val;  long long out_cap;
  // multi-GPU: this launch covers slice `slice_rank` of `slice_world` of the hit list and
  // records the slice length in slice_count (null when the whole list is covered)
  int slice_rank, slice_world; int *slice_count;
};

int launch_gather(const LoopGeom &g, const GatherArgs &a, int nloops, int nsig, cudaStream_t st);
int launch_select(const SelectArgs &a, int nrows, int nsig, cudaStream_t st);
int launch_vote(const LoopGeom &g, const VoteArgs &a, int nsig, cudaStream_t st);
int launch_estimate(const LoopGeom &g, const EstimateArgs &a, int nsig, long long max_per_sig,
                    cudaStream_t st);
// ---- v2 structured estimation ---------------------------------------------------
// v2's result list is {jj*W + r : r approved, jj < n/W} (cf12.cc:505-512).  For a fixed
// loop j and residue r, ai_j*(jj*W + r) walks the buckets of ONE residue class modulo
// q = W/(n/B) with a constant offset inside the bucket, visiting each of the n/W buckets
// of that class exactly once; restricted to jj == c (mod 2^s) it walks the n/W/2^s buckets
// of one class modulo q*2^s.  So instead of 16.4 M x 20 random 16-byte L2 reads
// (request-rate bound) a CTA owns one (r, c) tile of hits, copies the L bucket groups it
// needs -- contiguous in a class-major copy `xt` of the spectra -- into shared memory, and
// finishes divisions and medians in registers.  Same arithmetic, same results.
struct V2StructArgs {
  const int *perm;                 // per signal: a[loops], ai[loops]
  const cplx *xs;                  // bucket spectra            [nsig][x_samp_size]
  cplx *xt;                        // class-major copy of them   [nsig][x_samp_size]
  unsigned char *run_unsafe;       // per run of 2^logT elements of xt: needs real divisions
  unsigned *tile_counter;          // [nsig] dynamic tile scheduler
  const cplx *fwin[2]; int fw_half[2]; const double2 *fdr[2];
  const int *approved; long long approved_stride; const int *num_comb;
  int logW;                        // W_Comb = 2^logW
  int logT;                        // hits per tile = 2^logT (v2_struct_log_tile)
  int *out_loc; cplx *out_val; long long out_cap;
  int slice_rank, slice_world; int *slice_count;   // slices are whole tiles
};
bool v2_struct_supported(const LoopGeom &g, int logW);
int v2_struct_log_tile(const LoopGeom &g, int logW);
int launch_v2_struct(const LoopGeom &g, const V2StructArgs &a, int max_comb, int nsig, cudaStream_t st);

int launch_filter_den(const cplx *fwin, int len, double2 *fdr, cudaStream_t st);
long long run_div_check(unsigned long long seed, long long count);

// v2: xs[c][bitrev(i)] = x[offset_c + i*sigma]   (cf12.cc:61-67)
int launch_comb_sample(const cplx *x, const unsigned long long *x_indirect, long long x_stride,
                       const int *comb_off, int comb_loops, int logW, int logn, cplx *cxs,
                       long long cxs_stride, int nsig, cudaStream_t st);
// v2: union of the per-loop selections -> sorted residue list, its size, and the
// size of the pre-filled hit list (cf12.cc:492-512)
int launch_comb_merge(const unsigned *loop_bitmaps, int comb_loops, int W, int n_over_W,
                      unsigned *approved_bitmap, int *approved, int *num_comb, int *count,
                      int nsig, cudaStream_t st);

// legacy dense output: out[loc] = val (after a memset)   (sfft.cc:121-123, cf12.cc:413)
int launch_scatter(const int *loc, const cplx *val, const int *count, cplx *out, cudaStream_t st);

}  // namespace sfftb
