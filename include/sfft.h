/*
 * sfft.h -- C ABI of the B200-native sparse-FFT engine (libsfft.so).
 *
 * Part 1 is the drop-in boundary: the six entry points and the plan/enum types of
 * the reference library's public header (reference src/sfft.h:37-50 and :158-167),
 * in plain C (the reference header is C++-only because of <tr1/unordered_map>,
 * src/sfft.h:26; that typedef is not part of the call surface and is dropped).
 *
 * Part 2 is the device-resident extension this engine adds beside them: the
 * signal stays in HBM, the result comes back as a sparse (location, value) list
 * in HBM.  The six legacy symbols are thin wrappers over it (H2D copy -> device
 * transform -> densify -> D2H copy).
 *
 * There is no CPU implementation behind any of these: every entry point that
 * computes needs a CUDA device (sm_100a) and fails loudly (NULL / negative
 * return, message via sfftb_last_error()) without one.
 */
#ifndef SFFT_B200_SFFT_H
#define SFFT_B200_SFFT_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------ */
/* Part 1 -- drop-in boundary (reference src/sfft.h)                         */
/* ------------------------------------------------------------------------ */

/* 16-byte interleaved (re, im) double pair; layout-identical to the reference's
 * `typedef double complex complex_t` (src/fft.h:43). */
#if defined(__cplusplus) || defined(SFFT_NO_C99_COMPLEX)
typedef struct sfft_complex { double re, im; } sfft_complex;
#else
typedef double _Complex sfft_complex;
#endif

/* reference src/sfft.h:37-42 (the Python binding passes version-1, python/sfft/sfft.py:74) */
typedef enum sfft_version {
  SFFT_VERSION_1 = 0,
  SFFT_VERSION_2 = 1,
  SFFT_VERSION_3 = 2
} sfft_version;

/* reference src/sfft.h:44-50.  Callers only ever hold the pointer. */
typedef struct sfft_plan {
  sfft_version version;
  unsigned int n;
  unsigned int k;
  void *data;
} sfft_plan;

/* fftw_optimization values the reference accepts (fftw3.h via python/sfft/sfft.py:7-8);
 * accepted and ignored here: there is no FFTW behind this library. */
#define SFFT_FFTW_MEASURE 0
#define SFFT_FFTW_ESTIMATE 64
/* Opt-in bit for the fftw_optimization word of sfft_make_plan (FFTW's planner flags end at
 * bit 21): look the tuned parameters of a k > 50 plan up BY k, as src/parameters.cc:282-513
 * was written for (tuned at n = 2^22).  The reference passes n as the key
 * (src/sfft.cc:316-321), so the lookup never matches and such plans run on the defaults;
 * without this bit this library does exactly the same.  Parity claims hold without it. */
#define SFFTB_PLAN_TUNED_BY_K (1 << 24)

/* reference src/sfft.cc:60-69 (_mm_malloc(s,16)).  Here: page-locked host memory
 * (cudaHostAlloc) so the H2D/D2H legs of sfft_exec run at PCIe speed; falls back
 * to 64-byte aligned pageable memory when no CUDA device is present. */
void *sfft_malloc(size_t s);
void sfft_free(void *p);

/* reference src/sfft.cc:71-101.  NULL on unknown version, allocation failure,
 * no CUDA device, or an (n,k) the reference would reject by assert at plan time
 * (src/utils.cc:134 via src/computefourier-1.0-2.0.cc:78,301; src/filters.cc:111-112). */
sfft_plan *sfft_make_plan(int n, int k, sfft_version version, int fftw_optimization);

/* reference src/sfft.cc:103-117.  Frees everything (the reference leaks the plan
 * struct and the filters). */
void sfft_free_plan(sfft_plan *plan);

/* reference src/sfft.cc:119-137.  `in` and `out` are HOST arrays of plan->n
 * elements; `out` is fully overwritten (zeros + recovered coefficients).
 * Consumes libc random()/drand48() on the calling thread in the reference's
 * order (src/computefourier-1.0-2.0.cc:465-474,:62; src/computefourier-3.0.cc:800-810). */
void sfft_exec(sfft_plan *plan, sfft_complex *in, sfft_complex *out);

/* reference src/sfft.cc:139-147.  Permutations are drawn on the calling thread in
 * signal order, so results do not depend on a thread schedule. */
void sfft_exec_many(sfft_plan *plan, int num, sfft_complex **in, sfft_complex **out);

/* ------------------------------------------------------------------------ */
/* Part 2 -- device-resident extension                                       */
/* ------------------------------------------------------------------------ */

#define SFFTB_MAX_LOOPS 64
#define SFFTB_MAX_COMB_LOOPS 16

/* Derived plan parameters (reference struct sfft_v1v2_data / sfft_v3_data,
 * src/sfft.h:78-103,121-154; derivation src/sfft.cc:298-353,506-560). */
typedef struct sfftb_info {
  int version;            /* 1, 2, 3 */
  int n;                  /* floor_to_pow2 of the requested n */
  int k;
  int device;             /* CUDA ordinal the plan lives on */
  /* v1/v2 */
  int B_loc, B_est, B_thresh, W_Comb, Comb_loops;
  int loops_loc, loops_thresh, loops_est;
  int w_loc, w_est, b_loc, b_est;
  long long x_samp_size;
  /* v3 */
  int B_g1, w_g1, B_g2, w_g2, W_Man;
  /* capacity of the sparse result list */
  long long max_hits;
  /* bytes of signal the windowed gathers read per transform (16 B per sample) and
   * bytes of window taps (16 B per tap per distinct filter) */
  long long gather_samples;
  long long gather_tap_bytes;
} sfftb_info;

/* The random draw of one transform, in the reference's order
 * (src/computefourier-1.0-2.0.cc:465-474, :62; src/computefourier-3.0.cc:800-810). */
typedef struct sfftb_draw {
  int loops;
  int a[SFFTB_MAX_LOOPS];
  int ai[SFFTB_MAX_LOOPS];
  int comb_offset[SFFTB_MAX_COMB_LOOPS];
  int v3_a, v3_ai, v3_b, v3_init_offset, v3_init_G_offset;
} sfftb_draw;

/* Sparse result living in device memory owned by the plan; valid until the next
 * transform on the same plan.  Order of entries is unspecified. */
typedef struct sfftb_result {
  const int *d_loc;             /* [count] frequency indices */
  const sfft_complex *d_val;    /* [count] coefficients */
  const int *d_count;           /* device scalar */
  long long count;              /* host copy (filled when `sync` was requested) */
} sfftb_result;

const char *sfftb_last_error(void);
int sfftb_device_count(void);

int sfftb_plan_info(const sfft_plan *plan, sfftb_info *info);

/* Use the caller's CUDA stream (cudaStream_t passed as void*) for all work of
 * this plan; NULL selects the plan's own stream. */
int sfftb_set_stream(sfft_plan *plan, void *cuda_stream);

/* Fill `draw` from libc random()/drand48() exactly as the reference would for one
 * transform of this plan. */
int sfftb_draw_random(const sfft_plan *plan, sfftb_draw *draw);

/* One transform; d_in is a DEVICE array of n complex doubles.  draw==NULL draws
 * from libc.  sync!=0 waits and fills result->count. */
int sfftb_exec_device(sfft_plan *plan, const void *d_in, const sfftb_draw *draw,
                      sfftb_result *result, int sync);

/* Batch of `num` independent signals, signal i at d_in + i*stride_elems complex
 * elements (device memory).  draws==NULL draws from libc in signal order.
 * Results: counts[i] entries at (d_loc + i*max_hits, d_val + i*max_hits). */
int sfftb_exec_many_device(sfft_plan *plan, int num, const void *d_in,
                           long long stride_elems, const sfftb_draw *draws,
                           sfftb_result *result, long long *counts, int sync);

/* Zero d_out[0..n) and scatter the sparse result of signal `which` into it (device). */
int sfftb_densify(sfft_plan *plan, int which, void *d_out);

/* Wait until everything queued on the plan's stream has finished. */
int sfftb_synchronize(sfft_plan *plan);

/* Copy the sparse result of signal `which` to host arrays (capacity entries). */
long long sfftb_fetch_result(sfft_plan *plan, int which, int *loc, sfft_complex *val,
                             long long capacity);

/* ---- multi-GPU sharding of ONE v1/v2 transform (one process per GPU) --------
 * (reference: the loops of outer_loop, src/computefourier-1.0-2.0.cc:438-541, are
 * independent until voting and again until the median)
 * The signal is resident on every GPU.  Rank r of `world` bucketises only its own
 * block of loops (gather + bucket FFT, the bandwidth-heavy part); one exchange then
 * makes the bucket spectra complete everywhere.  Selection and voting are replicated
 * (cheap and deterministic); estimation covers every hit (v1) or this rank's slice of
 * the pre-filled list (v2, where that list is the bulk of the work; the result stays
 * distributed: rank r holds entries [offset, offset+count) of the single-GPU list,
 * see sfftb_shard_slice).
 * The draw must be the same on every rank: pass the same sfftb_draw, or pass NULL and
 * seed libc (srand / srand48) identically on every rank.
 * v3 has no loop structure to shard: replicas only.
 *
 * (1) NVLink peer exchange -- the fast path.  Buffers are mapped across processes with
 * CUDA IPC; ranks store their rows into their peers' buffers over NVLink/NVSwitch and
 * signal with flags in peer memory.  No NCCL call, no host synchronisation; the whole
 * transform replays from one CUDA graph per rank.
 *
 *   sfftb_shard_export(plan, &mine);            // every rank
 *   <all-gather the sfftb_peer_handle structs (plain bytes) by any means>
 *   sfftb_shard_attach(plan, rank, world, all); // then a barrier between the ranks
 *   sfftb_shard_exec(plan, d_in, draw, &result, sync);   // any number of times
 *   <barrier>  sfftb_shard_detach(plan);
 *
 * A peer that never arrives makes the waiting kernels give up after 4 s (counted in
 * sfftb_shard_status) instead of hanging the GPU. */
#define SFFTB_IPC_HANDLE_BYTES 64
typedef struct sfftb_peer_handle {
  unsigned char spectra[SFFTB_IPC_HANDLE_BYTES];   /* cudaIpcMemHandle_t of the bucket-spectra buffer */
  unsigned char flags[SFFTB_IPC_HANDLE_BYTES];     /* cudaIpcMemHandle_t of the flag block */
} sfftb_peer_handle;
int sfftb_shard_export(sfft_plan *plan, sfftb_peer_handle *mine);
int sfftb_shard_attach(sfft_plan *plan, int rank, int world, const sfftb_peer_handle *all);
int sfftb_shard_detach(sfft_plan *plan);
int sfftb_shard_exec(sfft_plan *plan, const void *d_in, const sfftb_draw *draw, sfftb_result *result,
                     int sync);
/* transforms completed since attach, and flag waits that timed out (must be 0) */
int sfftb_shard_status(sfft_plan *plan, long long *epoch, long long *timeouts);
/* the part of the single-GPU result list that `rank` of `world` produced in the last
 * sharded transform: entries [offset, offset + count) of that list, in its order
 * (v1: the whole list on every rank) */
int sfftb_shard_slice(sfft_plan *plan, int rank, int world, long long *offset, long long *count);

/* (2) caller-side collective -- the portable fallback.  The caller sums the buffer
 * returned by sfftb_shard_spectra() over ranks (rows a rank does not own are zero, so
 * the sum is exact), e.g. torch.distributed.all_reduce over NCCL:
 *
 *   sfftb_shard_bucketize(plan, d_in, draw, rank, world);
 *   sfftb_shard_spectra(plan, &ptr, &count);  all_reduce(ptr, count doubles, SUM);
 *   sfftb_shard_finish(plan, rank, world, &result, sync);
 *
 * All three run on the plan's stream (sfftb_set_stream): the collective must be ordered
 * after bucketize and before finish on that stream.  The pointer returned by
 * sfftb_shard_spectra is invalidated when a later batch call grows the plan's scratch. */
int sfftb_shard_bucketize(sfft_plan *plan, const void *d_in, const sfftb_draw *draw, int rank,
                          int world);
int sfftb_shard_spectra(sfft_plan *plan, void **d_spectra, long long *n_doubles);
int sfftb_shard_finish(sfft_plan *plan, int rank, int world, sfftb_result *result, int sync);
/* loops [begin, end) of the plan's `loops` that `rank` owns (block partition) */
int sfftb_shard_loops(const sfft_plan *plan, int rank, int world, int *begin, int *end);

/* ---- plan cache ----------------------------------------------
 * A plan is (n, k, version, flags) plus its two filters; everything else is derived.
 * sfftb_save_plan writes them to a file; sfftb_load_plan re-creates the plan from it
 * without running the filter builder (seconds at n >= 2^26; the reference needs minutes,
 * src/filters.cc:70-160).  The loaded plan is bit-identical to the saved one.
 * Returns 0 / a plan, or -1 / NULL with sfftb_last_error() set. */
int sfftb_save_plan(const sfft_plan *plan, const char *path);
sfft_plan *sfftb_load_plan(const char *path);

/* ---- plan-builder hooks (parity injection / plan cache) ------------------ */
/* which: 0 = location filter, 1 = estimation filter (v1/v2); 0/1 = first/second
 * Gaussian-style filter (v3).  time has w taps; freq_window has fw_len entries
 * centred on frequency 0: freq_window[m] = freq[(m - fw_len/2) mod n]. */
int sfftb_filter_sizes(const sfft_plan *plan, int which, int *w, int *fw_len);
int sfftb_get_filter(const sfft_plan *plan, int which, sfft_complex *time,
                     sfft_complex *freq_window);
int sfftb_set_filter(sfft_plan *plan, int which, const sfft_complex *time,
                     const sfft_complex *freq_window);

/* ---- stage hooks for parity tests (copy a device scratch array to host) ---
 * what: "x_sampt" (folded samples, bit-reversed bucket order), "x_samp" (bucket
 * spectra), "J" (selected buckets, loops_loc x num), "hits", "vals",
 * "comb_approved", "voted", "twiddle", "perm_a", "perm_ai".
 * Returns the number of bytes copied, or -1. */
long long sfftb_debug_fetch(sfft_plan *plan, const char *what, void *dst, size_t capacity);

/* standalone kernels on host data (staged through the device) */
int sfftb_debug_fft(const sfft_complex *in, sfft_complex *out, int log2n, int batch, int sign,
                    int table_twiddles);
int sfftb_debug_select(const double *mags, int B, int num, int batch, int *out_J);
int sfftb_debug_dft_any(const sfft_complex *in, sfft_complex *out, int n);
/* number of (a,b) pairs, out of `count` pseudo-random ones, for which the engine's
 * reciprocal-based exact division differs from IEEE division (must be 0) */
long long sfftb_debug_div_check(unsigned long long seed, long long count);

/* per-stage device timings (ms) of the last transform when timing was enabled */
int sfftb_enable_stage_timing(sfft_plan *plan, int on);
int sfftb_stage_times(sfft_plan *plan, float *ms, const char **names, int capacity);

/* kernels launched by this library since load (for bench.py's gpu_launches) */
long long sfftb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
