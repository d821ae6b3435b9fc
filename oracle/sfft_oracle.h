/*
 * oracle/sfft_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99) of the reference sparse-FFT transform path:
 * plan derivation, window construction, and the v1/v2/v3 transforms, with every
 * intermediate array kept so the CUDA engine can be compared stage by stage.
 * Each function cites the reference file:line it restates.  The DFTs, which the
 * reference delegates to FFTW 3 (absent, unpinned), are pinned to oracle/fft_ref.c.
 *
 * PINNING: this restatement is checked bit-for-bit (filters, bucket spectra,
 * selected buckets, scores, outputs) against oracle/_ref/libsfft_ref_parity.so,
 * i.e. the reference's own sources compiled over the same DFT, by
 * tests/test_oracle_vs_ref.py, and against tests/golden/ fixtures generated from
 * that build by tests/golden/make_golden.py.  The reference ships no golden
 * vectors of its own.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use
 * anything under oracle/.
 */
#ifndef SFFT_ORACLE_H
#define SFFT_ORACLE_H

#include "fft_ref.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int version;                 /* 1, 2 or 3 */
  int n_requested, n, k;       /* n = floor_to_pow2(n_requested) */

  /* ---- v1 / v2 (reference struct sfft_v1v2_data, src/sfft.h:78-103) ---- */
  int with_comb;
  int B_loc, B_est, B_thresh, W_Comb, Comb_loops;
  int loops_loc, loops_thresh, loops_est;
  int w_loc, w_est, b_loc, b_est;
  double tolerance_loc, tolerance_est, lobefrac_loc, lobefrac_est;
  ocplx *time_loc, *freq_loc;  /* w_loc taps, n-point response */
  ocplx *time_est, *freq_est;
  long x_samp_size;

  /* per-exec state, kept for inspection */
  int *a, *ai;                 /* loops entries */
  ocplx *x_sampt, *x_samp;     /* x_samp_size each: folded samples, bucket spectra */
  double *mag;                 /* x_samp_size */
  int *J;                      /* loops * B_thresh: selected buckets, every loop */
  int *score;                  /* n */
  int *hits; long hits_found;  /* n */
  long hits_prefill;           /* v2: number of comb pre-filled entries */
  int *comb_approved; int num_comb; int *comb_offsets;
  ocplx *comb_spec;            /* Comb_loops * W_Comb */

  /* ---- v3 (reference struct sfft_v3_data, src/sfft.h:121-154) ---- */
  int B_g1, w_g1, B_g2, w_g2, W_Man;
  ocplx *filtert1, *filterf1, *filtert2, *filterf2;
  ocplx *man_samp, *gauss_samp, *gauss_perm_samp, *perm_x;
  int v3_a, v3_ai, v3_b, v3_shift, v3_init_offset, v3_init_G_offset;
  int *v3_keys; ocplx *v3_vals; int v3_count, v3_cap;   /* insertion-ordered result */
  int v3_rounds;

  ocplx *tw; long tw_n;        /* shared twiddle table */
} orc_plan;

/* src/utils.cc:243-248 */
int orc_floor_to_pow2(double x);
/* src/utils.cc:41-46, :85-100 */
int orc_gcd(int a, int b);
int orc_mod_inverse(int a, int n);

/* src/filters.cc:70-86 -- returns malloc'd w taps (real parts only are non-zero) */
ocplx *orc_dolph_chebyshev(double lobefrac, double tolerance, int *w_out);
/* frequency-domain samples before the w-point DFT (src/filters.cc:77-80) */
void orc_dolph_chebyshev_samples(double lobefrac, double tolerance, int w, double *out);
/* src/filters.cc:109-160 -- rewrites taps[0..w) in place, returns malloc'd n-point response */
ocplx *orc_make_multiple(ocplx *taps, int w, int n, int b);

/* src/utils.cc:131-158 */
void orc_find_largest_indices(int *out, int num, const double *samples, int n);

/* src/sfft.cc:71-101, :298-392, :506-579.  Returns NULL where the reference
 * returns NULL or would trip an assert that does not depend on NDEBUG-able
 * state. */
orc_plan *orc_make_plan(int n, int k, int version);
void orc_free_plan(orc_plan *p);

/* src/sfft.cc:119-137: out[0..n) dense.  Consumes libc random()/drand48()
 * exactly as the reference does (src/computefourier-1.0-2.0.cc:465-474,:62;
 * src/computefourier-3.0.cc:800-810). */
void orc_exec(orc_plan *p, const ocplx *x, ocplx *out);

/* stage entry points (v1/v2), usable one at a time by tests */
void orc_draw_permutations(orc_plan *p);                       /* cf12.cc:465-474 */
void orc_comb_stage(orc_plan *p, const ocplx *x);              /* cf12.cc:49-82,:483-512 */
void orc_bucketize(orc_plan *p, const ocplx *x);               /* cf12.cc:213-261 */
void orc_bucket_ffts(orc_plan *p);                             /* cf12.cc:270-289 */
void orc_select_and_vote(orc_plan *p);                         /* cf12.cc:292-324,:92-184 */
void orc_estimate(orc_plan *p, ocplx *out);                    /* cf12.cc:341-419 */

/* simulation input (src/simulation.cc:104-111): caller seeds drand48 */
void orc_generate_input(int n, int k, ocplx *x_time, ocplx *x_freq);
/* src/utils.cc:250-280 */
double orc_awgn(ocplx *x, int n, double std_noise);

#ifdef __cplusplus
}
#endif
#endif
