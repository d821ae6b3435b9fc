/*
 * oracle/ref_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin C-linkage window onto the UNMODIFIED reference library (compiled from
 * /root/reference/src by oracle/Makefile into oracle/_ref/).  It only calls the
 * reference's public API (src/sfft.h:158-167) and reads the plan structures the
 * reference's own header exposes (src/sfft.h:44-154) so that tests can compare
 * intermediate arrays (filters, bucket spectra, scores) and not only the output.
 *
 * Input synthesis follows src/simulation.cc:95-112 / src/timing_many.cc:98-121
 * with a caller-supplied drand48 seed instead of time^pid.
 */
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <omp.h>

#include "sfft.h"
#include "common.h"
#include "fftw.h"
#include "utils.h"

extern "C" {

void ref_set_threads(int t) { omp_set_num_threads(t); }
int ref_get_max_threads(void) { return omp_get_max_threads(); }

void ref_seed(unsigned s, long s48) { srand(s); srand48(s48); }

void *ref_make_plan(int n, int k, int version)
{
  return (void *)sfft_make_plan(n, k, (sfft_version)version, FFTW_ESTIMATE);
}

static void fix_globals(sfft_plan *p)
{
  /* src/common.cc:22-23 are process globals that sfft_v1 never resets
   * (src/sfft.cc:582); pin them to the plan's own mode so plans of different
   * versions can coexist in one test process. */
  ALGORITHM1 = true;
  WITH_COMB = (p->version == SFFT_VERSION_2);
}

void ref_exec(void *plan, void *in, void *out)
{
  sfft_plan *p = (sfft_plan *)plan;
  fix_globals(p);
  sfft_exec(p, (complex_t *)in, (complex_t *)out);
}

void ref_exec_many(void *plan, int num, void **in, void **out)
{
  sfft_plan *p = (sfft_plan *)plan;
  fix_globals(p);
  sfft_exec_many(p, num, (complex_t **)in, (complex_t **)out);
}

void ref_free_plan(void *plan) { sfft_free_plan((sfft_plan *)plan); }

/* ---- v1/v2 plan introspection (src/sfft.h:78-103) ---- */
int ref_v12_params(void *plan, int *out /* 12 ints */)
{
  sfft_plan *p = (sfft_plan *)plan;
  if (p->version == SFFT_VERSION_3) return -1;
  sfft_v1v2_data *d = (sfft_v1v2_data *)p->data;
  out[0] = d->B_loc;  out[1] = d->B_est;  out[2] = d->B_thresh;
  out[3] = d->W_Comb; out[4] = d->Comb_loops;
  out[5] = d->loops_loc; out[6] = d->loops_thresh; out[7] = d->loops_est;
  out[8] = d->filter.sizet; out[9] = d->filter_est.sizet;
  out[10] = (int)d->x_samp_size; out[11] = (int)d->threads;
  return 0;
}
void *ref_v12_filter_time(void *plan, int est)
{
  sfft_v1v2_data *d = (sfft_v1v2_data *)((sfft_plan *)plan)->data;
  return est ? d->filter_est.time : d->filter.time;
}
void *ref_v12_filter_freq(void *plan, int est)
{
  sfft_v1v2_data *d = (sfft_v1v2_data *)((sfft_plan *)plan)->data;
  return est ? d->filter_est.freq : d->filter.freq;
}
void *ref_v12_x_samp(void *plan)
{ return ((sfft_v1v2_data *)((sfft_plan *)plan)->data)->threadlocal_data[0].x_samp; }
void *ref_v12_x_sampt(void *plan)
{ return ((sfft_v1v2_data *)((sfft_plan *)plan)->data)->threadlocal_data[0].inner_loop_locate_x_sampt; }
int *ref_v12_score(void *plan)
{ return ((sfft_v1v2_data *)((sfft_plan *)plan)->data)->threadlocal_data[0].score; }
int *ref_v12_hits(void *plan)
{ return ((sfft_v1v2_data *)((sfft_plan *)plan)->data)->threadlocal_data[0].hits; }
int *ref_v12_permute(void *plan)
{ return ((sfft_v1v2_data *)((sfft_plan *)plan)->data)->threadlocal_data[0].permute; }
int *ref_v12_comb_approved(void *plan)
{ return ((sfft_v1v2_data *)((sfft_plan *)plan)->data)->threadlocal_data[0].Comb_Approved; }
int *ref_v12_J(void *plan)
{ return ((sfft_v1v2_data *)((sfft_plan *)plan)->data)->threadlocal_data[0].J; }

/* ---- v3 plan introspection (src/sfft.h:110-154) ---- */
int ref_v3_params(void *plan, int *out /* 8 ints */)
{
  sfft_plan *p = (sfft_plan *)plan;
  if (p->version != SFFT_VERSION_3) return -1;
  sfft_v3_data *d = (sfft_v3_data *)p->data;
  out[0] = d->B_g1; out[1] = d->w_g1; out[2] = d->B_g2; out[3] = d->w_g2;
  out[4] = d->W_Man; out[5] = d->Gauss_loops; out[6] = d->Gauss2_loops;
  out[7] = d->Man_loops;
  return 0;
}
void *ref_v3_filter(void *plan, int which /*0 t1,1 f1,2 t2,3 f2*/)
{
  sfft_v3_data *d = (sfft_v3_data *)((sfft_plan *)plan)->data;
  switch (which) {
    case 0: return d->filtert1; case 1: return d->filterf1;
    case 2: return d->filtert2; default: return d->filterf2;
  }
}
void *ref_v3_samples(void *plan, int which /*0 man,1 gauss,2 gauss_perm*/)
{
  sfft_v3_threadlocal_data *t = ((sfft_v3_data *)((sfft_plan *)plan)->data)->threadlocal_data;
  switch (which) {
    case 0: return t->man_samples; case 1: return t->gauss_samples;
    default: return t->gauss_perm_samples;
  }
}

/* ---- helpers from the reference's own utility layer ---- */
int ref_floor_to_pow2(double x) { return floor_to_pow2(x); }
int ref_mod_inverse(int a, int n) { return mod_inverse(a, n); }
void ref_fftw_dft(void *out, int n, void *in, int backwards)
{ fftw_dft((complex_t *)out, n, (complex_t *)in, backwards); }
void ref_find_largest_indices(int *output, int num, double *samples, int n, double *tmp)
{ find_largest_indices(output, num, samples, n, tmp); }
double ref_awgn(void *x, int n, double std_noise)
{ return AWGN((complex_t *)x, n, std_noise); }

/* k unit spikes at floor(drand48()*n), x = unnormalised inverse DFT
 * (src/simulation.cc:104-111); the caller seeds drand48 first. */
void ref_generate_input(int n, int k, void *x_time, void *x_freq)
{
  complex_t *xf = (complex_t *)x_freq;
  memset(xf, 0, (size_t)n * sizeof(complex_t));
  for (int i = 0; i < k; i++) {
    unsigned f = (unsigned)floor(drand48() * n);
    xf[f] = 1.0;
  }
  fftw_dft((complex_t *)x_time, n, xf, 1);
}

} /* extern "C" */
