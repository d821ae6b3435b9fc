/*
 * oracle/sfft_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See sfft_oracle.h.
 *
 * Plain-C restatement of the reference's sparse-FFT path.  Build with
 * -ffp-contract=off and without -ffast-math (oracle/Makefile): every product and
 * sum below is one IEEE-754 rounding, which is what the reference's SSE2
 * intrinsics do (src/computefourier-1.0-2.0.cc:243-256) and what the CUDA engine
 * reproduces with __dmul_rn/__dadd_rn.
 */
#include "sfft_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* small integer helpers                                                      */
/* ------------------------------------------------------------------------- */

/* src/utils.cc:243-248: largest power of two <= x */
int orc_floor_to_pow2(double x)
{
  unsigned int p = 1;
  while (p <= x) p <<= 1;
  return (int)(p / 2);
}

/* src/utils.cc:41-46 */
int orc_gcd(int a, int b)
{
  while (a % b != 0) { int r = a % b; a = b; b = r; }
  return b;
}

/* src/utils.cc:85-100: extended Euclid, result in [0, n) */
int orc_mod_inverse(int a, int n)
{
  int rem = n, prev = 0, cur = 1;
  while (a > 0) {
    int q = rem / a, old_a = a;
    a = rem % old_a;
    rem = old_a;
    int next = prev - q * cur;
    prev = cur;
    cur = next;
  }
  prev %= n;
  if (prev < 0) prev = (prev + n) % n;
  return prev;
}

static inline int mulmod(int x, int a, int n)
{
  /* src/computefourier-1.0-2.0.cc:43-46 */
  return (int)(((long long)x * a) % n);
}

/* ------------------------------------------------------------------------- */
/* windows                                                                    */
/* ------------------------------------------------------------------------- */

/* src/filters.cc:62-68: Chebyshev polynomial of (real) degree m */
static double cheb_poly(double m, double x)
{
  if (fabs(x) <= 1) return cos(m * acos(x));
  return creal(ccosh(m * cacosh(x)));
}

static int dolph_width(double lobefrac, double tolerance)
{
  /* src/filters.cc:72-74 */
  int w = (int)((1 / M_PI) * (1 / lobefrac) * acosh(1. / tolerance));
  if (!(w % 2)) w--;
  return w;
}

void orc_dolph_chebyshev_samples(double lobefrac, double tolerance, int w, double *out)
{
  (void)lobefrac;
  /* src/filters.cc:76-80 */
  double t0 = cosh(acosh(1 / tolerance) / (w - 1));
  for (int i = 0; i < w; i++)
    out[i] = cheb_poly(w - 1, t0 * cos(M_PI * i / w)) * tolerance;
}

ocplx *orc_dolph_chebyshev(double lobefrac, double tolerance, int *w_out)
{
  int w = dolph_width(lobefrac, tolerance);
  *w_out = w;
  double *s = (double *)malloc((size_t)w * sizeof(double));
  orc_dolph_chebyshev_samples(lobefrac, tolerance, w, s);
  ocplx *x = (ocplx *)malloc((size_t)w * sizeof(ocplx));
  for (int i = 0; i < w; i++) { x[i].re = s[i]; x[i].im = 0.0; }
  free(s);
  /* src/filters.cc:81: w-point forward DFT */
  orc_fft_any(x, w, -1);
  /* src/filters.cc:82-84 with src/utils.cc:28-38: rotate right by w/2, keep real part */
  ocplx *y = (ocplx *)malloc((size_t)w * sizeof(ocplx));
  int r = w / 2;
  for (int i = 0; i < w; i++) {
    y[(i + r) % w].re = x[i].re;
    y[(i + r) % w].im = 0.0;
  }
  free(x);
  return y;
}

ocplx *orc_make_multiple(ocplx *taps, int w, int n, int b)
{
  /* src/filters.cc:109-160 */
  ocplx *g = (ocplx *)calloc((size_t)n, sizeof(ocplx));
  ocplx *h = (ocplx *)malloc((size_t)n * sizeof(ocplx));
  ocplx *tw = orc_twiddle_table(n);
  /* :113-114 centre the window on index 0 */
  memcpy(g, taps + w / 2, (size_t)(w - w / 2) * sizeof(ocplx));
  memcpy(g + n - w / 2, taps, (size_t)(w / 2) * sizeof(ocplx));
  orc_fft_pow2(g, n, -1, tw, n);                       /* :115 */
  /* :116-130 boxcar of width b by running sum; track the peak magnitude */
  double sr = 0, si = 0;
  for (int i = 0; i < b; i++) { sr += g[i].re; si += g[i].im; }
  double peak = 0;
  int off = b / 2;
  for (int i = 0; i < n; i++) {
    ocplx *dst = &h[(i + n + off) % n];
    dst->re = sr; dst->im = si;
    double m = cabs(CMPLX(sr, si));
    if (m > peak) peak = m;
    const ocplx *in = &g[(i + b) % n], *outg = &g[i];
    double dr = in->re - outg->re, di = in->im - outg->im;
    sr = sr + dr; si = si + di;
  }
  for (int i = 0; i < n; i++) { h[i].re /= peak; h[i].im /= peak; }   /* :131-132 */
  /* :134-140 phase ramp by running product */
  double complex step = cexp(-2 * M_PI * I * (w / 2) / n);
  double cr = 1, ci = 0, stepr = creal(step), stepi = cimag(step);
  for (int i = 0; i < n; i++) {
    double hr = h[i].re, hi = h[i].im;
    h[i].re = hr * cr - hi * ci;
    h[i].im = hr * ci + hi * cr;
    double nr = cr * stepr - ci * stepi;
    double ni = cr * stepi + ci * stepr;
    cr = nr; ci = ni;
  }
  /* :141-142 back to time, keep the first w samples; :153-154 scale by 1/n */
  memcpy(g, h, (size_t)n * sizeof(ocplx));
  orc_fft_pow2(g, n, +1, tw, n);
  for (int i = 0; i < w; i++) { taps[i].re = g[i].re / n; taps[i].im = g[i].im / n; }
  free(g); free(tw);
  return h;
}

/* ------------------------------------------------------------------------- */
/* top-num selection                                                          */
/* ------------------------------------------------------------------------- */

static int cmp_double(const void *a, const void *b)
{
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}
static int cmp_int(const void *a, const void *b)
{
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* src/utils.cc:131-158: indices of the num largest, ascending; the cutoff is the
 * (num+1)-th largest value; ties at the cutoff are admitted in index order. */
void orc_find_largest_indices(int *out, int num, const double *samples, int n)
{
  double *tmp = (double *)malloc((size_t)n * sizeof(double));
  memcpy(tmp, samples, (size_t)n * sizeof(double));
  qsort(tmp, (size_t)n, sizeof(double), cmp_double);
  double cutoff = tmp[n - num - 1];
  free(tmp);
  int count = 0;
  for (int i = 0; i < n; i++)
    if (samples[i] > cutoff) out[count++] = i;
  if (count < num) {
    for (int i = 0; i < n && count < num; i++)
      if (samples[i] == cutoff) out[count++] = i;
    qsort(out, (size_t)count, sizeof(int), cmp_int);
  }
}

/* ------------------------------------------------------------------------- */
/* plan                                                                       */
/* ------------------------------------------------------------------------- */

typedef struct {
  int by_k, with_comb, key;
  double Bcst_loc, Bcst_est, Comb_cst;
  int loc_loops, est_loops, threshold_loops, comb_loops;
  double tolerance_loc, tolerance_est;
} param_row;

static const param_row PARAM_ROWS[] = {
#include "param_table.inc"
};

static void lookup_params(int by_k, int with_comb, int key, param_row *r)
{
  /* src/parameters.cc:24-312 (by N), :314-513 (by K); no match keeps defaults */
  size_t cnt = sizeof(PARAM_ROWS) / sizeof(PARAM_ROWS[0]);
  for (size_t i = 0; i < cnt; i++) {
    const param_row *t = &PARAM_ROWS[i];
    if (t->by_k == by_k && t->with_comb == with_comb && t->key == key) {
      r->Bcst_loc = t->Bcst_loc; r->Bcst_est = t->Bcst_est; r->Comb_cst = t->Comb_cst;
      r->loc_loops = t->loc_loops; r->est_loops = t->est_loops;
      r->threshold_loops = t->threshold_loops; r->comb_loops = t->comb_loops;
      r->tolerance_loc = t->tolerance_loc; r->tolerance_est = t->tolerance_est;
      return;
    }
  }
}

static orc_plan *plan_v12(int n_req, int k, int with_comb)
{
  /* src/sfft.cc:298-392 */
  param_row pr;
  pr.Bcst_loc = 1; pr.Bcst_est = 1; pr.Comb_cst = 2;
  pr.loc_loops = 4; pr.est_loops = 16; pr.threshold_loops = 3; pr.comb_loops = 1;
  pr.tolerance_loc = 1.e-8; pr.tolerance_est = 1.e-8;
  /* :316-327 -- for k > 50 the by-K table is consulted with n as the key */
  if ((unsigned)k > 50) lookup_params(1, with_comb, n_req, &pr);
  else lookup_params(0, with_comb, n_req, &pr);

  unsigned n = (unsigned)orc_floor_to_pow2(n_req);
  if (n < 2) return NULL;

  double BB_loc = (unsigned)(pr.Bcst_loc * sqrt((double)(int)n * (unsigned)k / (log2(n))));
  double BB_est = (unsigned)(pr.Bcst_est * sqrt((double)(int)n * (unsigned)k / (log2(n))));
  if (BB_loc < 1 || BB_est < 1) return NULL;

  orc_plan *p = (orc_plan *)calloc(1, sizeof(orc_plan));
  p->version = with_comb ? 2 : 1;
  p->with_comb = with_comb;
  p->n_requested = n_req; p->n = (int)n; p->k = k;
  p->lobefrac_loc = 0.5 / BB_loc;
  p->lobefrac_est = 0.5 / BB_est;
  p->b_loc = (int)(1.2 * 1.1 * ((double)n / BB_loc));
  p->b_est = (int)(1.4 * 1.1 * ((double)n / BB_est));
  p->B_loc = orc_floor_to_pow2(BB_loc);
  p->B_thresh = 2 * k;
  p->B_est = orc_floor_to_pow2(BB_est);
  p->W_Comb = orc_floor_to_pow2(pr.Comb_cst * n / p->B_loc);
  p->Comb_loops = pr.comb_loops;
  p->loops_loc = pr.loc_loops;
  p->loops_thresh = pr.threshold_loops;
  p->loops_est = pr.est_loops;
  p->tolerance_loc = pr.tolerance_loc;
  p->tolerance_est = pr.tolerance_est;

  p->time_loc = orc_dolph_chebyshev(p->lobefrac_loc, p->tolerance_loc, &p->w_loc);
  if (p->w_loc > (int)n || p->b_loc > (int)n) { free(p->time_loc); free(p); return NULL; }
  p->freq_loc = orc_make_multiple(p->time_loc, p->w_loc, (int)n, p->b_loc);
  p->time_est = orc_dolph_chebyshev(p->lobefrac_est, p->tolerance_est, &p->w_est);
  if (p->w_est > (int)n || p->b_est > (int)n) { free(p->time_est); free(p); return NULL; }
  p->freq_est = orc_make_multiple(p->time_est, p->w_est, (int)n, p->b_est);

  int loops = p->loops_loc + p->loops_est;
  p->x_samp_size = (long)p->loops_loc * p->B_loc + (long)p->loops_est * p->B_est;
  p->a = (int *)calloc((size_t)loops, sizeof(int));
  p->ai = (int *)calloc((size_t)loops, sizeof(int));
  p->x_sampt = (ocplx *)calloc((size_t)p->x_samp_size, sizeof(ocplx));
  p->x_samp = (ocplx *)calloc((size_t)p->x_samp_size, sizeof(ocplx));
  p->mag = (double *)calloc((size_t)p->x_samp_size, sizeof(double));
  p->J = (int *)calloc((size_t)loops * p->B_thresh, sizeof(int));
  p->score = (int *)calloc((size_t)n, sizeof(int));
  p->hits = (int *)calloc((size_t)n + 1, sizeof(int));
  p->comb_approved = (int *)calloc((size_t)p->Comb_loops * p->B_thresh + 1, sizeof(int));
  p->comb_offsets = (int *)calloc((size_t)p->Comb_loops + 1, sizeof(int));
  p->comb_spec = (ocplx *)calloc((size_t)p->Comb_loops * p->W_Comb + 1, sizeof(ocplx));

  long tw_n = p->B_loc > p->B_est ? p->B_loc : p->B_est;
  if (with_comb && p->W_Comb > tw_n) tw_n = p->W_Comb;
  p->tw_n = tw_n;
  p->tw = orc_twiddle_table(tw_n);
  return p;
}

static orc_plan *plan_v3(int n_req, int k);   /* below */

orc_plan *orc_make_plan(int n, int k, int version)
{
  switch (version) {
    case 1: return plan_v12(n, k, 0);
    case 2: return plan_v12(n, k, 1);
    case 3: return plan_v3(n, k);
    default: return NULL;          /* src/sfft.cc:87-88 */
  }
}

void orc_free_plan(orc_plan *p)
{
  if (!p) return;
  free(p->time_loc); free(p->freq_loc); free(p->time_est); free(p->freq_est);
  free(p->a); free(p->ai); free(p->x_sampt); free(p->x_samp); free(p->mag);
  free(p->J); free(p->score); free(p->hits); free(p->comb_approved);
  free(p->comb_offsets); free(p->comb_spec);
  free(p->filtert1); free(p->filterf1); free(p->filtert2); free(p->filterf2);
  free(p->man_samp); free(p->gauss_samp); free(p->gauss_perm_samp); free(p->perm_x);
  free(p->v3_keys); free(p->v3_vals);
  free(p->tw);
  free(p);
}

/* ------------------------------------------------------------------------- */
/* v1 / v2 transform                                                          */
/* ------------------------------------------------------------------------- */

static long loop_offset(const orc_plan *p, int j)
{
  /* src/computefourier-1.0-2.0.cc:228-230 */
  int lo = j < p->loops_loc ? j : p->loops_loc;
  int hi = j > p->loops_loc ? j - p->loops_loc : 0;
  return (long)lo * p->B_loc + (long)hi * p->B_est;
}

void orc_draw_permutations(orc_plan *p)
{
  /* src/computefourier-1.0-2.0.cc:465-474: a odd by rejection, b = 0 */
  int loops = p->loops_loc + p->loops_est;
  for (int i = 0; i < loops; i++) {
    int a = 0;
    while (orc_gcd(a, p->n) != 1) a = (int)(random() % p->n);
    p->a[i] = a;
    p->ai[i] = orc_mod_inverse(a, p->n);
  }
}

void orc_comb_stage(orc_plan *p, const ocplx *x)
{
  p->num_comb = p->B_thresh;                       /* cf12.cc:480 */
  p->hits_found = 0;
  p->hits_prefill = 0;
  if (!p->with_comb) return;
  const int n = p->n, W = p->W_Comb, num = p->B_thresh;
  double *mag = (double *)malloc((size_t)W * sizeof(double));
  for (int c = 0; c < p->Comb_loops; c++) {
    /* src/computefourier-1.0-2.0.cc:61-79 */
    int sigma = n / W;
    int offset = (int)(unsigned)floor(drand48() * sigma);
    p->comb_offsets[c] = offset;
    ocplx *s = p->comb_spec + (long)c * W;
    for (int i = 0; i < W; i++) s[i] = x[offset + i * sigma];
    orc_fft_pow2(s, W, -1, p->tw, p->tw_n);
    for (int i = 0; i < W; i++) mag[i] = s[i].re * s[i].re + s[i].im * s[i].im;
    orc_find_largest_indices(p->comb_approved + (long)c * num, num, mag, W);
  }
  free(mag);
  if (p->Comb_loops > 1) {
    /* cf12.cc:492-502: sort, then drop repeats */
    int total = p->Comb_loops * num;
    qsort(p->comb_approved, (size_t)total, sizeof(int), cmp_int);
    int last = 0;
    for (int i = 1; i < total; i++)
      if (p->comb_approved[i] != p->comb_approved[last])
        p->comb_approved[++last] = p->comb_approved[i];
    p->num_comb = last + 1;
  }
  /* cf12.cc:505-512: every index congruent to an approved residue is pre-listed */
  long cnt = 0;
  for (int j = 0; j < n / W; j++)
    for (int i = 0; i < p->num_comb; i++)
      p->hits[cnt++] = j * W + p->comb_approved[i];
  p->hits_found = cnt;
  p->hits_prefill = cnt;
}

void orc_bucketize(orc_plan *p, const ocplx *x)
{
  /* src/computefourier-1.0-2.0.cc:213-261 */
  const int n = p->n, loops = p->loops_loc + p->loops_est;
  memset(p->x_sampt, 0, (size_t)p->x_samp_size * sizeof(ocplx));
  for (int j = 0; j < loops; j++) {
    const int is_loc = j < p->loops_loc;
    const ocplx *taps = is_loc ? p->time_loc : p->time_est;
    const int w = is_loc ? p->w_loc : p->w_est;
    const int B = is_loc ? p->B_loc : p->B_est;
    ocplx *dst = p->x_sampt + loop_offset(p, j);
    unsigned idx = 0;                                /* b = 0 */
    const unsigned ai = (unsigned)p->ai[j];
    for (int i = 0; i < w; i++) {
      const double xr = x[idx].re, xi = x[idx].im;
      const double fr = taps[i].re, fi = taps[i].im;
      const double p0 = xr * fr, p1 = xi * fi, p2 = xr * fi, p3 = xi * fr;
      const double pr = p0 - p1, pi = p2 + p3;
      ocplx *b = &dst[i & (B - 1)];
      b->re = b->re + pr;
      b->im = b->im + pi;
      idx = (idx + ai) & (unsigned)(n - 1);
    }
  }
}

void orc_bucket_ffts(orc_plan *p)
{
  /* src/computefourier-1.0-2.0.cc:270-289 */
  const int loops = p->loops_loc + p->loops_est;
  memcpy(p->x_samp, p->x_sampt, (size_t)p->x_samp_size * sizeof(ocplx));
  for (int j = 0; j < loops; j++) {
    const int B = j < p->loops_loc ? p->B_loc : p->B_est;
    orc_fft_pow2(p->x_samp + loop_offset(p, j), B, -1, p->tw, p->tw_n);
  }
  for (long i = 0; i < p->x_samp_size; i++) {
    const double r = p->x_samp[i].re, m = p->x_samp[i].im;
    const double rr = r * r, mm = m * m;
    p->mag[i] = rr + mm;
  }
}

/* src/computefourier-1.0-2.0.cc:92-116 */
static void vote_regular(orc_plan *p, const int *J, int B, int a)
{
  const int n = p->n, num = p->B_thresh;
  for (int i = 0; i < num; i++) {
    int low = ((int)ceil((J[i] - 0.5) * n / B) + n) % n;
    int high = ((int)ceil((J[i] + 0.5) * n / B) + n) % n;
    int loc = mulmod(low, a, n);
    for (int j = low; j != high; j = (j + 1) % n) {
      p->score[loc]++;
      if (p->score[loc] == p->loops_thresh) p->hits[p->hits_found++] = loc;
      loc = (loc + a) % n;
    }
  }
}

typedef struct { int first, second; } int_pair;
static int cmp_pair(const void *a, const void *b)
{
  const int_pair *x = (const int_pair *)a, *y = (const int_pair *)b;
  if (x->first != y->first) return (x->first > y->first) - (x->first < y->first);
  return (x->second > y->second) - (x->second < y->second);
}

/* src/computefourier-1.0-2.0.cc:126-184 */
static void vote_comb(orc_plan *p, const int *J, int B, int a, int ai)
{
  const int n = p->n, num = p->B_thresh, W = p->W_Comb, nc = p->num_comb;
  int_pair *pa = (int_pair *)malloc((size_t)nc * sizeof(int_pair));
  for (int m = 0; m < nc; m++) {
    int prev = mulmod(p->comb_approved[m], ai, W);
    pa[m].first = prev;
    pa[m].second = mulmod(prev, a, n);
  }
  qsort(pa, (size_t)nc, sizeof(int_pair), cmp_pair);
  for (int i = 0; i < num; i++) {
    int low = ((int)ceil((J[i] - 0.5) * n / B) + n) % n;
    int high = ((int)ceil((J[i] + 0.5) * n / B) + n) % n;
    /* first entry with (first, second) > (low % W, -1) */
    int key = low % W, index = 0;
    while (index < nc && pa[index].first < key) index++;
    int location = low - (low % W);
    int locinv = mulmod(location, a, n);
    for (int j = index;; j++) {
      if (j == nc) {
        j -= nc;
        location = (location + W) % n;
        locinv = mulmod(location, a, n);
      }
      int approved_loc = location + pa[j].first;
      if ((low < high && (approved_loc >= high || approved_loc < low)) ||
          (low > high && (approved_loc >= high && approved_loc < low)))
        break;
      int loc = (locinv + pa[j].second) % n;
      p->score[loc]++;
      if (p->score[loc] == p->loops_thresh) p->hits[p->hits_found++] = loc;
    }
  }
  free(pa);
}

void orc_select_and_vote(orc_plan *p)
{
  /* src/computefourier-1.0-2.0.cc:292-324 */
  const int loops = p->loops_loc + p->loops_est, num = p->B_thresh;
  for (int j = 0; j < loops; j++) {
    const int is_loc = j < p->loops_loc;
    const int B = is_loc ? p->B_loc : p->B_est;
    int *J = p->J + (long)j * num;
    orc_find_largest_indices(J, num, p->mag + loop_offset(p, j), B);
    if (!is_loc) continue;
    if (!p->with_comb) vote_regular(p, J, B, p->a[j]);
    else vote_comb(p, J, B, p->a[j], p->ai[j]);
  }
}

void orc_estimate(orc_plan *p, ocplx *out)
{
  /* src/computefourier-1.0-2.0.cc:341-419.  NOTE the sign of the imaginary
   * part: the reference multiplies (a*d, b*c) by (+1, -1) before the horizontal
   * add (:388-392), i.e. it forms (a*d - b*c)/(c^2+d^2), the CONJUGATE of the
   * quotient's imaginary part.  Restated as executed, not as intended. */
  const int n = p->n, loops = p->loops_loc + p->loops_est;
  double *vr = (double *)malloc((size_t)loops * sizeof(double));
  double *vi = (double *)malloc((size_t)loops * sizeof(double));
  const int mid = (loops - 1) / 2;                     /* :406 */
  for (long h = 0; h < p->hits_found; h++) {
    const int loc = p->hits[h];
    for (int j = 0; j < loops; j++) {
      const int is_loc = j < p->loops_loc;
      const int B = is_loc ? p->B_loc : p->B_est;
      const ocplx *freq = is_loc ? p->freq_loc : p->freq_est;
      const int seg = n / B;
      int pos = mulmod(p->ai[j], loc, n);              /* :370, permute[j] = ai[j] */
      int bucket = pos / seg;
      int dist = pos % seg;
      if (dist > seg / 2) { bucket = (bucket + 1) % B; dist -= seg; }
      dist = (n - dist) % n;
      const ocplx s = p->x_samp[loop_offset(p, j) + bucket];
      const ocplx f = freq[dist];
      const double ac = s.re * f.re, bd = s.im * f.im;
      const double ad = s.re * f.im, bc = s.im * f.re;
      const double cc = f.re * f.re, dd = f.im * f.im;
      const double den = cc + dd;
      const double num_re = ac + bd;
      const double num_im = ad + (bc * -1.0);
      vr[j] = num_re / den;
      vi[j] = num_im / den;
    }
    qsort(vr, (size_t)loops, sizeof(double), cmp_double);
    qsort(vi, (size_t)loops, sizeof(double), cmp_double);
    out[loc].re = vr[mid];
    out[loc].im = vi[mid];
  }
  free(vr); free(vi);
}

static void exec_v12(orc_plan *p, const ocplx *x, ocplx *out)
{
  /* src/computefourier-1.0-2.0.cc:438-541 */
  memset(p->score, 0, (size_t)p->n * sizeof(int));
  orc_draw_permutations(p);
  orc_comb_stage(p, x);
  orc_bucketize(p, x);
  orc_bucket_ffts(p);
  orc_select_and_vote(p);
  orc_estimate(p, out);
}

/* ------------------------------------------------------------------------- */
/* v3 (exact-sparse): see sfft_oracle_v3.inc                                  */
/* ------------------------------------------------------------------------- */
#include "sfft_oracle_v3.inc"

void orc_exec(orc_plan *p, const ocplx *x, ocplx *out)
{
  /* src/sfft.cc:119-137 */
  for (int i = 0; i < p->n_requested; i++) { out[i].re = 0; out[i].im = 0; }
  if (p->version == 3) exec_v3(p, x, out);
  else exec_v12(p, x, out);
}

/* ------------------------------------------------------------------------- */
/* input synthesis                                                            */
/* ------------------------------------------------------------------------- */

void orc_generate_input(int n, int k, ocplx *x_time, ocplx *x_freq)
{
  /* src/simulation.cc:104-111 */
  memset(x_freq, 0, (size_t)n * sizeof(ocplx));
  for (int i = 0; i < k; i++) {
    unsigned f = (unsigned)floor(drand48() * n);
    x_freq[f].re = 1.0; x_freq[f].im = 0.0;
  }
  memcpy(x_time, x_freq, (size_t)n * sizeof(ocplx));
  orc_fft_any(x_time, n, +1);
}

double orc_awgn(ocplx *x, int n, double std_noise)
{
  /* src/utils.cc:250-280 */
  if (std_noise == 0) return 1000000000;
  double sig_power = 0, noise_power = 0;
  for (int h = 0; h < n; h++) {
    double m = cabs(CMPLX(x[h].re, x[h].im));
    sig_power += m * m;
    double u = drand48();
    double v = drand48();
    double complex gn = std_noise * sqrt(-2 * log(u)) * cexp(2 * M_PI * I * v);
    noise_power += -2 * log(u);
    x[h].re += creal(gn);
    x[h].im += cimag(gn);
  }
  noise_power = noise_power * std_noise * std_noise;
  return sig_power / noise_power;
}
