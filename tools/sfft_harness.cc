// sfft_harness.cc -- the reference's three drivers re-created over this library:
//   sfft-timing        (reference src/timing.cc:27-38)          TIME <seconds>
//   sfft-verification  (reference src/verification.cc:26-62)    OK / ERROR, exit 0/1/2
//   sfft-timing_many   (reference src/timing_many.cc:135-198)   -i inputs, -s one plan per input
// Same CLI (-n -k -r -v -o -h, plus -i -s), same synthetic input
// (src/simulation.cc:95-112: k unit spikes at floor(drand48()*n), x = unnormalised
// inverse DFT), but with DETERMINISTIC seeds (-S seed, default 12345) instead of
// time^pid, and the input synthesised by table lookup instead of an n-point FFTW call.
//
// Build: g++ -O2 -DHARNESS_MODE=<0|1|2> tools/sfft_harness.cc -Iinclude -Lsfft_b200 -lsfft
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <vector>

#include "sfft.h"

#ifndef HARNESS_MODE
#define HARNESS_MODE 0   // 0 timing, 1 verification, 2 timing_many
#endif

static void usage(const char *base)
{
  printf("Usage: %s [OPTIONS]\n"
         "Options:\n"
         "  -n num     Set the problem size to num\n"
         "  -k num     Set the number of frequencies to num\n"
         "  -r num     Perform num repetitions of the experiment\n"
         "  -i num     Perform the sFFT on num inputs (timing_many)\n"
         "  -s         Simple parallelism: one plan per input (timing_many)\n"
         "  -o         Use FFTW_MEASURE instead of FFTW_ESTIMATE (accepted, ignored)\n"
         "  -v version Use specific sFFT version (valid arguments: 1, 2 or 3)\n"
         "  -S seed    srand48 seed of the synthetic input (default 12345)\n"
         "  -h         Print this help message\n", base);
}

// The reference's generator (src/simulation.cc:104-111): k draws f = floor(drand48() * n), X[f] = 1
// (collisions allowed), x = unnormalised inverse DFT of X.  The inverse DFT runs on the device
// through the library's stand-alone FFT hook, so BASELINE sizes (n = 2^27) take milliseconds;
// round 1 summed k complex exponentials per sample, O(n k).
static int generate(int n, int k, sfft_complex *x, std::vector<int> &freqs)
{
  std::vector<sfft_complex> xf((size_t)n);
  for (int t = 0; t < n; t++) { xf[(size_t)t].re = 0; xf[(size_t)t].im = 0; }
  freqs.clear();
  for (int i = 0; i < k; i++) {
    const int f = (int)(unsigned)floor(drand48() * n);
    if (xf[(size_t)f].re == 0) freqs.push_back(f);
    xf[(size_t)f].re = 1.0;
  }
  int log2n = 0;
  while ((1 << log2n) < n) log2n++;
  if (sfftb_debug_fft(xf.data(), x, log2n, 1, +1, 0)) {
    fprintf(stderr, "input synthesis failed: %s\n", sfftb_last_error());
    return -1;
  }
  return 0;
}

int main(int argc, char **argv)
{
  int n = HARNESS_MODE == 2 ? (1 << 18) : 16384, k = HARNESS_MODE == 2 ? 100 : 50;
  int repetitions = 1, version = 1, num_inputs = HARNESS_MODE == 2 ? 100 : 1, fftw_opt = SFFT_FFTW_ESTIMATE;
  bool simple = false;
  long seed = 12345;
  int ch;
  while ((ch = getopt(argc, argv, "htosi:n:k:r:v:S:")) != -1) {
    switch (ch) {
      case 'n': n = atoi(optarg); break;
      case 'k': k = atoi(optarg); break;
      case 'r': repetitions = atoi(optarg); break;
      case 'v': version = atoi(optarg); break;
      case 'i': num_inputs = atoi(optarg); break;
      case 's': simple = true; break;
      case 'o': fftw_opt = SFFT_FFTW_MEASURE; break;
      case 'S': seed = atol(optarg); break;
      default: usage(argv[0]); return 1;
    }
  }
  if (version < 1 || version > 3) { usage(argv[0]); return 1; }

  std::vector<sfft_plan *> plans;
  const int nplans = (HARNESS_MODE == 2 && simple) ? num_inputs : 1;
  for (int i = 0; i < nplans; i++) {
    sfft_plan *p = sfft_make_plan(n, k, (sfft_version)(version - 1), fftw_opt);
    if (!p) { fprintf(stderr, "sfft_make_plan failed: %s\n", sfftb_last_error()); usage(argv[0]); return 1; }
    plans.push_back(p);
  }

  srand(17);            // src/simulation.cc:100
  srand48(seed);        // deterministic stand-in for time^pid (:101)
  std::vector<sfft_complex *> in((size_t)num_inputs), out((size_t)num_inputs);
  std::vector<std::vector<int> > freqs((size_t)num_inputs);
  for (int s = 0; s < num_inputs; s++) {
    in[(size_t)s] = (sfft_complex *)sfft_malloc(sizeof(sfft_complex) * (size_t)n);
    out[(size_t)s] = (sfft_complex *)sfft_malloc(sizeof(sfft_complex) * (size_t)n);
    if (generate(n, k, in[(size_t)s], freqs[(size_t)s])) return 1;
  }

  timespec ts, te;
  clock_gettime(CLOCK_REALTIME, &ts);
  if (HARNESS_MODE == 2) {
    if (simple) for (int s = 0; s < num_inputs; s++) sfft_exec(plans[(size_t)s], in[(size_t)s], out[(size_t)s]);
    else sfft_exec_many(plans[0], num_inputs, in.data(), out.data());
  } else {
    for (int r = 0; r < repetitions; r++) sfft_exec(plans[0], in[0], out[0]);
  }
  clock_gettime(CLOCK_REALTIME, &te);
  const double t = (te.tv_sec + 1e-9 * te.tv_nsec) - (ts.tv_sec + 1e-9 * ts.tv_nsec);

  int rc = 0;
  if (HARNESS_MODE == 1) {
    // src/verification.cc:39-56: every planted frequency recovered to within 0.1
    const double ERROR_THRESHOLD = 0.1;
    for (size_t q = 0; q < freqs[0].size() && rc == 0; q++) {
      const int f = freqs[0][q];
      const sfft_complex a = out[0][f];
      if (a.re == 0 && a.im == 0) { printf("ERROR: Frequency %d was not recovered!\n", f); rc = 1; }
      else if (hypot(a.re - 1.0, a.im) > ERROR_THRESHOLD) {
        printf("ERROR: Error of frequency %d is too big:\n  Expected: (1,0)\n  Actual  : (%g,%g)\n", f, a.re, a.im);
        rc = 2;
      }
    }
    if (rc == 0) printf("OK\n");
  } else {
    printf("TIME %g\n", t);
  }
  for (size_t i = 0; i < plans.size(); i++) sfft_free_plan(plans[i]);
  for (int s = 0; s < num_inputs; s++) { sfft_free(in[(size_t)s]); sfft_free(out[(size_t)s]); }
  return rc;
}
