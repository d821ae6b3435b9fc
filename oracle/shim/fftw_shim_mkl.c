/*
 * oracle/shim/fftw_shim_mkl.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * FFTW-API stand-in backed by Intel MKL's DFTI, for the CPU TIMING baseline only
 * (the radix-2 shim overstates FFT cost; FFTW itself cannot be installed here).
 * MKL is not installed as a library either: its DFTI entry points are exported by the
 * libtorch_cpu.so that ships with PyTorch, so they are resolved at run time with dlopen
 * from the path in $SFFT_REF_MKL_LIB (oracle/ref.py sets it).  No MKL header exists in the
 * image; the few DFTI constants used are declared below (values of mkl_dfti.h) and the
 * whole shim is validated against numpy.fft by tests/test_oracle_vs_ref.py before it times
 * anything.  This is "MKL DFTI, not FFTW": the bench line says so.
 * The parity oracle stays on oracle/fft_ref.c (shim/fftw_shim.c); rounding differs here.
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { DFTI_FORWARD_SCALE = 4, DFTI_BACKWARD_SCALE = 5, DFTI_NUMBER_OF_TRANSFORMS = 7, DFTI_PLACEMENT = 11,
       DFTI_INPUT_STRIDES = 12, DFTI_OUTPUT_STRIDES = 13, DFTI_INPUT_DISTANCE = 14, DFTI_OUTPUT_DISTANCE = 15,
       DFTI_THREAD_LIMIT = 27 };
enum { DFTI_COMPLEX = 32, DFTI_DOUBLE = 36, DFTI_INPLACE = 43, DFTI_NOT_INPLACE = 44 };

typedef void *dfti_desc;
static long (*p_create)(dfti_desc *, int, long);               /* DftiCreateDescriptor_d_1d(desc, domain, n) */
static long (*p_set)(dfti_desc, int, ...);
static long (*p_commit)(dfti_desc);
static long (*p_fwd)(dfti_desc, void *, ...);
static long (*p_bwd)(dfti_desc, void *, ...);
static long (*p_free)(dfti_desc *);
static char *(*p_errmsg)(long);

static int mkl_bind(void)
{
  static int state = 0;      /* 0 untried, 1 ok, -1 failed */
  if (state) return state;
  const char *path = getenv("SFFT_REF_MKL_LIB");
  void *h = path ? dlopen(path, RTLD_NOW | RTLD_GLOBAL) : NULL;
  if (!h) {
    fprintf(stderr, "[fftw_shim_mkl] cannot dlopen $SFFT_REF_MKL_LIB (%s): %s\n", path ? path : "unset", dlerror());
    state = -1;
    return state;
  }
  p_create = (long (*)(dfti_desc *, int, long))dlsym(h, "DftiCreateDescriptor_d_1d");
  p_set = (long (*)(dfti_desc, int, ...))dlsym(h, "DftiSetValue");
  p_commit = (long (*)(dfti_desc))dlsym(h, "DftiCommitDescriptor");
  p_fwd = (long (*)(dfti_desc, void *, ...))dlsym(h, "DftiComputeForward");
  p_bwd = (long (*)(dfti_desc, void *, ...))dlsym(h, "DftiComputeBackward");
  p_free = (long (*)(dfti_desc *))dlsym(h, "DftiFreeDescriptor");
  p_errmsg = (char *(*)(long))dlsym(h, "DftiErrorMessage");
  state = (p_create && p_set && p_commit && p_fwd && p_bwd && p_free) ? 1 : -1;
  if (state < 0) fprintf(stderr, "[fftw_shim_mkl] DFTI symbols missing in %s\n", path);
  return state;
}

typedef struct mkl_fftw_plan_s {
  dfti_desc desc;
  void *in, *out;
  int sign;
  long n;
  int howmany;
} mkl_plan;

static void check(long st, const char *what)
{
  if (st != 0) {
    fprintf(stderr, "[fftw_shim_mkl] %s failed: %s\n", what, p_errmsg ? p_errmsg(st) : "?");
    abort();
  }
}

static void *make(long n, int howmany, void *in, long istride, long idist, void *out, long ostride, long odist,
                  int sign)
{
  if (mkl_bind() < 0) abort();
  mkl_plan *p = (mkl_plan *)calloc(1, sizeof *p);
  p->in = in; p->out = out; p->sign = sign; p->n = n; p->howmany = howmany;
  check(p_create(&p->desc, DFTI_COMPLEX, n), "DftiCreateDescriptor");
  long is[2] = {0, istride}, os[2] = {0, ostride};
  check(p_set(p->desc, DFTI_PLACEMENT, in == out ? DFTI_INPLACE : DFTI_NOT_INPLACE), "DFTI_PLACEMENT");
  check(p_set(p->desc, DFTI_NUMBER_OF_TRANSFORMS, (long)howmany), "DFTI_NUMBER_OF_TRANSFORMS");
  check(p_set(p->desc, DFTI_INPUT_STRIDES, is), "DFTI_INPUT_STRIDES");
  check(p_set(p->desc, DFTI_OUTPUT_STRIDES, os), "DFTI_OUTPUT_STRIDES");
  check(p_set(p->desc, DFTI_INPUT_DISTANCE, idist), "DFTI_INPUT_DISTANCE");
  check(p_set(p->desc, DFTI_OUTPUT_DISTANCE, odist), "DFTI_OUTPUT_DISTANCE");
  /* the reference's FFTW plans are single-threaded (no fftw_init_threads anywhere in src/) */
  check(p_set(p->desc, DFTI_THREAD_LIMIT, 1L), "DFTI_THREAD_LIMIT");
  check(p_commit(p->desc), "DftiCommitDescriptor");
  return p;
}

void *fftw_plan_dft_1d(int n, void *in, void *out, int sign, unsigned flags)
{
  (void)flags;
  return make(n, 1, in, 1, n, out, 1, n, sign);
}

void *fftw_plan_many_dft(int rank, const int *n, int howmany, void *in, const int *inembed, int istride, int idist,
                         void *out, const int *onembed, int ostride, int odist, int sign, unsigned flags)
{
  (void)inembed; (void)onembed; (void)flags;
  if (rank != 1) return NULL;
  return make(n[0], howmany, in, istride, idist, out, ostride, odist, sign);
}

void fftw_execute(void *vp)
{
  mkl_plan *p = (mkl_plan *)vp;
  long st;
  /* FFTW_FORWARD = -1: exponent sign -1 = DFTI forward; both unnormalised by default */
  if (p->in == p->out) st = p->sign < 0 ? p_fwd(p->desc, p->in) : p_bwd(p->desc, p->in);
  else st = p->sign < 0 ? p_fwd(p->desc, p->in, p->out) : p_bwd(p->desc, p->in, p->out);
  check(st, "DftiCompute");
}

void fftw_destroy_plan(void *vp)
{
  mkl_plan *p = (mkl_plan *)vp;
  if (!p) return;
  if (p->desc) p_free(&p->desc);
  free(p);
}

void *fftw_malloc(size_t n)
{
  void *p = NULL;
  if (posix_memalign(&p, 64, n ? n : 64)) return NULL;
  return p;
}

void fftw_free(void *p) { free(p); }

void fftw_flops(void *vp, double *add, double *mul, double *fmas)
{
  mkl_plan *p = (mkl_plan *)vp;
  double lg = log2((double)p->n);
  *add = 3.0 * p->n * lg * p->howmany; *mul = 2.0 * p->n * lg * p->howmany; *fmas = 0;
}
