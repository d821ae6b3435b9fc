// api.cu -- the C ABI of libsfft.so (include/sfft.h).
//
// Part 1 replaces the reference's public entry points (src/sfft.cc:60-147);
// part 2 is the device-resident extension they are built on.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "fft.cuh"
#include "plan.cuh"

namespace sfftb {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const std::string &msg)
{
  t_error = msg;
  if (getenv("SFFTB_VERBOSE")) fprintf(stderr, "[libsfft] %s\n", msg.c_str());
}

void timer_begin(PlanImpl *p)
{
  StageTimer &t = p->timer;
  if (!t.enabled) return;
  if (!t.created) {
    for (int i = 0; i <= kMaxStages; i++) cudaEventCreate(&t.ev[i]);
    t.created = true;
  }
  t.count = 0;
  cudaEventRecord(t.ev[0], p->stream);
}

void timer_mark(PlanImpl *p, const char *name)
{
  StageTimer &t = p->timer;
  if (!t.enabled || t.count >= kMaxStages) return;
  t.names[t.count] = name;
  t.count++;
  cudaEventRecord(t.ev[t.count], p->stream);
}

static std::mutex g_pin_mu;
static std::unordered_set<void *> g_pinned;

static PlanImpl *impl(const sfft_plan *plan) { return plan ? (PlanImpl *)plan->data : nullptr; }

static int bind_device(const PlanImpl *p)
{
  SFFTB_CUDA(cudaSetDevice(p->device));
  return 0;
}

}  // namespace sfftb

using namespace sfftb;

extern "C" {

const char *sfftb_last_error(void) { return t_error.c_str(); }

int sfftb_device_count(void)
{
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
  return c;
}

long long sfftb_launch_count(void) { return g_launches.load(); }

/* ------------------------------------------------------------------------ */
/* Part 1: drop-in boundary                                                  */
/* ------------------------------------------------------------------------ */

void *sfft_malloc(size_t s)
{
  void *p = nullptr;
  if (sfftb_device_count() > 0 && cudaHostAlloc(&p, s ? s : 16, cudaHostAllocDefault) == cudaSuccess) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pinned.insert(p);
    return p;
  }
  cudaGetLastError();
  if (posix_memalign(&p, 64, s ? s : 16)) return nullptr;
  return p;
}

void sfft_free(void *p)
{
  if (!p) return;
  bool pinned = false;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    pinned = g_pinned.erase(p) > 0;
  }
  if (pinned) cudaFreeHost(p);
  else free(p);
}

sfft_plan *sfft_make_plan(int n, int k, sfft_version version, int fftw_optimization)
{
  // FFTW planner flags are accepted and ignored: no FFTW here (python/sfft/sfft.py:74 passes
  // them through); SFFTB_PLAN_TUNED_BY_K is this library's own opt-in bit
  if (version != SFFT_VERSION_1 && version != SFFT_VERSION_2 && version != SFFT_VERSION_3) {
    set_error("sfft_make_plan: unknown version (reference returns NULL, sfft.cc:87-88)");
    return nullptr;
  }
  if (sfftb_device_count() <= 0) {
    set_error("sfft_make_plan: no CUDA device; this library has no CPU path");
    fprintf(stderr, "[libsfft] FATAL: no CUDA device visible; libsfft.so has no CPU fallback\n");
    return nullptr;
  }
  cudaGetLastError();   // drop any stale, non-sticky error left by the caller's earlier CUDA calls
  if (const char *e = getenv("SFFTB_L2_FETCH")) {
    // A/B knob (profiles/r02_gather_ab.md): device-wide L2 fetch granularity hint, 32 / 64 / 128 bytes
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
    cudaGetLastError();
  }
  PlanImpl *p = new PlanImpl();
  if (cudaGetDevice(&p->device) != cudaSuccess) { delete p; return nullptr; }
  if (cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("sfft_make_plan: cannot create a stream");
    delete p;
    return nullptr;
  }
  p->stream = p->own_stream;
  p->version = (int)version + 1;
  p->flags = fftw_optimization;
  int rc;
  if (version == SFFT_VERSION_3) {
    rc = v3_build(p, n, k);
  } else {
    rc = v12_derive(p, n, k, version == SFFT_VERSION_2, (fftw_optimization & SFFTB_PLAN_TUNED_BY_K) != 0);
    if (!rc) rc = v12_build(p);
  }
  if (rc) {
    if (version == SFFT_VERSION_3) v3_free(p); else v12_free(p);
    cudaStreamDestroy(p->own_stream);
    delete p;
    return nullptr;
  }
  sfft_plan *plan = (sfft_plan *)malloc(sizeof(sfft_plan));
  plan->version = version;
  plan->n = (unsigned)n;
  plan->k = (unsigned)k;
  plan->data = p;
  return plan;
}

void sfft_free_plan(sfft_plan *plan)
{
  if (!plan) return;
  PlanImpl *p = impl(plan);
  if (p) {
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    if (p->version == 3) v3_free(p); else v12_free(p);
    cudaFree(p->d_in);
    cudaFree(p->d_out);
    cudaFree(p->d_zero);
    if (p->h_loc) cudaFreeHost(p->h_loc);
    if (p->h_val) cudaFreeHost(p->h_val);
    if (p->zero_stream) cudaStreamDestroy(p->zero_stream);
    if (p->timer.created)
      for (int i = 0; i <= kMaxStages; i++) cudaEventDestroy(p->timer.ev[i]);
    cudaStreamDestroy(p->own_stream);
    delete p;
  }
  free(plan);
}

static int ensure_io(PlanImpl *p, long long in_elems)
{
  if (p->d_in_elems < in_elems) {
    SFFTB_CUDA(cudaStreamSynchronize(p->stream));
    cudaFree(p->d_in);
    p->d_in = nullptr;
    SFFTB_CUDA(cudaMalloc(&p->d_in, sizeof(cplx) * in_elems));
    p->d_in_elems = in_elems;
  }
  return 0;
}

static int ensure_dense_out(PlanImpl *p)
{
  if (!p->d_out) SFFTB_CUDA(cudaMalloc(&p->d_out, sizeof(cplx) * (long long)p->n));
  return 0;
}

constexpr long long kHostScatterCap = 1ll << 18;   // coefficients the host scatters itself
constexpr long long kZeroElems = 4ll << 20;         // 64 MiB of device zeros, copied repeatedly

static int ensure_sparse_out(PlanImpl *p)
{
  if (p->d_zero) return 0;
  p->zero_elems = p->n < kZeroElems ? p->n : kZeroElems;
  SFFTB_CUDA(cudaStreamCreateWithFlags(&p->zero_stream, cudaStreamNonBlocking));
  SFFTB_CUDA(cudaMalloc(&p->d_zero, sizeof(cplx) * p->zero_elems));
  SFFTB_CUDA(cudaMemset(p->d_zero, 0, sizeof(cplx) * p->zero_elems));
  SFFTB_CUDA(cudaHostAlloc(&p->h_loc, sizeof(int) * kHostScatterCap, cudaHostAllocDefault));
  SFFTB_CUDA(cudaHostAlloc(&p->h_val, sizeof(cplx) * kHostScatterCap, cudaHostAllocDefault));
  return 0;
}

// The legacy entry points (host arrays in, dense zero-filled host arrays out;
// src/sfft.cc:119-147).  v1 and v3 return a handful of coefficients, so the dense output is
// zeros plus a host-side scatter: the zeros are DMA-ed into `out` on a second stream WHILE
// the input is DMA-ed in (PCIe is full duplex), and only the sparse list comes back after the
// transform.  v2's result is dense (cf12.cc:505-512): it is densified on the device and
// copied back whole.
static int exec_host(sfft_plan *plan, int num, sfft_complex **in, sfft_complex **out)
{
  PlanImpl *p = impl(plan);
  if (!p) { set_error("sfft_exec: null plan"); return -1; }
  if (bind_device(p)) return -1;
  const long long n = p->n;
  const bool sparse_out = p->version != 2;
  if (sparse_out && ensure_sparse_out(p)) return -1;
  // batch size bounded by ~4 GiB of staged input
  long long chunk = (4ll << 30) / (16 * n);
  if (chunk < 1) chunk = 1;
  if (chunk > num) chunk = num;
  if (chunk > 1024) chunk = 1024;
  if (ensure_io(p, chunk * n)) return -1;
  std::vector<sfftb_draw> draws((size_t)chunk);
  for (int base = 0; base < num; base += (int)chunk) {
    const int cnt = num - base < chunk ? num - base : (int)chunk;
    // draws in signal order on the calling thread (reference: sfft.cc:142-146 draws
    // inside each OpenMP iteration, i.e. in a schedule-dependent order)
    for (int s = 0; s < cnt; s++) {
      if (sfftb_draw_random(plan, &draws[(size_t)s])) return -1;
      SFFTB_CUDA(cudaMemcpyAsync(p->d_in + s * n, in[base + s], sizeof(cplx) * n,
                                 cudaMemcpyHostToDevice, p->stream));
      if (sparse_out)
        for (long long off = 0; off < n; off += p->zero_elems) {
          const long long len = n - off < p->zero_elems ? n - off : p->zero_elems;
          SFFTB_CUDA(cudaMemcpyAsync(reinterpret_cast<cplx *>(out[base + s]) + off, p->d_zero, sizeof(cplx) * len,
                                     cudaMemcpyDeviceToHost, p->zero_stream));
        }
    }
    sfftb_result res;
    if (sfftb_exec_many_device(plan, cnt, p->d_in, n, draws.data(), &res, nullptr, 0)) return -1;
    if (sparse_out) SFFTB_CUDA(cudaStreamSynchronize(p->zero_stream));
    for (int s = 0; s < cnt; s++) {
      if (sparse_out) {
        const long long c = sfftb_fetch_result(plan, s, p->h_loc, reinterpret_cast<sfft_complex *>(p->h_val),
                                               kHostScatterCap);
        if (c < 0) return -1;
        if (c <= kHostScatterCap) {
          cplx *o = reinterpret_cast<cplx *>(out[base + s]);
          for (long long i = 0; i < c; i++) o[p->h_loc[i]] = p->h_val[i];            // sfft.cc:121-123, cf12.cc:413
          continue;
        }
      }
      if (ensure_dense_out(p) || sfftb_densify(plan, s, p->d_out)) return -1;
      SFFTB_CUDA(cudaMemcpyAsync(out[base + s], p->d_out, sizeof(cplx) * n, cudaMemcpyDeviceToHost,
                                 p->stream));
    }
    SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  }
  return 0;
}

void sfft_exec(sfft_plan *plan, sfft_complex *in, sfft_complex *out)
{
  if (exec_host(plan, 1, &in, &out)) {
    fprintf(stderr, "[libsfft] sfft_exec failed: %s\n", sfftb_last_error());
    abort();   // the reference's failure mode is assert -> abort
  }
}

void sfft_exec_many(sfft_plan *plan, int num, sfft_complex **in, sfft_complex **out)
{
  if (num <= 0) return;
  if (exec_host(plan, num, in, out)) {
    fprintf(stderr, "[libsfft] sfft_exec_many failed: %s\n", sfftb_last_error());
    abort();
  }
}

/* ------------------------------------------------------------------------ */
/* Part 2: device-resident extension                                         */
/* ------------------------------------------------------------------------ */

int sfftb_plan_info(const sfft_plan *plan, sfftb_info *info)
{
  const PlanImpl *p = impl(plan);
  if (!p || !info) { set_error("sfftb_plan_info: null argument"); return -1; }
  memset(info, 0, sizeof(*info));
  info->version = p->version;
  info->n = p->n;
  info->k = p->k;
  info->device = p->device;
  if (p->version == 3) return v3_info(p, info);
  const PlanV12 &v = p->v12;
  info->B_loc = v.B_loc; info->B_est = v.B_est; info->B_thresh = v.B_thresh;
  info->W_Comb = v.W_Comb; info->Comb_loops = v.Comb_loops;
  info->loops_loc = v.loops_loc; info->loops_thresh = v.loops_thresh; info->loops_est = v.loops_est;
  info->w_loc = v.filt[0].w; info->w_est = v.filt[1].w;
  info->b_loc = v.b_loc; info->b_est = v.b_est;
  info->x_samp_size = v.x_samp_size;
  info->max_hits = v.max_hits;
  info->gather_samples = (long long)v.loops_loc * v.filt[0].w + (long long)v.loops_est * v.filt[1].w +
                         (v.with_comb ? (long long)v.Comb_loops * v.W_Comb : 0);
  info->gather_tap_bytes = 16ll * (v.filt[0].w + v.filt[1].w);
  return 0;
}

int sfftb_set_stream(sfft_plan *plan, void *cuda_stream)
{
  PlanImpl *p = impl(plan);
  if (!p) { set_error("sfftb_set_stream: null plan"); return -1; }
  if (bind_device(p)) return -1;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  p->stream = cuda_stream ? (cudaStream_t)cuda_stream : p->own_stream;
  return 0;
}

int sfftb_draw_random(const sfft_plan *plan, sfftb_draw *draw)
{
  const PlanImpl *p = impl(plan);
  if (!p || !draw) { set_error("sfftb_draw_random: null argument"); return -1; }
  memset(draw, 0, sizeof(*draw));
  return p->version == 3 ? v3_draw(p, draw) : v12_draw(p, draw);
}

static int result_view(PlanImpl *p, int which, const int **loc, const cplx **val, const int **count,
                       long long *cap)
{
  if (p->version == 3) {
    if (v3_result(p, loc, val, count, cap)) return -1;
    *loc += (long long)which * *cap;
    *val += (long long)which * *cap;
    *count += which;
    return 0;
  }
  PlanV12 &v = p->v12;
  *cap = v.max_hits;
  *loc = v.d_hit_loc + (long long)which * v.max_hits;
  *val = v.d_hit_val + (long long)which * v.max_hits;
  *count = (v.with_comb ? v.d_count : v.d_voted_count) + which;
  return 0;
}

int sfftb_exec_many_device(sfft_plan *plan, int num, const void *d_in, long long stride_elems,
                           const sfftb_draw *draws, sfftb_result *result, long long *counts, int sync)
{
  PlanImpl *p = impl(plan);
  if (!p || !d_in || num <= 0) { set_error("sfftb_exec_many_device: bad argument"); return -1; }
  if (bind_device(p)) return -1;
  cudaGetLastError();   // stale errors of the caller must not be mistaken for launch failures
  std::vector<sfftb_draw> local;
  if (!draws) {
    local.resize((size_t)num);
    for (int s = 0; s < num; s++)
      if (sfftb_draw_random(plan, &local[(size_t)s])) return -1;
    draws = local.data();
  }
  int rc = p->version == 3 ? v3_exec(p, (const cplx *)d_in, stride_elems, num, draws)
                           : v12_exec(p, (const cplx *)d_in, stride_elems, num, draws);
  if (rc) return -1;
  const int *loc; const cplx *val; const int *cnt; long long cap;
  if (result_view(p, 0, &loc, &val, &cnt, &cap)) return -1;
  if (result) {
    result->d_loc = loc;
    result->d_val = (const sfft_complex *)val;
    result->d_count = cnt;
    result->count = -1;
  }
  if (sync) {
    std::vector<int> h((size_t)num);
    SFFTB_CUDA(cudaMemcpyAsync(h.data(), cnt, sizeof(int) * num, cudaMemcpyDeviceToHost, p->stream));
    SFFTB_CUDA(cudaStreamSynchronize(p->stream));
    for (int s = 0; s < num; s++) {
      long long c = h[(size_t)s];
      if (c > cap) c = cap;
      if (counts) counts[s] = c;
    }
    if (result) result->count = h[0] > cap ? cap : h[0];
  }
  return 0;
}

int sfftb_exec_device(sfft_plan *plan, const void *d_in, const sfftb_draw *draw, sfftb_result *result,
                      int sync)
{
  const PlanImpl *p = impl(plan);
  if (!p) { set_error("sfftb_exec_device: null plan"); return -1; }
  return sfftb_exec_many_device(plan, 1, d_in, p->n, draw, result, nullptr, sync);
}

int sfftb_densify(sfft_plan *plan, int which, void *d_out)
{
  PlanImpl *p = impl(plan);
  if (!p || !d_out) { set_error("sfftb_densify: null argument"); return -1; }
  if (bind_device(p)) return -1;
  if (which < 0 || which >= p->last_nsig) { set_error("sfftb_densify: no such signal in the last batch"); return -1; }
  const int *loc; const cplx *val; const int *cnt; long long cap;
  if (result_view(p, which, &loc, &val, &cnt, &cap)) return -1;
  // sfft.cc:121-123: the legacy API hands back a dense, zero-filled spectrum
  SFFTB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(cplx) * (long long)p->n, p->stream));
  return launch_scatter(loc, val, cnt, (cplx *)d_out, p->stream);
}

int sfftb_synchronize(sfft_plan *plan)
{
  PlanImpl *p = impl(plan);
  if (!p) { set_error("sfftb_synchronize: null plan"); return -1; }
  if (bind_device(p)) return -1;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  return 0;
}

long long sfftb_fetch_result(sfft_plan *plan, int which, int *loc_out, sfft_complex *val_out,
                             long long capacity)
{
  PlanImpl *p = impl(plan);
  if (!p) { set_error("sfftb_fetch_result: null plan"); return -1; }
  if (bind_device(p)) return -1;
  if (which < 0 || which >= p->last_nsig) { set_error("sfftb_fetch_result: no such signal"); return -1; }
  const int *loc; const cplx *val; const int *cnt; long long cap;
  if (result_view(p, which, &loc, &val, &cnt, &cap)) return -1;
  int c = 0;
  SFFTB_CUDA(cudaMemcpyAsync(&c, cnt, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  long long count = c > cap ? cap : c;
  const long long take = count < capacity ? count : capacity;
  if (take > 0 && loc_out) SFFTB_CUDA(cudaMemcpyAsync(loc_out, loc, sizeof(int) * take, cudaMemcpyDeviceToHost, p->stream));
  if (take > 0 && val_out) SFFTB_CUDA(cudaMemcpyAsync(val_out, val, sizeof(cplx) * take, cudaMemcpyDeviceToHost, p->stream));
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  return count;
}

/* ---- plan-builder hooks ---- */
static DeviceFilter *filter_of(PlanImpl *p, int which)
{
  if (which < 0 || which > 1) return nullptr;
  return p->version == 3 ? v3_filter(p, which) : &p->v12.filt[which];
}

int sfftb_filter_sizes(const sfft_plan *plan, int which, int *w, int *fw_len)
{
  PlanImpl *p = impl(plan);
  DeviceFilter *f = p ? filter_of(p, which) : nullptr;
  if (!f) { set_error("sfftb_filter_sizes: bad argument"); return -1; }
  if (w) *w = f->w;
  if (fw_len) *fw_len = 2 * f->fw_half + 1;
  return 0;
}

int sfftb_get_filter(const sfft_plan *plan, int which, sfft_complex *time, sfft_complex *freq_window)
{
  PlanImpl *p = impl(plan);
  DeviceFilter *f = p ? filter_of(p, which) : nullptr;
  if (!f) { set_error("sfftb_get_filter: bad argument"); return -1; }
  if (bind_device(p)) return -1;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  if (time) SFFTB_CUDA(cudaMemcpy(time, f->time, sizeof(cplx) * f->w, cudaMemcpyDeviceToHost));
  if (freq_window)
    SFFTB_CUDA(cudaMemcpy(freq_window, f->fwin, sizeof(cplx) * (2ll * f->fw_half + 1), cudaMemcpyDeviceToHost));
  return 0;
}

int sfftb_set_filter(sfft_plan *plan, int which, const sfft_complex *time, const sfft_complex *freq_window)
{
  PlanImpl *p = impl(plan);
  DeviceFilter *f = p ? filter_of(p, which) : nullptr;
  if (!f) { set_error("sfftb_set_filter: bad argument"); return -1; }
  if (bind_device(p)) return -1;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  if (time) SFFTB_CUDA(cudaMemcpy(f->time, time, sizeof(cplx) * f->w, cudaMemcpyHostToDevice));
  if (freq_window) {
    SFFTB_CUDA(cudaMemcpy(f->fwin, freq_window, sizeof(cplx) * (2ll * f->fw_half + 1), cudaMemcpyHostToDevice));
    if (filter_refresh(f, p->stream)) return -1;
    SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  }
  return 0;
}

/* ---- plan cache ---- */
namespace {
struct PlanFileHeader {
  char magic[8];
  int version, n, k, flags, nfilt;
  int w[2], fw_half[2];
};
const char kPlanMagic[8] = {'S', 'F', 'F', 'T', 'B', 'P', 'L', '1'};
}

int sfftb_save_plan(const sfft_plan *plan, const char *path)
{
  PlanImpl *p = impl(plan);
  if (!p || !path) { set_error("sfftb_save_plan: null argument"); return -1; }
  PlanFileHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, kPlanMagic, 8);
  h.version = p->version - 1; h.n = p->n; h.k = p->k; h.flags = p->flags; h.nfilt = 2;
  std::vector<std::vector<cplx>> time(2), fwin(2);
  for (int f = 0; f < 2; f++) {
    int w = 0, len = 0;
    if (sfftb_filter_sizes(plan, f, &w, &len)) return -1;
    h.w[f] = w; h.fw_half[f] = (len - 1) / 2;
    time[f].resize((size_t)w); fwin[f].resize((size_t)len);
    if (sfftb_get_filter(plan, f, (sfft_complex *)time[f].data(), (sfft_complex *)fwin[f].data())) return -1;
  }
  FILE *fp = fopen(path, "wb");
  if (!fp) { set_error(std::string("sfftb_save_plan: cannot open ") + path); return -1; }
  bool ok = fwrite(&h, sizeof h, 1, fp) == 1;
  for (int f = 0; f < 2 && ok; f++) {
    ok = fwrite(time[f].data(), sizeof(cplx), time[f].size(), fp) == time[f].size() &&
         fwrite(fwin[f].data(), sizeof(cplx), fwin[f].size(), fp) == fwin[f].size();
  }
  ok = fclose(fp) == 0 && ok;
  if (!ok) { set_error(std::string("sfftb_save_plan: short write to ") + path); return -1; }
  return 0;
}

sfft_plan *sfftb_load_plan(const char *path)
{
  if (!path) { set_error("sfftb_load_plan: null path"); return nullptr; }
  FILE *fp = fopen(path, "rb");
  if (!fp) { set_error(std::string("sfftb_load_plan: cannot open ") + path); return nullptr; }
  PlanFileHeader h;
  bool ok = fread(&h, sizeof h, 1, fp) == 1 && memcmp(h.magic, kPlanMagic, 8) == 0 && h.nfilt == 2 &&
            h.version >= 0 && h.version <= 2;
  std::vector<std::vector<cplx>> time(2), fwin(2);
  for (int f = 0; f < 2 && ok; f++) {
    ok = h.w[f] > 0 && h.fw_half[f] >= 0 && h.w[f] <= h.n && h.fw_half[f] < h.n;
    if (!ok) break;
    time[f].resize((size_t)h.w[f]);
    fwin[f].resize((size_t)(2ll * h.fw_half[f] + 1));
    ok = fread(time[f].data(), sizeof(cplx), time[f].size(), fp) == time[f].size() &&
         fread(fwin[f].data(), sizeof(cplx), fwin[f].size(), fp) == fwin[f].size();
  }
  fclose(fp);
  if (!ok) { set_error(std::string("sfftb_load_plan: not a plan file, or truncated: ") + path); return nullptr; }
  PresetFilters pre;
  pre.count = 2;
  for (int f = 0; f < 2; f++) {
    pre.w[f] = h.w[f]; pre.fw_half[f] = h.fw_half[f];
    pre.time[f] = time[f].data(); pre.fwin[f] = fwin[f].data();
  }
  set_preset_filters(&pre);
  sfft_plan *plan = sfft_make_plan(h.n, h.k, (sfft_version)h.version, h.flags);
  set_preset_filters(nullptr);     // in case plan creation failed before reaching the builder
  return plan;
}

/* ---- stage hooks ---- */
long long sfftb_debug_fetch(sfft_plan *plan, const char *what, void *dst, size_t capacity)
{
  PlanImpl *p = impl(plan);
  if (!p || !what || !dst) { set_error("sfftb_debug_fetch: null argument"); return -1; }
  if (bind_device(p)) return -1;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  if (p->version == 3) return v3_debug_fetch(p, what, dst, capacity);
  PlanV12 &v = p->v12;
  const void *src = nullptr;
  long long bytes = 0;
  const std::string w(what);
  if ((w == "J" || w == "bitmap" || w == "voted") && v12_locate_on_demand(p)) return -1;
  const int loops = v.geom.loops;
  if (w == "x_samp" || w == "x_sampt") { src = v.d_xs; bytes = sizeof(cplx) * v.x_samp_size; }
  else if (w == "J") { src = v.d_J; bytes = sizeof(int) * (long long)v.loops_loc * v.B_thresh; }
  else if (w == "bitmap") { src = v.d_bitmap; bytes = sizeof(unsigned) * (long long)v.loops_loc * (v.B_loc >= 32 ? v.B_loc / 32 : 1); }
  else if (w == "perm_a") { src = v.d_stage; bytes = sizeof(int) * loops; }
  else if (w == "perm_ai") { src = v.d_stage + loops; bytes = sizeof(int) * loops; }
  else if (w == "twiddle") { src = v.d_tw; bytes = sizeof(cplx) * ((1ll << v.log_twN) > 1 ? (1ll << v.log_twN) - 1 : 1); }
  else if (w == "voted" || w == "hits" || w == "vals" || w == "comb_approved") {
    int c = 0;
    const int *cp = (w == "voted") ? v.d_voted_count
                    : (w == "comb_approved") ? v.d_num_comb
                    : (v.with_comb ? v.d_count : v.d_voted_count);
    if (!cp) { set_error("sfftb_debug_fetch: array not present for this plan"); return -1; }
    SFFTB_CUDA(cudaMemcpy(&c, cp, sizeof(int), cudaMemcpyDeviceToHost));
    if (w == "voted") { if (c > v.max_voted) c = (int)v.max_voted; src = v.d_voted; bytes = sizeof(int) * (long long)c; }
    else if (w == "comb_approved") { src = v.d_approved; bytes = sizeof(int) * (long long)c; }
    else {
      if (c > v.max_hits) c = (int)v.max_hits;
      if (w == "hits") { src = v.d_hit_loc; bytes = sizeof(int) * (long long)c; }
      else { src = v.d_hit_val; bytes = sizeof(cplx) * (long long)c; }
    }
  } else if (w == "comb_spec") {
    src = v.d_comb_xs; bytes = sizeof(cplx) * (long long)v.Comb_loops * v.W_Comb;
  } else {
    set_error("sfftb_debug_fetch: unknown array name");
    return -1;
  }
  if (bytes > (long long)capacity) { set_error("sfftb_debug_fetch: destination too small"); return -1; }
  if (bytes > 0) {
    if (!src) { set_error("sfftb_debug_fetch: array not present for this plan"); return -1; }
    SFFTB_CUDA(cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost));
  }
  return bytes;
}

int sfftb_debug_fft(const sfft_complex *in, sfft_complex *out, int log2n, int batch, int sign,
                    int table_twiddles)
{
  if (sfftb_device_count() <= 0) { set_error("sfftb_debug_fft: no CUDA device"); return -1; }
  const long long n = 1ll << log2n, total = n * batch;
  cplx *d_a = nullptr, *d_b = nullptr, *d_tw = nullptr;
  ScratchGuard guard;
  SFFTB_CUDA(cudaMalloc(&d_a, sizeof(cplx) * total));
  guard.track(d_a);
  SFFTB_CUDA(cudaMalloc(&d_b, sizeof(cplx) * total));
  guard.track(d_b);
  SFFTB_CUDA(cudaMemcpy(d_a, in, sizeof(cplx) * total, cudaMemcpyHostToDevice));
  for (int b = 0; b < batch; b++)
    if (bitrev_permute(d_a + b * n, d_b + b * n, log2n, 0)) return -1;
  if (table_twiddles) {
    std::vector<cplx> tw((size_t)(n > 1 ? n - 1 : 1));
    host_twiddle_levels(n, tw.data());
    SFFTB_CUDA(cudaMalloc(&d_tw, sizeof(cplx) * tw.size()));
    guard.track(d_tw);
    SFFTB_CUDA(cudaMemcpy(d_tw, tw.data(), sizeof(cplx) * tw.size(), cudaMemcpyHostToDevice));
  }
  if (fft_dit_inplace(d_b, log2n, batch, n, 1, total, d_tw, log2n, sign, 0)) return -1;
  SFFTB_CUDA(cudaDeviceSynchronize());
  SFFTB_CUDA(cudaMemcpy(out, d_b, sizeof(cplx) * total, cudaMemcpyDeviceToHost));
  cudaFree(d_a); cudaFree(d_b); cudaFree(d_tw);
  guard.dismiss();
  return 0;
}

int sfftb_debug_select(const double *mags, int B, int num, int batch, int *out_J)
{
  if (sfftb_device_count() <= 0) { set_error("sfftb_debug_select: no CUDA device"); return -1; }
  if (B < num + 1 || (B & (B - 1))) { set_error("sfftb_debug_select: need power-of-two B >= num+1"); return -1; }
  // feed magnitudes as (sqrt-free) complex values: re = sqrt(m) would round, so
  // place m itself through a value whose square is exact is impossible in
  // general; instead the hook takes |.|^2 = re^2 + 0 with re chosen by the caller.
  // The caller passes `mags` as the REAL PARTS; the kernel squares them.
  const long long total = (long long)B * batch;
  std::vector<cplx> h((size_t)total);
  for (long long i = 0; i < total; i++) h[(size_t)i] = make_double2(mags[i], 0.0);
  cplx *d_x = nullptr; int *d_J = nullptr; unsigned *d_bm = nullptr; unsigned long long *d_k = nullptr;
  const int words = B >= 32 ? B / 32 : 1;
  ScratchGuard guard;
  SFFTB_CUDA(cudaMalloc(&d_x, sizeof(cplx) * total));
  guard.track(d_x);
  SFFTB_CUDA(cudaMalloc(&d_J, sizeof(int) * (long long)num * batch));
  guard.track(d_J);
  SFFTB_CUDA(cudaMalloc(&d_bm, sizeof(unsigned) * (long long)words * batch));
  guard.track(d_bm);
  if (B > 16384) {
    SFFTB_CUDA(cudaMalloc(&d_k, sizeof(unsigned long long) * select_gkeys_per_row(B) * batch));
    guard.track(d_k);
  }
  SFFTB_CUDA(cudaMemcpy(d_x, h.data(), sizeof(cplx) * total, cudaMemcpyHostToDevice));
  SelectArgs sa;
  sa.xs = d_x; sa.xs_stride = 0; sa.row_stride = B; sa.logB = ilog2((unsigned)B); sa.num = num;
  sa.J = d_J; sa.J_sig_stride = 0; sa.bitmap = d_bm; sa.bm_sig_stride = 0;
  sa.gkeys = d_k; sa.gk_sig_stride = 0; sa.gk_scratch_off = total; sa.row_begin = 0; sa.row_step = 1;
  if (launch_select(sa, batch, 1, 0)) return -1;
  SFFTB_CUDA(cudaDeviceSynchronize());
  SFFTB_CUDA(cudaMemcpy(out_J, d_J, sizeof(int) * (long long)num * batch, cudaMemcpyDeviceToHost));
  cudaFree(d_x); cudaFree(d_J); cudaFree(d_bm); cudaFree(d_k);
  guard.dismiss();
  return 0;
}

int sfftb_debug_dft_any(const sfft_complex *in, sfft_complex *out, int n)
{
  if (sfftb_device_count() <= 0) { set_error("sfftb_debug_dft_any: no CUDA device"); return -1; }
  cplx *d_x = nullptr, *d_y = nullptr;
  ScratchGuard guard;
  SFFTB_CUDA(cudaMalloc(&d_x, sizeof(cplx) * n));
  guard.track(d_x);
  SFFTB_CUDA(cudaMalloc(&d_y, sizeof(cplx) * n));
  guard.track(d_y);
  SFFTB_CUDA(cudaMemcpy(d_x, in, sizeof(cplx) * n, cudaMemcpyHostToDevice));
  if (bluestein_forward(d_x, n, d_y, 0)) return -1;
  SFFTB_CUDA(cudaMemcpy(out, d_y, sizeof(cplx) * n, cudaMemcpyDeviceToHost));
  cudaFree(d_x); cudaFree(d_y);
  guard.dismiss();
  return 0;
}

long long sfftb_debug_div_check(unsigned long long seed, long long count)
{
  if (sfftb_device_count() <= 0) { set_error("sfftb_debug_div_check: no CUDA device"); return -1; }
  return run_div_check(seed, count);
}

int sfftb_enable_stage_timing(sfft_plan *plan, int on)
{
  PlanImpl *p = impl(plan);
  if (!p) { set_error("sfftb_enable_stage_timing: null plan"); return -1; }
  p->timer.enabled = on != 0;
  return 0;
}

int sfftb_stage_times(sfft_plan *plan, float *ms, const char **names, int capacity)
{
  PlanImpl *p = impl(plan);
  if (!p) { set_error("sfftb_stage_times: null plan"); return -1; }
  if (bind_device(p)) return -1;
  StageTimer &t = p->timer;
  if (!t.enabled || !t.created) return 0;
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  int cnt = t.count < capacity ? t.count : capacity;
  for (int i = 0; i < cnt; i++) {
    float v = 0;
    SFFTB_CUDA(cudaEventElapsedTime(&v, t.ev[i], t.ev[i + 1]));
    ms[i] = v;
    if (names) names[i] = t.names[i];
  }
  return cnt;
}

/* ---- multi-GPU loop sharding: see shard.cu ---- */

}  // extern "C"
