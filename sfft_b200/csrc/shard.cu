// shard.cu -- C ABI of the multi-GPU sharding of one v1/v2 transform (include/sfft.h).
// One process per GPU; the only exchange is a sum of the bucket-spectra buffer over
// ranks, done by the caller (torch.distributed all_reduce over NCCL / NVLink).
#include "plan.cuh"

using namespace sfftb;

static PlanImpl *impl(const sfft_plan *plan) { return plan ? (PlanImpl *)plan->data : nullptr; }

extern "C" {

int sfftb_shard_loops(const sfft_plan *plan, int rank, int world, int *begin, int *end)
{
  const PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || world < 1 || rank < 0 || rank >= world || !begin || !end) {
    set_error("sfftb_shard_loops: bad argument (v3 has no loops to shard)");
    return -1;
  }
  v12_shard_loops(p, rank, world, begin, end);
  return 0;
}

int sfftb_shard_bucketize(sfft_plan *plan, const void *d_in, const sfftb_draw *draw, int rank, int world)
{
  PlanImpl *p = impl(plan);
  if (!p || !d_in || !draw || world < 1 || rank < 0 || rank >= world) {
    set_error("sfftb_shard_bucketize: bad argument");
    return -1;
  }
  if (p->version == 3) { set_error("sfftb_shard_bucketize: v3 does not shard (replicas only)"); return -1; }
  SFFTB_CUDA(cudaSetDevice(p->device));
  cudaGetLastError();
  return v12_shard_bucketize(p, (const cplx *)d_in, draw, rank, world);
}

int sfftb_shard_spectra(sfft_plan *plan, void **d_spectra, long long *n_doubles)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || !d_spectra || !n_doubles) { set_error("sfftb_shard_spectra: bad argument"); return -1; }
  *d_spectra = p->v12.d_xs;
  *n_doubles = 2 * p->v12.x_samp_size;
  return 0;
}

int sfftb_shard_finish(sfft_plan *plan, int rank, int world, sfftb_result *result, int sync)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || world < 1 || rank < 0 || rank >= world) {
    set_error("sfftb_shard_finish: bad argument");
    return -1;
  }
  SFFTB_CUDA(cudaSetDevice(p->device));
  if (v12_shard_finish(p, rank, world)) return -1;
  PlanV12 &v = p->v12;
  const int *cnt = v.with_comb ? v.d_count : v.d_voted_count;
  if (result) {
    result->d_loc = v.d_hit_loc;
    result->d_val = (const sfft_complex *)v.d_hit_val;
    result->d_count = cnt;
    result->count = -1;
  }
  if (sync) {
    int h = 0;
    SFFTB_CUDA(cudaMemcpyAsync(&h, cnt, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    SFFTB_CUDA(cudaStreamSynchronize(p->stream));
    if (result) result->count = h > v.max_hits ? v.max_hits : h;
  }
  return 0;
}

}
