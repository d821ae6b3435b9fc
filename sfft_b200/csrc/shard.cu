// shard.cu -- multi-GPU sharding of one v1/v2 transform (include/sfft.h), one process per GPU.
//
// Two exchanges are offered for the one data-path step that crosses GPUs (making every
// rank's bucket spectra complete):
//
//  * peer exchange (sfftb_shard_export / _attach / _exec): every rank maps every other
//    rank's bucket-spectra buffer and a small flag block through CUDA IPC.  After its own
//    loops are bucketised, a rank STORES its rows straight into its peers' buffers over
//    NVLink / NVSwitch (plain 16-byte stores from one kernel, all peers at once) and raises
//    a flag in each peer; a one-block kernel on every rank waits for the flags of all
//    peers.  No NCCL call, no host synchronisation, no memset, and the whole transform --
//    exchange included -- replays from ONE CUDA graph per rank.  Reuse of the buffers by the
//    next transform is guarded by a second flag set ("done reading"), waited on just before
//    the stores, i.e. after the next gather: it never costs wall time in steady state.
//
//  * caller-side collective (sfftb_shard_bucketize / _spectra / _finish): the caller sums
//    the spectra buffer over ranks (torch.distributed all_reduce); kept as the portable
//    fallback and as the baseline the peer exchange is measured against.
#include "plan.cuh"

using namespace sfftb;

namespace sfftb {

namespace {

constexpr unsigned long long kSpinTimeoutNs = 4000000000ull;   // a dead peer must not hang the GPU

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= want; on timeout count an error and give up (results are then garbage
// but the stream drains; sfftb_shard_status reports it)
__device__ __forceinline__ void spin_until(const unsigned *p, unsigned want, unsigned *err)
{
  if ((int)(ld_acquire_sys(p) - want) >= 0) return;
  const unsigned long long t0 = globaltimer_ns();
  while ((int)(ld_acquire_sys(p) - want) < 0) {
    __nanosleep(64);
    if (globaltimer_ns() - t0 > kSpinTimeoutNs) {
      atomicAdd(err, 1u);
      return;
    }
  }
}

struct PushArgs {
  const cplx *src;
  cplx *dst[kMaxPeers];
  unsigned *peer_flags[kMaxPeers];
  unsigned *flags;
  long long off, cnt;       // elements this rank owns in the spectra buffer
  int rank, world;
};

// grid (blocks, world-1): block column q stores this rank's rows into peer (rank+1+q) % world.
// Before the first store it waits until that peer is done reading the previous transform's
// spectra; after the last store of the last block to finish, the peer's READY flag is raised.
__global__ void __launch_bounds__(256)
shard_push_kernel(PushArgs a)
{
  const int p = (a.rank + 1 + (int)blockIdx.y) % a.world;
  const unsigned epoch = a.flags[SH_EPOCH];
  if (threadIdx.x == 0) spin_until(a.flags + SH_DONE + p, epoch, a.flags + SH_ERR);
  __syncthreads();
  const double2 *__restrict__ src = a.src + a.off;
  double2 *dst = a.dst[p] + a.off;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.cnt;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(a.flags + SH_CTR + p, 1u);
    if (done == gridDim.x - 1) {
      a.flags[SH_CTR + p] = 0u;
      __threadfence_system();
      st_release_sys(a.peer_flags[p] + SH_READY + a.rank, epoch + 1u);
    }
  }
}

// one block: thread p waits until peer p has delivered its rows of this transform
__global__ void shard_wait_ready_kernel(unsigned *flags, int rank, int world)
{
  const int p = threadIdx.x;
  if (p < world && p != rank) spin_until(flags + SH_READY + p, flags[SH_EPOCH] + 1u, flags + SH_ERR);
}

struct DoneArgs {
  unsigned *peer_flags[kMaxPeers];
  unsigned *flags;
  int rank, world;
};
// one block, last kernel of a transform: tell every peer this rank no longer reads its
// spectra buffer, and advance the local epoch
__global__ void shard_done_kernel(DoneArgs a)
{
  const int p = threadIdx.x;
  const unsigned epoch = a.flags[SH_EPOCH];
  if (p < a.world && p != a.rank) st_release_sys(a.peer_flags[p] + SH_DONE + a.rank, epoch + 1u);
  __syncthreads();
  if (p == 0) a.flags[SH_EPOCH] = epoch + 1u;
}

}  // namespace

int launch_shard_push(const ShardPeers &sh, long long elem_off, long long elem_cnt, cudaStream_t st)
{
  if (sh.world <= 1 || elem_cnt <= 0) return 0;
  PushArgs a;
  a.src = sh.peer_xs[sh.rank];
  for (int p = 0; p < kMaxPeers; p++) { a.dst[p] = sh.peer_xs[p]; a.peer_flags[p] = sh.peer_flags[p]; }
  a.flags = sh.d_flags;
  a.off = elem_off; a.cnt = elem_cnt;
  a.rank = sh.rank; a.world = sh.world;
  // enough blocks per peer to fill the machine with 16-byte stores in flight, few enough
  // that the per-peer completion counter stays cheap
  long long blocks = (elem_cnt + 256 * 4 - 1) / (256 * 4);
  const long long cap = (148 * 4) / (sh.world - 1) > 8 ? (148 * 4) / (sh.world - 1) : 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  shard_push_kernel<<<dim3((unsigned)blocks, (unsigned)(sh.world - 1)), 256, 0, st>>>(a);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

int launch_shard_wait_ready(const ShardPeers &sh, cudaStream_t st)
{
  if (sh.world <= 1) return 0;
  shard_wait_ready_kernel<<<1, 32, 0, st>>>(sh.d_flags, sh.rank, sh.world);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

int launch_shard_done(const ShardPeers &sh, cudaStream_t st)
{
  if (sh.world <= 1) return 0;
  DoneArgs a;
  for (int p = 0; p < kMaxPeers; p++) a.peer_flags[p] = sh.peer_flags[p];
  a.flags = sh.d_flags;
  a.rank = sh.rank; a.world = sh.world;
  shard_done_kernel<<<1, 32, 0, st>>>(a);
  SFFTB_LAUNCH_CHECK();
  return 0;
}

}  // namespace sfftb

static PlanImpl *impl(const sfft_plan *plan) { return plan ? (PlanImpl *)plan->data : nullptr; }

static_assert(sizeof(cudaIpcMemHandle_t) == SFFTB_IPC_HANDLE_BYTES, "sfftb_peer_handle layout");

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

static int shard_result(PlanImpl *p, sfftb_result *result, int sync)
{
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

int sfftb_shard_finish(sfft_plan *plan, int rank, int world, sfftb_result *result, int sync)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || world < 1 || rank < 0 || rank >= world) {
    set_error("sfftb_shard_finish: bad argument");
    return -1;
  }
  SFFTB_CUDA(cudaSetDevice(p->device));
  if (v12_shard_finish(p, rank, world)) return -1;
  return shard_result(p, result, sync);
}

/* ---- NVLink peer exchange ---- */

int sfftb_shard_export(sfft_plan *plan, sfftb_peer_handle *mine)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || !mine) { set_error("sfftb_shard_export: bad argument (v3 does not shard)"); return -1; }
  SFFTB_CUDA(cudaSetDevice(p->device));
  PlanV12 &v = p->v12;
  if (v12_ensure_capacity(p, 1)) return -1;
  ShardPeers &sh = v.shard;
  if (!sh.d_flags) {
    SFFTB_CUDA(cudaMalloc(&sh.d_flags, sizeof(unsigned) * SH_WORDS));
    SFFTB_CUDA(cudaMemset(sh.d_flags, 0, sizeof(unsigned) * SH_WORDS));
  }
  memset(mine, 0, sizeof(*mine));
  SFFTB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)mine->spectra, v.d_xs));
  SFFTB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)mine->flags, sh.d_flags));
  return 0;
}

int sfftb_shard_attach(sfft_plan *plan, int rank, int world, const sfftb_peer_handle *all)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || (world > 1 && !all)) {
    set_error("sfftb_shard_attach: bad argument (at most 16 ranks; v3 does not shard)");
    return -1;
  }
  SFFTB_CUDA(cudaSetDevice(p->device));
  PlanV12 &v = p->v12;
  ShardPeers &sh = v.shard;
  if (sh.attached) { set_error("sfftb_shard_attach: already attached (detach first)"); return -1; }
  if (!sh.d_flags) {
    sfftb_peer_handle tmp;
    if (sfftb_shard_export(plan, &tmp)) return -1;
  }
  sh.rank = rank; sh.world = world;
  sh.n_opened = 0;
  for (int q = 0; q < kMaxPeers; q++) { sh.peer_xs[q] = nullptr; sh.peer_flags[q] = nullptr; }
  sh.peer_xs[rank] = v.d_xs;
  sh.peer_flags[rank] = sh.d_flags;
  for (int q = 0; q < world; q++) {
    if (q == rank) continue;
    void *xs = nullptr, *fl = nullptr;
    cudaIpcMemHandle_t hx, hf;
    memcpy(&hx, all[q].spectra, sizeof hx);
    memcpy(&hf, all[q].flags, sizeof hf);
    cudaError_t e = cudaIpcOpenMemHandle(&xs, hx, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) { sh.opened[sh.n_opened++] = xs; e = cudaIpcOpenMemHandle(&fl, hf, cudaIpcMemLazyEnablePeerAccess); }
    if (e != cudaSuccess) {
      set_error(std::string("sfftb_shard_attach: cannot map a peer's buffers through CUDA IPC: ") + cudaGetErrorString(e));
      cudaGetLastError();
      for (int i = 0; i < sh.n_opened; i++) cudaIpcCloseMemHandle(sh.opened[i]);
      sh.n_opened = 0;
      return -1;
    }
    sh.opened[sh.n_opened++] = fl;
    sh.peer_xs[q] = (cplx *)xs;
    sh.peer_flags[q] = (unsigned *)fl;
  }
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  SFFTB_CUDA(cudaMemset(sh.d_flags, 0, sizeof(unsigned) * SH_WORDS));
  sh.attached = true;
  sh.plain_execs = 0;
  return 0;
}

int sfftb_shard_detach(sfft_plan *plan)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3) { set_error("sfftb_shard_detach: bad argument"); return -1; }
  SFFTB_CUDA(cudaSetDevice(p->device));
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  v12_shard_release(p);
  return 0;
}

int sfftb_shard_exec(sfft_plan *plan, const void *d_in, const sfftb_draw *draw, sfftb_result *result, int sync)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || !d_in) { set_error("sfftb_shard_exec: bad argument (v3 does not shard)"); return -1; }
  if (!p->v12.shard.attached) { set_error("sfftb_shard_exec: call sfftb_shard_attach first"); return -1; }
  SFFTB_CUDA(cudaSetDevice(p->device));
  cudaGetLastError();
  sfftb_draw local;
  if (!draw) {
    // every rank draws from its own libc state: the caller seeds them identically
    if (sfftb_draw_random(plan, &local)) return -1;
    draw = &local;
  }
  if (v12_shard_exec(p, (const cplx *)d_in, draw)) return -1;
  return shard_result(p, result, sync);
}

int sfftb_shard_status(sfft_plan *plan, long long *epoch, long long *timeouts)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || !p->v12.shard.d_flags) { set_error("sfftb_shard_status: not a sharded plan"); return -1; }
  SFFTB_CUDA(cudaSetDevice(p->device));
  unsigned h[2] = {0, 0};
  SFFTB_CUDA(cudaMemcpyAsync(h, p->v12.shard.d_flags + SH_EPOCH, sizeof h, cudaMemcpyDeviceToHost, p->stream));
  SFFTB_CUDA(cudaStreamSynchronize(p->stream));
  if (epoch) *epoch = h[0];
  if (timeouts) *timeouts = h[1];
  return 0;
}

int sfftb_shard_slice(sfft_plan *plan, int rank, int world, long long *offset, long long *count)
{
  PlanImpl *p = impl(plan);
  if (!p || p->version == 3 || world < 1 || rank < 0 || rank >= world || !offset || !count) {
    set_error("sfftb_shard_slice: bad argument");
    return -1;
  }
  SFFTB_CUDA(cudaSetDevice(p->device));
  return v12_shard_slice(p, rank, world, offset, count);
}

}
