// shard.cu -- multi-GPU loop sharding of one v1/v2 transform (one process per GPU).
// Placeholder until the phase kernels land.
#include "plan.cuh"

using namespace sfftb;

extern "C" {

int sfftb_shard_phase1(sfft_plan *, const void *, const sfftb_draw *, int, int, int **, long long *)
{ set_error("loop sharding not built yet"); return -1; }
int sfftb_shard_phase2(sfft_plan *, int, int, long long *, double **, double **)
{ set_error("loop sharding not built yet"); return -1; }
int sfftb_shard_phase3(sfft_plan *, sfftb_result *, int)
{ set_error("loop sharding not built yet"); return -1; }

}
