// v3.cu -- sFFT v3 (exact-sparse) on the device.  Placeholder until the kernels land.
#include "plan.cuh"

namespace sfftb {

struct PlanV3 { int dummy; };

int v3_build(PlanImpl *, int, int) { set_error("sFFT v3 is not built yet in this library"); return -1; }
void v3_free(PlanImpl *) {}
int v3_draw(const PlanImpl *, sfftb_draw *) { set_error("v3 not built"); return -1; }
int v3_exec(PlanImpl *, const cplx *, long long, int, const sfftb_draw *) { set_error("v3 not built"); return -1; }
int v3_info(const PlanImpl *, sfftb_info *) { set_error("v3 not built"); return -1; }
int v3_result(PlanImpl *, const int **, const cplx **, const int **, long long *) { set_error("v3 not built"); return -1; }
int v3_filter_sizes(const PlanImpl *, int, int *, int *) { return -1; }
DeviceFilter *v3_filter(PlanImpl *, int) { return nullptr; }
long long v3_debug_fetch(PlanImpl *, const char *, void *, size_t) { return -1; }

}  // namespace sfftb
