// plan_builder.cuh -- device-side window/filter construction (see plan_builder.cu).
#pragma once

#include "common.cuh"

namespace sfftb {

// What the transform keeps of a reference `Filter` (src/filters.h:27-32): the w
// time-domain taps and a window of the n-point response around frequency 0
// (only n/B+1 entries of filter.freq are ever read by v1/v2, cf12.cc:371-385;
// 3n/B by v3, computefourier-3.0.cc:375-379).
struct DeviceFilter {
  int w = 0;
  int fw_half = 0;       // fwin[m] = freq[(m - fw_half) mod n], m in [0, 2*fw_half]
  cplx *time = nullptr;  // [w]
  cplx *fwin = nullptr;  // [2*fw_half + 1]
  double2 *fdr = nullptr; // [2*fw_half + 1]: (|f|^2, RN(1/|f|^2)), refreshed by filter_refresh()
};

struct FilterSpec {
  double lobefrac, tolerance;
  int b;         // boxcar width in frequency bins
  int fw_half;   // half-width of the response window to keep
};

// w for (lobefrac, tolerance)   (src/filters.cc:72-74)
int filter_width(double lobefrac, double tolerance);

// Dolph-Chebyshev window of main-lobe half-width lobefrac widened by a boxcar of b
// frequency bins (src/filters.cc:70-86 then :109-160).
int build_filter(int logn, double lobefrac, double tolerance, int b, int fw_half, DeviceFilter *out,
                 cudaStream_t st);
// all (one or two) filters of a plan at once: shared windows are built once and the
// sequential chains of every filter run concurrently
int build_filters(int logn, int count, const FilterSpec *specs, DeviceFilter **outs, cudaStream_t st);
void free_filter(DeviceFilter *f);
// Plan cache: the next build_filters() call on this thread uploads these host arrays instead
// of running the builder (sizes must match what the specs derive).  Cleared by that call.
struct PresetFilters {
  int count = 0;
  int w[2] = {0, 0}, fw_half[2] = {0, 0};
  const cplx *time[2] = {nullptr, nullptr}, *fwin[2] = {nullptr, nullptr};
};
void set_preset_filters(const PresetFilters *preset);
// recompute the derived tables after fwin changed (plan build, sfftb_set_filter)
int filter_refresh(DeviceFilter *f, cudaStream_t st);

// forward DFT of arbitrary length (device in/out)
int bluestein_forward(const cplx *d_x, int w, cplx *d_out, cudaStream_t st);

}  // namespace sfftb
