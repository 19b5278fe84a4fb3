// tabcorr_b200 -- spline interpolation kernel (Interpolator) and the DMMA peak microbenchmark.
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md).
#pragma once

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// interpolation kernel (spline_interpolate for B draws)
// ------------------------------------------------------------------------------------------
constexpr int kMaxDims = 8;

struct InterpDev {
  int n_dims;
  int n_tables;
  int n_knots[kMaxDims];
  int knot_off[kMaxDims];    // offset of axis d in knots
  int a_off[kMaxDims];       // offset of axis d in a
  const double* knots;
  const double* a;
  const int* grid_to_table;  // [n_tables]
};

struct InterpArgs {
  InterpDev it;
  const double* x;      // [B, n_dims]
  long long n_draws;
  const double* data;   // [B, T, n_cols]
  int n_cols;
  double* out;          // [B, n_cols]
  int extrapolate;
  int* flag;
  int sum_knots;
};

__global__ void __launch_bounds__(128) interp_kernel(const InterpArgs args) {
  extern __shared__ double ism[];
  const InterpDev& it = args.it;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* wd = ism + (size_t)warp * (args.sum_knots + it.n_tables);  // per-axis knot weights
  double* wt = wd + args.sum_knots;                                    // per-table weights
  const long long draw = (long long)blockIdx.x * 4 + warp;
  if (draw >= args.n_draws) return;
  bool outside = false;
  for (int d = 0; d < it.n_dims; d++) {
    const int nk = it.n_knots[d];
    const double* xp = it.knots + it.knot_off[d];
    const double x = args.x[draw * it.n_dims + d];
    int seg = -1;
    for (int k = 0; k < nk; k++) seg += xp[k] <= x ? 1 : 0;  // digitize(x, xp) - 1
    if (x == xp[nk - 1]) seg = nk - 2;
    if (seg < 0 || seg >= nk - 1 || !(x == x)) {
      outside = true;
      seg = min(max(seg, 0), nk - 2);
    }
    const double* a = it.a + it.a_off[d] + (size_t)seg * 4 * nk;
    const double x2 = x * x, x3 = x2 * x;
    for (int k = lane; k < nk; k += 32)
      wd[it.knot_off[d] + k] = a[k] + a[nk + k] * x + a[2 * nk + k] * x2 + a[3 * nk + k] * x3;
  }
  __syncwarp();
  for (int gpos = lane; gpos < it.n_tables; gpos += 32) {
    int rem = gpos;
    double w = 1.0;
    for (int d = it.n_dims - 1; d >= 0; d--) {
      const int k = rem % it.n_knots[d];
      rem /= it.n_knots[d];
      w *= wd[it.knot_off[d] + k];
    }
    wt[it.grid_to_table[gpos]] = w;
  }
  __syncwarp();
  const bool bad = outside && !args.extrapolate;
  if (bad && lane == 0) atomicOr(args.flag, 1);
  const double* data = args.data + (size_t)draw * it.n_tables * args.n_cols;
  for (int c = lane; c < args.n_cols; c += 32) {
    double s = 0.0;
    for (int t = 0; t < it.n_tables; t++) s = fma(wt[t], data[(size_t)t * args.n_cols + c], s);
    args.out[draw * args.n_cols + c] = bad ? CUDART_NAN : s;
  }
}

// ------------------------------------------------------------------------------------------
// DMMA peak microbenchmark (roofline denominator)
// ------------------------------------------------------------------------------------------
__global__ void dmma_peak_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - 1e-9 * threadIdx.x;
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma884(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 ALU (DFMA) peak: 8 independent chains per thread.  The roofline denominator of the paths
// bound by the occupation arithmetic (cross tables, mean_occupation_batch).
__global__ void dfma_peak_kernel(double* out, int iters) {
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9 * threadIdx.x;
  double c[8];
#pragma unroll
  for (int i = 0; i < 8; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace
