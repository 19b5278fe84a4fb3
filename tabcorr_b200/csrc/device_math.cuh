// tabcorr_b200 -- table-driven FP64 erf / log / exp for the occupation phase.
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md).
#pragma once

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// table-driven double-precision math for the occupation phase
//
// The FP64 pipe is shared with DMMA, and the CUDA math library's erf/log/exp spend most of their
// issue slots on constant loads and range branches (ncu: 211 warp instructions per evaluation).
// The occupation functions only need ~1e-14 accuracy (parity bar: rtol 1e-10 on ngal, xi), so the
// kernel uses branch-free piecewise polynomials with coefficients in shared memory:
//   cen:  0.5 (1 + erf(x))  degree-13 polynomial on 25 intervals of width 0.5 covering [-6.25, 6.25]
//         (absolute error < 2e-15; exactly the 1e-16-level noise 1 + erf(x) has in the reference);
//   sat:  t^alpha = exp(alpha log t) with a 128-entry log table (degree-7 log1p) and a 32-entry
//         2^(j/32) table (degree-6 exp); relative error < 3e-14 over the reachable range.
// The tables are computed on the host in long double when the library first touches a device.
// (Tried: degree 7 on 193 intervals of width 1/16 -- same accuracy, 8 instead of 14 coefficient
// loads per evaluation.  Not faster: with finer intervals the lanes of a warp hit more distinct
// table columns, so every load costs more shared-memory wavefronts; standalone occupation kernel
// 0.536 vs 0.503 ms per 1e5 draws, fused kernel unchanged.  The small table also leaves room for
// wider draw tiles.  Also tried: high and low words of the coefficients in separate 32-word rows,
// two conflict-free LDS.32 instead of one conflicting LDS.64 -- bank conflicts 4x lower, time
// unchanged (0.490 ms): ncu shows the occupation code at 56 % issue, 54 % LSU, 35 % FP64 pipe
// utilisation with 77 warp instructions per 32 evaluations, bound by no single unit.)
// ------------------------------------------------------------------------------------------
constexpr int kErfDeg = 13;
constexpr int kErfIntervals = 27;                               // 25 polynomial + 2 saturated
constexpr int kErfStride = 32;                                  // doubles per coefficient row
constexpr int kErfDoubles = (kErfDeg + 1) * kErfStride;         // 448
constexpr int kLogEntries = 128;                                // (1 / c_i, ln c_i) pairs
constexpr int kExpEntries = 32;
constexpr int kTabLog = kErfDoubles;
constexpr int kTabExp = kTabLog + 2 * kLogEntries;
constexpr int kTabDoubles = kTabExp + kExpEntries;              // 736 doubles = 5888 bytes
constexpr double kRoundMagic = 6755399441055744.0;              // 2^52 + 2^51: round-to-nearest int

__device__ double g_math_tables[kTabDoubles];

// Wide erf table of the leauthaud11 kernel (leauthaud11.cuh): the same intervals of width 0.5
// (29 of them, centred at -7, -6.5, ..., 7, and two saturated columns), but every polynomial is a
// degree-19 fit over centre +- 0.75, so that the quadrature nodes a lane evaluates together -- a
// few hundredths apart in x -- share ONE column of coefficients: 20 coefficient loads per group
// of nodes instead of 14 per node (the kernel is bound by shared-memory wavefronts).  Absolute
// error < 2e-16 (rounding level).  Followed by a copy of the 2^(j/32) table.
constexpr int kErfWDeg = 19;
constexpr int kErfWPoly = 29;                                   // polynomial columns
constexpr int kErfWStride = 32;
constexpr double kErfWHalf = 0.75;                              // validity |x - centre| <= 0.75
constexpr int kErfWDoubles = (kErfWDeg + 1) * kErfWStride;      // 640
constexpr int kL11TabDoubles = kErfWDoubles + kExpEntries;      // 672 doubles = 5376 bytes
__device__ double g_l11_tables[kL11TabDoubles];

__device__ __forceinline__ void load_math_tables(double* tab) {
  for (int i = threadIdx.x; i < kTabDoubles; i += blockDim.x) tab[i] = g_math_tables[i];
}

// 0.5 (1 + erf(x)).  Interval i = rint(2 x + 12) is centred at x = -6 + i / 2; intervals below 0 /
// above 24 map to two extra table columns holding the constants 0 and 1, so the range clamp is
// two integer min/max instead of double-precision ones (7 instructions each).  The rounding trick
// needs |x| < 2^49: beyond |x| >= 16 (sigma_logM -> 0, infinities) the result is the step 0 / 1
// that erf gives there, selected on the integer pipe from the sign and exponent bits.
__device__ __forceinline__ double half_erfc_neg(double x, const double* __restrict__ tab) {
  const double v = fma(x, 2.0, 12.0 + kRoundMagic);
  const int i = __double2loint(v);
  const double t = fma(x, 2.0, 12.0 - (v - kRoundMagic));   // in [-0.5, 0.5]
  const double* c = tab + (min(max(i, -1), kErfIntervals - 2) + 1);
  double p = c[kErfDeg * kErfStride];
#pragma unroll
  for (int k = kErfDeg - 1; k >= 0; k--) p = fma(p, t, c[k * kErfStride]);
  const int hx = __double2hiint(x);
  const unsigned ax = hx & 0x7fffffff;
  const bool huge = ax >= 0x40300000u;                      // |x| >= 16, infinity, NaN
  const bool nan = ax > 0x7ff00000u || (ax == 0x7ff00000u && __double2loint(x) != 0);
  const int step_hi = nan ? 0x7ff80000 : hx < 0 ? 0 : 0x3ff00000;   // NaN, 0.0 or 1.0
  return huge ? __hiloint2double(step_hi, 0) : p;
}

// ln t for t > 0 (normal double): 128-entry table of (1 / c_i, ln c_i) + degree-7 log1p
__device__ __forceinline__ double log_pos(double t, const double* __restrict__ tab) {
  const int hi = __double2hiint(t);
  const int i = (hi >> 13) & (kLogEntries - 1);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(t));
  const double2 lc = *reinterpret_cast<const double2*>(tab + kTabLog + 2 * i);
  const double r = fma(m, lc.x, -1.0);       // |r| < 2^-8
  double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
  p = fma(p, r, 1.0 / 5.0);
  p = fma(p, r, -1.0 / 4.0);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  p = fma(p, r, 1.0);
  const double e = (double)((hi >> 20) - 1023);
  return fma(e, 0.6931471805599453094, fma(p, r, lc.y));
}

// e^y for |y| < 2^26: 32-entry 2^(j/32) table + degree-6 polynomial; |y| beyond ~690 saturates
// instead of overflowing
__device__ __forceinline__ double exp_scaled_with(double y, const double* __restrict__ exp_tab) {
  const double v = fma(y, 46.16624130844682903551 /* 32 / ln 2 */, kRoundMagic);
  const double kf = v - kRoundMagic;
  double q = fma(-kf, 0.0216608493924982895 /* hi(ln2 / 32) */, y);
  q = fma(-kf, 1.4168872360403518e-18 /* lo */, q);
  double w = fma(q, 1.0 / 720.0, 1.0 / 120.0);
  w = fma(w, q, 1.0 / 24.0);
  w = fma(w, q, 1.0 / 6.0);
  w = fma(w, q, 0.5);
  w = fma(w, q, 1.0);
  w = fma(w, q, 1.0);
  const int k = __double2loint(v);
  const double res = exp_tab[k & (kExpEntries - 1)] * w;   // in [1, 2) * (1 +- 0.011)
  const int scale = min(max(k >> 5, -1000), 1000);
  return __hiloint2double(__double2hiint(res) + (scale << 20), __double2loint(res));
}
__device__ __forceinline__ double exp_scaled(double y, const double* __restrict__ tab) {
  return exp_scaled_with(y, tab + kTabExp);
}

// 0.5 (1 + erf(x)) of U arguments that lie close together, through ONE column of the wide table
// (`wt`, kErfWDoubles doubles): the column is chosen by the midpoint of the first and the last
// argument; an argument further than kErfWHalf from the column's centre sends the lane through
// the one-column-per-argument path.  Arguments are clamped to [-16, 16] (the saturated columns
// give the exact 0 / 1 erf gives beyond; NaN arguments are the caller's business).
template <int U>
__device__ __forceinline__ void half_erfc_neg_group(double (&x)[U], const double* __restrict__ wt) {
  constexpr double kOffset = 14.0;   // column i is centred at -7 + i / 2
  // only the midpoint is clamped here: arguments near a clamped midpoint pass the distance test
  // below, all others (and NaN) take the per-argument path, which clamps each of them
  const double xm = fmin(fmax(0.5 * (x[0] + x[U - 1]), -16.0), 16.0);
  const double v = fma(xm, 2.0, kOffset + kRoundMagic);
  const double xc = fma(v - kRoundMagic, 0.5, -0.5 * kOffset);
  const double* c = wt + (min(max(__double2loint(v), -1), kErfWPoly) + 1);
  double t[U], p[U];
  bool near = true;
#pragma unroll
  for (int u = 0; u < U; u++) {
    t[u] = x[u] - xc;
    near = near && fabs(t[u]) <= kErfWHalf;
  }
  if (near) {
    const double top = c[kErfWDeg * kErfWStride];
#pragma unroll
    for (int u = 0; u < U; u++) p[u] = top;
#pragma unroll
    for (int k = kErfWDeg - 1; k >= 0; k--) {
      const double ck = c[k * kErfWStride];
#pragma unroll
      for (int u = 0; u < U; u++) p[u] = fma(p[u], t[u], ck);
    }
#pragma unroll
    for (int u = 0; u < U; u++) x[u] = p[u];
  } else {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const double xu = fmin(fmax(x[u], -16.0), 16.0);
      const double vu = fma(xu, 2.0, kOffset + kRoundMagic);
      const double tu = xu - fma(vu - kRoundMagic, 0.5, -0.5 * kOffset);   // in [-0.25, 0.25]
      const double* cu = wt + (min(max(__double2loint(vu), -1), kErfWPoly) + 1);
      double q = cu[kErfWDeg * kErfWStride];
#pragma unroll
      for (int k = kErfWDeg - 1; k >= 0; k--) q = fma(q, tu, cu[k * kErfWStride]);
      x[u] = q;
    }
  }
}

// t^alpha for t > 0 (normal double); |alpha ln t| beyond ~690 saturates instead of overflowing
__device__ __forceinline__ double pow_pos(double t, double alpha, const double* __restrict__ tab) {
  return exp_scaled(alpha * log_pos(t, tab), tab);
}

// element-wise evaluation of the table-driven math, for the accuracy tests (tc_debug_math)
// kind 2: the wide-table erf of the leauthaud11 kernel on the pair (x[i], y[i]); out[i] = value at
// x[i], and if out has room (kind 3) the value at y[i]
__global__ void debug_math_kernel(int kind, const double* x, const double* y, double* out,
                                  long long n) {
  __shared__ double tab[kTabDoubles];
  __shared__ double wide[kL11TabDoubles];
  load_math_tables(tab);
  for (int i = threadIdx.x; i < kL11TabDoubles; i += blockDim.x) wide[i] = g_l11_tables[i];
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (kind >= 2) {
      double pair[2] = {x[i], y[i]};
      half_erfc_neg_group<2>(pair, wide);
      out[i] = pair[kind - 2];
    } else {
      out[i] = kind == 0 ? half_erfc_neg(x[i], tab) : pow_pos(x[i], y[i], tab);
    }
  }
}

}  // namespace
