// tabcorr_b200 -- sm_100a kernels and C ABI for TabCorr's prediction hot path.
//
// What is computed (reference johannesulf/TabCorr v1.2.0, tabcorr/tabcorr.py:465-683): for each of
// B parameter draws, the Gauss-Legendre averaged mean occupation of every tabulated halo bin
// (:537-578), the tracer weights w = occ * n_h (:623), ngal = sum w and, per radial bin r, the
// quadratic form xi_r = w^T M_r w / ngal^2 over the symmetric tracer-pair table (:641-647) or the
// linear form M_r . w / ngal for cross-correlation tables (:648-649); optionally split by galaxy
// type (:652-683).  Interpolator.predict (tabcorr/interpolator.py:124-216) runs this for every table
// of a parameter grid and applies a tensor-product cubic spline (:275-331).
//
// How it is mapped to B200 (see DESIGN.md for the full account):
//  * FP64 has no tcgen05/UMMA kind; the FP64 tensor path on sm_100a is the warp-level DMMA
//    (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4).  Measured: 37.05 TFLOP/s per B200, and DFMA shares
//    the same pipe (tools/fp64_peaks.cu), so the occupation arithmetic competes with the
//    contraction for issue slots -- the kernel is built to minimise FP64 instructions outside DMMA.
//  * The draws are the GEMM "n" dimension: one CTA owns a tile of 8*NT draws whose weights W live
//    in shared memory in DMMA B-fragment order for the whole tile.  The table is the "A" operand:
//    at load time M_r is rewritten as a lower-triangular matrix with doubled off-diagonal terms
//    (exactly the reference's packed prefactor-2 sum) and re-tiled into a DMMA A-fragment stream,
//    so that each warp streams its tiles from L2 with one coalesced 16-byte load per lane and
//    k-step, with no shared-memory staging and no block-level synchronisation in the main loop.
//    Only the lower triangle is multiplied: half the flops of the dense form.
//  * Work inside a CTA is a list of chunks (radial bin, range of 16-row tiles) that the 12 warps
//    take dynamically; each chunk ends in a register row-dot against W and a fixed-order shuffle
//    reduction, and writes its partial sums to a scratch slot, so results are bitwise
//    reproducible whatever the schedule or the number of GPUs.
//
// This file holds no CPU implementation of the path: without a CUDA device every entry point
// that computes returns TC_ECUDA.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tabcorr_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------
thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define TC_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t err__ = (expr);                                                            \
    if (err__ != cudaSuccess) {                                                            \
      (void)cudaGetLastError();                                                            \
      return fail(TC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));        \
    }                                                                                      \
  } while (0)

// ------------------------------------------------------------------------------------------
// constants shared by host and device
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 384;        // 12 warps, 3 per SM sub-partition, <= 170 registers each
constexpr int kWarps = kThreads / 32;
constexpr int kOccUnroll = 5;        // quadrature nodes in flight per lane (n_gauss_prim = 10 default)
constexpr int kGroupRows = 2;        // rows (secondary-percentile bins) sharing one mass bin
constexpr int kSmemLimit = 227 * 1024;

// One unit of contraction work.  Auto mode: radial bin `r`, 16-row tiles [mt0, mt1); for tile mt
// the k-steps (4 table rows each) [k_begin, min(4 (mt + 1), k_cap)) are multiplied.  Cross mode:
// 16-radial-bin tile `r`, k-steps [k_begin, k_cap).  `part_row` is where the result goes.
struct Chunk {
  int r, mt0, mt1, k_begin, k_cap, part_row, pad0, pad1;
};

struct OccPlan {       // device pointers, one per (layout, n_gauss)
  int n_groups;
  int n_cen_groups;         // groups are ordered centrals first
  int n_gauss;
  int n_gauss_pad;          // n_gauss rounded up to a multiple of `unroll`; padding nodes have zero weight
  int unroll;               // nodes evaluated per iteration: kOccUnroll when it divides n_gauss, else 2
  int zero_row;             // index of an all-zero row of row_c (second row of 1-row groups)
  const double* node_logm;  // [n_groups, G]  log10 of the node masses
  const double* node_m;     // [n_groups, G]  node masses
  const double* node_inv_m; // [n_groups, G]  1 / node mass
  const int* grp_rows;      // [n_groups, kGroupRows] padded row index or -1
  const int* grp_is_sat;    // [n_groups]
  const double* row_c;      // [n_pad, G] normalised quadrature weights
  const double* row_nh;     // [n_pad]
  const double* row_pct;    // [n_pad] secondary-property percentile of the row
};

struct LayoutDev {
  int n_rows;          // N of the table
  int n_pad;           // padded rows, multiple of 16
  int nc_pad;          // first satellite row in padded order
  int n_parts;         // scratch rows per draw tile
  int n_chunks;
  int n_out;           // outputs per draw: Reff * n_comp
  long long ks_per_r;  // k-steps per radial bin (auto) / per 16-bin tile (cross) in the A stream
  const double2* afrag;
  const float4* afrag32;   // 3xTF32 mode: per k8-step a 32-lane block of high parts, then of low parts
  long long ks8_per_r;     // k8-steps per radial bin in afrag32
  const Chunk* chunks;
  const long long* chunk_cost_prefix;  // [n_chunks + 1] cumulative cost of the sorted chunks
  const int* out_ptr;    // [n_out + 1] CSR: which scratch rows sum to output o
  const int* out_parts;
  const int* pad_to_row;  // [n_pad] reference row index or -1
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// internal kernel mode: auto-correlation table contracted in 3xTF32 (tc_predict_batch precision 1)
constexpr int kModeAutoTf32 = 2;

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float4& a, float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
        "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// d = a * b with a fresh (zero) accumulator
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const float4& a, float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
        "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.0f));
}

__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// 3xTF32 mode: index (in floats) of the HIGH part of (padded row i, draw b) in the shared W tile;
// the low part is two floats further.  m16n8k8 B-fragment order: per k8-step and n-tile the lane
// holding B[k = i % 4 (+ 4)][n = b % 8] reads one float4 {hi(k), hi(k + 4), lo(k), lo(k + 4)}.
// The tile has the same size as the FP64 one (8 bytes per weight).
template <int NT>
__device__ __forceinline__ int widx32(int i, int b) {
  return (((((i >> 3) * NT + (b >> 3)) << 5) + ((b & 7) << 2) + (i & 3)) << 2) + ((i >> 2) & 1);
}

// index of (padded row i, draw b) in the shared W tile: DMMA B-fragment order, so that the lane
// holding B[k = i % 4][n = b % 8] of k-step i / 4 and n-tile b / 8 reads consecutive doubles.
template <int NT>
__device__ __forceinline__ int widx(int i, int b) {
  return (((i >> 2) * NT + (b >> 3)) << 5) + ((b & 7) << 2) + (i & 3);
}

// store / load one tracer weight of the shared W tile in the representation of the mode
template <int NT, int MODE>
__device__ __forceinline__ void store_weight(double* Ws, int row, int b, double w) {
  if (MODE == kModeAutoTf32) {
    float* wf = reinterpret_cast<float*>(Ws) + widx32<NT>(row, b);
    const float hi = to_tf32((float)w);
    wf[0] = hi;
    wf[2] = to_tf32((float)(w - (double)hi));
  } else {
    Ws[widx<NT>(row, b)] = w;
  }
}
template <int NT, int MODE>
__device__ __forceinline__ double load_weight(const double* Ws, int row, int b) {
  if (MODE == kModeAutoTf32) {
    const float* wf = reinterpret_cast<const float*>(Ws) + widx32<NT>(row, b);
    return (double)wf[0] + (double)wf[2];
  }
  return Ws[widx<NT>(row, b)];
}


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

__device__ __forceinline__ void load_math_tables(double* tab) {
  for (int i = threadIdx.x; i < kTabDoubles; i += blockDim.x) tab[i] = g_math_tables[i];
}

// 0.5 (1 + erf(x)) for |x| < 2^49.  Interval i = rint(2 x + 12) is centred at x = -6 + i / 2;
// intervals below 0 / above 24 map to two extra table columns holding the constants 0 and 1, so
// the range clamp is two integer min/max instead of double-precision ones (7 instructions each).
__device__ __forceinline__ double half_erfc_neg(double x, const double* __restrict__ tab) {
  const double v = fma(x, 2.0, 12.0 + kRoundMagic);
  const int i = __double2loint(v);
  const double t = fma(x, 2.0, 12.0 - (v - kRoundMagic));   // in [-0.5, 0.5]
  const double* c = tab + (min(max(i, -1), kErfIntervals - 2) + 1);
  double p = c[kErfDeg * kErfStride];
#pragma unroll
  for (int k = kErfDeg - 1; k >= 0; k--) p = fma(p, t, c[k * kErfStride]);
  return p;
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
__device__ __forceinline__ double exp_scaled(double y, const double* __restrict__ tab) {
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
  const double res = tab[kTabExp + (k & (kExpEntries - 1))] * w;   // in [1, 2) * (1 +- 0.011)
  const int scale = min(max(k >> 5, -1000), 1000);
  return __hiloint2double(__double2hiint(res) + (scale << 20), __double2loint(res));
}

// t^alpha for t > 0 (normal double); |alpha ln t| beyond ~690 saturates instead of overflowing
__device__ __forceinline__ double pow_pos(double t, double alpha, const double* __restrict__ tab) {
  return exp_scaled(alpha * log_pos(t, tab), tab);
}

struct DrawParams {
  double logMmin, inv_sigma, m0, inv_m1, alpha, a_cen, a_sat;
  // node arrays handed to baseline_occupation as (first, second); zheng07 satellites need the
  // mass only, unless they are modulated by the central occupation (log10 mass, mass)
  static __device__ __forceinline__ const double* first_nodes(const OccPlan& plan, bool sat,
                                                              bool modulate) {
    return sat && !modulate ? plan.node_m : plan.node_logm;
  }
  static __device__ __forceinline__ const double* second_nodes(const OccPlan& plan) {
    return plan.node_m;
  }
  static __device__ __forceinline__ bool needs_second(bool sat, bool modulate) {
    return sat && modulate;
  }
  __device__ __forceinline__ void begin_group(double) {}
};

// theta points at the draw's first parameter; consecutive parameters are `ps` doubles apart
// (1 for the [B, TC_N_THETA] layout, the leading dimension for the [TC_N_THETA, ld] layout)
__device__ __forceinline__ DrawParams load_draw(const double* __restrict__ theta, long long ps) {
  DrawParams p;
  p.logMmin = theta[0];
  p.inv_sigma = 1.0 / theta[ps];
  p.m0 = exp10(theta[2 * ps]);
  p.inv_m1 = 1.0 / exp10(theta[3 * ps]);
  p.alpha = theta[4 * ps];
  p.a_cen = fmin(fmax(theta[5 * ps], -1.0), 1.0);
  p.a_sat = fmin(fmax(theta[6 * ps], -1.0), 1.0);
  return p;
}

// Heaviside assembly bias (halotools HeavisideAssembias, call site tabcorr.py:556-563): haloes above
// the split percentile get +delta, the others -delta (1 - s) / s; delta is the strength A times the
// largest perturbation that keeps both sub-populations inside [lo, hi]:
//   A > 0:  delta = A min(hi - f, s / (1 - s) (f - lo))
//   A <= 0: delta = -A max(lo - f, s / (1 - s) (f - hi)) = A min(f - lo, s / (1 - s) (hi - f))
// Returns delta (0 where the baseline is on a bound or the split is degenerate); lo = 0.
__device__ __forceinline__ double assembias_delta(double f, double strength, double ratio,
                                                  double hi, bool split_ok) {
  const double up = hi - f, down = f;
  const bool positive = strength > 0.0;
  const double p = positive ? up : down;
  const double q = ratio * (positive ? down : up);
  const double m = p < q ? p : q;
  const bool inside = split_ok && f > 0.0 && f < hi;
  return inside ? strength * m : 0.0;
}

// Baseline occupation of one quadrature node.
template <bool SAT, bool MODULATE>
__device__ __forceinline__ double baseline_occupation(double logm, double mass,
                                                      const DrawParams& p,
                                                      const double* __restrict__ tab) {
  if (!SAT) {
    // Zheng07Cens: 0.5 (1 + erf((log10 M - logMmin) / sigma_logM))
    return half_erfc_neg((logm - p.logMmin) * p.inv_sigma, tab);
  }
  // Zheng07Sats: ((M - M0) / M1)^alpha for M > M0, else 0
  const double d = mass - p.m0;
  const bool pos = d > 0.0;
  double f = pow_pos(pos ? d * p.inv_m1 : 1.0, p.alpha, tab);
  f = pos ? f : 0.0;
  if (MODULATE) f *= half_erfc_neg((logm - p.logMmin) * p.inv_sigma, tab);
  return f;
}

// One mass-bin group: the baseline occupation at each quadrature node is evaluated once (U nodes
// per iteration as independent dependency chains -- beside DMMA warps a dependent DFMA gets an
// issue turn only every ~24-32 cycles, tools/fp64_mix.cu; the plan pads G to a multiple of U with
// zero-weight nodes) and accumulated into the two rows (secondary-percentile bins) of the group.
// A group with a single row points its second row at an all-zero weight row.
template <bool SAT, bool DECORATED, bool MODULATE, int U, typename Params>
__device__ __forceinline__ void occupation_group(const OccPlan& plan, int grp, Params& p,
                                                 double split, const double* __restrict__ tab,
                                                 double& occ0, double& occ1) {
  const int G = plan.n_gauss_pad;
  const int row0 = plan.grp_rows[grp * kGroupRows], row1 = plan.grp_rows[grp * kGroupRows + 1];
  const double* c0 = plan.row_c + (size_t)row0 * G;
  const double* c1 = plan.row_c + (size_t)(row1 >= 0 ? row1 : plan.zero_row) * G;
  const double* node = Params::first_nodes(plan, SAT, MODULATE) + (size_t)grp * G;
  const double* node2 = Params::second_nodes(plan) + (size_t)grp * G;
  p.begin_group(node[0]);
  double k0 = 0.0, k1 = 0.0, ratio = 0.0;
  bool split_ok = false;
  if (DECORATED) {
    split_ok = split > 0.0 && split < 1.0;
    ratio = split / (1.0 - split);
    const double down = -(1.0 - split) / split;
    k0 = plan.row_pct[row0] > split ? 1.0 : down;
    k1 = (row1 >= 0 && plan.row_pct[row1] > split) ? 1.0 : down;
  }
  const double hi = SAT ? CUDART_INF : 1.0;
  const double strength = SAT ? p.a_sat : p.a_cen;
  double a0 = 0.0, a1 = 0.0;
  for (int g = 0; g < G; g += U) {
    double f[U];
#pragma unroll
    for (int u = 0; u < U; u++)
      f[u] = baseline_occupation<SAT, MODULATE>(
          node[g + u], Params::needs_second(SAT, MODULATE) ? node2[g + u] : node[g + u], p, tab);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (DECORATED) {
        const double d = assembias_delta(f[u], strength, ratio, hi, split_ok);
        a0 = fma(c0[g + u], fma(k0, d, f[u]), a0);
        a1 = fma(c1[g + u], fma(k1, d, f[u]), a1);
      } else {
        a0 = fma(c0[g + u], f[u], a0);
        a1 = fma(c1[g + u], f[u], a1);
      }
    }
  }
  occ0 = a0;
  occ1 = a1;
}

// Occupation work item of one warp: the 8 draws of one n-tile (lane & 7) times the mass-bin groups
// [g_begin, g_end), four groups in flight per warp (lane >> 3).  A range never mixes centrals and
// satellites (groups are ordered centrals first), so the galaxy type is warp-uniform.
// store(padded_row, occ, n_h) receives the Gauss-Legendre averaged occupation of each row.
template <bool DECORATED, bool MODULATE, int U, typename Store>
__device__ __forceinline__ void occupation_item_impl(const OccPlan& plan, const tc_model& model,
                                                     const double* __restrict__ theta_row,
                                                     long long theta_ps, int g_begin, int g_end,
                                                     const double* __restrict__ tab, Store store) {
  DrawParams p = load_draw(theta_row, theta_ps);
  if (!model.decorated) p.a_cen = p.a_sat = 0.0;  // strengths are ignored unless decorated
  const bool sat = g_begin >= plan.n_cen_groups;
  for (int grp = g_begin + (threadIdx.x >> 3 & 3); grp < g_end; grp += 4) {
    double occ0, occ1;
    if (sat)
      occupation_group<true, DECORATED, MODULATE, U>(plan, grp, p, model.split, tab, occ0, occ1);
    else
      occupation_group<false, DECORATED, false, U>(plan, grp, p, model.split, tab, occ0, occ1);
    const int row0 = plan.grp_rows[grp * kGroupRows], row1 = plan.grp_rows[grp * kGroupRows + 1];
    store(row0, occ0, plan.row_nh[row0]);
    if (row1 >= 0) store(row1, occ1, plan.row_nh[row1]);
  }
}

template <typename Store>
__device__ __forceinline__ void occupation_item(const OccPlan& plan, const tc_model& model,
                                                const double* __restrict__ theta_row,
                                                long long theta_ps, int g_begin, int g_end,
                                                const double* __restrict__ tab, Store store) {
  if (model.modulate_with_cenocc) {   // rare: keep one generic instantiation
    occupation_item_impl<true, true, 2>(plan, model, theta_row, theta_ps, g_begin, g_end, tab, store);
  } else if (plan.unroll == 2 * kOccUnroll) {
    if (model.decorated)
      occupation_item_impl<true, false, 2 * kOccUnroll>(plan, model, theta_row, theta_ps, g_begin, g_end, tab, store);
    else
      occupation_item_impl<false, false, 2 * kOccUnroll>(plan, model, theta_row, theta_ps, g_begin, g_end, tab, store);
  } else if (plan.unroll == kOccUnroll) {
    if (model.decorated)
      occupation_item_impl<true, false, kOccUnroll>(plan, model, theta_row, theta_ps, g_begin, g_end, tab, store);
    else
      occupation_item_impl<false, false, kOccUnroll>(plan, model, theta_row, theta_ps, g_begin, g_end, tab, store);
  } else if (model.decorated) {
    occupation_item_impl<true, false, 2>(plan, model, theta_row, theta_ps, g_begin, g_end, tab, store);
  } else {
    occupation_item_impl<false, false, 2>(plan, model, theta_row, theta_ps, g_begin, g_end, tab, store);
  }
}

// Group range q of n_ranges = n_ranges_cen + n_ranges_sat: each galaxy type's groups are cut into
// pieces whose length is a multiple of 4 (the groups a warp evaluates at a time).
__device__ __forceinline__ void occupation_range(const OccPlan& plan, int n_ranges_cen,
                                                 int n_ranges_sat, int q, int& g_begin,
                                                 int& g_end) {
  const bool sat = q >= n_ranges_cen;
  const int first = sat ? plan.n_cen_groups : 0;
  const int count = sat ? plan.n_groups - plan.n_cen_groups : plan.n_cen_groups;
  const int pieces = sat ? n_ranges_sat : n_ranges_cen;
  const int piece = sat ? q - n_ranges_cen : q;
  const int units = (count + 3) >> 2;
  g_begin = first + min(count, 4 * (int)((long long)units * piece / pieces));
  g_end = first + min(count, 4 * (int)((long long)units * (piece + 1) / pieces));
}

// ------------------------------------------------------------------------------------------
// pipeline flags in shared memory: monotonically increasing counters, so a waiter can never be
// lapped (a parity-based mbarrier can: a warp that only ran occupation items of a tile may meet
// that tile's barrier one or two phases later)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void flag_wait(const int* counter, int target) {
  const volatile int* c = counter;
  while (*c < target) __nanosleep(40);
  __threadfence_block();   // acquire: order the W reads / writes that follow after the flag read
}
// all lanes call it after their last shared-memory access of the item
__device__ __forceinline__ void flag_signal(int* counter, int lane) {
  __threadfence_block();   // release: this lane's W accesses before the flag update
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicAdd(counter, 1);
  }
}

// ------------------------------------------------------------------------------------------
// fused occupation + contraction kernel
// ------------------------------------------------------------------------------------------
struct PredictArgs {
  LayoutDev lay;
  OccPlan plan;
  tc_model model;
  const double* theta;   // parameter draws or nullptr: theta[draw * theta_ds + k * theta_ps]
  long long theta_ds, theta_ps;
  const double* occ;     // [B, n_rows] or nullptr
  int theta_is_inline;   // one draw whose parameters travel in the launch arguments
  double theta_inline[TC_N_THETA];
  long long n_draws;
  long long n_tiles;
  double* parts;         // [n_tiles, n_parts, BM]
  double* ngal_tile;     // [n_tiles, 2, BM]  centrals / satellites number density
  int n_buf;             // W tiles in shared memory: 2 = occupation of tile t + 1 overlaps tile t
  int n_ranges_cen;      // occupation items per n-tile: group ranges of centrals ...
  int n_ranges_sat;      // ... and of satellites
  int occ_stride;        // every occ_stride-th slot of a tile's work list is an occupation item
  int tf32_segment;      // 3xTF32 mode: k8-steps per FP32 accumulation chain
};

struct PredictCtrl {
  int full[2];                  // occupation items finished, per W buffer (n_occ per tile)
  int empty[2];                 // warps that left a tile's work list, per W buffer (kWarps per tile)
  int next;                     // work-list cursor
  int first_lo, last_hi;        // chunk range of the CTA's first / last tile
  int n_local;                  // tiles this CTA works on
  long long tile_first;
  double theta_inline[TC_N_THETA];   // shared-memory copy of the inline parameters
};

// One contraction chunk by one warp.  W is the draw tile in B-fragment order.
template <int NT, int MODE>
__device__ __forceinline__ void run_chunk(const LayoutDev& lay, const Chunk& ch,
                                          const double* __restrict__ Ws,
                                          double* __restrict__ parts, int lane) {
  constexpr int BM = 8 * NT;
  const int g = lane >> 2, tig = lane & 3;
  if (MODE == TC_MODE_AUTO) {
    double sums[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) sums[nt][0] = sums[nt][1] = 0.0;
    for (int mt = ch.mt0; mt < ch.mt1; mt++) {
      const int k_tile = 4 * (mt + 1);                 // k-steps of the full lower-triangular tile
      const int k_end = min(k_tile, ch.k_cap);
      // the upper 8 rows of the tile are zero in its last two k-steps: skip their DMMAs
      const int k_both = min(k_end, k_tile - 2);
      const double2* ap = lay.afrag +
          ((size_t)ch.r * lay.ks_per_r + 2 * (size_t)mt * (mt + 1) + ch.k_begin) * 32 + lane;
      const double* wk = Ws + (size_t)ch.k_begin * NT * 32 + lane;
      double acc[2][NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; nt++)
        acc[0][nt][0] = acc[0][nt][1] = acc[1][nt][0] = acc[1][nt][1] = 0.0;
      double2 a_next = ld_stream(ap);
      int ks = ch.k_begin;
      for (; ks < k_both; ks++) {
        const double2 a = a_next;
        ap += 32;
        a_next = ld_stream(ap);  // the stream is padded by one k-step, always safe
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          const double b = wk[nt * 32];
          dmma884(acc[0][nt], a.x, b);
          dmma884(acc[1][nt], a.y, b);
        }
        wk += NT * 32;
      }
      for (; ks < k_end; ks++) {
        const double2 a = a_next;
        ap += 32;
        a_next = ld_stream(ap);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[1][nt], a.y, wk[nt * 32]);
        wk += NT * 32;
      }
      // row-dot: acc[h][nt][e] = (M' W)[row = 16 mt + 8 h + g][draw = 8 nt + 2 tig + e]
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int row = 16 * mt + 8 * h + g;
        const double* wr = Ws + (size_t)(row >> 2) * NT * 32 + (row & 3) + tig * 8;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          sums[nt][0] = fma(acc[h][nt][0], wr[nt * 32], sums[nt][0]);
          sums[nt][1] = fma(acc[h][nt][1], wr[nt * 32 + 4], sums[nt][1]);
        }
      }
    }
    // fixed-order butterfly over the 8 row groups of the warp (lane xor 16, 8, 4): every lane ends
    // with the full sums; the lanes of row group 0 store them
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        double v = sums[nt][e];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        sums[nt][e] = v;
      }
    }
    if (g == 0) {
#pragma unroll
      for (int nt = 0; nt < NT; nt++)
        *reinterpret_cast<double2*>(parts + (size_t)ch.part_row * BM + 8 * nt + 2 * tig) =
            make_double2(sums[nt][0], sums[nt][1]);
    }
  } else {
    // cross mode: a 16-radial-bin tile times a k-range of W; the product is the output
    const double2* ap = lay.afrag + ((size_t)ch.r * lay.ks_per_r + ch.k_begin) * 32 + lane;
    const double* wk = Ws + (size_t)ch.k_begin * NT * 32 + lane;
    double acc[2][NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
      acc[0][nt][0] = acc[0][nt][1] = acc[1][nt][0] = acc[1][nt][1] = 0.0;
    double2 a_next = ld_stream(ap);
    for (int ks = ch.k_begin; ks < ch.k_cap; ks++) {
      const double2 a = a_next;
      ap += 32;
      a_next = ld_stream(ap);
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const double b = wk[nt * 32];
        dmma884(acc[0][nt], a.x, b);
        dmma884(acc[1][nt], a.y, b);
      }
      wk += NT * 32;
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++)
        *reinterpret_cast<double2*>(parts + (size_t)(ch.part_row + 8 * h + g) * BM + 8 * nt +
                                    2 * tig) = make_double2(acc[h][nt][0], acc[h][nt][1]);
  }
}

// One contraction chunk in 3xTF32: every table entry and every weight is split into a TF32 high
// part and a TF32 low part (22 significant bits together); hi*hi + lo*hi + hi*lo are accumulated in
// FP32 by the warp-level m16n8k8 MMA (k8-steps of 8 table columns, 16-row tiles as in the FP64
// path), the row-dot and everything after it stay in FP64.  Relative error ~1e-7 of the sum of the
// term magnitudes (tests: 1e-6).
constexpr int kTf32Segment = 1 << 20;   // k8-steps per FP32 running sum (default: the whole tile)

template <int NT>
__device__ __forceinline__ void run_chunk_tf32(const LayoutDev& lay, const Chunk& ch,
                                               const double* __restrict__ Ws,
                                               double* __restrict__ parts, int lane,
                                               int segment) {
  constexpr int BM = 8 * NT;
  const int g = lane >> 2, tig = lane & 3;
  const float* Wf = reinterpret_cast<const float*>(Ws);
  const int k_begin = ch.k_begin >> 1, k_cap = ch.k_cap >> 1;   // k4-steps -> k8-steps
  double sums[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) sums[nt][0] = sums[nt][1] = 0.0;
  for (int mt = ch.mt0; mt < ch.mt1; mt++) {
    const int k_end = min(2 * (mt + 1), k_cap);
    const float4* ap = lay.afrag32 +
        (((size_t)ch.r * lay.ks8_per_r + (size_t)mt * (mt + 1) + k_begin) * 2) * 32 + lane;
    const float4* wk = reinterpret_cast<const float4*>(Wf) + (size_t)k_begin * NT * 32 + lane;
    float4 hi_next = ld_stream4(ap), lo_next = ld_stream4(ap + 32);
    int ks = k_begin;
    while (ks < k_end) {
      // optional: cut the FP32 running sum every `segment` k8-steps (row-dot into the FP64 sums)
      const int seg_end = min(ks + segment, k_end);
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
      for (; ks < seg_end; ks++) {
        const float4 a_hi = hi_next, a_lo = lo_next;
        ap += 64;
        hi_next = ld_stream4(ap);        // the stream is padded by one k8-step, always safe
        lo_next = ld_stream4(ap + 32);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          const float4 w = wk[nt * 32];  // {hi(k), hi(k + 4), lo(k), lo(k + 4)}
          // The tensor core truncates its FP32 accumulator after every MMA -- up to one ulp of the
          // RUNNING sum, always towards zero: chained over a 16-row tile's 30 k8-steps that is a
          // bias of 1e-6 (measured, tools/tf32_error.py).  So each k8-step starts from a zero
          // accumulator (its truncations are relative to the small increment) and is added to the
          // running sum with round-to-nearest FADDs, which are unbiased and nearly free.
          float d[4];
          mma_tf32_zero(d, a_lo, w.x, w.y);   // small terms first
          mma_tf32(d, a_hi, w.z, w.w);
          mma_tf32(d, a_hi, w.x, w.y);
#pragma unroll
          for (int j = 0; j < 4; j++) acc[nt][j] += d[j];
        }
        wk += NT * 32;
      }
      // row-dot: acc[nt][2 h + e] = (M' W)[row = 16 mt + 8 h + g][draw = 8 nt + 2 tig + e]
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int row = 16 * mt + 8 * h + g;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const float* wf = Wf + widx32<NT>(row, 8 * nt + 2 * tig + e);
            const float a = acc[nt][2 * h + e];
            sums[nt][e] += (double)fmaf(a, wf[2], a * wf[0]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      double v = sums[nt][e];
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      sums[nt][e] = v;
    }
  }
  if (g == 0) {
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
      *reinterpret_cast<double2*>(parts + (size_t)ch.part_row * BM + 8 * nt + 2 * tig) =
          make_double2(sums[nt][0], sums[nt][1]);
  }
}

// The kernel is a barrier-free software pipeline over the CTA's draw tiles.  Work is a sequence of
// per-tile lists of S slots that the 12 warps take from one shared cursor:
//   slot 0                          number densities of tile j (one warp, sequential row order)
//   every occ_stride-th next slot   occupation item (n-tile, group range) of tile j + n_buf - 1,
//                                   written into the other W buffer
//   the remaining slots             contraction chunks of tile j, longest first
// Dependencies always point backwards in that sequence, so taking slots in order cannot deadlock:
// a chunk waits until full[buf] counts all occupation items of its tile, an occupation item until
// empty[buf] counts every warp having left the list of the tile that used its buffer before.
template <int NT, int MODE>
__global__ void __launch_bounds__(kThreads, 1) predict_kernel(const PredictArgs args) {
  constexpr int BM = 8 * NT;
  extern __shared__ __align__(16) double smem[];
  const LayoutDev& lay = args.lay;
  const int n_buf = args.n_buf;
  const size_t tile_doubles = (size_t)lay.n_pad * BM;
  double* tab = smem + n_buf * tile_doubles;                               // math tables
  PredictCtrl* ctrl = reinterpret_cast<PredictCtrl*>(tab + kTabDoubles);
  const int tid = threadIdx.x, lane = tid & 31;

  for (size_t i = tid; i < n_buf * tile_doubles; i += kThreads) smem[i] = 0.0;  // padding rows stay 0
  load_math_tables(tab);

  const int n_occ = NT * (args.n_ranges_cen + args.n_ranges_sat);
  {
    // The grid cuts the total COST (n_tiles x per-tile chunk cost) into equal contiguous ranges, so
    // that every CTA gets the same amount of DMMA work whatever the number of draws; a CTA
    // recomputes the weights of the (at most two) tiles it shares with its neighbours.
    const long long tile_cost = lay.chunk_cost_prefix[lay.n_chunks];
    const long long total_cost = tile_cost * args.n_tiles;
    const long long cost_lo = total_cost / gridDim.x * blockIdx.x +
                              total_cost % gridDim.x * blockIdx.x / gridDim.x;
    const long long cost_hi = total_cost / gridDim.x * (blockIdx.x + 1) +
                              total_cost % gridDim.x * (blockIdx.x + 1) / gridDim.x;
    long long tile_first = cost_lo / tile_cost;
    long long tile_last = min((cost_hi + tile_cost - 1) / tile_cost, args.n_tiles);  // exclusive
    // first chunk whose start cost is >= the range boundary = number of prefix entries below it
    // (neighbouring CTAs count against the same value, so their chunk ranges meet exactly).  The
    // whole CTA counts in parallel: one round trip to L2 instead of a dependent binary search, which
    // is what a one-draw call waits for.
    const long long bound_lo = cost_lo - tile_first * tile_cost;
    const long long bound_hi = cost_hi - (tile_last - 1) * tile_cost;
    int below_lo = 0, below_hi = 0;
    for (int base = 0; base < lay.n_chunks; base += kThreads) {
      const int i = base + tid;
      const long long start = i < lay.n_chunks ? lay.chunk_cost_prefix[i] : tile_cost;
      below_lo += __syncthreads_count(i < lay.n_chunks && start < bound_lo);
      below_hi += __syncthreads_count(i < lay.n_chunks && start < bound_hi);
    }
    if (tid == 0) {
      int first_lo = bound_lo <= 0 ? 0 : bound_lo >= tile_cost ? lay.n_chunks : below_lo;
      if (first_lo >= lay.n_chunks) { tile_first++; first_lo = 0; }
      int last_hi = lay.n_chunks;
      if (tile_last > tile_first) {
        last_hi = bound_hi <= 0 ? 0 : bound_hi >= tile_cost ? lay.n_chunks : below_hi;
        if (last_hi <= (tile_last - 1 == tile_first ? first_lo : 0)) { tile_last--; last_hi = lay.n_chunks; }
      }
      ctrl->tile_first = tile_first;
      ctrl->n_local = (int)max(tile_last - tile_first, 0LL);
      ctrl->first_lo = first_lo;
      ctrl->last_hi = last_hi;
      ctrl->next = 0;
      ctrl->full[0] = ctrl->full[1] = ctrl->empty[0] = ctrl->empty[1] = 0;
    }
    if (args.theta_is_inline && tid < TC_N_THETA) ctrl->theta_inline[tid] = args.theta_inline[tid];
  }
  __syncthreads();
  // parameters of the draws: device (or mapped host) memory, or the launch arguments of a
  // one-draw call (a mapped-host read costs every CTA a PCIe round trip: 14 us per call)
  const double* theta_base = args.theta_is_inline ? ctrl->theta_inline : args.theta;
  const int n_local = ctrl->n_local;
  const long long tile_first = ctrl->tile_first;
  const int first_lo = ctrl->first_lo, last_hi = ctrl->last_hi;
  const int occ_ahead = n_buf - 1;
  const int S = 1 + lay.n_chunks + n_occ;
  const int stride = args.occ_stride;

  int left = -occ_ahead;   // lists [.., left) have been left behind by this warp
  int full_seen = -1;      // newest tile whose W this warp has seen complete
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(&ctrl->next, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    const int list = i / S - occ_ahead;          // tile (local index) whose list the slot is in
    const int s = i - (list + occ_ahead) * S;
    // this warp has finished everything it took from earlier lists: release those tiles
    const int upto = min(list, n_local);
    for (int t = max(left, 0); t < upto; t++) flag_signal(&ctrl->empty[t % n_buf], lane);
    left = max(left, upto);
    if (list >= n_local) break;

    int kind = 0, idx = 0;                       // 0 ngal, 1 occupation, 2 chunk
    if (s > 0) {
      const int u = s - 1, q = u / stride;
      if (u - q * stride == stride - 1 && q < n_occ) { kind = 1; idx = q; }
      else { kind = 2; idx = u - min(n_occ, q); }
    }

    if (kind == 1) {
      // ---- occupation item idx of tile list + occ_ahead -> W[(list + occ_ahead) % n_buf] ------
      const int j = list + occ_ahead;
      if (j >= n_local) continue;
      const int buf = j % n_buf;
      if (j >= n_buf) flag_wait(&ctrl->empty[buf], (j / n_buf) * kWarps);
      double* Ws = smem + buf * tile_doubles;
      const int nt = idx % NT, q = idx / NT;
      const int b = 8 * nt + (lane & 7);
      long long draw = (tile_first + j) * BM + b;
      if (draw >= args.n_draws) draw = args.n_draws - 1;  // tail tile: recompute the last draw
      if (theta_base != nullptr) {
        int g_begin, g_end;
        occupation_range(args.plan, args.n_ranges_cen, args.n_ranges_sat, q, g_begin, g_end);
        occupation_item(args.plan, args.model, theta_base + draw * args.theta_ds, args.theta_ps,
                        g_begin, g_end, tab,
                        [&](int row, double occ, double nh) {
                          store_weight<NT, MODE>(Ws, row, b, occ * nh);
                        });
      } else {
        const int n_q = args.n_ranges_cen + args.n_ranges_sat;
        const int r_begin = (int)((long long)lay.n_pad * q / n_q);
        const int r_end = (int)((long long)lay.n_pad * (q + 1) / n_q);
        for (int row = r_begin + (lane >> 3); row < r_end; row += 4) {
          const int src = lay.pad_to_row[row];
          if (src >= 0)
            store_weight<NT, MODE>(Ws, row, b,
                                   args.occ[draw * lay.n_rows + src] * args.plan.row_nh[row]);
        }
      }
      flag_signal(&ctrl->full[buf], lane);
      continue;
    }

    if (list < 0) continue;                      // prologue lists hold occupation items only
    const int c_lo = list == 0 ? first_lo : 0;
    const int c_hi = list == n_local - 1 ? last_hi : lay.n_chunks;
    if (kind == 0 ? c_lo != 0 : (idx < c_lo || idx >= c_hi)) continue;
    const int buf = list % n_buf;
    if (full_seen < list) {
      flag_wait(&ctrl->full[buf], (list / n_buf + 1) * n_occ);
      full_seen = list;
    }
    const double* Ws = smem + buf * tile_doubles;
    const long long tile = tile_first + list;

    if (kind == 0) {
      // ---- number densities (by the CTA that owns the tile's first chunk) ---------------------
      for (int b = lane; b < BM; b += 32) {
        double nc = 0.0, ns = 0.0;
        for (int r = 0; r < lay.nc_pad; r++) nc += load_weight<NT, MODE>(Ws, r, b);
        for (int r = lay.nc_pad; r < lay.n_pad; r++) ns += load_weight<NT, MODE>(Ws, r, b);
        args.ngal_tile[(tile * 2 + 0) * BM + b] = nc;
        args.ngal_tile[(tile * 2 + 1) * BM + b] = ns;
      }
    } else {
      const Chunk ch = lay.chunks[idx];
      if constexpr (MODE == kModeAutoTf32)
        run_chunk_tf32<NT>(lay, ch, Ws, args.parts + (size_t)tile * lay.n_parts * BM, lane,
                           args.tf32_segment);
      else
        run_chunk<NT, MODE>(lay, ch, Ws, args.parts + (size_t)tile * lay.n_parts * BM, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------
// finalize: sum the scratch rows of every output in fixed order and normalise by ngal
// ------------------------------------------------------------------------------------------
struct FinalizeArgs {
  LayoutDev lay;
  const double* parts;
  const double* ngal_tile;
  long long n_draws;
  int bm;
  int mode;
  int separate;
  int n_tables;
  double* ngal_out;
  long long ngal_stride;
  double* xi_out;
  long long xi_stride;
};

__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeArgs args) {
  const int bm = args.bm;
  const long long tile = blockIdx.x;
  const int b = threadIdx.x % bm;
  const long long draw = tile * bm + b;
  if (draw >= args.n_draws) return;
  const double nc = args.ngal_tile[(tile * 2 + 0) * bm + b];
  const double ns = args.ngal_tile[(tile * 2 + 1) * bm + b];
  const double ngal = nc + ns;
  const double norm = args.mode == TC_MODE_AUTO ? ngal * ngal : ngal;
  const int o_step = (blockDim.x / bm) * gridDim.y;
  const int o_first = threadIdx.x / bm + (blockDim.x / bm) * blockIdx.y;
  if (o_first == 0) {
    for (int t = 0; t < args.n_tables; t++) {
      if (args.separate) {
        args.ngal_out[draw * args.ngal_stride + 2 * t + 0] = nc;
        args.ngal_out[draw * args.ngal_stride + 2 * t + 1] = ns;
      } else {
        args.ngal_out[draw * args.ngal_stride + t] = ngal;
      }
    }
  }
  const double* parts = args.parts + (size_t)tile * args.lay.n_parts * bm + b;
  for (int o = o_first; o < args.lay.n_out; o += o_step) {
    double s = 0.0;
    for (int j = args.lay.out_ptr[o]; j < args.lay.out_ptr[o + 1]; j++)
      s += parts[(size_t)args.lay.out_parts[j] * bm];
    args.xi_out[draw * args.xi_stride + o] = s / norm;
  }
}

// ------------------------------------------------------------------------------------------
// standalone occupation kernel (TabCorr.mean_occupation)
// ------------------------------------------------------------------------------------------
struct OccArgs {
  OccPlan plan;
  tc_model model;
  const double* theta;
  long long theta_ds, theta_ps;
  long long n_draws;
  int n_rows;
  int n_ranges_cen, n_ranges_sat;
  const int* pad_to_row;
  double* occ_out;
};

__global__ void __launch_bounds__(kThreads, 1) occupation_kernel(const OccArgs args) {
  __shared__ double tab[kTabDoubles];
  load_math_tables(tab);
  __syncthreads();
  // one warp per item = 8 draws x one group range; four groups in flight per warp
  const int lane = threadIdx.x & 31;
  const int n_ranges = args.n_ranges_cen + args.n_ranges_sat;
  const long long n_blocks = (args.n_draws + 7) / 8;
  const long long n_items = n_blocks * n_ranges;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  for (long long item = warp0; item < n_items; item += (long long)gridDim.x * kWarps) {
    const long long block = item / n_ranges;
    const int q = (int)(item - block * n_ranges);
    const long long draw = block * 8 + (lane & 7);
    const bool live = draw < args.n_draws;
    int g_begin, g_end;
    occupation_range(args.plan, args.n_ranges_cen, args.n_ranges_sat, q, g_begin, g_end);
    occupation_item(args.plan, args.model,
                    args.theta + (live ? draw : args.n_draws - 1) * args.theta_ds, args.theta_ps,
                    g_begin, g_end, tab,
                    [&](int row, double occ, double) {
                      const int dst = args.pad_to_row[row];
                      if (live && dst >= 0) args.occ_out[draw * args.n_rows + dst] = occ;
                    });
  }
}

// ------------------------------------------------------------------------------------------
// family 1: Leauthaud11 occupations (halotools Leauthaud11Cens / Leauthaud11Sats over the
// Behroozi10SmHm stellar-to-halo-mass relation; call sites tabcorr.py:556-563)
//
//   <N_cen>(M) = 0.5 (1 - erf((log10 M*_thr - log10 M*(M)) / (sqrt(2) sigma_logM*)))
//   <N_sat>(M) = exp(-M_cut / (M h)) (M h / M_sat)^alphasat  [x <N_cen>(M) if modulate_with_cenocc]
//   M_sat = 1e12 bsat (M_knee / 1e12)^betasat,  M_cut = 1e12 bcut (M_knee / 1e12)^betacut,
//   M_knee = h 10^(log10 M_h(M*_thr)),  h = 0.7
// log10 M_h(log10 M*) is Behroozi et al. (2010) eq. 21 with parameters x_0 + x_a (a - 1) at the
// model redshift.  halotools inverts it numerically: it tabulates log10 M_h on the 100 knots
// log10 M* = linspace(8.5, 12.5, 100) and evaluates the interpolating cubic spline (scipy
// InterpolatedUnivariateSpline, k = 3: not-a-knot end conditions, cubic extrapolation) of log10 M*
// over log10 M_h.  Parity means reproducing that spline, not the exact inverse, so every draw
// builds the same table and solves the same not-a-knot system (Thomas algorithm) in shared
// memory; a mass bin is then one 7-step binary search, and a node a short walk and one cubic.  A draw whose table is not
// strictly increasing (halotools raises there) gets NaN occupations.
//
// This family does not run inside the fused kernel (its per-draw spline does not fit beside the W
// tiles): occupation_l11_kernel writes occ[B, N] and the contraction runs on the occupation
// input of predict_kernel.  Restated from memory of halotools -- parity unpinned, see DESIGN.md.
// ------------------------------------------------------------------------------------------
constexpr int kL11Knots = 100;
constexpr int kL11DrawsPerBlock = 64;     // 64 x 3.2 KB of spline tables + math tables < 227 KB
constexpr double kL11LittleH = 0.7;
constexpr double kL11LogMsLo = 8.5, kL11LogMsHi = 12.5;
constexpr double kLn10 = 2.302585092994045684;

struct L11Draw {
  // knot k: x = log10 M_h of the knot; y, z, w = c1, c3, c2 of the cubic on [knot k, knot k + 1):
  // log10 M* = s_k + t (c1 + t (c2 + t c3)), t = log10 M - x
  double4 knot[kL11Knots];
  double inv_scatter;     // 1 / (sqrt(2) sigma)
  double neg_mcut_h;      // -M_cut / h
  double ln_h_over_msat;  // ln(h / M_sat)
  double alphasat;
  double a_cen, a_sat;    // assembly-bias strengths (0 unless decorated)
  double bad;             // NaN if the table is not strictly increasing, else 0
  double pad[3];          // sizeof = 25 x 128 + 80 bytes: the same knot of 8 consecutive draws
                          // falls into 8 different 16-byte bank groups
};
static_assert(sizeof(L11Draw) % 128 == 80, "L11Draw stride chosen against bank conflicts");

struct L11Params {
  const L11Draw* d;
  double threshold;
  double a_cen, a_sat;
  int hint;   // knot interval of the group's first node: the nodes of a mass bin ascend from it
  double next_x[3];   // abscissae of the next three knots (+inf past the last interval)
  // baseline_occupation receives (log10 mass, 1 / mass)
  static __device__ __forceinline__ const double* first_nodes(const OccPlan& plan, bool, bool) {
    return plan.node_logm;
  }
  static __device__ __forceinline__ const double* second_nodes(const OccPlan& plan) {
    return plan.node_inv_m;
  }
  static __device__ __forceinline__ bool needs_second(bool sat, bool) { return sat; }
  __device__ __forceinline__ void begin_group(double logm) {
    int i = 0;   // largest knot index in [0, kL11Knots - 2] with x_i <= logm (0 if none)
#pragma unroll
    for (int step = 64; step >= 1; step >>= 1) {
      const int j = i + step;
      if (j <= kL11Knots - 2 && d->knot[j].x <= logm) i = j;
    }
    hint = i;
#pragma unroll
    for (int j = 0; j < 3; j++)
      next_x[j] = i + 1 + j <= kL11Knots - 2 ? d->knot[i + 1 + j].x : CUDART_INF;
  }
};

__device__ __forceinline__ double l11_knot_logms(int k) {
  // numpy.linspace(8.5, 12.5, 100): arange(100) * step + start, last element set to stop
  return k == kL11Knots - 1 ? kL11LogMsHi
                            : (double)k * ((kL11LogMsHi - kL11LogMsLo) / (kL11Knots - 1)) + kL11LogMsLo;
}

// Behroozi10SmHm.mean_log_halo_mass: log10 M_h [h = 1 units] of log10 M* [h = 1 units].  halotools
// converts M* -> M* / h^2 and M_h -> M_h h through 10** and log10; here the conversions are added
// in log space (differences at the 1e-16 level).
__device__ __forceinline__ double l11_log_halo_mass(double log_ms, double logm0, double logm1,
                                                    double beta, double delta, double gamma) {
  const double log_h = -0.15490195998574316929;   // log10(0.7)
  const double lr = log_ms - 2.0 * log_h - logm0;  // log10(M* / M0) in h = 0.7 units
  return logm1 + beta * lr + exp10(delta * lr) / (1.0 + exp10(-gamma * lr)) - 0.5 + log_h;
}

__device__ __forceinline__ double l11_log_mstar(double logm, const L11Params& p) {
  // the nodes of a group ascend from the hinted interval and a mass bin spans few knots: three
  // branch-free steps against the cached abscissae, then (rarely) a linear walk
  int i = p.hint + (p.next_x[0] <= logm ? 1 : 0) + (p.next_x[1] <= logm ? 1 : 0) +
          (p.next_x[2] <= logm ? 1 : 0);
  if (p.next_x[2] <= logm)
    while (i < kL11Knots - 2 && p.d->knot[i + 1].x <= logm) i++;
  const double4 c = p.d->knot[i];
  const double t = logm - c.x;
  return fma(t, fma(t, fma(t, c.z, c.w), c.y), l11_knot_logms(i));
}

template <bool SAT, bool MODULATE>
__device__ __forceinline__ double baseline_occupation(double logm, double inv_mass,
                                                      const L11Params& p,
                                                      const double* __restrict__ tab) {
  if (!SAT) {
    const double x = (l11_log_mstar(logm, p) - p.threshold) * p.d->inv_scatter;
    return half_erfc_neg(x, tab) + p.d->bad;
  }
  // exp(-M_cut / (M h)) (M h / M_sat)^alphasat = exp(alphasat (ln M + ln(h / M_sat)) - M_cut / (M h))
  double y = fma(p.d->alphasat, fma(logm, kLn10, p.d->ln_h_over_msat), p.d->neg_mcut_h * inv_mass);
  y = fmin(fmax(y, -800.0), 800.0);
  double f = exp_scaled(y, tab);
  if (MODULATE) {
    const double x = (l11_log_mstar(logm, p) - p.threshold) * p.d->inv_scatter;
    f *= half_erfc_neg(x, tab);
  }
  return f + p.d->bad;
}

// Spline tables and per-draw constants of one block of draws, by the whole CTA.
__device__ void l11_prepare_block(L11Draw* draws, int n_block, long long draw0,
                                  long long n_draws, const double* __restrict__ theta,
                                  long long theta_ds, long long theta_ps, const tc_model& model) {
  const double a1 = 1.0 / (1.0 + model.redshift) - 1.0;   // a - 1
  // (1) knot abscissae and per-draw constants
  for (int idx = threadIdx.x; idx < n_block * (kL11Knots + 1); idx += blockDim.x) {
    const int b = idx / (kL11Knots + 1), k = idx - b * (kL11Knots + 1);
    const long long draw = min(draw0 + b, n_draws - 1);
    const double* th = theta + draw * theta_ds;
    const double logm0 = fma(th[1 * theta_ps], a1, th[0]);
    const double logm1 = fma(th[3 * theta_ps], a1, th[2 * theta_ps]);
    const double beta = fma(th[5 * theta_ps], a1, th[4 * theta_ps]);
    const double delta = fma(th[7 * theta_ps], a1, th[6 * theta_ps]);
    const double gamma = fma(th[9 * theta_ps], a1, th[8 * theta_ps]);
    if (k < kL11Knots) {
      draws[b].knot[k].x = l11_log_halo_mass(l11_knot_logms(k), logm0, logm1, beta, delta, gamma);
    } else {
      // Leauthaud11Sats._update_satellite_params: knee = h M_h(threshold) / 1e12
      const double log_knee = l11_log_halo_mass(model.threshold, logm0, logm1, beta, delta, gamma) +
                              log10(kL11LittleH) - 12.0;
      const double msat = 1e12 * th[12 * theta_ps] * exp10(th[15 * theta_ps] * log_knee);
      const double mcut = 1e12 * th[13 * theta_ps] * exp10(th[14 * theta_ps] * log_knee);
      draws[b].inv_scatter = 1.0 / (1.4142135623730951 * th[10 * theta_ps]);
      draws[b].neg_mcut_h = -mcut / kL11LittleH;
      draws[b].ln_h_over_msat = log(kL11LittleH / msat);
      draws[b].alphasat = th[11 * theta_ps];
      draws[b].a_cen = model.decorated ? fmin(fmax(th[16 * theta_ps], -1.0), 1.0) : 0.0;
      draws[b].a_sat = model.decorated ? fmin(fmax(th[17 * theta_ps], -1.0), 1.0) : 0.0;
    }
  }
  __syncthreads();
  // (2) second derivatives m_k of the not-a-knot spline s(x): one thread per draw (Thomas).
  // interior equations h_{i-1} m_{i-1} + 2 (h_{i-1} + h_i) m_i + h_i m_{i+1} = 6 (d_i - d_{i-1}),
  // d_i = (s_{i+1} - s_i) / h_i, with m_0 and m_{n-1} eliminated through the continuity of the
  // third derivative at knots 1 and n - 2.  Scratch: knot[i].y = modified upper diagonal,
  // knot[i].z = modified right-hand side, knot[i].w = m_i.  The draws are spread over the warps
  // (one lane group per SM sub-partition) because the recurrences are latency bound.
  {
    const int lanes = (n_block + kWarps - 1) / kWarps;              // draws per warp
    const int b = (threadIdx.x >> 5) * lanes + (threadIdx.x & 31);
    if ((threadIdx.x & 31) < lanes && b < n_block) {
      L11Draw& D = draws[b];
      constexpr int n = kL11Knots;
      double h_prev = D.knot[1].x - D.knot[0].x;            // h_0
      bool increasing = h_prev > 0.0;
      double d_prev = (l11_knot_logms(1) - l11_knot_logms(0)) / h_prev;
      double cp = 0.0, rp = 0.0;                            // c'_{i-1}, r'_{i-1}
      for (int i = 1; i <= n - 2; i++) {
        const double h = D.knot[i + 1].x - D.knot[i].x;     // h_i
        increasing = increasing && h > 0.0;
        const double d = (l11_knot_logms(i + 1) - l11_knot_logms(i)) / h;
        double lower = h_prev, diag = 2.0 * (h_prev + h), upper = h;
        if (i == 1) {
          lower = 0.0;
          diag = 3.0 * h_prev + 2.0 * h + h_prev * h_prev / h;
          upper = h - h_prev * h_prev / h;
        }
        if (i == n - 2) {
          lower = h_prev - h * h / h_prev;
          diag = 2.0 * h_prev + 3.0 * h + h * h / h_prev;
          upper = 0.0;
        }
        const double rhs = 6.0 * (d - d_prev);
        const double inv = 1.0 / (diag - lower * cp);
        cp = upper * inv;
        rp = (rhs - lower * rp) * inv;
        D.knot[i].y = cp;
        D.knot[i].z = rp;
        h_prev = h;
        d_prev = d;
      }
      double m_next = 0.0;
      for (int i = n - 2; i >= 1; i--) {
        const double m = D.knot[i].z - D.knot[i].y * m_next;
        D.knot[i].w = m;
        m_next = m;
      }
      const double h0 = D.knot[1].x - D.knot[0].x, h1 = D.knot[2].x - D.knot[1].x;
      D.knot[0].w = D.knot[1].w - h0 / h1 * (D.knot[2].w - D.knot[1].w);
      const double ha = D.knot[n - 1].x - D.knot[n - 2].x, hb = D.knot[n - 2].x - D.knot[n - 3].x;
      D.knot[n - 1].w = D.knot[n - 2].w + ha / hb * (D.knot[n - 2].w - D.knot[n - 3].w);
      D.bad = increasing ? 0.0 : CUDART_NAN;
    }
  }
  __syncthreads();
  // (3) cubic coefficients per interval: c1 and c3 go to the (now dead) y and z slots -- only x
  // and w (= m) of knots k, k + 1 are read in this pass -- then w becomes c2 = m / 2
  for (int idx = threadIdx.x; idx < n_block * (kL11Knots - 1); idx += blockDim.x) {
    const int b = idx / (kL11Knots - 1), k = idx - b * (kL11Knots - 1);
    const double x0 = draws[b].knot[k].x, x1 = draws[b].knot[k + 1].x;
    const double m0 = draws[b].knot[k].w, m1 = draws[b].knot[k + 1].w;
    const double h = x1 - x0;
    const double d = (l11_knot_logms(k + 1) - l11_knot_logms(k)) / h;
    draws[b].knot[k].y = d - h * (2.0 * m0 + m1) * (1.0 / 6.0);
    draws[b].knot[k].z = (m1 - m0) / (6.0 * h);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n_block * (kL11Knots - 1); idx += blockDim.x) {
    const int b = idx / (kL11Knots - 1), k = idx - b * (kL11Knots - 1);
    draws[b].knot[k].w *= 0.5;
  }
  __syncthreads();
}

template <bool DECORATED, bool MODULATE, int U, typename Store>
__device__ __forceinline__ void occupation_item_l11(const OccPlan& plan, const tc_model& model,
                                                    const L11Draw* d, int g_begin, int g_end,
                                                    const double* __restrict__ tab, Store store) {
  L11Params p;
  p.d = d;
  p.threshold = model.threshold;
  p.a_cen = d->a_cen;
  p.a_sat = d->a_sat;
  p.hint = 0;
  const bool sat = g_begin >= plan.n_cen_groups;
  for (int grp = g_begin + (threadIdx.x >> 3 & 3); grp < g_end; grp += 4) {
    double occ0, occ1;
    if (sat)
      occupation_group<true, DECORATED, MODULATE, U>(plan, grp, p, model.split, tab, occ0, occ1);
    else
      occupation_group<false, DECORATED, false, U>(plan, grp, p, model.split, tab, occ0, occ1);
    const int row0 = plan.grp_rows[grp * kGroupRows], row1 = plan.grp_rows[grp * kGroupRows + 1];
    store(row0, occ0, plan.row_nh[row0]);
    if (row1 >= 0) store(row1, occ1, plan.row_nh[row1]);
  }
}

__global__ void __launch_bounds__(kThreads, 1) occupation_l11_kernel(const OccArgs args) {
  extern __shared__ __align__(16) double l11_smem[];
  double* tab = l11_smem;
  L11Draw* draws = reinterpret_cast<L11Draw*>(l11_smem + kTabDoubles);
  load_math_tables(tab);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_ranges = args.n_ranges_cen + args.n_ranges_sat;
  const long long n_blocks = (args.n_draws + kL11DrawsPerBlock - 1) / kL11DrawsPerBlock;
  for (long long block = blockIdx.x; block < n_blocks; block += gridDim.x) {
    const long long draw0 = block * kL11DrawsPerBlock;
    const int n_block = (int)min((long long)kL11DrawsPerBlock, args.n_draws - draw0);
    __syncthreads();   // the previous block's tables are no longer read
    l11_prepare_block(draws, n_block, draw0, args.n_draws, args.theta, args.theta_ds,
                      args.theta_ps, args.model);
    // one warp per item = 8 draws x one group range; four groups in flight per warp
    const int n_items = (kL11DrawsPerBlock / 8) * n_ranges;
    for (int item = warp; item < n_items; item += kWarps) {
      const int sub = item / n_ranges, q = item - sub * n_ranges;
      const int b = min(sub * 8 + (lane & 7), n_block - 1);
      const long long draw = draw0 + sub * 8 + (lane & 7);
      const bool live = draw < args.n_draws;
      int g_begin, g_end;
      occupation_range(args.plan, args.n_ranges_cen, args.n_ranges_sat, q, g_begin, g_end);
      auto store = [&](int row, double occ, double) {
        const int dst = args.pad_to_row[row];
        if (live && dst >= 0) args.occ_out[draw * args.n_rows + dst] = occ;
      };
      const bool u5 = args.plan.unroll == kOccUnroll;
      const bool mod = args.model.modulate_with_cenocc != 0;
      if (args.model.decorated) {
        if (mod) { if (u5) occupation_item_l11<true, true, kOccUnroll>(args.plan, args.model, draws + b, g_begin, g_end, tab, store);
                   else occupation_item_l11<true, true, 2>(args.plan, args.model, draws + b, g_begin, g_end, tab, store); }
        else     { if (u5) occupation_item_l11<true, false, kOccUnroll>(args.plan, args.model, draws + b, g_begin, g_end, tab, store);
                   else occupation_item_l11<true, false, 2>(args.plan, args.model, draws + b, g_begin, g_end, tab, store); }
      } else {
        if (mod) { if (u5) occupation_item_l11<false, true, kOccUnroll>(args.plan, args.model, draws + b, g_begin, g_end, tab, store);
                   else occupation_item_l11<false, true, 2>(args.plan, args.model, draws + b, g_begin, g_end, tab, store); }
        else     { if (u5) occupation_item_l11<false, false, kOccUnroll>(args.plan, args.model, draws + b, g_begin, g_end, tab, store);
                   else occupation_item_l11<false, false, 2>(args.plan, args.model, draws + b, g_begin, g_end, tab, store); }
      }
    }
  }
}

// element-wise evaluation of the table-driven math, for the accuracy tests (tc_debug_math)
__global__ void debug_math_kernel(int kind, const double* x, const double* y, double* out,
                                  long long n) {
  __shared__ double tab[kTabDoubles];
  load_math_tables(tab);
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = kind == 0 ? half_erfc_neg(x[i], tab) : pow_pos(x[i], y[i], tab);
}

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

// ------------------------------------------------------------------------------------------
// host-side table preparation
// ------------------------------------------------------------------------------------------
template <typename T>
int upload(const std::vector<T>& host, T** dev) {
  *dev = nullptr;
  size_t bytes = std::max<size_t>(host.size(), 1) * sizeof(T);
  TC_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), bytes));
  if (!host.empty())
    TC_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return TC_OK;
}

struct PlanHost {
  OccPlan dev{};
  std::vector<void*> allocations;
};

struct Layout {
  bool built = false;
  bool built32 = false;   // afrag32 (3xTF32 mode) is built on first use
  LayoutDev dev{};
  std::vector<int> row_to_pad;
  std::vector<void*> allocations;
  std::map<int, PlanHost> plans;  // by n_gauss
};

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Tuning knobs for experiments (tools/bench_variants.py): TC_TUNE_<NAME>=<int> in the environment.
int tune(const char* name, int fallback) {
  const char* v = std::getenv((std::string("TC_TUNE_") + name).c_str());
  return v && *v ? std::atoi(v) : fallback;
}

}  // namespace

struct tc_table {
  int device = 0;
  int mode = 0;
  int n_rows = 0, n_r = 0, n_tables = 0, n_cen = 0;
  std::vector<double> n_h, log_min, log_max, pct, dist;
  bool has_dist = false;
  std::vector<int> is_sat;
  std::vector<std::vector<double>> matrices;  // host copies, kept to build the split layout lazily
  std::map<int, std::pair<std::vector<double>, std::vector<double>>> rules;  // n_gauss -> (x, w)
  Layout layouts[2];  // [separate]
  std::mutex mutex;
};

struct tc_interp {
  int device = 0;
  InterpDev dev{};
  int sum_knots = 0;
  std::vector<void*> allocations;
};

namespace {

// Build the padded row order and the A-fragment stream of one layout.
int build_layout(tc_table* t, int separate) {
  Layout& L = t->layouts[separate];
  if (L.built) return TC_OK;
  const int N = t->n_rows, R = t->n_r, T = t->n_tables;
  const int Reff = R * T;
  const int n_cen = t->n_cen, n_sat = N - n_cen;
  // centrals first (stable); in the split layout the satellite block starts on a 16-row tile
  const int nc_pad = separate ? round_up(n_cen, 16) : n_cen;
  const int n_pad = std::max(16, round_up(nc_pad + n_sat, 16));
  L.row_to_pad.assign(N, -1);
  std::vector<int> pad_to_row(n_pad, -1);
  {
    int ic = 0, is = nc_pad;
    for (int i = 0; i < N; i++) {
      int p = t->is_sat[i] ? is++ : ic++;
      L.row_to_pad[i] = p;
      pad_to_row[p] = i;
    }
  }
  const int T16 = n_pad / 16;
  std::vector<Chunk> chunks;
  std::vector<std::vector<int>> out_lists;
  std::vector<double2> afrag;
  long long ks_per_r = 0;
  int n_parts = 0;

  if (t->mode == TC_MODE_AUTO) {
    ks_per_r = 2LL * T16 * (T16 + 1);
    afrag.assign((size_t)Reff * ks_per_r * 32 + 32, make_double2(0.0, 0.0));
    // M'[i][j] (j <= i) = M[i][j] (i == j) or 2 M[i][j]: the reference's packed prefactor sum
    // (tabcorr.py:638-642) written as a lower-triangular matrix product.
    for (int tb = 0; tb < T; tb++) {
      const double* packed = t->matrices[tb].data();
      const size_t P = (size_t)N * (N + 1) / 2;
      for (int r = 0; r < R; r++) {
        const double* m = packed + (size_t)r * P;
        double2* dst = afrag.data() + (size_t)(tb * R + r) * ks_per_r * 32;
        for (int i = 0; i < N; i++) {
          const int pi = L.row_to_pad[i];
          for (int j = 0; j <= i; j++) {
            const int pj = L.row_to_pad[j];
            double val = m[(size_t)i * (i + 1) / 2 + j];
            if (i != j) val *= 2.0;
            const int hi = std::max(pi, pj), lo = std::min(pi, pj);
            const int mt = hi / 16, rr = hi % 16, ks = lo / 4, tg = lo % 4;
            double2& d = dst[((size_t)2 * mt * (mt + 1) + ks) * 32 + (rr % 8) * 4 + tg];
            if (rr < 8) d.x = val; else d.y = val;
          }
        }
      }
    }
    // chunks: per radial bin, the tile range cut into pieces of similar cost
    const int n_comp = separate ? 3 : 1;
    out_lists.assign((size_t)Reff * n_comp, {});
    const int c16 = nc_pad / 16;  // first satellite tile (split layout)
    int pieces = std::max(1, std::min(T16, (tune("CHUNKS", 6 * kWarps) + Reff - 1) / Reff));
    auto add_range = [&](int r, int mt_lo, int mt_hi, int k_begin, int k_cap, int comp, int np) {
      // cost of tile mt ~ number of k-steps
      auto cost = [&](int mt) { return std::max(0, std::min(4 * (mt + 1), k_cap) - k_begin); };
      long long total = 0;
      for (int mt = mt_lo; mt < mt_hi; mt++) total += cost(mt);
      if (total == 0) return;
      np = std::max(1, std::min(np, mt_hi - mt_lo));
      long long acc = 0;
      int start = mt_lo, piece = 0;
      for (int mt = mt_lo; mt < mt_hi; mt++) {
        acc += cost(mt);
        bool last = mt == mt_hi - 1;
        if (last || acc * np >= total * (piece + 1)) {
          Chunk c{};
          c.r = r; c.mt0 = start; c.mt1 = mt + 1; c.k_begin = k_begin; c.k_cap = k_cap;
          c.part_row = n_parts++;
          chunks.push_back(c);
          out_lists[(size_t)r * n_comp + comp].push_back(c.part_row);
          start = mt + 1;
          piece++;
        }
      }
    };
    const int kinf = 1 << 28;
    for (int r = 0; r < Reff; r++) {
      if (!separate) {
        add_range(r, 0, T16, 0, kinf, 0, pieces);
      } else {
        add_range(r, 0, c16, 0, kinf, 0, pieces);              // centrals-centrals
        add_range(r, c16, T16, 0, nc_pad / 4, 1, pieces);       // centrals-satellites
        add_range(r, c16, T16, nc_pad / 4, kinf, 2, pieces);    // satellites-satellites
      }
    }
  } else {
    const int n_rt = (Reff + 15) / 16;
    ks_per_r = n_pad / 4;
    afrag.assign((size_t)n_rt * ks_per_r * 32 + 32, make_double2(0.0, 0.0));
    for (int tb = 0; tb < T; tb++) {
      const double* m = t->matrices[tb].data();
      for (int r = 0; r < R; r++) {
        const int re = tb * R + r, rt = re / 16, rr = re % 16;
        for (int i = 0; i < N; i++) {
          const int pi = L.row_to_pad[i];
          double2& d = afrag[((size_t)rt * ks_per_r + pi / 4) * 32 + (rr % 8) * 4 + pi % 4];
          if (rr < 8) d.x = m[(size_t)r * N + i]; else d.y = m[(size_t)r * N + i];
        }
      }
    }
    const int n_comp = separate ? 2 : 1;
    out_lists.assign((size_t)Reff * n_comp, {});
    const int ks_total = n_pad / 4;
    // k-ranges: split at the centrals/satellites boundary (split layout) and into pieces
    std::vector<std::pair<int, int>> segs;
    if (separate) {
      segs.push_back({0, nc_pad / 4});
      segs.push_back({nc_pad / 4, ks_total});
    } else {
      segs.push_back({0, ks_total});
    }
    const int want = std::max(1, (4 * kWarps + n_rt - 1) / n_rt / (int)segs.size());
    for (int rt = 0; rt < n_rt; rt++) {
      for (size_t sg = 0; sg < segs.size(); sg++) {
        const int lo = segs[sg].first, hi = segs[sg].second;
        if (hi <= lo) continue;
        const int np = std::max(1, std::min(want, (hi - lo + 7) / 8));
        for (int pc = 0; pc < np; pc++) {
          Chunk c{};
          c.r = rt;
          c.k_begin = lo + (int)((long long)(hi - lo) * pc / np);
          c.k_cap = lo + (int)((long long)(hi - lo) * (pc + 1) / np);
          if (c.k_cap <= c.k_begin) continue;
          c.part_row = n_parts;
          n_parts += 16;
          chunks.push_back(c);
          for (int rr = 0; rr < 16; rr++) {
            const int re = rt * 16 + rr;
            if (re < Reff) out_lists[(size_t)re * n_comp + sg].push_back(c.part_row + rr);
          }
        }
      }
    }
  }
  auto chunk_cost = [&](const Chunk& c) {
    if (t->mode != TC_MODE_AUTO) return (long long)(c.k_cap - c.k_begin);
    long long s = 0;
    for (int mt = c.mt0; mt < c.mt1; mt++)
      s += std::max(0, std::min(4 * (mt + 1), c.k_cap) - c.k_begin);
    return s;
  };
  // longest chunks first (only the end of a CTA's last tile is sensitive to the order)
  std::stable_sort(chunks.begin(), chunks.end(),
                   [&](const Chunk& a, const Chunk& b) { return chunk_cost(a) > chunk_cost(b); });

  std::vector<long long> cost_prefix(chunks.size() + 1, 0);
  for (size_t c = 0; c < chunks.size(); c++)
    cost_prefix[c + 1] = cost_prefix[c] + std::max<long long>(1, chunk_cost(chunks[c]));

  std::vector<int> out_ptr(out_lists.size() + 1, 0), out_parts;
  for (size_t o = 0; o < out_lists.size(); o++) {
    out_ptr[o + 1] = out_ptr[o] + (int)out_lists[o].size();
    out_parts.insert(out_parts.end(), out_lists[o].begin(), out_lists[o].end());
  }

  double2* d_afrag; Chunk* d_chunks; int *d_out_ptr, *d_out_parts, *d_pad_to_row;
  long long* d_cost_prefix;
  int rc;
  if ((rc = upload(cost_prefix, &d_cost_prefix))) return rc;
  L.allocations.push_back(d_cost_prefix);
  if ((rc = upload(afrag, &d_afrag))) return rc;
  L.allocations.push_back(d_afrag);
  if ((rc = upload(chunks, &d_chunks))) return rc;
  L.allocations.push_back(d_chunks);
  if ((rc = upload(out_ptr, &d_out_ptr))) return rc;
  L.allocations.push_back(d_out_ptr);
  if ((rc = upload(out_parts, &d_out_parts))) return rc;
  L.allocations.push_back(d_out_parts);
  if ((rc = upload(pad_to_row, &d_pad_to_row))) return rc;
  L.allocations.push_back(d_pad_to_row);

  L.dev.n_rows = N;
  L.dev.n_pad = n_pad;
  L.dev.nc_pad = nc_pad;
  L.dev.n_parts = std::max(n_parts, 1);
  L.dev.n_chunks = (int)chunks.size();
  L.dev.n_out = (int)out_lists.size();
  L.dev.ks_per_r = ks_per_r;
  L.dev.afrag = d_afrag;
  L.dev.chunks = d_chunks;
  L.dev.chunk_cost_prefix = d_cost_prefix;
  L.dev.out_ptr = d_out_ptr;
  L.dev.out_parts = d_out_parts;
  L.dev.pad_to_row = d_pad_to_row;
  L.built = true;
  return TC_OK;
}

// round an FP32 value to TF32 (10 explicit mantissa bits), ties to even
float tf32_round_host(float x) {
  uint32_t u;
  std::memcpy(&u, &x, sizeof(u));
  u += 0xfffu + ((u >> 13) & 1u);
  u &= 0xffffe000u;
  std::memcpy(&x, &u, sizeof(u));
  return x;
}

// A-fragment stream of the 3xTF32 mode: the same lower-triangular M' as build_layout, in m16n8k8
// fragments (16-row tiles x k8-steps of 8 columns), every entry split into TF32 high and low part.
int build_afrag32(tc_table* t, int separate) {
  Layout& L = t->layouts[separate];
  if (L.built32) return TC_OK;
  if (t->mode != TC_MODE_AUTO)
    return fail(TC_EUNSUPPORTED, "the 3xTF32 mode exists for auto-correlation tables only");
  const int N = t->n_rows, R = t->n_r, T = t->n_tables, Reff = R * T;
  const int T16 = L.dev.n_pad / 16;
  const long long ks8_per_r = (long long)T16 * (T16 + 1);
  std::vector<float4> frag((size_t)Reff * ks8_per_r * 64 + 64, make_float4(0.f, 0.f, 0.f, 0.f));
  const size_t P = (size_t)N * (N + 1) / 2;
  for (int tb = 0; tb < T; tb++) {
    const double* packed = t->matrices[tb].data();
    for (int r = 0; r < R; r++) {
      const double* m = packed + (size_t)r * P;
      float4* dst = frag.data() + (size_t)(tb * R + r) * ks8_per_r * 64;
      for (int i = 0; i < N; i++) {
        const int pi = L.row_to_pad[i];
        for (int j = 0; j <= i; j++) {
          const int pj = L.row_to_pad[j];
          double val = m[(size_t)i * (i + 1) / 2 + j];
          if (i != j) val *= 2.0;
          const int hi_r = std::max(pi, pj), lo_c = std::min(pi, pj);
          const int mt = hi_r / 16, rr = hi_r % 16, ks = lo_c / 8, kk = lo_c % 8;
          const int lane = (rr % 8) * 4 + (kk % 4), reg = (rr / 8) + 2 * (kk / 4);
          const float hi = tf32_round_host((float)val);
          const float lo = tf32_round_host((float)(val - (double)hi));
          float4* block = dst + ((size_t)mt * (mt + 1) + ks) * 64;
          reinterpret_cast<float*>(&block[lane])[reg] = hi;
          reinterpret_cast<float*>(&block[32 + lane])[reg] = lo;
        }
      }
    }
  }
  float4* d_frag;
  int rc;
  if ((rc = upload(frag, &d_frag))) return rc;
  L.allocations.push_back(d_frag);
  L.dev.afrag32 = d_frag;
  L.dev.ks8_per_r = ks8_per_r;
  L.built32 = true;
  return TC_OK;
}

// Quadrature plan of a layout for one Gauss-Legendre rule (tabcorr.py:543-552,568-578).
int build_plan(tc_table* t, int separate, int n_gauss) {
  Layout& L = t->layouts[separate];
  if (L.plans.count(n_gauss)) return TC_OK;
  auto rule = t->rules.find(n_gauss);
  if (rule == t->rules.end())
    return fail(TC_EINVAL, "no quadrature rule registered for n_gauss=" + std::to_string(n_gauss) +
                               " (call tc_table_plan first)");
  const std::vector<double>& x01 = rule->second.first;
  const std::vector<double>& wq = rule->second.second;
  const int N = t->n_rows, G = n_gauss, n_pad = L.dev.n_pad;

  struct Group { double lo, hi; int sat; std::vector<int> rows; };
  std::vector<Group> groups;
  for (int pass = 0; pass < 2; pass++) {  // centrals groups first
    for (int i = 0; i < N; i++) {
      if ((t->is_sat[i] != 0) != (pass == 1)) continue;
      bool placed = false;
      for (auto& gq : groups) {
        if (gq.sat == pass && gq.lo == t->log_min[i] && gq.hi == t->log_max[i] &&
            (int)gq.rows.size() < kGroupRows) {
          gq.rows.push_back(i);
          placed = true;
          break;
        }
      }
      if (!placed) groups.push_back(Group{t->log_min[i], t->log_max[i], pass, {i}});
    }
  }
  const int n_groups = (int)groups.size();
  // the kernel evaluates `unroll` nodes per iteration
  const int want_unroll = tune("OCC_UNROLL", kOccUnroll);
  const int unroll = want_unroll == 2 * kOccUnroll && G % (2 * kOccUnroll) == 0 ? 2 * kOccUnroll
                     : want_unroll >= kOccUnroll && G % kOccUnroll == 0     ? kOccUnroll
                                                                              : 2;
  const int GP = round_up(G, unroll);
  std::vector<double> node_logm((size_t)n_groups * GP), node_m((size_t)n_groups * GP);
  std::vector<int> grp_rows((size_t)n_groups * kGroupRows, -1), grp_is_sat(n_groups);
  std::vector<double> row_c((size_t)(n_pad + 1) * GP, 0.0), row_nh(n_pad, 0.0), row_pct(n_pad, 0.0);
  for (int q = 0; q < n_groups; q++) {
    const Group& gq = groups[q];
    grp_is_sat[q] = gq.sat;
    for (int k = 0; k < G; k++) {
      // prim_haloprop = 10**(log_min + d_log * x) and halotools' log10(prim_haloprop)
      double m = std::pow(10.0, gq.lo + (gq.hi - gq.lo) * x01[k]);
      node_m[(size_t)q * GP + k] = m;
      node_logm[(size_t)q * GP + k] = std::log10(m);
    }
    for (int k = G; k < GP; k++) {  // zero-weight padding node
      node_m[(size_t)q * GP + k] = node_m[(size_t)q * GP];
      node_logm[(size_t)q * GP + k] = node_logm[(size_t)q * GP];
    }
    for (size_t s = 0; s < gq.rows.size(); s++) {
      const int i = gq.rows[s], p = L.row_to_pad[i];
      grp_rows[(size_t)q * kGroupRows + s] = p;
      row_nh[p] = t->n_h[i];
      row_pct[p] = t->pct[i];
      const double n = t->has_dist ? t->dist[i] + 1.0 : 0.0;  // tabcorr.py:568-574
      double norm = 0.0;
      for (int k = 0; k < G; k++) norm += wq[k] * std::pow(node_m[(size_t)q * GP + k], n);
      for (int k = 0; k < G; k++)
        row_c[(size_t)p * GP + k] = wq[k] * std::pow(node_m[(size_t)q * GP + k], n) / norm;
    }
  }
  PlanHost ph;
  double *d_logm, *d_m, *d_c, *d_nh, *d_pct; int *d_rows, *d_sat;
  int rc;
  if ((rc = upload(node_logm, &d_logm))) return rc; ph.allocations.push_back(d_logm);
  if ((rc = upload(node_m, &d_m))) return rc; ph.allocations.push_back(d_m);
  {
    std::vector<double> node_inv_m(node_m.size());
    for (size_t k = 0; k < node_m.size(); k++) node_inv_m[k] = 1.0 / node_m[k];
    double* d_inv;
    if ((rc = upload(node_inv_m, &d_inv))) return rc;
    ph.allocations.push_back(d_inv);
    ph.dev.node_inv_m = d_inv;
  }
  if ((rc = upload(grp_rows, &d_rows))) return rc; ph.allocations.push_back(d_rows);
  if ((rc = upload(grp_is_sat, &d_sat))) return rc; ph.allocations.push_back(d_sat);
  if ((rc = upload(row_c, &d_c))) return rc; ph.allocations.push_back(d_c);
  if ((rc = upload(row_nh, &d_nh))) return rc; ph.allocations.push_back(d_nh);
  if ((rc = upload(row_pct, &d_pct))) return rc; ph.allocations.push_back(d_pct);
  ph.dev.n_groups = n_groups;
  ph.dev.n_cen_groups = 0;
  for (int q = 0; q < n_groups; q++) ph.dev.n_cen_groups += grp_is_sat[q] ? 0 : 1;
  ph.dev.n_gauss = G;
  ph.dev.n_gauss_pad = GP;
  ph.dev.unroll = unroll;
  ph.dev.zero_row = n_pad;
  ph.dev.node_logm = d_logm;
  ph.dev.node_m = d_m;
  ph.dev.grp_rows = d_rows;
  ph.dev.grp_is_sat = d_sat;
  ph.dev.row_c = d_c;
  ph.dev.row_nh = d_nh;
  ph.dev.row_pct = d_pct;
  L.plans[n_gauss] = ph;
  return TC_OK;
}

size_t predict_smem_bytes(int n_pad, int nt, int n_buf) {
  return ((size_t)n_buf * n_pad * 8 * nt + kTabDoubles) * sizeof(double) + sizeof(PredictCtrl);
}

// Draw-tile width (8 nt draws) and number of W buffers.  Two buffers let the occupation of the
// next tile overlap the contraction of the current one, but halve the tile width that fits in
// shared memory, and the width is what the table stream from L2 is amortised over: measured on
// B200, N = 240: 2 x 56 draws = 1 x 64 draws (4.03 ms per 1e5 draws), N = 500: 1 x 48 draws beats
// 2 x 24 draws (16.9 vs 18.1 ms).  So two buffers are used while they leave at least 40 draws;
// within a buffer count the widest tile that fits, narrower only while the batch is too small to
// give every SM a tile.
void pick_tile(int n_pad, long long n_draws, int n_sm, int* nt_out, int* n_buf_out) {
  *nt_out = 0;
  *n_buf_out = 0;
  const int forced = tune("NBUF", 0);
  auto widest = [&](int n_buf) {
    for (int nt = 8; nt >= 1; nt--)
      if (predict_smem_bytes(n_pad, nt, n_buf) <= (size_t)kSmemLimit) return nt;
    return 0;
  };
  int n_buf = widest(2) >= 5 ? 2 : 1;
  if (forced == 1 || forced == 2) n_buf = forced;
  if (widest(n_buf) == 0) n_buf = 1;
  int best = 0;
  for (int nt = 8; nt >= 1; nt--) {
    if (predict_smem_bytes(n_pad, nt, n_buf) > (size_t)kSmemLimit) continue;
    best = nt;  // the largest that fits, shrinking while the grid would not fill the device
    if ((n_draws + 8 * nt - 1) / (8 * nt) >= n_sm) break;
  }
  if (best) {
    *nt_out = best;
    *n_buf_out = n_buf;
  }
}

struct Workspace {
  size_t parts_bytes, ngal_bytes, total;
  long long n_tiles;
  int nt, n_buf;
};

Workspace plan_workspace(const Layout& L, long long n_draws, int n_sm) {
  Workspace w{};
  pick_tile(L.dev.n_pad, n_draws, n_sm, &w.nt, &w.n_buf);
  if (w.nt == 0) return w;
  const int bm = 8 * w.nt;
  w.n_tiles = (n_draws + bm - 1) / bm;
  w.parts_bytes = (size_t)w.n_tiles * L.dev.n_parts * bm * sizeof(double);
  w.ngal_bytes = (size_t)w.n_tiles * 2 * bm * sizeof(double);
  w.total = w.parts_bytes + w.ngal_bytes;
  return w;
}

// Occupation items per n-tile: about kOccItemsPerTile / nt group ranges, split between centrals and
// satellites in proportion to their groups (at least one each where the type exists).
constexpr int kOccItemsPerTile = 14;

void pick_ranges(const OccPlan& plan, int nt, int* n_cen, int* n_sat) {
  const int cen = plan.n_cen_groups, sat = plan.n_groups - plan.n_cen_groups;
  const int want = std::max(2, (tune("OCC_ITEMS", kOccItemsPerTile) + nt - 1) / nt);
  auto share = [&](int count) {
    if (count == 0) return 0;
    const int units = (count + 3) / 4;
    return std::max(1, std::min(units, (int)std::lround((double)want * count / (cen + sat))));
  };
  *n_cen = share(cen);
  *n_sat = share(sat);
  if (*n_cen + *n_sat == 0) *n_cen = 1;  // table without rows cannot happen; keep n_occ > 0
}

// Coefficient tables of the occupation math (see half_erfc_neg / pow_pos), computed in long double.
int ensure_math_tables(int device) {
  static std::mutex m;
  static std::map<int, bool> done;
  std::lock_guard<std::mutex> lock(m);
  if (done[device]) return TC_OK;
  std::vector<double> tab(kTabDoubles, 0.0);
  const int n = kErfDeg + 1;
  const long double pi = 3.14159265358979323846264338327950288L;
  tab[kErfIntervals - 1] = 1.0;  // saturated columns: 0 below the first, 1 above the last interval
  for (int i = 0; i < kErfIntervals - 2; i++) {
    const long double xc = -6.0L + 0.5L * i;  // x = xc + s / 4 with s in [-1, 1]; t = s / 2
    std::vector<long double> fs(n), sn(n);
    for (int j = 0; j < n; j++) {
      sn[j] = cosl(pi * (j + 0.5L) / n);
      fs[j] = 0.5L * erfcl(-(xc + 0.25L * sn[j]));
    }
    // Chebyshev coefficients of the interpolant, then Chebyshev -> monomial in s
    std::vector<long double> a(n, 0.0L);
    for (int k = 0; k < n; k++) {
      long double sum = 0.0L;
      for (int j = 0; j < n; j++) sum += fs[j] * cosl(k * pi * (j + 0.5L) / n);
      a[k] = (k == 0 ? 1.0L : 2.0L) * sum / n;
    }
    std::vector<long double> mono(n, 0.0L), t0(n, 0.0L), t1(n, 0.0L), t2(n, 0.0L);
    t0[0] = 1.0L;                 // T_0
    t1[1] = 1.0L;                 // T_1
    for (int d = 0; d < n; d++) mono[d] += a[0] * t0[d] + (n > 1 ? a[1] * t1[d] : 0.0L);
    for (int k = 2; k < n; k++) {  // T_k = 2 s T_{k-1} - T_{k-2}
      for (int d = 0; d < n; d++) t2[d] = (d > 0 ? 2.0L * t1[d - 1] : 0.0L) - t0[d];
      for (int d = 0; d < n; d++) mono[d] += a[k] * t2[d];
      t0 = t1;
      t1 = t2;
    }
    long double scale = 1.0L;    // s = 2 t
    for (int d = 0; d < n; d++) {
      tab[(size_t)d * kErfStride + i + 1] = (double)(mono[d] * scale);
      scale *= 2.0L;
    }
  }
  for (int i = 0; i < kLogEntries; i++) {
    const double inv_c = (double)(1.0L / (1.0L + (i + 0.5L) / kLogEntries));
    tab[kTabLog + 2 * i] = inv_c;
    tab[kTabLog + 2 * i + 1] = (double)(-logl((long double)inv_c));
  }
  for (int j = 0; j < kExpEntries; j++) tab[kTabExp + j] = (double)exp2l((long double)j / kExpEntries);
  TC_CUDA(cudaMemcpyToSymbol(g_math_tables, tab.data(), sizeof(double) * kTabDoubles));
  done[device] = true;
  return TC_OK;
}

int device_sms(int device, int* n_sm) {
  static std::mutex m;
  static std::map<int, int> cache;
  std::lock_guard<std::mutex> lock(m);
  auto it = cache.find(device);
  if (it == cache.end()) {
    int v = 0;
    TC_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    it = cache.emplace(device, v).first;
  }
  *n_sm = it->second;
  return TC_OK;
}

template <int NT, int MODE>
int launch_predict(const PredictArgs& args, dim3 grid, size_t smem, cudaStream_t stream) {
  static std::mutex m;
  static std::map<int, bool> configured;
  int dev = 0;
  TC_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lock(m);
    if (!configured[dev]) {
      TC_CUDA(cudaFuncSetAttribute(predict_kernel<NT, MODE>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
      configured[dev] = true;
    }
  }
  predict_kernel<NT, MODE><<<grid, kThreads, smem, stream>>>(args);
  TC_CUDA(cudaGetLastError());
  return TC_OK;
}

// optional per-kernel timing for bench.py (tc_profile_enable / tc_profile_read)
struct Profile {
  bool enabled = false;
  bool recorded = false;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};
Profile g_profile;

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; (void)cudaGetLastError(); return; }
    if (prev != device && cudaSetDevice(device) != cudaSuccess) { ok = false; (void)cudaGetLastError(); }
  }
  ~DeviceGuard() { if (prev >= 0) (void)cudaSetDevice(prev); }
};

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* tc_last_error(void) { return g_last_error.c_str(); }

int tc_version(void) { return TC_VERSION; }

int tc_model_n_theta(const tc_model* model) {
  if (!model) return fail(TC_EINVAL, "tc_model_n_theta: model is NULL");
  if (model->family == TC_FAMILY_ZHENG07) return TC_N_THETA;
  if (model->family == TC_FAMILY_LEAUTHAUD11) return TC_N_THETA_LEAUTHAUD11;
  return fail(TC_EUNSUPPORTED, "tc_model_n_theta: unknown model family");
}

int tc_table_create(tc_table** out, int mode, int n_rows, int n_r, int n_tables,
                    const double* n_h, const double* log_min, const double* log_max,
                    const double* sec_pct, const double* dist_index, const int32_t* is_sat,
                    const double* const* tpcf_matrix, int device) {
  if (out == nullptr) return fail(TC_EINVAL, "tc_table_create: out is NULL");
  *out = nullptr;
  if (mode != TC_MODE_AUTO && mode != TC_MODE_CROSS)
    return fail(TC_EINVAL, "tc_table_create: mode must be TC_MODE_AUTO or TC_MODE_CROSS");
  if (n_rows <= 0 || n_r <= 0 || n_tables <= 0)
    return fail(TC_EINVAL, "tc_table_create: n_rows, n_r and n_tables must be positive");
  if (!n_h || !log_min || !log_max || !sec_pct || !is_sat || !tpcf_matrix)
    return fail(TC_EINVAL, "tc_table_create: NULL input array");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_table_create: cannot select CUDA device " +
                                           std::to_string(device) + " (no CUDA device available?)");
  tc_table* t = new (std::nothrow) tc_table();
  if (!t) return fail(TC_ENOMEM, "tc_table_create: out of host memory");
  t->device = device;
  t->mode = mode;
  t->n_rows = n_rows;
  t->n_r = n_r;
  t->n_tables = n_tables;
  t->n_h.assign(n_h, n_h + n_rows);
  t->log_min.assign(log_min, log_min + n_rows);
  t->log_max.assign(log_max, log_max + n_rows);
  t->pct.assign(sec_pct, sec_pct + n_rows);
  t->has_dist = dist_index != nullptr;
  if (t->has_dist) t->dist.assign(dist_index, dist_index + n_rows);
  t->is_sat.assign(is_sat, is_sat + n_rows);
  t->n_cen = 0;
  for (int i = 0; i < n_rows; i++) t->n_cen += t->is_sat[i] ? 0 : 1;
  const size_t cols = mode == TC_MODE_AUTO ? (size_t)n_rows * (n_rows + 1) / 2 : (size_t)n_rows;
  t->matrices.resize(n_tables);
  for (int tb = 0; tb < n_tables; tb++) {
    if (!tpcf_matrix[tb]) { delete t; return fail(TC_EINVAL, "tc_table_create: NULL matrix"); }
    t->matrices[tb].assign(tpcf_matrix[tb], tpcf_matrix[tb] + (size_t)n_r * cols);
  }
  int rc = ensure_math_tables(device);
  if (rc == TC_OK) rc = build_layout(t, 0);
  if (rc != TC_OK) { tc_table_destroy(t); return rc; }
  *out = t;
  return TC_OK;
}

int tc_table_destroy(tc_table* t) {
  if (!t) return TC_OK;
  DeviceGuard guard(t->device);
  for (Layout& L : t->layouts) {
    for (void* p : L.allocations) (void)cudaFree(p);
    for (auto& kv : L.plans)
      for (void* p : kv.second.allocations) (void)cudaFree(p);
  }
  delete t;
  return TC_OK;
}

int tc_table_n_rows(const tc_table* t) { return t ? t->n_rows : TC_EINVAL; }
int tc_table_n_r(const tc_table* t) { return t ? t->n_r : TC_EINVAL; }
int tc_table_n_tables(const tc_table* t) { return t ? t->n_tables : TC_EINVAL; }

int tc_table_plan(tc_table* t, int n_gauss, const double* x01, const double* w) {
  if (!t || !x01 || !w || n_gauss <= 0) return fail(TC_EINVAL, "tc_table_plan: bad argument");
  std::lock_guard<std::mutex> lock(t->mutex);
  DeviceGuard guard(t->device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_table_plan: cannot select the table's CUDA device");
  if (!t->rules.count(n_gauss))
    t->rules[n_gauss] = {std::vector<double>(x01, x01 + n_gauss), std::vector<double>(w, w + n_gauss)};
  return build_plan(t, 0, n_gauss);
}

int tc_occupation_batch(tc_table* t, const tc_model* model, int n_gauss, const double* theta,
                        int64_t theta_ld, int64_t n_draws, double* occ, void* stream) {
  if (!t || !model || !theta || !occ) return fail(TC_EINVAL, "tc_occupation_batch: NULL argument");
  if (theta_ld != 0 && theta_ld < n_draws)
    return fail(TC_EINVAL, "tc_occupation_batch: theta_ld must be 0 or >= n_draws");
  if (model->family != TC_FAMILY_ZHENG07 && model->family != TC_FAMILY_LEAUTHAUD11)
    return fail(TC_EUNSUPPORTED, "tc_occupation_batch: unknown model family");
  if (n_draws <= 0) return TC_OK;
  std::lock_guard<std::mutex> lock(t->mutex);
  DeviceGuard guard(t->device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_occupation_batch: cannot select the table's CUDA device");
  int rc = build_plan(t, 0, n_gauss);
  if (rc != TC_OK) return rc;
  int n_sm = 0;
  if ((rc = device_sms(t->device, &n_sm))) return rc;
  const int n_theta = model->family == TC_FAMILY_LEAUTHAUD11 ? TC_N_THETA_LEAUTHAUD11 : TC_N_THETA;
  OccArgs args{};
  args.plan = t->layouts[0].plans[n_gauss].dev;
  args.model = *model;
  args.theta = theta;
  args.theta_ds = theta_ld ? 1 : n_theta;
  args.theta_ps = theta_ld ? theta_ld : 1;
  args.n_draws = n_draws;
  args.n_rows = t->n_rows;
  args.pad_to_row = t->layouts[0].dev.pad_to_row;
  args.occ_out = occ;
  pick_ranges(args.plan, 1, &args.n_ranges_cen, &args.n_ranges_sat);
  if (model->family == TC_FAMILY_LEAUTHAUD11) {
    const size_t smem = kTabDoubles * sizeof(double) + kL11DrawsPerBlock * sizeof(L11Draw);
    static std::mutex m;
    static std::map<int, bool> configured;
    {
      std::lock_guard<std::mutex> lock_attr(m);
      if (!configured[t->device]) {
        TC_CUDA(cudaFuncSetAttribute(occupation_l11_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[t->device] = true;
      }
    }
    const long long n_blocks = (n_draws + kL11DrawsPerBlock - 1) / kL11DrawsPerBlock;
    const int grid = (int)std::max<long long>(1, std::min<long long>(n_blocks, n_sm));
    occupation_l11_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(args);
    TC_CUDA(cudaGetLastError());
    return TC_OK;
  }
  const long long n_items = (n_draws + 7) / 8 * (args.n_ranges_cen + args.n_ranges_sat);
  int grid = (int)std::max<long long>(1, std::min<long long>((n_items + kWarps - 1) / kWarps,
                                                             (long long)n_sm));
  occupation_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(args);
  TC_CUDA(cudaGetLastError());
  return TC_OK;
}

size_t tc_predict_workspace_bytes(const tc_table* t, int64_t n_draws, int separate) {
  if (!t || n_draws <= 0) return 0;
  tc_table* tt = const_cast<tc_table*>(t);
  std::lock_guard<std::mutex> lock(tt->mutex);
  DeviceGuard guard(t->device);
  if (!guard.ok) return 0;
  if (build_layout(tt, separate ? 1 : 0) != TC_OK) return 0;
  int n_sm = 0;
  if (device_sms(t->device, &n_sm) != TC_OK) return 0;
  return plan_workspace(t->layouts[separate ? 1 : 0], n_draws, n_sm).total;
}

}  // extern "C"

namespace {

// tc_predict_batch / tc_predict_one.  theta_inline: host pointer to the TC_N_THETA parameters of a
// single draw, passed to the kernel in its launch arguments (theta and occ are NULL then).
int predict_impl(tc_table* t, const tc_model* model, int n_gauss, const double* theta,
                 int64_t theta_ld, const double* occ, const double* theta_inline, int64_t n_draws,
                 int separate, int precision, double* ngal, int64_t ngal_stride, double* xi,
                 int64_t xi_stride, void* workspace, size_t workspace_bytes, void* stream_) {
  if (!t || !ngal || !xi) return fail(TC_EINVAL, "tc_predict_batch: NULL argument");
  if (theta_inline) {
    if (theta || occ || n_draws != 1)
      return fail(TC_EINVAL, "tc_predict_one: one draw with host parameters only");
    theta = theta_inline;   // validated like device parameters below; never dereferenced on the device
  }
  if ((theta == nullptr) == (occ == nullptr))
    return fail(TC_EINVAL, "tc_predict_batch: exactly one of theta_dev and occ_dev must be given");
  if (theta && !model) return fail(TC_EINVAL, "tc_predict_batch: model is NULL");
  if (theta && theta_ld != 0 && theta_ld < n_draws)
    return fail(TC_EINVAL, "tc_predict_batch: theta_ld must be 0 or >= n_draws");
  if (theta && model->family != TC_FAMILY_ZHENG07)
    return fail(TC_EUNSUPPORTED, "tc_predict_batch: the fused occupation phase implements "
                                 "TC_FAMILY_ZHENG07 only; evaluate tc_occupation_batch and pass "
                                 "its result as occ_dev");
  if (precision != TC_PRECISION_FP64 && precision != TC_PRECISION_3XTF32)
    return fail(TC_EINVAL, "tc_predict_batch: precision must be TC_PRECISION_FP64 or _3XTF32");
  if (precision == TC_PRECISION_3XTF32 && t->mode != TC_MODE_AUTO)
    return fail(TC_EUNSUPPORTED, "tc_predict_batch: the 3xTF32 mode exists for auto-correlation "
                                 "tables only (cross tables are bound by the occupation phase)");
  if (n_draws <= 0) return TC_OK;
  separate = separate ? 1 : 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  std::lock_guard<std::mutex> lock(t->mutex);
  DeviceGuard guard(t->device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_predict_batch: cannot select the table's CUDA device");
  int rc = build_layout(t, separate);
  if (rc != TC_OK) return rc;
  if (precision == TC_PRECISION_3XTF32 && (rc = build_afrag32(t, separate))) return rc;
  Layout& L = t->layouts[separate];
  if (!theta && t->rules.empty()) {
    // the occupation branch needs n_h per padded row only; any plan carries it
    const double x = 0.5, w = 2.0;
    t->rules[1] = {std::vector<double>(1, x), std::vector<double>(1, w)};
  }
  const int plan_g = theta ? n_gauss : t->rules.begin()->first;
  if ((rc = build_plan(t, separate, plan_g))) return rc;
  int n_sm = 0;
  if ((rc = device_sms(t->device, &n_sm))) return rc;
  Workspace ws = plan_workspace(L, n_draws, n_sm);
  if (ws.nt == 0)
    return fail(TC_EUNSUPPORTED, "tc_predict_batch: table too large for the shared-memory tile (" +
                                     std::to_string(L.dev.n_pad) + " padded rows)");
  if (!workspace || workspace_bytes < ws.total)
    return fail(TC_EINVAL, "tc_predict_batch: workspace too small, need " +
                               std::to_string(ws.total) + " bytes");
  const int n_comp_ngal = separate ? 2 : 1;
  const int n_out = L.dev.n_out;
  if (ngal_stride < (int64_t)t->n_tables * n_comp_ngal || xi_stride < n_out)
    return fail(TC_EINVAL, "tc_predict_batch: output stride smaller than one draw's outputs");

  PredictArgs args{};
  args.lay = L.dev;
  args.plan = L.plans[plan_g].dev;
  if (model) args.model = *model;
  args.theta = theta_inline ? nullptr : theta;
  args.theta_ds = theta_ld ? 1 : TC_N_THETA;
  args.theta_ps = theta_ld ? theta_ld : 1;
  if (theta_inline) {
    args.theta_is_inline = 1;
    args.theta_ds = 0;
    args.theta_ps = 1;
    for (int k = 0; k < TC_N_THETA; k++) args.theta_inline[k] = theta_inline[k];
  }
  args.occ = occ;
  args.n_draws = n_draws;
  args.n_tiles = ws.n_tiles;
  args.parts = static_cast<double*>(workspace);
  args.ngal_tile = reinterpret_cast<double*>(static_cast<char*>(workspace) + ws.parts_bytes);

  args.n_buf = ws.n_buf;
  pick_ranges(args.plan, ws.nt, &args.n_ranges_cen, &args.n_ranges_sat);
  {
    // occupation items take every occ_stride-th slot of the first 70 % of a tile's work list (an
    // item runs for tens of microseconds beside DMMA warps that starve its scalar FP64, so the
    // last one must be taken well before the list ends: 70 % measured 2 % faster than 90 %);
    // with a single W buffer they must all come before the chunks that wait for them
    const int n_occ = ws.nt * (args.n_ranges_cen + args.n_ranges_sat);
    const int slots = L.dev.n_chunks + n_occ;
    // (every occupation slot must exist: stride * n_occ <= slots)
    const int spread = std::min(100, std::max(1, tune("OCC_SPREAD", 70)));
    args.occ_stride = ws.n_buf == 1 ? 1 : std::max(1, (int)((long long)spread * slots / (100LL * n_occ)));
  }
  args.tf32_segment = std::max(1, tune("TF32_SEG", kTf32Segment));
  const int bm = 8 * ws.nt;
  const size_t smem = predict_smem_bytes(L.dev.n_pad, ws.nt, ws.n_buf);
  const int gx = (int)std::min<long long>(ws.n_tiles * L.dev.n_chunks, n_sm);
  dim3 grid(gx, 1);
#define TC_LAUNCH(NT_)                                                                        \
  rc = precision == TC_PRECISION_3XTF32                                                       \
           ? launch_predict<NT_, kModeAutoTf32>(args, grid, smem, stream)                     \
           : t->mode == TC_MODE_AUTO ? launch_predict<NT_, TC_MODE_AUTO>(args, grid, smem, stream) \
                                     : launch_predict<NT_, TC_MODE_CROSS>(args, grid, smem, stream)
  if (g_profile.enabled) TC_CUDA(cudaEventRecord(g_profile.ev[0], stream));
  switch (ws.nt) {
    case 8: TC_LAUNCH(8); break;
    case 7: TC_LAUNCH(7); break;
    case 6: TC_LAUNCH(6); break;
    case 5: TC_LAUNCH(5); break;
    case 4: TC_LAUNCH(4); break;
    case 3: TC_LAUNCH(3); break;
    case 2: TC_LAUNCH(2); break;
    default: TC_LAUNCH(1); break;
  }
#undef TC_LAUNCH
  if (rc != TC_OK) return rc;
  if (g_profile.enabled) TC_CUDA(cudaEventRecord(g_profile.ev[1], stream));

  FinalizeArgs fa{};
  fa.lay = L.dev;
  fa.parts = args.parts;
  fa.ngal_tile = args.ngal_tile;
  fa.n_draws = n_draws;
  fa.bm = bm;
  fa.mode = t->mode;
  fa.separate = separate;
  fa.n_tables = t->n_tables;
  fa.ngal_out = ngal;
  fa.ngal_stride = ngal_stride;
  fa.xi_out = xi;
  fa.xi_stride = xi_stride;
  const int outs_per_block = 256 / bm;
  int fy = std::max(1, std::min(64, (n_out + outs_per_block - 1) / outs_per_block));
  if (ws.n_tiles > 4LL * n_sm) fy = 1;
  dim3 fgrid((unsigned)ws.n_tiles, fy);
  finalize_kernel<<<fgrid, 256, 0, stream>>>(fa);
  TC_CUDA(cudaGetLastError());
  if (g_profile.enabled) {
    TC_CUDA(cudaEventRecord(g_profile.ev[2], stream));
    g_profile.recorded = true;
  }
  return TC_OK;
}

}  // namespace

extern "C" {

int tc_predict_batch(tc_table* t, const tc_model* model, int n_gauss, const double* theta,
                     int64_t theta_ld, const double* occ, int64_t n_draws, int separate,
                     int precision, double* ngal, int64_t ngal_stride, double* xi,
                     int64_t xi_stride, void* workspace, size_t workspace_bytes, void* stream) {
  return predict_impl(t, model, n_gauss, theta, theta_ld, occ, nullptr, n_draws, separate,
                      precision, ngal, ngal_stride, xi, xi_stride, workspace, workspace_bytes,
                      stream);
}

int tc_predict_one(tc_table* t, const tc_model* model, int n_gauss, const double* theta_host,
                   int separate, int precision, double* ngal, int64_t ngal_stride, double* xi,
                   int64_t xi_stride, void* workspace, size_t workspace_bytes, void* stream) {
  if (!theta_host) return fail(TC_EINVAL, "tc_predict_one: theta_host is NULL");
  return predict_impl(t, model, n_gauss, nullptr, 0, nullptr, theta_host, 1, separate, precision,
                      ngal, ngal_stride, xi, xi_stride, workspace, workspace_bytes, stream);
}

int tc_profile_enable(int on) {
  if (on && !g_profile.ev[0]) {
    for (auto& e : g_profile.ev) TC_CUDA(cudaEventCreate(&e));
  }
  g_profile.enabled = on != 0;
  g_profile.recorded = false;
  return TC_OK;
}

int tc_profile_read(float* predict_ms, float* finalize_ms) {
  if (!predict_ms || !finalize_ms) return fail(TC_EINVAL, "tc_profile_read: NULL output");
  if (!g_profile.recorded) return fail(TC_EINVAL, "tc_profile_read: nothing recorded");
  TC_CUDA(cudaEventSynchronize(g_profile.ev[2]));
  TC_CUDA(cudaEventElapsedTime(predict_ms, g_profile.ev[0], g_profile.ev[1]));
  TC_CUDA(cudaEventElapsedTime(finalize_ms, g_profile.ev[1], g_profile.ev[2]));
  return TC_OK;
}

int tc_interp_create(tc_interp** out, int n_dims, const int32_t* n_knots, const double* knots,
                     const double* a, const int32_t* grid_to_table, int device) {
  if (!out) return fail(TC_EINVAL, "tc_interp_create: out is NULL");
  *out = nullptr;
  if (n_dims <= 0 || n_dims > kMaxDims)
    return fail(TC_EUNSUPPORTED, "tc_interp_create: between 1 and 8 interpolation axes supported");
  if (!n_knots || !knots || !a || !grid_to_table)
    return fail(TC_EINVAL, "tc_interp_create: NULL argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_interp_create: cannot select CUDA device " +
                                           std::to_string(device));
  tc_interp* it = new (std::nothrow) tc_interp();
  if (!it) return fail(TC_ENOMEM, "tc_interp_create: out of host memory");
  it->device = device;
  it->dev.n_dims = n_dims;
  long long n_tables = 1;
  int knot_off = 0, a_off = 0;
  for (int d = 0; d < n_dims; d++) {
    if (n_knots[d] < 4) {
      delete it;
      return fail(TC_EINVAL, "tc_interp_create: at least 4 knots per axis are required");
    }
    it->dev.n_knots[d] = n_knots[d];
    it->dev.knot_off[d] = knot_off;
    it->dev.a_off[d] = a_off;
    knot_off += n_knots[d];
    a_off += (n_knots[d] - 1) * 4 * n_knots[d];
    n_tables *= n_knots[d];
  }
  if (n_tables > 8192) {
    delete it;
    return fail(TC_EUNSUPPORTED, "tc_interp_create: more than 8192 grid tables");
  }
  it->dev.n_tables = (int)n_tables;
  it->sum_knots = knot_off;
  std::vector<double> hk(knots, knots + knot_off), ha(a, a + a_off);
  std::vector<int> hg(grid_to_table, grid_to_table + n_tables);
  double *dk, *da; int* dg;
  int rc;
  if ((rc = upload(hk, &dk))) { delete it; return rc; }
  it->allocations.push_back(dk);
  if ((rc = upload(ha, &da))) { tc_interp_destroy(it); return rc; }
  it->allocations.push_back(da);
  if ((rc = upload(hg, &dg))) { tc_interp_destroy(it); return rc; }
  it->allocations.push_back(dg);
  it->dev.knots = dk;
  it->dev.a = da;
  it->dev.grid_to_table = dg;
  *out = it;
  return TC_OK;
}

int tc_interp_destroy(tc_interp* it) {
  if (!it) return TC_OK;
  DeviceGuard guard(it->device);
  for (void* p : it->allocations) (void)cudaFree(p);
  delete it;
  return TC_OK;
}

int tc_interp_apply_batch(tc_interp* it, const double* x, int64_t n_draws, const double* data,
                          int n_cols, double* out, int extrapolate, int32_t* flag, void* stream) {
  if (!it || !x || !data || !out || !flag || n_cols <= 0)
    return fail(TC_EINVAL, "tc_interp_apply_batch: bad argument");
  if (n_draws <= 0) return TC_OK;
  DeviceGuard guard(it->device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_interp_apply_batch: cannot select the CUDA device");
  InterpArgs args{};
  args.it = it->dev;
  args.x = x;
  args.n_draws = n_draws;
  args.data = data;
  args.n_cols = n_cols;
  args.out = out;
  args.extrapolate = extrapolate;
  args.flag = flag;
  args.sum_knots = it->sum_knots;
  size_t smem = (size_t)4 * (it->sum_knots + it->dev.n_tables) * sizeof(double);
  if (smem > 48 * 1024) {
    TC_CUDA(cudaFuncSetAttribute(interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  }
  unsigned grid = (unsigned)((n_draws + 3) / 4);
  interp_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(args);
  TC_CUDA(cudaGetLastError());
  return TC_OK;
}

int tc_debug_math(int kind, const double* x, const double* y, double* out, int64_t n,
                  void* stream) {
  if (!x || !out || (kind != 0 && !y) || n < 0) return fail(TC_EINVAL, "tc_debug_math: bad argument");
  if (n == 0) return TC_OK;
  int device = 0;
  TC_CUDA(cudaGetDevice(&device));
  int rc = ensure_math_tables(device);
  if (rc) return rc;
  debug_math_kernel<<<256, 256, 0, static_cast<cudaStream_t>(stream)>>>(kind, x, y, out, n);
  TC_CUDA(cudaGetLastError());
  return TC_OK;
}

int tc_measure_dmma_peak(int device, double* tflops) {
  if (!tflops) return fail(TC_EINVAL, "tc_measure_dmma_peak: NULL output");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_measure_dmma_peak: cannot select CUDA device");
  int n_sm = 0;
  int rc = device_sms(device, &n_sm);
  if (rc) return rc;
  double* scratch;
  TC_CUDA(cudaMalloc(reinterpret_cast<void**>(&scratch), (size_t)n_sm * 512 * sizeof(double)));
  cudaEvent_t e0, e1;
  TC_CUDA(cudaEventCreate(&e0));
  TC_CUDA(cudaEventCreate(&e1));
  const int iters = 1 << 14;
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    TC_CUDA(cudaEventRecord(e0));
    dmma_peak_kernel<<<n_sm, 512>>>(scratch, iters);
    TC_CUDA(cudaEventRecord(e1));
    TC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    TC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double tf = (double)n_sm * 16 * iters * 8 * 512.0 / (ms * 1e-3) * 1e-12;
    if (rep > 0) best = std::max(best, tf);
  }
  (void)cudaEventDestroy(e0);
  (void)cudaEventDestroy(e1);
  (void)cudaFree(scratch);
  *tflops = best;
  return TC_OK;
}

}  // extern "C"
