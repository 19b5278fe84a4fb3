// tabcorr_b200 -- zheng07-family occupations, Heaviside assembly bias, Gauss-Legendre averaging
//   per mass bin, standalone occupation kernel.
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md).
#pragma once

#include "common.cuh"
#include "device_math.cuh"

namespace {

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
                                                     const double* __restrict__ tab, Store store,
                                                     int first, int stride) {
  DrawParams p = load_draw(theta_row, theta_ps);
  if (!model.decorated) p.a_cen = p.a_sat = 0.0;  // strengths are ignored unless decorated
  const bool sat = g_begin >= plan.n_cen_groups;
  for (int grp = g_begin + first; grp < g_end; grp += stride) {
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

// `first` / `stride`: which groups of the range this lane takes.  A batch gives a lane the draw
// lane & 7 and the groups (lane >> 3) + 4 k; a one-draw call (the latency path) has a single
// parameter set, so all 32 lanes spread over the groups (first = lane, stride = 32) instead of
// 24 of them recomputing the same draw.
template <typename Store>
__device__ __forceinline__ void occupation_item(const OccPlan& plan, const tc_model& model,
                                                const double* __restrict__ theta_row,
                                                long long theta_ps, int g_begin, int g_end,
                                                const double* __restrict__ tab, Store store,
                                                int first, int stride) {
#define TC_OCC_ITEM(DEC_, MOD_, U_)                                                              \
  occupation_item_impl<DEC_, MOD_, U_>(plan, model, theta_row, theta_ps, g_begin, g_end, tab,   \
                                       store, first, stride)
  if (model.modulate_with_cenocc) {   // rare: keep one generic instantiation
    TC_OCC_ITEM(true, true, 2);
  } else if (plan.unroll == kOccUnroll) {
    if (model.decorated) TC_OCC_ITEM(true, false, kOccUnroll);
    else TC_OCC_ITEM(false, false, kOccUnroll);
  } else if (model.decorated) {
    TC_OCC_ITEM(true, false, 2);
  } else {
    TC_OCC_ITEM(false, false, 2);
  }
#undef TC_OCC_ITEM
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
                    },
                    threadIdx.x >> 3 & 3, 4);
  }
}

}  // namespace
