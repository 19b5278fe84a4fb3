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
  double s_ord[TC_MAX_KNOTS];   // mass-dependent decoration: the strength ordinates of the item's type
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

// Mass-dependent decoration (tc_model.n_strength / n_split, include/tabcorr_b200.h)
__device__ __forceinline__ bool model_mass_dependent(const tc_model& m) {
  return m.decorated && (m.n_strength[0] > 1 || m.n_strength[1] > 1 || m.n_split[0] > 0 ||
                         m.n_split[1] > 0);
}
// doubles per draw of a zheng07-family model and where a type's strength ordinates start
__device__ __host__ __forceinline__ int zheng07_strength_count(const tc_model& m, int type) {
  return m.n_strength[type] > 1 ? m.n_strength[type] : 1;
}
__device__ __host__ __forceinline__ int zheng07_n_theta(const tc_model& m) {
  return TC_N_THETA_ZHENG07_BASE + zheng07_strength_count(m, 0) + zheng07_strength_count(m, 1);
}
// interpolating polynomial through n <= TC_MAX_KNOTS points (Lagrange form)
__device__ __forceinline__ double lagrange_eval(int n, const double* x, const double* y, double t) {
  double sum = 0.0;
#pragma unroll 1
  for (int k = 0; k < n; k++) {
    double w = y[k];
#pragma unroll 1
    for (int j = 0; j < n; j++)
      if (j != k) w *= (t - x[j]) / (x[k] - x[j]);
    sum += w;
  }
  return sum;
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

// The parameters one galaxy type needs (exp10 and the divisions are ~50 FP64 instructions each)
template <bool SAT>
__device__ __forceinline__ DrawParams load_draw_typed(const double* __restrict__ theta,
                                                      long long ps, bool everything) {
  DrawParams p{};
  if (!SAT || everything) {
    p.logMmin = theta[0];
    p.inv_sigma = 1.0 / theta[ps];
    p.a_cen = fmin(fmax(theta[5 * ps], -1.0), 1.0);
  }
  if (SAT || everything) {
    p.m0 = exp10(theta[2 * ps]);
    p.inv_m1 = 1.0 / exp10(theta[3 * ps]);
    p.alpha = theta[4 * ps];
    p.a_sat = fmin(fmax(theta[6 * ps], -1.0), 1.0);
  }
  return p;
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
  const bool sat = g_begin >= plan.n_cen_groups;
  // only what the item's galaxy type needs: exp10 and the divisions are ~50 FP64 instructions each
  DrawParams p = sat ? load_draw_typed<true>(theta_row, theta_ps, MODULATE)
                     : load_draw_typed<false>(theta_row, theta_ps, false);
  if (!model.decorated) p.a_cen = p.a_sat = 0.0;  // strengths are ignored unless decorated
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
// Series evaluation of the Gauss-Legendre averages (zheng07 family).
//
// The average of a mass bin is the finite sum  occ_i = sum_j c_ij f(node_j)  (tabcorr.py:568-578).
// Evaluating f at every node costs G table-driven erf / pow per group and draw; both functions are
// analytic over the (narrow) bin, so the same finite sum is rewritten -- exactly, up to a
// truncation bounded below 1e-14 -- as a short series in per-row MOMENTS of the nodes that the
// plan precomputes in long double (build_plan):
//
//   centrals   f = Phi(x), x_j = x_c + h tau_j (tau_j in [-1, 1], h = half bin width / sigma):
//              sum_j c_ij Phi(x_j) = mu_0 Phi(x_c)
//                                    + exp(-x_c^2) / sqrt(pi) sum_k H_{k-1}(x_c) (-h)^k... (below)
//              with Phi^(k)(x) = (-1)^(k-1) H_{k-1}(x) exp(-x^2) / sqrt(pi) (H: physicists'
//              Hermite polynomials) and mu_k = sum_j c_ij tau_j^k.  The terms
//              v_k = (-1)^(k-1) H_{k-1}(x_c) h^k obey v_{k+1} = -2 x_c h v_k - 2 (k - 1) h^2 v_{k-1};
//              the plan stores mu_k / k!.  One erf and one exp per GROUP instead of G erf.
//   satellites f = ((M - M0) / M1)^alpha, M_j = m_ref (1 + u_max s_j), all M_j > M0:
//              sum_j c_ij f(M_j) = f(m_ref) sum_k binom(alpha, k) y^k nu_k,  y = u_max m_ref /
//              (m_ref - M0), nu_k = sum_j c_ij s_j^k; the plan stores nu_k / k!.  One pow per group.
//
// The number of terms comes from rigorous tail bounds tabulated per plan (cen_terms: Cramer's
// bound on the Hermite functions, in buckets of h; sat_terms: the binomial coefficients for
// 0 <= alpha <= 4, in buckets of y).  The cost no longer depends on n_gauss_prim.
// Where the series does not apply -- sigma so small that h leaves the table, a bin that
// straddles M0 or lies just above it, the one bin where the Heaviside perturbation changes its
// branch, alpha outside [0, 4], modulate_with_cenocc -- the (draw, group) pair is queued in shared
// memory and the queue is worked off 32 pairs at a time by the node-by-node code above, so the
// hard pairs of the 8 draws of an item share dense warp iterations instead of stalling the others.
//
// Lane mapping: a warp iteration is ONE draw x 32 consecutive groups (sigma, hence the number of
// terms, is warp-uniform; plan arrays are read coalesced).
// ------------------------------------------------------------------------------------------
constexpr int kSerMaxTerms = 24;               // moments 0..24 per row
constexpr int kSerMom = kSerMaxTerms + 1;
constexpr int kSerQueue = 64;                  // queued (draw, group) pairs per warp
#ifndef TC_SER_UNROLL
#define TC_SER_UNROLL 4
#endif
#ifndef TC_SER_DRAWS
#define TC_SER_DRAWS 2
#endif
constexpr int kSerDraws = TC_SER_DRAWS;        // draws in flight per warp iteration
constexpr int kSerUnroll = TC_SER_UNROLL;      // unrolling of the term loops
constexpr int kSerBuckets = 64;
constexpr double kSerCenBucket = 64.0;         // h buckets of width 1/64 up to h = 1
constexpr double kSerSatBucket = 128.0;        // y buckets of width 1/128 up to y = 0.5
constexpr double kSerAlphaMax = 4.0;
constexpr double kInvSqrtPi = 0.56418958354775628695;

// 1 / x to full double precision for a normal positive x: hardware seed + two Newton steps
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// Node-by-node evaluation of one (draw, group) pair: the reference arithmetic, used for the pairs
// the series does not cover.  One copy per translation unit (not inlined into the 24 kernel
// instantiations).
// One (draw, group) pair of a model with mass-dependent strength / split: both are evaluated per
// quadrature node at log10 of the node's mass (halotools evaluates them per halo).
__device__ __forceinline__ void occupation_pair_nodes_massdep(const OccPlan& plan,
                                                              const tc_model& model, int grp,
                                                              bool sat, const DrawParams& p,
                                                              const double* __restrict__ tab,
                                                              double* occ) {
  const int G = plan.n_gauss, GP = plan.n_gauss_pad, type = sat ? 1 : 0;
  const int row0 = plan.grp_rows[grp * kGroupRows], row1 = plan.grp_rows[grp * kGroupRows + 1];
  const double* c0 = plan.row_c + (size_t)row0 * GP;
  const double* c1 = plan.row_c + (size_t)(row1 >= 0 ? row1 : plan.zero_row) * GP;
  const double* node_m = plan.node_m + (size_t)grp * GP;
  const double* node_logm = plan.node_logm + (size_t)grp * GP;
  const double pct0 = plan.row_pct[row0], pct1 = row1 >= 0 ? plan.row_pct[row1] : 0.0;
  const double hi = sat ? CUDART_INF : 1.0;
  const bool with_erf = !sat || model.modulate_with_cenocc;
  const int n_str = model.n_strength[type], n_split = model.n_split[type];
  double a0 = 0.0, a1 = 0.0;
#pragma unroll 1
  for (int g = 0; g < G; g++) {
    const double logm = node_logm[g];
    double f = 1.0;
    if (with_erf) f = half_erfc_neg((logm - p.logMmin) * p.inv_sigma, tab);
    if (sat) {
      const double d = node_m[g] - p.m0;
      const bool pos = d > 0.0;
      const double pw = pow_pos(pos ? d * p.inv_m1 : 1.0, p.alpha, tab);
      f = pos ? f * pw : 0.0;
    }
    double strength = p.s_ord[0];
    if (n_str > 1) strength = lagrange_eval(n_str, model.strength_abscissa[type], p.s_ord, logm);
    strength = fmin(fmax(strength, -1.0), 1.0);
    double split = model.split;
    if (n_split > 0)
      split = fmin(fmax(lagrange_eval(n_split, model.split_abscissa[type],
                                      model.split_ordinates[type], logm), 0.0), 1.0);
    const bool split_ok = split > 0.0 && split < 1.0;
    const double ratio = split_ok ? split / (1.0 - split) : 0.0;
    const double down = split_ok ? -(1.0 - split) / split : 0.0;
    const double dl = assembias_delta(f, strength, ratio, hi, split_ok);
    a0 = fma(c0[g], fma(pct0 > split ? 1.0 : down, dl, f), a0);
    a1 = fma(c1[g], fma((row1 >= 0 && pct1 > split) ? 1.0 : down, dl, f), a1);
  }
  occ[0] = a0;
  occ[1] = a1;
}

__device__ __noinline__ void occupation_pair_nodes(const OccPlan plan, int grp, bool sat,
                                                   bool decorated, bool modulate, DrawParams p,
                                                   double split, const double* __restrict__ tab,
                                                   double* occ) {
  double occ0, occ1;
#define TC_PAIR(SAT_, DEC_, MOD_, U_) \
  occupation_group<SAT_, DEC_, MOD_, U_>(plan, grp, p, split, tab, occ0, occ1)
  const bool wide = plan.unroll == kOccUnroll;
  if (sat) {
    if (modulate) TC_PAIR(true, true, true, 2);
    else if (decorated) { if (wide) TC_PAIR(true, true, false, kOccUnroll); else TC_PAIR(true, true, false, 2); }
    else { if (wide) TC_PAIR(true, false, false, kOccUnroll); else TC_PAIR(true, false, false, 2); }
  } else {
    if (decorated) { if (wide) TC_PAIR(false, true, false, kOccUnroll); else TC_PAIR(false, true, false, 2); }
    else { if (wide) TC_PAIR(false, false, false, kOccUnroll); else TC_PAIR(false, false, false, 2); }
  }
#undef TC_PAIR
  occ[0] = occ0;
  occ[1] = occ1;
}

// Occupation item of one warp in series mode: draws [0, n_b) of an 8-draw block (parameters of
// draw b at theta0 + b * theta_ds) times the groups [g_begin, g_end) of ONE galaxy type.
// `queue` points at kSerQueue ints of shared memory owned by this warp.
// store(b, group, padded_row, occ, n_h) receives the Gauss-Legendre averaged occupation of each row.
template <bool SAT, typename Store>
__device__ __forceinline__ void occupation_item_series(const OccPlan& plan, const tc_model& model,
                                                       const double* __restrict__ theta0,
                                                       long long theta_ds, long long theta_ps,
                                                       int n_b, int g_begin, int g_end,
                                                       const double* __restrict__ tab, int* queue,
                                                       Store store) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool modulate = SAT && model.modulate_with_cenocc != 0;
  const bool decorated = model.decorated != 0;
  // lane l holds the parameters of draw l & 7 (clamped to the item's draws)
  DrawParams mine = load_draw_typed<SAT>(theta0 + (long long)min(lane & 7, n_b - 1) * theta_ds,
                                         theta_ps, modulate);
  if (!decorated) mine.a_cen = mine.a_sat = 0.0;   // strengths are ignored unless decorated
  const double split = model.split;
  const bool split_ok = decorated && split > 0.0 && split < 1.0;
  const double ratio = split_ok ? split / (1.0 - split) : 0.0;
  const double k_down = split_ok ? -(1.0 - split) / split : 0.0;
  const int ng = plan.n_groups;
  int head = 0, count = 0;   // ring buffer of queued pairs (warp-uniform)

  auto drain = [&](int n) {  // node-by-node evaluation of the first n queued pairs
    __syncwarp();
    const int entry = queue[(head + min(lane, n - 1)) & (kSerQueue - 1)];
    const int b = entry >> 24, grp = entry & 0xffffff;
    DrawParams p;
    p.logMmin = __shfl_sync(full, mine.logMmin, b);
    p.inv_sigma = __shfl_sync(full, mine.inv_sigma, b);
    p.m0 = __shfl_sync(full, mine.m0, b);
    p.inv_m1 = __shfl_sync(full, mine.inv_m1, b);
    p.alpha = __shfl_sync(full, mine.alpha, b);
    p.a_cen = __shfl_sync(full, mine.a_cen, b);
    p.a_sat = __shfl_sync(full, mine.a_sat, b);
    if (lane < n) {
      double occ[2];
      occupation_pair_nodes(plan, grp, SAT, decorated, modulate, p, split, tab, occ);
      const int row0 = plan.grp_rows[grp * kGroupRows], row1 = plan.grp_rows[grp * kGroupRows + 1];
      store(b, grp, row0, occ[0], plan.row_nh[row0]);
      if (row1 >= 0) store(b, grp, row1, occ[1], plan.row_nh[row1]);
    }
    __syncwarp();
    head = (head + n) & (kSerQueue - 1);
    count -= n;
  };

  // kSerDraws draws in flight per warp iteration: the series of one (draw, group) pair is a chain
  // of ~100 dependent FP64 operations, and beside DMMA warps a dependent DFMA gets an issue turn
  // only every ~24-32 cycles (tools/fp64_mix.cu), so independent chains are what hides it
  for (int b0 = 0; b0 < n_b; b0 += kSerDraws) {
    // ---- per-draw constants (warp-uniform) -------------------------------------------------
    double strength[kSerDraws], logMmin[kSerDraws], inv_sigma[kSerDraws], m0[kSerDraws],
        inv_m1[kSerDraws], alpha[kSerDraws];
    bool all_queued[kSerDraws];
    int terms_of[kSerDraws];   // a draw only adds its own terms: results do not depend on which
    int n_terms = 2;           // draws share the iteration (bitwise batch invariance)
#pragma unroll
    for (int u = 0; u < kSerDraws; u++) {
      const int b = min(b0 + u, n_b - 1);
      strength[u] = __shfl_sync(full, SAT ? mine.a_sat : mine.a_cen, b);
      all_queued[u] = modulate;
      if (!SAT) {
        logMmin[u] = __shfl_sync(full, mine.logMmin, b);
        inv_sigma[u] = __shfl_sync(full, mine.inv_sigma, b);
        const double hb =
            fmin(fabs(plan.cen_d_max * inv_sigma[u]) * kSerCenBucket, kSerBuckets - 1.0);
        const int terms = plan.cen_terms[(int)hb];   // a NaN lands in the last bucket (fmin)
        terms_of[u] = 2;
        if (!(hb < kSerBuckets - 1.0) || terms == 255) all_queued[u] = true;
        else terms_of[u] = terms;
        n_terms = max(n_terms, terms_of[u]);
      } else {
        m0[u] = __shfl_sync(full, mine.m0, b);
        inv_m1[u] = __shfl_sync(full, mine.inv_m1, b);
        alpha[u] = __shfl_sync(full, mine.alpha, b);
        if (!(alpha[u] >= 0.0 && alpha[u] <= kSerAlphaMax)) all_queued[u] = true;
      }
    }
    for (int g0 = g_begin; g0 < g_end; g0 += 32) {
      const int grp = g0 + lane;
      const bool valid = grp < g_end;
      const int gsafe = valid ? grp : g_end - 1;
      const double4 gs = plan.grp_ser[gsafe];
      const double2* mom = plan.grp_mom + gsafe;
      const double2 mom0 = mom[0];
      const int row0 = plan.grp_rows[gsafe * kGroupRows], row1 = plan.grp_rows[gsafe * kGroupRows + 1];
      double k0 = 0.0, k1 = 0.0;
      if (split_ok) {
        k0 = plan.row_pct[row0] > split ? 1.0 : k_down;
        k1 = (row1 >= 0 && plan.row_pct[max(row1, 0)] > split) ? 1.0 : k_down;
      }
      double f0[kSerDraws], f1[kSerDraws];
      bool queued[kSerDraws];
      if (!SAT) {
        // Hermite series around the centre of the group's nodes
        double phi[kSerDraws], e[kSerDraws], h[kSerDraws], a[kSerDraws], b2[kSerDraws],
            v_prev[kSerDraws], v[kSerDraws], acc0[kSerDraws], acc1[kSerDraws];
        double2 m = mom[ng], m2 = mom[2 * ng];
#pragma unroll
        for (int u = 0; u < kSerDraws; u++) {
          const double xc = (gs.x - logMmin[u]) * inv_sigma[u];
          h[u] = gs.y * inv_sigma[u];
          phi[u] = half_erfc_neg(xc, tab);
          e[u] = exp_scaled(fmax(-xc * xc, -1400.0), tab) * kInvSqrtPi;
          a[u] = -2.0 * xc * h[u];
          b2[u] = -2.0 * h[u] * h[u];
          v_prev[u] = h[u];
          v[u] = a[u] * h[u];
          acc0[u] = fma(v[u], m2.x, v_prev[u] * m.x);
          acc1[u] = fma(v[u], m2.y, v_prev[u] * m.y);
        }
        // Terms 3 .. n_terms.  A draw adds only its own terms (results must not depend on which
        // draws share the iteration): when the smaller count is reached the finished chain's
        // recurrence is zeroed (0 * moment adds exactly nothing) and the other runs on.
        static_assert(kSerDraws <= 2, "the two-phase term loop handles at most two chains");
        const double2* mp = mom + 3 * ng;      // the plan pads the moment table by two rows
        double2 m_next = *mp;
        int k_first = n_terms;
#pragma unroll
        for (int u = 0; u < kSerDraws; u++) k_first = min(k_first, terms_of[u]);
        // (two passes over ONE copy of the loop: the compiler turns a test inside the loop into
        // selects on every term)
        int k = 2;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
          const int k_end = pass == 0 ? k_first : n_terms;
#pragma unroll kSerUnroll
          for (; k < k_end; k++) {   // term k + 1
            m = m_next;
            mp += ng;
            m_next = *mp;
            const double km1 = (double)(k - 1);
#pragma unroll
            for (int u = 0; u < kSerDraws; u++) {
              const double vn = fma(a[u], v[u], (b2[u] * km1) * v_prev[u]);
              acc0[u] = fma(vn, m.x, acc0[u]);
              acc1[u] = fma(vn, m.y, acc1[u]);
              v_prev[u] = v[u];
              v[u] = vn;
            }
          }
#pragma unroll
          for (int u = 0; u < kSerDraws; u++)
            if (terms_of[u] == k_first) v[u] = v_prev[u] = 0.0;
        }
#pragma unroll
        for (int u = 0; u < kSerDraws; u++) {
          queued[u] = all_queued[u];
          f0[u] = fma(e[u], acc0[u], mom0.x * phi[u]);
          f1[u] = fma(e[u], acc1[u], mom0.y * phi[u]);
          if (split_ok && strength[u] != 0.0) {
            // Heaviside perturbation (assembias_delta): strength * min(p, q) is linear in f as
            // long as every node of the group is on the same side of the branch point thr; f is
            // within h / sqrt(pi) of phi at every node (|Phi'| <= 1 / sqrt(pi))
            const bool positive = strength[u] > 0.0;
            const double thr = positive ? 1.0 - split : split;
            if (fabs(phi[u] - thr) <= h[u] * kInvSqrtPi * (1.0 + 1e-9) + 1e-13) queued[u] = true;
            const bool above = phi[u] > thr;
            const double d0 = positive ? (above ? mom0.x - f0[u] : ratio * f0[u])
                                       : (above ? ratio * (mom0.x - f0[u]) : f0[u]);
            const double d1 = positive ? (above ? mom0.y - f1[u] : ratio * f1[u])
                                       : (above ? ratio * (mom0.y - f1[u]) : f1[u]);
            f0[u] = fma(k0 * strength[u], d0, f0[u]);
            f1[u] = fma(k1 * strength[u], d1, f1[u]);
          }
        }
      } else {
        // binomial series around m_ref; gs = {m_ref, u_max, lowest node mass, highest node mass}
        double y[kSerDraws], ya[kSerDraws], fref[kSerDraws], p[kSerDraws], acc0[kSerDraws],
            acc1[kSerDraws];
        bool none[kSerDraws];
        int kt_of[kSerDraws];
        double2 m = mom[ng];
#pragma unroll
        for (int u = 0; u < kSerDraws; u++) {
          none[u] = !(gs.w > m0[u]);              // no node above M0: the occupation is 0
          const bool all_above = gs.z > m0[u];
          const double base = all_above ? gs.x - m0[u] : 1.0;
          y[u] = gs.y * gs.x * fast_rcp(base);
          const double yb = fmin(y[u] * kSerSatBucket, kSerBuckets - 1.0);
          const int terms = plan.sat_terms[(int)yb];
          queued[u] = all_queued[u] || !all_above || !(yb < kSerBuckets - 1.0) || terms == 255;
          if (none[u]) queued[u] = false;
          kt_of[u] = valid && !queued[u] && !none[u] ? terms : 0;
          fref[u] = pow_pos(base * inv_m1[u], alpha[u], tab);
          ya[u] = y[u] * alpha[u];
          p[u] = ya[u];
          acc0[u] = p[u] * m.x;
          acc1[u] = p[u] * m.y;
        }
        // Terms 2 .. kw[u], kw[u] = the largest count of draw u's lanes in this window of 32
        // groups (a property of the draw and the window, not of the batch; lanes needing fewer
        // terms add further valid ones).  Two chains as for the centrals.
        int kw[kSerDraws], k_last = 0, k_first = kSerMaxTerms;
#pragma unroll
        for (int u = 0; u < kSerDraws; u++) {
          kw[u] = __reduce_max_sync(full, kt_of[u]);
          k_last = max(k_last, kw[u]);
          k_first = min(k_first, kw[u]);
        }
        const double2* mp = mom + 2 * ng;
        double2 m_next = *mp;
        int k = 2;
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
          const int k_end = pass == 0 ? k_first : k_last;
#pragma unroll kSerUnroll
          for (; k <= k_end; k++) {
            m = m_next;
            mp += ng;
            m_next = *mp;
            const double mkm1 = -(double)(k - 1);
#pragma unroll
            for (int u = 0; u < kSerDraws; u++) {
              p[u] *= fma(mkm1, y[u], ya[u]);   // y (alpha - k + 1); the plan divides by k!
              acc0[u] = fma(p[u], m.x, acc0[u]);
              acc1[u] = fma(p[u], m.y, acc1[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < kSerDraws; u++)
            if (kw[u] == k_first) p[u] = 0.0;
        }
#pragma unroll
        for (int u = 0; u < kSerDraws; u++) {
          f0[u] = none[u] ? 0.0 : fref[u] * (mom0.x + acc0[u]);
          f1[u] = none[u] ? 0.0 : fref[u] * (mom0.y + acc1[u]);
          if (split_ok) {
            // satellites have no upper bound: min(p, q) is ratio * f (strength > 0) or f
            const double sr = strength[u] > 0.0 ? strength[u] * ratio : strength[u];
            f0[u] = fma(k0 * sr, f0[u], f0[u]);
            f1[u] = fma(k1 * sr, f1[u], f1[u]);
          }
        }
      }
      const double nh0 = plan.row_nh[row0], nh1 = plan.row_nh[max(row1, 0)];
#pragma unroll
      for (int u = 0; u < kSerDraws; u++) {
        const int b = b0 + u;
        if (b >= n_b) break;                 // warp-uniform
        const bool q = queued[u] && valid;
        if (valid && !q) {
          store(b, grp, row0, f0[u], nh0);
          if (row1 >= 0) store(b, grp, row1, f1[u], nh1);
        }
        const unsigned qm = __ballot_sync(full, q);
        if (qm) {
          if (q)
            queue[(head + count + __popc(qm & ((1u << lane) - 1u))) & (kSerQueue - 1)] =
                (b << 24) | grp;
          count += __popc(qm);
          if (count >= 32) drain(32);
        }
      }
    }
  }
  if (count > 0) drain(count);
}

// Items of an 8-draw block in series mode: item q < n_ranges_cen works on centrals, the others on
// satellites; a type's items are `pieces` draw pieces x (n_ranges / pieces) group ranges cut in
// units of 32 groups (one warp iteration).
struct SeriesItem {
  bool sat;
  int b_begin, n_b, g_begin, g_end;
};
__device__ __forceinline__ SeriesItem series_item(const OccPlan& plan, int n_ranges_cen,
                                                  int n_ranges_sat, int pieces_cen,
                                                  int pieces_sat, int q) {
  SeriesItem it;
  it.sat = q >= n_ranges_cen;
  const int first = it.sat ? plan.n_cen_groups : 0;
  const int count = it.sat ? plan.n_groups - plan.n_cen_groups : plan.n_cen_groups;
  const int pieces = it.sat ? pieces_sat : pieces_cen;
  const int ranges = (it.sat ? n_ranges_sat : n_ranges_cen) / pieces;
  const int local = it.sat ? q - n_ranges_cen : q;
  const int piece = local % pieces, range = local / pieces;
  it.b_begin = 8 * piece / pieces;
  it.n_b = 8 * (piece + 1) / pieces - it.b_begin;
  const int units = (count + 31) >> 5;
  it.g_begin = first + min(count, 32 * (int)((long long)units * range / ranges));
  it.g_end = first + min(count, 32 * (int)((long long)units * (range + 1) / ranges));
  return it;
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
  int n_ranges_cen, n_ranges_sat;   // items per 8-draw block and galaxy type
  int pieces_cen, pieces_sat;       // series mode: draw pieces per type (series_item)
  const int* pad_to_row;
  double* occ_out;
};

__global__ void __launch_bounds__(kThreads, 1) occupation_kernel(const OccArgs args) {
  __shared__ double tab[kTabDoubles];
  __shared__ int queue[kWarps][kSerQueue];
  load_math_tables(tab);
  __syncthreads();
  // one warp per item = a piece of an 8-draw block x a range of groups of one galaxy type
  const int n_ranges = args.n_ranges_cen + args.n_ranges_sat;
  const long long n_blocks = (args.n_draws + 7) / 8;
  const long long n_items = n_blocks * n_ranges;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  for (long long item = warp0; item < n_items; item += (long long)gridDim.x * kWarps) {
    const long long block = item / n_ranges;
    const int q = (int)(item - block * n_ranges);
    const SeriesItem it = series_item(args.plan, args.n_ranges_cen, args.n_ranges_sat,
                                      args.pieces_cen, args.pieces_sat, q);
    const long long draw0 = block * 8 + it.b_begin;
    const int n_b = (int)min((long long)it.n_b, args.n_draws - draw0);
    if (n_b <= 0 || it.g_end <= it.g_begin) continue;
    auto store = [&](int b, int, int row, double occ, double) {
      const int dst = args.pad_to_row[row];
      if (dst >= 0) args.occ_out[(draw0 + b) * args.n_rows + dst] = occ;
    };
    if (it.sat)
      occupation_item_series<true>(args.plan, args.model, args.theta + draw0 * args.theta_ds,
                                   args.theta_ds, args.theta_ps, n_b, it.g_begin, it.g_end, tab,
                                   queue[threadIdx.x >> 5], store);
    else
      occupation_item_series<false>(args.plan, args.model, args.theta + draw0 * args.theta_ds,
                                    args.theta_ds, args.theta_ps, n_b, it.g_begin, it.g_end, tab,
                                    queue[threadIdx.x >> 5], store);
  }
}

// Models with mass-dependent assembly-bias strength / split (tc_model.n_strength / n_split): one
// thread per (draw, group) pair, node by node (the series of occupation_item_series needs a
// strength and a split that are constant over the bin).
__global__ void __launch_bounds__(256) occupation_massdep_kernel(const OccArgs args) {
  __shared__ double tab[kTabDoubles];
  load_math_tables(tab);
  __syncthreads();
  const OccPlan& plan = args.plan;
  const long long n_pairs = args.n_draws * plan.n_groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n_pairs;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long draw = idx / plan.n_groups;
    const int grp = (int)(idx - draw * plan.n_groups);
    const bool sat = grp >= plan.n_cen_groups;
    const double* th = args.theta + draw * args.theta_ds;
    const long long ps = args.theta_ps;
    DrawParams p{};
    p.logMmin = th[0];
    p.inv_sigma = 1.0 / th[ps];
    p.m0 = exp10(th[2 * ps]);
    p.inv_m1 = 1.0 / exp10(th[3 * ps]);
    p.alpha = th[4 * ps];
    const int first = TC_N_THETA_ZHENG07_BASE + (sat ? zheng07_strength_count(args.model, 0) : 0);
    const int n_ord = zheng07_strength_count(args.model, sat ? 1 : 0);
    for (int k = 0; k < TC_MAX_KNOTS; k++) p.s_ord[k] = k < n_ord ? th[(first + k) * ps] : 0.0;
    double occ[2];
    occupation_pair_nodes_massdep(plan, args.model, grp, sat, p, tab, occ);
    const int row0 = plan.grp_rows[grp * kGroupRows], row1 = plan.grp_rows[grp * kGroupRows + 1];
    const int dst0 = args.pad_to_row[row0];
    if (dst0 >= 0) args.occ_out[draw * args.n_rows + dst0] = occ[0];
    if (row1 >= 0) {
      const int dst1 = args.pad_to_row[row1];
      if (dst1 >= 0) args.occ_out[draw * args.n_rows + dst1] = occ[1];
    }
  }
}

}  // namespace
