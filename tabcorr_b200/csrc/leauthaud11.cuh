// tabcorr_b200 -- leauthaud11 / hearin15 occupation kernel (family 1).
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md).
#pragma once

#include "common.cuh"
#include "device_math.cuh"
#include "occupation.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// family 1: Leauthaud11 occupations (halotools Leauthaud11Cens / Leauthaud11Sats over the
// Behroozi10SmHm stellar-to-halo-mass relation; call sites tabcorr.py:556-563)
//
//   <N_cen>(M) = 0.5 (1 - erf((log10 M*_thr - log10 M*(M)) / (sqrt(2) sigma_logM*)))
//   <N_sat>(M) = exp(-M_cut / (M h)) (M h / M_sat)^alphasat  [x <N_cen>(M) if modulate_with_cenocc]
//   M_sat = 1e12 bsat (M_knee / 1e12)^betasat,  M_cut = 1e12 bcut (M_knee / 1e12)^betacut,
//   M_knee = h 10^(log10 M_h(M*_thr)),  h = 0.72 in Leauthaud11Sats (its own littleh; the
//   stellar-to-halo-mass relation Behroozi10SmHm converts with h = 0.7)
// log10 M_h(log10 M*) is Behroozi et al. (2010) eq. 21 with parameters x_0 + x_a (a - 1) at the
// model redshift.  halotools inverts it numerically: it tabulates log10 M_h on the 100 knots
// log10 M* = linspace(8.5, 12.5, 100) and evaluates the interpolating cubic spline (scipy
// InterpolatedUnivariateSpline, k = 3: not-a-knot end conditions, cubic extrapolation) of log10 M*
// over log10 M_h.  Parity means reproducing that spline, not the exact inverse, so every draw
// builds the same table and solves the same not-a-knot system (Thomas algorithm) in shared
// memory; a mass bin is then one 7-step binary search, and a node a short walk and one cubic.
// A draw whose table is not strictly increasing (halotools raises there) gets NaN occupations.
//
// This family does not run inside the fused kernel (its per-draw spline does not fit beside the W
// tiles): occupation_l11_kernel writes occ[B, N] and the contraction runs on the occupation
// input of predict_kernel.  Restated from memory of halotools -- parity unpinned, see DESIGN.md.
// ------------------------------------------------------------------------------------------
constexpr int kL11Knots = 100;
constexpr int kL11DrawsPerBlock = 64;     // 64 x 3.2 KB of spline tables + math tables < 227 KB
constexpr double kL11LittleH = 0.7;        // Behroozi10SmHm.littleh
constexpr double kL11LittleHSats = 0.72;   // Leauthaud11Sats.littleh
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
                              log10(kL11LittleHSats) - 12.0;
      const double msat = 1e12 * th[12 * theta_ps] * exp10(th[15 * theta_ps] * log_knee);
      const double mcut = 1e12 * th[13 * theta_ps] * exp10(th[14 * theta_ps] * log_knee);
      draws[b].inv_scatter = 1.0 / (1.4142135623730951 * th[10 * theta_ps]);
      draws[b].neg_mcut_h = -mcut / kL11LittleHSats;
      draws[b].ln_h_over_msat = log(kL11LittleHSats / msat);
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

// quadrature nodes in flight per lane: kOccUnroll when it divides n_gauss, else 2 (see build_plan)
template <bool DECORATED, bool MODULATE, typename Store>
__device__ __forceinline__ void occupation_item_l11_any(const OccArgs& args, const L11Draw* d,
                                                        int g_begin, int g_end,
                                                        const double* __restrict__ tab,
                                                        Store store) {
  if (args.plan.unroll == kOccUnroll)
    occupation_item_l11<DECORATED, MODULATE, kOccUnroll>(args.plan, args.model, d, g_begin, g_end,
                                                         tab, store);
  else
    occupation_item_l11<DECORATED, MODULATE, 2>(args.plan, args.model, d, g_begin, g_end, tab,
                                                store);
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
      const bool mod = args.model.modulate_with_cenocc != 0;
      if (args.model.decorated) {
        if (mod) occupation_item_l11_any<true, true>(args, draws + b, g_begin, g_end, tab, store);
        else occupation_item_l11_any<true, false>(args, draws + b, g_begin, g_end, tab, store);
      } else {
        if (mod) occupation_item_l11_any<false, true>(args, draws + b, g_begin, g_end, tab, store);
        else occupation_item_l11_any<false, false>(args, draws + b, g_begin, g_end, tab, store);
      }
    }
  }
}

}  // namespace
