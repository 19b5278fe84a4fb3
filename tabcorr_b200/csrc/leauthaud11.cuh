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
// builds the same table and solves the same not-a-knot system (eliminated from both ends by two
// warps) in shared memory; a mass bin -- the centrals group and the satellites group over the same
// node masses, evaluated together -- is then one 7-step binary search, and a node a short walk,
// one cubic, one erf (five nodes share a column of a wide-interval table) and one exponential.
// A draw whose table is not strictly increasing (halotools raises there) gets NaN occupations.
// Strength / split of the Heaviside decoration and the stellar-mass scatter may depend on mass
// (tc_model.n_strength / n_split / n_scatter): occupation_l11_kernel<true>.
//
// This family does not run inside the fused kernel (its per-draw spline does not fit beside the W
// tiles): occupation_l11_kernel writes occ[B, N] and the contraction runs on the occupation
// input of predict_kernel.  Restated from memory of halotools -- parity unpinned, see DESIGN.md.
// ------------------------------------------------------------------------------------------
constexpr int kL11Knots = 100;
#ifndef TC_L11_THREADS
#define TC_L11_THREADS 192
#endif
#ifndef TC_L11_DRAWS
#define TC_L11_DRAWS 32
#endif
#ifndef TC_L11_MIN_BLOCKS
#define TC_L11_MIN_BLOCKS 2
#endif
// Two CTAs of 6 warps and 32 draws share an SM (2 x 109 KB of shared memory, <= 170 registers per
// thread): while one CTA is in the latency-bound tridiagonal solves the other evaluates nodes.
constexpr int kL11Threads = TC_L11_THREADS;
constexpr int kL11DrawsPerBlock = TC_L11_DRAWS;
constexpr int kL11MinBlocks = TC_L11_MIN_BLOCKS;
#ifndef TC_L11_LANE_BINS
#define TC_L11_LANE_BINS 4
#endif
constexpr int kL11LaneBins = TC_L11_LANE_BINS;        // mass bins per warp iteration (power of two)
constexpr int kL11LaneDraws = 32 / kL11LaneBins;      // draws per warp iteration
constexpr double kL11LittleH = 0.7;        // Behroozi10SmHm.littleh
constexpr double kL11LittleHSats = 0.72;   // Leauthaud11Sats.littleh
constexpr double kL11LogMsLo = 8.5, kL11LogMsHi = 12.5;
constexpr double kLn10 = 2.302585092994045684;

// doubles per draw: the 16 occupation parameters, then the assembly-bias strength ordinates of the
// centrals and of the satellites (one each unless the strength depends on mass)
constexpr int kL11Base = 16;
__device__ __host__ __forceinline__ int l11_n_theta(const tc_model& m) {
  return kL11Base + zheng07_strength_count(m, 0) + zheng07_strength_count(m, 1) +
         (m.n_scatter > 1 ? m.n_scatter - 1 : 0);
}
// models that run in occupation_l11_kernel<true>: strength, split or scatter depend on mass
__device__ __host__ __forceinline__ bool l11_mass_dependent(const tc_model& m) {
  return m.n_scatter > 1 || (m.decorated && (m.n_strength[0] > 1 || m.n_strength[1] > 1 ||
                                             m.n_split[0] > 0 || m.n_split[1] > 0));
}
constexpr int kL11OrdPerDraw = 3 * TC_MAX_KNOTS;   // centrals strengths | satellites strengths | scatter

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
// (MASSDEP: + the ordinates of a mass-dependent decoration / scatter, 3 x TC_MAX_KNOTS doubles per
// draw)
constexpr int kL11OrdDoubles = kL11DrawsPerBlock * kL11OrdPerDraw;
constexpr size_t l11_smem_bytes(bool massdep) {
  return (kL11TabDoubles + kL11Knots + (massdep ? kL11OrdDoubles : 0)) * sizeof(double) +
         kL11DrawsPerBlock * sizeof(L11Draw);
}
static_assert(kL11MinBlocks * (l11_smem_bytes(true) + 1024) <= 228 * 1024,
              "CTAs per SM vs shared memory");

__device__ __forceinline__ double l11_knot_logms(int k) {
  // numpy.linspace(8.5, 12.5, 100): arange(100) * step + start, last element set to stop
  return k == kL11Knots - 1 ? kL11LogMsHi
                            : (double)k * ((kL11LogMsHi - kL11LogMsLo) / (kL11Knots - 1)) + kL11LogMsLo;
}

__device__ __forceinline__ double l11_knot_logms_inner(int k) {   // the same for k < kL11Knots - 1
  return (double)k * ((kL11LogMsHi - kL11LogMsLo) / (kL11Knots - 1)) + kL11LogMsLo;
}

// Behroozi10SmHm.mean_log_halo_mass: log10 M_h [h = 1 units] of log10 M* [h = 1 units].  halotools
// converts M* -> M* / h^2 and M_h -> M_h h through 10** and log10; here the conversions are added
// in log space (differences at the 1e-16 level).  FAST: 10^z through the exp table (relative
// error ~1e-16 like the library's, a fifth of its instructions) -- the 100 knots of every draw.
template <bool FAST>
__device__ __forceinline__ double l11_log_halo_mass(double log_ms, double logm0, double logm1,
                                                    double beta, double delta, double gamma,
                                                    const double* __restrict__ tab) {
  const double log_h = -0.15490195998574316929;   // log10(0.7)
  const double lr = log_ms - 2.0 * log_h - logm0;  // log10(M* / M0) in h = 0.7 units
  const double up = FAST ? exp_scaled_with(delta * lr * kLn10, tab + kErfWDoubles) : exp10(delta * lr);
  const double dn = FAST ? exp_scaled_with(-gamma * lr * kLn10, tab + kErfWDoubles) : exp10(-gamma * lr);
  // (the fast reciprocal: two Newton steps on the hardware seed, full precision for normal x)
  return logm1 + beta * lr + (FAST ? up * fast_rcp(1.0 + dn) : up / (1.0 + dn)) - 0.5 + log_h;
}

// Spline tables and per-draw constants of one block of draws, by the whole CTA.  `sk`: the knot
// ordinates log10 M* (shared by all draws), `tab`: the math tables.
template <bool MASSDEP>
__device__ __forceinline__ void l11_prepare_block(L11Draw* draws, double* __restrict__ ords,
                                  const double* __restrict__ sk,
                                  const double* __restrict__ tab, int n_block, long long draw0,
                                  long long n_draws, const double* __restrict__ theta,
                                  long long theta_ds, long long theta_ps, const tc_model& model) {
  constexpr int n = kL11Knots;
  // (0) one thread per draw: relation parameters at the model redshift (parked in the y, z, w
  // slots of knots 0 and 1, free until (2)) and the constants of the occupation functions
  for (int b = threadIdx.x; b < n_block; b += blockDim.x) {
    const double a1 = 1.0 / (1.0 + model.redshift) - 1.0;   // a - 1
    const double* th = theta + min(draw0 + b, n_draws - 1) * theta_ds;
    const double logm0 = fma(th[1 * theta_ps], a1, th[0]);
    const double logm1 = fma(th[3 * theta_ps], a1, th[2 * theta_ps]);
    const double beta = fma(th[5 * theta_ps], a1, th[4 * theta_ps]);
    const double delta = fma(th[7 * theta_ps], a1, th[6 * theta_ps]);
    const double gamma = fma(th[9 * theta_ps], a1, th[8 * theta_ps]);
    L11Draw& D = draws[b];
    D.knot[0].y = logm0;
    D.knot[0].z = logm1;
    D.knot[0].w = beta;
    D.knot[1].y = delta;
    D.knot[1].z = gamma;
    // Leauthaud11Sats._update_satellite_params: knee = h M_h(threshold) / 1e12
    const double log_knee =
        l11_log_halo_mass<false>(model.threshold, logm0, logm1, beta, delta, gamma, tab) +
        log10(kL11LittleHSats) - 12.0;
    const double msat = 1e12 * th[12 * theta_ps] * exp10(th[15 * theta_ps] * log_knee);
    const double mcut = 1e12 * th[13 * theta_ps] * exp10(th[14 * theta_ps] * log_knee);
    D.inv_scatter = 1.0 / (1.4142135623730951 * th[10 * theta_ps]);
    D.neg_mcut_h = -mcut / kL11LittleHSats;
    D.ln_h_over_msat = log(kL11LittleHSats / msat);
    D.alphasat = th[11 * theta_ps];
    // strengths: ordinates [centrals 0..3 | satellites 0..3] (the first of each is the constant
    // strength of the plain decoration; clipped per node when they depend on mass)
    const int n_cen_ord = zheng07_strength_count(model, 0), n_sat_ord = zheng07_strength_count(model, 1);
    if (MASSDEP) {
      double* o = ords + b * kL11OrdPerDraw;
      for (int k = 0; k < TC_MAX_KNOTS; k++) {
        o[k] = model.decorated && k < n_cen_ord ? th[(kL11Base + k) * theta_ps] : 0.0;
        o[TC_MAX_KNOTS + k] =
            model.decorated && k < n_sat_ord ? th[(kL11Base + n_cen_ord + k) * theta_ps] : 0.0;
        o[2 * TC_MAX_KNOTS + k] =
            k == 0 ? th[10 * theta_ps]
                   : k < model.n_scatter ? th[(kL11Base + n_cen_ord + n_sat_ord + k - 1) * theta_ps]
                                         : 0.0;
      }
    }
    D.a_cen = model.decorated ? fmin(fmax(th[kL11Base * theta_ps], -1.0), 1.0) : 0.0;
    D.a_sat = model.decorated ? fmin(fmax(th[(kL11Base + n_cen_ord) * theta_ps], -1.0), 1.0) : 0.0;
    // the erf argument is clamped below (NaN would be lost): a parameter that is not finite makes
    // every occupation of the draw NaN through `bad`
    double sum = 0.0;
    for (int k = 0; k < l11_n_theta(model); k++)   // (the strengths only if they are used)
      if (model.decorated || k < kL11Base || k >= kL11Base + n_cen_ord + n_sat_ord)
        sum += th[k * theta_ps];
    D.bad = sum - sum;
  }
  __syncthreads();
  // (1) knot abscissae
  for (int idx = threadIdx.x; idx < n_block * n; idx += blockDim.x) {
    const int b = idx / n, k = idx - b * n;
    const L11Draw& D = draws[b];
    draws[b].knot[k].x = l11_log_halo_mass<true>(sk[k], D.knot[0].y, D.knot[0].z, D.knot[0].w,
                                                 D.knot[1].y, D.knot[1].z, tab);
  }
  __syncthreads();
  // (2) reciprocal interval widths (y slot), monotonicity
  for (int idx = threadIdx.x; idx < n_block * (n - 1); idx += blockDim.x) {
    const int b = idx / (n - 1), k = idx - b * (n - 1);
    const double h = draws[b].knot[k + 1].x - draws[b].knot[k].x;
    draws[b].knot[k].y = fast_rcp(h);   // (h <= 0 or NaN: the draw is flagged below)
    if (!(h > 0.0)) draws[b].bad = CUDART_NAN;
  }
  __syncthreads();
  // (3) HALVED second derivatives c2_k = m_k / 2 of the not-a-knot spline s(x).  Interior equations
  // h_{i-1} m_{i-1} + 2 (h_{i-1} + h_i) m_i + h_i m_{i+1} = 6 (d_i - d_{i-1}), d_i = (s_{i+1} -
  // s_i) / h_i, with m_0 and m_{n-1} eliminated through the continuity of the third derivative at
  // knots 1 and n - 2; the right-hand side is halved, so the solution is exactly m / 2.
  // The recurrences are latency bound (one lane per draw, one reciprocal per row on the critical
  // path) and the rest of the CTA waits for them, so the system is eliminated FROM BOTH ENDS
  // (twisted factorisation): warp 0 runs the Thomas elimination down rows 1 .. kMid, warp 1 its
  // mirror image up rows n - 2 .. kMid + 1; the two meet in a 2 x 2 system for rows kMid and
  // kMid + 1 and substitute outwards.  Half the sequential length of a one-sided solve.
  // Scratch: knot[i].z = modified off-diagonal, knot[i].w = modified right-hand side, then c2_i.
  constexpr int kMid = (n - 2) / 2;
  static_assert(kL11DrawsPerBlock <= 32 && kL11Threads >= 64, "one lane per draw, two warps");
  const int solver_warp = threadIdx.x >> 5, b_lane = threadIdx.x & 31;
  const bool solving = solver_warp < 2 && b_lane < n_block;
  auto row = [&](int i, double h_prev, double h, double ih_prev, double ih, double& lower,
                 double& diag, double& upper) {
    lower = h_prev;
    diag = 2.0 * (h_prev + h);
    upper = h;
    if (i == 1) {
      const double e = h_prev * h_prev * ih;
      lower = 0.0;
      diag = 3.0 * h_prev + 2.0 * h + e;
      upper = h - e;
    }
    if (i == n - 2) {
      const double e = h * h * ih_prev;
      lower = h_prev - e;
      diag = 2.0 * h_prev + 3.0 * h + e;
      upper = 0.0;
    }
  };
  if (solving && solver_warp == 0) {
    L11Draw& D = draws[b_lane];
    double x_cur = D.knot[1].x;
    double h_prev = x_cur - D.knot[0].x, ih_prev = D.knot[0].y;   // h_0, 1 / h_0
    double d_prev = (sk[1] - sk[0]) * ih_prev;
    double cp = 0.0, rp = 0.0;                                    // c'_{i-1}, r'_{i-1}
#pragma unroll 2
    for (int i = 1; i <= kMid; i++) {
      const double x_next = D.knot[i + 1].x;
      const double h = x_next - x_cur, ih = D.knot[i].y;          // h_i, 1 / h_i
      const double d = (sk[i + 1] - sk[i]) * ih;
      double lower, diag, upper;
      row(i, h_prev, h, ih_prev, ih, lower, diag, upper);
      const double inv = fast_rcp(diag - lower * cp);
      cp = upper * inv;
      rp = (3.0 * (d - d_prev) - lower * rp) * inv;
      D.knot[i].z = cp;
      D.knot[i].w = rp;
      x_cur = x_next;
      h_prev = h;
      ih_prev = ih;
      d_prev = d;
    }
  } else if (solving) {
    L11Draw& D = draws[b_lane];
    double x_cur = D.knot[n - 2].x;
    double h = D.knot[n - 1].x - x_cur, ih = D.knot[n - 2].y;     // h_{n-2}, 1 / h_{n-2}
    double d = (sk[n - 1] - sk[n - 2]) * ih;
    double bq = 0.0, br = 0.0;                                    // mirrored c'_{i+1}, r'_{i+1}
#pragma unroll 2
    for (int i = n - 2; i >= kMid + 1; i--) {
      const double x_before = D.knot[i - 1].x;
      const double h_prev = x_cur - x_before, ih_prev = D.knot[i - 1].y;   // h_{i-1}, 1 / h_{i-1}
      const double d_prev = (sk[i] - sk[i - 1]) * ih_prev;
      double lower, diag, upper;
      row(i, h_prev, h, ih_prev, ih, lower, diag, upper);
      const double inv = fast_rcp(diag - upper * bq);
      bq = lower * inv;
      br = (3.0 * (d - d_prev) - upper * br) * inv;
      D.knot[i].z = bq;
      D.knot[i].w = br;
      x_cur = x_before;
      h = h_prev;
      ih = ih_prev;
      d = d_prev;
    }
  }
  __syncthreads();
  double c2_mid = 0.0, c2_mid1 = 0.0;   // rows kMid and kMid + 1: x_k + c' x_{k+1} = r', q x_k + x_{k+1} = r
  if (solving) {
    const L11Draw& D = draws[b_lane];
    const double cp = D.knot[kMid].z, rp = D.knot[kMid].w;
    const double bq = D.knot[kMid + 1].z, br = D.knot[kMid + 1].w;
    c2_mid1 = (br - bq * rp) / (1.0 - bq * cp);
    c2_mid = rp - cp * c2_mid1;
  }
  __syncthreads();   // both warps have read the four values before either overwrites its own
  if (solving && solver_warp == 0) {
    L11Draw& D = draws[b_lane];
    D.knot[kMid].w = c2_mid;
    double m_next = c2_mid;
    for (int i = kMid - 1; i >= 1; i--) {
      const double m = D.knot[i].w - D.knot[i].z * m_next;
      D.knot[i].w = m;
      m_next = m;
    }
    const double h0 = D.knot[1].x - D.knot[0].x;
    D.knot[0].w = D.knot[1].w - h0 * D.knot[1].y * (D.knot[2].w - D.knot[1].w);
  } else if (solving) {
    L11Draw& D = draws[b_lane];
    D.knot[kMid + 1].w = c2_mid1;
    double m_prev = c2_mid1;
    for (int i = kMid + 2; i <= n - 2; i++) {
      const double m = D.knot[i].w - D.knot[i].z * m_prev;
      D.knot[i].w = m;
      m_prev = m;
    }
    const double ha = D.knot[n - 1].x - D.knot[n - 2].x;
    D.knot[n - 1].w =
        D.knot[n - 2].w + ha * D.knot[n - 3].y * (D.knot[n - 2].w - D.knot[n - 3].w);
  }
  __syncthreads();
  // (4) cubic coefficients per interval: c1 and c3 replace 1 / h (y) and the scratch (z); w = c2
  // of knots k, k + 1 is only read
  for (int idx = threadIdx.x; idx < n_block * (n - 1); idx += blockDim.x) {
    const int b = idx / (n - 1), k = idx - b * (n - 1);
    const double4 k0 = draws[b].knot[k];
    const double x1 = draws[b].knot[k + 1].x, m1 = draws[b].knot[k + 1].w;
    const double h = x1 - k0.x, ih = k0.y;
    const double d = (sk[k + 1] - sk[k]) * ih;
    draws[b].knot[k].y = d - h * (2.0 * k0.w + m1) * (1.0 / 3.0);
    draws[b].knot[k].z = (m1 - k0.w) * ih * (1.0 / 3.0);
  }
  __syncthreads();
}

// Heaviside perturbation of one node when the strength and / or the split of galaxy type `type`
// depend on mass (tc_model.n_strength / n_split): both are the interpolating polynomial of their
// control points at log10 of the node mass, clipped to [-1, 1] / [0, 1].  Returns delta and the
// factors of the two rows (+1 above the split percentile, -(1 - s) / s below).
__device__ __forceinline__ double l11_massdep_delta(const tc_model& model, int type,
                                                    const double* __restrict__ ord, double logm,
                                                    double f, double hi, int row_a, int row_b,
                                                    double pct_a, double pct_b, double& ka,
                                                    double& kb) {
  const int n_str = model.n_strength[type], n_split = model.n_split[type];
  double strength = ord[0];
  if (n_str > 1) {
    double o[TC_MAX_KNOTS];
#pragma unroll
    for (int k = 0; k < TC_MAX_KNOTS; k++) o[k] = ord[k];
    strength = lagrange_eval(n_str, model.strength_abscissa[type], o, logm);
  }
  strength = fmin(fmax(strength, -1.0), 1.0);
  double split = model.split;
  if (n_split > 0)
    split = fmin(fmax(lagrange_eval(n_split, model.split_abscissa[type],
                                    model.split_ordinates[type], logm), 0.0), 1.0);
  const bool split_ok = split > 0.0 && split < 1.0;
  const double ratio = split_ok ? split / (1.0 - split) : 0.0;
  const double down = split_ok ? -(1.0 - split) / split : 0.0;
  ka = (row_a >= 0 && pct_a > split) ? 1.0 : down;
  kb = (row_b >= 0 && pct_b > split) ? 1.0 : down;
  return assembias_delta(f, strength, ratio, hi, split_ok);
}

// One lane: one draw x one mass bin = the centrals group `cen` and / or the satellites group
// `sat` over the same node masses (-1: absent).  The spline and the erf of a node serve both
// galaxy types (Leauthaud11Sats is modulated by <N_cen> by default); the Gauss-Legendre weights
// and the Heaviside decoration are those of occupation_group (occupation.cuh).
// DEC: 0 no decoration, 1 constant strength and split, 2 the general model: strength, split and / or
// the stellar-mass scatter depend on mass (evaluated per node at log10 of the node mass, like
// occupation_pair_nodes_massdep; `ords`: the draw's strength and scatter ordinates; an undecorated
// model has zero strengths).
template <int U, int DEC>
__device__ __forceinline__ void l11_bin(const OccArgs& args, const L11Draw* __restrict__ d,
                                        const double* __restrict__ ords,
                                        const double* __restrict__ tab,
                                        const L11Bin* __restrict__ bin_ptr,
                                        double* __restrict__ out) {
  const OccPlan& plan = args.plan;
  const int G = plan.n_gauss_pad;
  const L11Bin bin = *bin_ptr;
  const int cen = bin.cen, sat = bin.sat;
  const int lead = cen >= 0 ? cen : sat;
  const double* __restrict__ node = plan.node_logm + (size_t)lead * G;
  const double* __restrict__ node_inv = plan.node_inv_m + (size_t)lead * G;
  const bool has_cen = cen >= 0, has_sat = sat >= 0, same_w = bin.same_w != 0;
  const bool modulate = args.model.modulate_with_cenocc != 0;
  const double* __restrict__ w0 = plan.row_c + (size_t)(bin.row[0] >= 0 ? bin.row[0] : plan.zero_row) * G;
  const double* __restrict__ w1 = plan.row_c + (size_t)(bin.row[1] >= 0 ? bin.row[1] : plan.zero_row) * G;
  const double* __restrict__ w2 = plan.row_c + (size_t)(bin.row[2] >= 0 ? bin.row[2] : plan.zero_row) * G;
  const double* __restrict__ w3 = plan.row_c + (size_t)(bin.row[3] >= 0 ? bin.row[3] : plan.zero_row) * G;
  constexpr bool DECORATED = DEC == 1;
  double k0 = 0.0, k1 = 0.0, k2 = 0.0, k3 = 0.0, ratio = 0.0;
  bool split_ok = false;
  if (DECORATED) {
    const double split = args.model.split;
    split_ok = split > 0.0 && split < 1.0;
    ratio = split / (1.0 - split);
    const double down = -(1.0 - split) / split;
    k0 = (bin.row[0] >= 0 && bin.pct[0] > split) ? 1.0 : down;
    k1 = (bin.row[1] >= 0 && bin.pct[1] > split) ? 1.0 : down;
    k2 = (bin.row[2] >= 0 && bin.pct[2] > split) ? 1.0 : down;
    k3 = (bin.row[3] >= 0 && bin.pct[3] > split) ? 1.0 : down;
  }
  // knot interval of the first node (7-step binary search); the nodes of a bin ascend from it and
  // a bin spans few knots: per node three branch-free steps against the cached abscissae, then
  // (rarely) a linear walk
  int hint = 0;   // largest knot index in [0, kL11Knots - 2] with x_i <= node[0] (0 if none)
  {
    const double first = bin.first_logm;
#pragma unroll
    for (int step = 64; step >= 1; step >>= 1) {
      const int j = hint + step;
      if (j <= kL11Knots - 2 && d->knot[j].x <= first) hint = j;
    }
  }
  const double threshold = args.model.threshold, inv_scatter = d->inv_scatter;
  const double alphasat = d->alphasat, ln_h_over_msat = d->ln_h_over_msat;
  const double neg_mcut_h = d->neg_mcut_h;
  const double a_cen = d->a_cen, a_sat = d->a_sat;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
  for (int g = 0; g < G; g += U) {
    // The U nodes of an iteration advance through every stage together (straight-line code, U
    // independent dependency chains): interval, cubic, erf, then the two galaxy types.
    double next_x[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
      next_x[j] = hint + 1 + j <= kL11Knots - 2 ? d->knot[hint + 1 + j].x : CUDART_INF;
    double e[U], logm[U];
    int idx[U];
    bool far = false;
#pragma unroll
    for (int u = 0; u < U; u++) {
      logm[u] = node[g + u];
      idx[u] = hint + (next_x[0] <= logm[u] ? 1 : 0) + (next_x[1] <= logm[u] ? 1 : 0) +
               (next_x[2] <= logm[u] ? 1 : 0);
      far = far || next_x[2] <= logm[u];
    }
    if (far) {   // rare: a node more than three knots beyond the hint
#pragma unroll
      for (int u = 0; u < U; u++)
        while (idx[u] < kL11Knots - 2 && d->knot[idx[u] + 1].x <= logm[u]) idx[u]++;
    }
    hint = idx[U - 1];   // the nodes ascend (zero-weight padding nodes repeat the first one)
#pragma unroll
    for (int u = 0; u < U; u++) {
      const double4 c = d->knot[idx[u]];
      const double t = logm[u] - c.x;
      const double s = fma(t, fma(t, fma(t, c.z, c.w), c.y), l11_knot_logms_inner(idx[u]));
      e[u] = (s - threshold) * inv_scatter;
      if (DEC == 2 && args.model.n_scatter > 1) {   // sigma(log10 M): polynomial through the control points
        double o[TC_MAX_KNOTS];
#pragma unroll
        for (int k = 0; k < TC_MAX_KNOTS; k++) o[k] = ords[2 * TC_MAX_KNOTS + k];
        const double sigma = lagrange_eval(args.model.n_scatter, args.model.scatter_abscissa, o, logm[u]);
        e[u] = (s - threshold) / (1.4142135623730951 * sigma);
      }
    }
    half_erfc_neg_group<U>(e, tab);
    double wa[U], wb[U];   // weights of the centrals rows (the satellites rows' too if same_w)
    if (has_cen) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        wa[u] = w0[g + u];
        wb[u] = w1[g + u];
        if (DEC == 2) {
          double ka, kb;
          const double dl = l11_massdep_delta(args.model, 0, ords, logm[u], e[u], 1.0, bin.row[0],
                                              bin.row[1], bin.pct[0], bin.pct[1], ka, kb);
          acc0 = fma(wa[u], fma(ka, dl, e[u]), acc0);
          acc1 = fma(wb[u], fma(kb, dl, e[u]), acc1);
        } else if (DECORATED) {
          const double dl = assembias_delta(e[u], a_cen, ratio, 1.0, split_ok);
          acc0 = fma(wa[u], fma(k0, dl, e[u]), acc0);
          acc1 = fma(wb[u], fma(k1, dl, e[u]), acc1);
        } else {
          acc0 = fma(wa[u], e[u], acc0);
          acc1 = fma(wb[u], e[u], acc1);
        }
      }
    }
    if (has_sat) {
      double f[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        // exp(-M_cut / (M h)) (M h / M_sat)^alphasat = exp(alphasat (ln M + ln(h / M_sat)) - M_cut / (M h))
        double y = fma(alphasat, fma(logm[u], kLn10, ln_h_over_msat), neg_mcut_h * node_inv[g + u]);
        y = y < -800.0 ? -800.0 : y;   // (plain selects: NaN parameters are reported through `bad`)
        y = y > 800.0 ? 800.0 : y;
        f[u] = exp_scaled_with(y, tab + kErfWDoubles) * (modulate ? e[u] : 1.0);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const double wc = same_w ? wa[u] : w2[g + u], wd = same_w ? wb[u] : w3[g + u];
        if (DEC == 2) {
          double ka, kb;
          const double dl = l11_massdep_delta(args.model, 1, ords + TC_MAX_KNOTS, logm[u], f[u],
                                              CUDART_INF, bin.row[2], bin.row[3], bin.pct[2],
                                              bin.pct[3], ka, kb);
          acc2 = fma(wc, fma(ka, dl, f[u]), acc2);
          acc3 = fma(wd, fma(kb, dl, f[u]), acc3);
        } else {   // (DEC == 1: the satellites' perturbation is applied to the sums below)
          acc2 = fma(wc, f[u], acc2);
          acc3 = fma(wd, f[u], acc3);
        }
      }
    }
  }
  if (DECORATED) {
    // Unbounded above, the satellites' Heaviside perturbation is linear in the baseline
    // (assembias_delta with hi = inf: A > 0: delta = A s / (1 - s) f, A <= 0: delta = A f), so it
    // scales the Gauss-Legendre sums instead of every node
    const double c = split_ok ? a_sat * (a_sat > 0.0 ? ratio : 1.0) : 0.0;
    acc2 = fma(k2 * c, acc2, acc2);
    acc3 = fma(k3 * c, acc3, acc3);
  }
  const double bad = d->bad;   // NaN: table not increasing, or a parameter that is not finite
  const double acc[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
  for (int r = 0; r < 4; r++)
    if (bin.dst[r] >= 0) out[bin.dst[r]] = acc[r] + bad;
}

// MASSDEP: the kernel of models whose assembly-bias strength / split depend on mass (a kernel of
// its own, so that the plain models' kernel does not carry its code and shared memory).
template <bool MASSDEP>
__global__ void __launch_bounds__(kL11Threads, kL11MinBlocks)
occupation_l11_kernel(const OccArgs args) {
  extern __shared__ __align__(16) double l11_smem[];
  double* tab = l11_smem;
  double* sk = l11_smem + kL11TabDoubles;
  double* ords = sk + kL11Knots;   // [draw][centrals ordinates | satellites ordinates]
  L11Draw* draws = reinterpret_cast<L11Draw*>(ords + (MASSDEP ? kL11OrdDoubles : 0));
  for (int i = threadIdx.x; i < kL11TabDoubles; i += blockDim.x) tab[i] = g_l11_tables[i];
  for (int k = threadIdx.x; k < kL11Knots; k += blockDim.x) sk[k] = l11_knot_logms(k);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long n_blocks = (args.n_draws + kL11DrawsPerBlock - 1) / kL11DrawsPerBlock;
  const int n_bins = args.plan.n_l11_bins, n_quads = (n_bins + kL11LaneBins - 1) / kL11LaneBins;
  for (long long block = blockIdx.x; block < n_blocks; block += gridDim.x) {
    const long long draw0 = block * kL11DrawsPerBlock;
    const int n_block = (int)min((long long)kL11DrawsPerBlock, args.n_draws - draw0);
    __syncthreads();   // the previous block's tables are no longer read; tab and sk are filled
    l11_prepare_block<MASSDEP>(draws, ords, sk, tab, n_block, draw0, args.n_draws, args.theta, args.theta_ds,
                      args.theta_ps, args.model);
    // one warp per unit = kL11LaneDraws draws x kL11LaneBins consecutive mass bins; the lanes of a
    // draw are adjacent (neighbouring bins of one draw read the same or neighbouring knots and
    // erf columns: fewer shared-memory wavefronts than with the draws adjacent)
    const int n_units = ((n_block + kL11LaneDraws - 1) / kL11LaneDraws) * n_quads;
    for (int unit = warp; unit < n_units; unit += n_warps) {
      const int sub = unit / n_quads, quad = unit - sub * n_quads;
      const int b = sub * kL11LaneDraws + lane / kL11LaneBins;
      const int bin = quad * kL11LaneBins + (lane & (kL11LaneBins - 1));
      if (b >= n_block || bin >= n_bins) continue;
      const L11Bin* bin_ptr = args.plan.l11_bins + bin;
      double* out = args.occ_out + (draw0 + b) * args.n_rows;
      const double* ord = ords + b * kL11OrdPerDraw;
      if (MASSDEP) {   // rare: one instantiation, a node at a time
        l11_bin<1, 2>(args, draws + b, ord, tab, bin_ptr, out);
      } else if (args.model.decorated) {
        if (args.plan.unroll == kOccUnroll)
          l11_bin<kOccUnroll, 1>(args, draws + b, ord, tab, bin_ptr, out);
        else
          l11_bin<2, 1>(args, draws + b, ord, tab, bin_ptr, out);
      } else {
        if (args.plan.unroll == kOccUnroll)
          l11_bin<kOccUnroll, 0>(args, draws + b, ord, tab, bin_ptr, out);
        else
          l11_bin<2, 0>(args, draws + b, ord, tab, bin_ptr, out);
      }
    }
  }
}

}  // namespace
