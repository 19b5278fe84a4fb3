/* tabcorr_b200 -- C ABI of the B200 (sm_100a) implementation of TabCorr's prediction hot path.
 *
 * The reference (johannesulf/TabCorr v1.2.0) is pure Python and has no FFI layer; its boundary for
 * this path is the Python API (tabcorr/tabcorr.py:374-416,465-683, tabcorr/interpolator.py:14-216).
 * The entry points below are what a binding for that path needs: tabcorr_b200/_lib.py binds them
 * with ctypes, and INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * TC_E* code (tc_last_error() then holds a message for the calling thread); nothing throws.
 * Pointers named *_host are host memory, *_dev device memory of the table's device.  Launches are
 * asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).  Device memory
 * is allocated only inside tc_*_create / tc_table_plan; per-call scratch is supplied by the caller
 * (tc_predict_workspace_bytes).
 */
#ifndef TABCORR_B200_H
#define TABCORR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TC_VERSION 109

#define TC_OK 0
#define TC_EINVAL (-1)      /* bad argument */
#define TC_ECUDA (-2)       /* CUDA runtime error, see tc_last_error() */
#define TC_EUNSUPPORTED (-3) /* shape outside what the kernels support */
#define TC_ENOMEM (-4)

#define TC_MODE_AUTO 0  /* xi_r = w^T M_r w / (sum w)^2, tabcorr/tabcorr.py:641-647 */
#define TC_MODE_CROSS 1 /* xi_r = M_r . w / sum w,        tabcorr/tabcorr.py:648-649 */

/* Arithmetic of the auto-mode contraction.  FP64: DMMA tensor cores, parity with the reference
 * to rtol 1e-10.  3XTF32 (optional, auto tables only): table entries and tracer weights split into
 * TF32 high + low parts, the products that matter to 22 bits accumulated in FP32 on the TF32 tensor
 * cores -- tcgen05.mma with TMEM accumulators and TMA-staged operands for batches on tables of at
 * most 256 padded rows (csrc/tcgen05_contract.cuh), the warp-level m16n8k8 MMA otherwise -- row-dot
 * and normalisation in FP64; relative error ~1e-7 of the term magnitudes (tested 1e-6).
 * The occupation arithmetic is FP64 in both. */
#define TC_PRECISION_FP64 0
#define TC_PRECISION_3XTF32 1

/* Occupation families of the occupation kernel (tc_model.family) and the number of doubles per
 * parameter draw each consumes (tc_model_n_theta):
 *   TC_FAMILY_ZHENG07 (7):  logMmin, sigma_logM, logM0, logM1, alpha, A_cen, A_sat
 *   TC_FAMILY_LEAUTHAUD11 (18): smhm_m0_0, smhm_m0_a, smhm_m1_0, smhm_m1_a, smhm_beta_0,
 *       smhm_beta_a, smhm_delta_0, smhm_delta_a, smhm_gamma_0, smhm_gamma_a, scatter_model_param1,
 *       alphasat, bsat, bcut, betacut, betasat, A_cen, A_sat
 * (halotools param_dict names; A_cen / A_sat are the mean_occupation_{centrals,satellites}_
 * assembias_param1 strengths, ignored unless decorated).  With a mass-dependent strength
 * (tc_model.n_strength > 1) a family's base parameters (5 / 16) are followed by the n_strength[0]
 * ordinates of the centrals and the n_strength[1] ordinates of the satellites; leauthaud11 with a
 * mass-dependent scatter (tc_model.n_scatter = n > 1) appends scatter_model_param2..n. */
#define TC_FAMILY_ZHENG07 0
#define TC_FAMILY_LEAUTHAUD11 1
#define TC_N_THETA 7              /* zheng07 */
#define TC_N_THETA_LEAUTHAUD11 18
#define TC_N_THETA_MAX 27             /* leauthaud11 with 4 + 4 strength and 4 scatter ordinates */
#define TC_N_THETA_ZHENG07_BASE 5   /* logMmin .. alpha; the strengths follow */

typedef struct tc_table tc_table;   /* device-resident table group (one gal_type, >=1 matrices) */
typedef struct tc_interp tc_interp; /* tensor-product cubic spline over a table grid */

/* Occupation model evaluated by the occupation kernel; replaces the halotools calls at
 * tabcorr/tabcorr.py:556-563 (Zheng07Cens/Zheng07Sats or Leauthaud11Cens/Leauthaud11Sats,
 * optionally decorated with HeavisideAssembias). */
#define TC_MAX_KNOTS 4   /* control points of a mass-dependent assembly-bias strength / split */

typedef struct tc_model {
  int32_t family;               /* TC_FAMILY_* */
  int32_t decorated;            /* 1 = Heaviside assembly bias on centrals and satellites */
  int32_t modulate_with_cenocc; /* 1 = <N_sat> is multiplied by the baseline <N_cen> */
  int32_t n_scatter;            /* leauthaud11: control points of a mass-dependent stellar-mass
                                 * scatter (0 / 1: the constant scatter_model_param1), see below */
  double split;                 /* percentile split of the decoration (halotools default 0.5) */
  double threshold;             /* leauthaud11: log10 of the stellar-mass threshold */
  double redshift;              /* leauthaud11: redshift of the stellar-to-halo-mass relation */
  /* Mass-dependent decoration (halotools HeavisideAssembias with assembias_strength_abscissa /
   * split_abscissa; zheng07 family only).  Index 0: centrals, 1: satellites.
   * n_strength[t] <= 1: one strength per draw (the layouts above).  n_strength[t] = n in
   * 2..TC_MAX_KNOTS: the draw carries n strength ordinates for type t (theta: logMmin,
   * sigma_logM, logM0, logM1, alpha, then the centrals' ordinates, then the satellites'), and
   * the strength of a halo is the interpolating polynomial of degree n - 1 through
   * (strength_abscissa[t][k], ordinate k) evaluated at log10(prim_haloprop), clipped to [-1, 1]
   * (halotools: custom_spline(..., k=3), a spline of degree min(3, n - 1) through n points).
   * n_split[t] = 0: `split` above; n in 1..TC_MAX_KNOTS: the splitting percentile of type t is
   * the same kind of polynomial through (split_abscissa[t][k], split_ordinates[t][k]), clipped to
   * [0, 1] (fixed at model construction, not part of the draw). */
  int32_t n_strength[2];
  int32_t n_split[2];
  double strength_abscissa[2][TC_MAX_KNOTS];
  double split_abscissa[2][TC_MAX_KNOTS];
  double split_ordinates[2][TC_MAX_KNOTS];
  /* leauthaud11, n_scatter = n in 2..TC_MAX_KNOTS (halotools LogNormalScatterModel with
   * scatter_abscissa / scatter_ordinates): the log-normal scatter in stellar mass of a halo is the
   * interpolating polynomial through (scatter_abscissa[k], scatter_model_param<k+1>) at
   * log10(prim_haloprop); the draw carries scatter_model_param1 in its usual place and
   * scatter_model_param2..n behind the strength ordinates. */
  double scatter_abscissa[TC_MAX_KNOTS];
} tc_model;

const char* tc_last_error(void);
int tc_version(void);

/* Doubles per parameter draw of the model's family, or TC_EUNSUPPORTED. */
int tc_model_n_theta(const tc_model* model);

/* TabCorr.read (tabcorr/tabcorr.py:374-416) for `n_tables` tables that share one gal_type table
 * (n_tables > 1 is an Interpolator group, tabcorr/interpolator.py:63-70).  All inputs are host
 * arrays in the reference's row order: the gal_type columns n_h, log_prim_haloprop_min/max,
 * sec_haloprop_percentile, prim_haloprop_dist_index (NULL for legacy tables without that column,
 * tabcorr.py:568-574), is_sat[i] = (gal_type[i] != 'centrals') (tabcorr.py:555), and per table
 * the matrix exactly as stored: [n_r, n_rows (n_rows + 1) / 2] packed lower triangle
 * (tabcorr.py:770-806) for TC_MODE_AUTO, [n_r, n_rows] for TC_MODE_CROSS. */
int tc_table_create(tc_table** out, int mode, int n_rows, int n_r, int n_tables,
                    const double* n_h_host, const double* log_min_host,
                    const double* log_max_host, const double* sec_pct_host,
                    const double* dist_index_host, const int32_t* is_sat_host,
                    const double* const* tpcf_matrix_host, int device);
int tc_table_destroy(tc_table* table);

/* Shape queries. */
int tc_table_n_rows(const tc_table* table);
int tc_table_n_r(const tc_table* table);
int tc_table_n_tables(const tc_table* table);

/* Register the Gauss-Legendre rule used by mean_occupation (tabcorr.py:543-552,568-578):
 * nodes x in [0, 1] and weights of numpy.polynomial.legendre.leggauss(n_gauss).  Builds (once per
 * n_gauss, cached in the table) the node masses and normalised quadrature weights on the device. */
int tc_table_plan(tc_table* table, int n_gauss, const double* x01_host, const double* w_host);

/* Layout of the parameter draws theta_dev (n_theta = tc_model_n_theta), selected by theta_ld:
 *   theta_ld == 0:        [B, n_theta], one row per draw;
 *   theta_ld >= n_draws:  [n_theta, theta_ld], one contiguous column per parameter (what a
 *                         sampler that keeps one array per parameter -- model.param_dict keys --
 *                         hands over without a transpose). */

/* TabCorr.mean_occupation (tabcorr.py:465-578) for B draws: occ_dev[B, n_rows] in reference row
 * order.  All families. */
int tc_occupation_batch(tc_table* table, const tc_model* model, int n_gauss,
                        const double* theta_dev, int64_t theta_ld, int64_t n_draws,
                        double* occ_dev, void* stream);

/* Scratch bytes tc_predict_batch needs for n_draws (separate = separate_gal_type). */
size_t tc_predict_workspace_bytes(const tc_table* table, int64_t n_draws, int separate);

/* The same for a given precision: TC_PRECISION_3XTF32 on eligible tables (auto mode, total
 * predictions, at most 256 padded rows) adds the operand images of the tcgen05 contraction; with
 * the smaller FP64 workspace that mode runs on the warp-level TF32 MMA instead. */
size_t tc_predict_workspace_bytes_for(const tc_table* table, int64_t n_draws, int separate,
                                      int precision);

/* TabCorr.predict (tabcorr.py:580-683) for B draws, fused occupation + contraction.
 * Exactly one of theta_dev (layout above; evaluated with `model` and the n_gauss plan) and
 * occ_dev ([B, n_rows] precomputed occupations, the ndarray branch tabcorr.py:616-621) is non-NULL.
 * The fused occupation phase implements TC_FAMILY_ZHENG07; for the other families the call
 * returns TC_EUNSUPPORTED and the caller chains tc_occupation_batch -> occ_dev (what
 * tabcorr_b200.DeviceTableGroup.predict_into does).
 * Outputs, with T = n_tables, R = n_r:
 *   separate == 0: ngal_dev[b * ngal_stride + t], xi_dev[b * xi_stride + t * R + r]
 *   separate == 1: ngal_dev[b * ngal_stride + t * 2 + q], q = centrals, satellites;
 *                  xi_dev[b * xi_stride + (t * R + r) * C + p] with C = 3, p = cen-cen, cen-sat,
 *                  sat-sat (auto) or C = 2, p = centrals, satellites (cross); tabcorr.py:652-683.
 * Strides are in doubles, so that several table groups can fill one [B, T_total, ...] buffer. */
int tc_predict_batch(tc_table* table, const tc_model* model, int n_gauss, const double* theta_dev,
                     int64_t theta_ld, const double* occ_dev, int64_t n_draws, int separate,
                     int precision, double* ngal_dev,
                     int64_t ngal_stride, double* xi_dev, int64_t xi_stride, void* workspace_dev,
                     size_t workspace_bytes, void* stream);

/* TabCorr.predict(model) (tabcorr.py:580-650) for ONE parameter set whose TC_N_THETA parameters
 * are read from HOST memory at call time and travel to the kernel in its launch arguments -- the
 * latency path of the reference's MCMC idiom (README.md:72-74): no parameter upload, and no
 * mapped-host reads by every thread block.  Outputs and workspace as in tc_predict_batch with
 * n_draws = 1 (they may live in mapped pinned host memory).  TC_FAMILY_ZHENG07 only. */
int tc_predict_one(tc_table* table, const tc_model* model, int n_gauss, const double* theta_host,
                   int separate, int precision, double* ngal_dev, int64_t ngal_stride,
                   double* xi_dev, int64_t xi_stride, void* workspace_dev, size_t workspace_bytes,
                   void* stream);

/* Interpolator.__init__ (tabcorr/interpolator.py:39-61): n_dims axes with n_knots[d] sorted knots
 * (concatenated in knots_host) and the spline matrices a[d] of shape [n_knots[d]-1, 4, n_knots[d]]
 * (spline_interpolation_matrix, interpolator.py:219-272; concatenated in a_host).
 * grid_to_table_host[g] is the index, in the caller's table order, of the table at row-major grid
 * position g (first axis slowest), i.e. the lexicographic sort of interpolator.py:59-61. */
int tc_interp_create(tc_interp** out, int n_dims, const int32_t* n_knots_host,
                     const double* knots_host, const double* a_host,
                     const int32_t* grid_to_table_host, int device);
int tc_interp_destroy(tc_interp* interp);

/* spline_interpolate (interpolator.py:275-331) for B draws: out_dev[b, c] = sum over grid tables
 * of w_t(x_b) data_dev[b, t, c], t in the caller's table order, n_cols values per table.
 * x_dev is [B, n_dims].  Without extrapolate, draws outside the knot hull get NaN outputs and
 * *flag_dev (int32, device) is set to 1 so that the host can raise the reference's ValueError
 * (interpolator.py:322-326); with extrapolate the end segments are used (:327-328). */
int tc_interp_apply_batch(tc_interp* interp, const double* x_dev, int64_t n_draws,
                          const double* data_dev, int n_cols, double* out_dev, int extrapolate,
                          int32_t* flag_dev, void* stream);

/* Peer-visible result slabs for the multi-GPU batch (the one collective of the path is
 * the collection of the [B_r, 1 + R] result rows on one rank).  tc_peer_alloc allocates `bytes` of
 * device memory on `device` and returns its 64-byte CUDA IPC handle; another process of the node
 * maps it with tc_peer_open (peer access over NVLink is enabled lazily) and passes row ranges of
 * the mapping as ngal_dev / xi_dev of tc_predict_batch: finalize_kernel then stores every rank's
 * results straight into the destination GPU's memory -- the gather is fused into the kernel's
 * epilogue and only a barrier remains.  tc_peer_close unmaps, tc_peer_free releases. */
int tc_peer_alloc(int device, size_t bytes, void** ptr_out, unsigned char* handle_out);
int tc_peer_open(int device, const unsigned char* handle, void** ptr_out);
int tc_peer_close(int device, void* ptr);
int tc_peer_free(int device, void* ptr);

/* Live FP64 tensor (DMMA) peak of the device in TFLOP/s, the roofline denominator bench.py
 * reports (MEASURED_PEAKS.json carries no FP64 figure). */
int tc_measure_dmma_peak(int device, double* tflops_out);

/* Live FP64 ALU (DFMA) peak of the device in TFLOP/s: the roofline denominator of the paths that
 * are bound by the occupation arithmetic (cross tables tabcorr.py:648-649, mean_occupation
 * tabcorr.py:465-578), where bench.py reports node evaluations/s. */
int tc_measure_dfma_peak(int device, double* tflops_out);

/* Halo-bin reductions of the tabulation side (tabcorr/tabcorr.py:194-227, the n_h histogram and
 * the per-cell mean primary property behind prim_haloprop_dist_index; sort_into_bins :676-737).
 * Inputs are device arrays of n_halos doubles: log_prim (what np.histogram2d / np.digitize bin,
 * computed by the caller exactly as the reference does, np.log10), sec_pct, prim (what is averaged);
 * prim_edges [n_prim + 1] and sec_edges [n_sec + 1] are HOST arrays of ascending bin edges.
 * Outputs (HOST arrays of n_sec * n_prim doubles, cell index = sec * n_prim + prim, i.e. the
 * reference's ravel(order='F')): n_h_out = counts with np.histogram2d semantics (values on the
 * last edge fall into the last bin), n_members_out and mean_out = number and mean of prim over the
 * members of np.digitize(right=False) (last edge excluded), mean_out = NaN for empty cells.
 * Counts are exact; the mean is accumulated in fixed point (52 fractional bits of the position
 * inside the cell) and is bit-reproducible.  Synchronises the stream. */
int tc_halo_bins(int device, const double* log_prim_dev, const double* sec_pct_dev,
                 const double* prim_dev, int64_t n_halos, const double* prim_edges, int n_prim,
                 const double* sec_edges, int n_sec, double* n_h_out, double* n_members_out,
                 double* mean_out, void* stream);

/* Element-wise evaluation of the occupation kernel's table-driven math on the current device, for
 * accuracy tests: kind 0: out = 0.5 (1 + erf(x)); kind 1: out = x^y for x > 0; kinds 2 / 3: the
 * leauthaud11 kernel's grouped erf (one coefficient column for nodes that lie close together)
 * of the pair (x, y), out = 0.5 (1 + erf(x)) / 0.5 (1 + erf(y)). */
int tc_debug_math(int kind, const double* x_dev, const double* y_dev, double* out_dev, int64_t n,
                  void* stream);

/* Per-kernel device timing of the most recent tc_predict_batch (CUDA events recorded on its
 * stream around the fused kernel and around the finalize kernel); used by bench.py for the
 * roofline of the dominant kernel.  Not thread safe; off by default. */
int tc_profile_enable(int on);
int tc_profile_read(float* predict_ms_out, float* finalize_ms_out);

#ifdef __cplusplus
}
#endif
#endif /* TABCORR_B200_H */
