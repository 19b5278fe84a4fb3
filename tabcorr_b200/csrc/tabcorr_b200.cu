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
//
// Layout of the translation unit: common.cuh (structs, MMA helpers) -> device_math.cuh ->
// occupation.cuh -> predict_kernel.cuh / leauthaud11.cuh / aux_kernels.cuh -> host_tables.cuh
// (layouts, plans, heuristics) -> the C ABI below.

#include "host_tables.cuh"

namespace {

// the mass-dependent decoration fields of tc_model (include/tabcorr_b200.h)
int check_model(const tc_model* model) {
  for (int type = 0; type < 2; type++) {
    const int ns = model->n_strength[type], np = model->n_split[type];
    if (ns < 0 || ns > TC_MAX_KNOTS || np < 0 || np > TC_MAX_KNOTS)
      return fail(TC_EUNSUPPORTED, "tc_model: at most " + std::to_string(TC_MAX_KNOTS) +
                                       " control points per assembly-bias strength / split");
    for (int k = 1; k < ns; k++)
      if (!(model->strength_abscissa[type][k] > model->strength_abscissa[type][k - 1]))
        return fail(TC_EINVAL, "tc_model: strength_abscissa must increase strictly");
    for (int k = 1; k < np; k++)
      if (!(model->split_abscissa[type][k] > model->split_abscissa[type][k - 1]))
        return fail(TC_EINVAL, "tc_model: split_abscissa must increase strictly");
  }
  if (model->n_scatter < 0 || model->n_scatter > TC_MAX_KNOTS)
    return fail(TC_EUNSUPPORTED, "tc_model: at most " + std::to_string(TC_MAX_KNOTS) +
                                     " control points of the stellar-mass scatter");
  if (model->n_scatter > 1 && model->family != TC_FAMILY_LEAUTHAUD11)
    return fail(TC_EINVAL, "tc_model: n_scatter belongs to TC_FAMILY_LEAUTHAUD11");
  for (int k = 1; k < model->n_scatter; k++)
    if (!(model->scatter_abscissa[k] > model->scatter_abscissa[k - 1]))
      return fail(TC_EINVAL, "tc_model: scatter_abscissa must increase strictly");
  return TC_OK;
}

bool host_mass_dependent(const tc_model* m) {
  return m && m->decorated && (m->n_strength[0] > 1 || m->n_strength[1] > 1 || m->n_split[0] > 0 ||
                               m->n_split[1] > 0);
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* tc_last_error(void) { return g_last_error.c_str(); }

int tc_version(void) { return TC_VERSION; }

int tc_model_n_theta(const tc_model* model) {
  if (!model) return fail(TC_EINVAL, "tc_model_n_theta: model is NULL");
  if (model->family == TC_FAMILY_ZHENG07) {
    int rc = check_model(model);
    return rc != TC_OK ? rc : zheng07_n_theta(*model);
  }
  if (model->family == TC_FAMILY_LEAUTHAUD11) {
    int rc = check_model(model);
    return rc != TC_OK ? rc : l11_n_theta(*model);
  }
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
  if ((rc = check_model(model))) return rc;
  const int n_theta = model->family == TC_FAMILY_LEAUTHAUD11 ? l11_n_theta(*model)
                                                              : zheng07_n_theta(*model);
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
  if (model->family == TC_FAMILY_ZHENG07)
    // enough items for every warp of the device, at most 8 draw pieces per type
    pick_series_ranges(args.plan, 1, n_draws,
                       (int)std::min<long long>(16, std::max<long long>(
                           2, (long long)n_sm * kWarps / ((n_draws + 7) / 8))),
                       &args.n_ranges_cen, &args.n_ranges_sat, &args.pieces_cen, &args.pieces_sat);
  if (model->family == TC_FAMILY_LEAUTHAUD11) {
    const bool massdep = l11_mass_dependent(*model);
    const size_t smem = l11_smem_bytes(massdep);
    static std::mutex m;
    static std::map<std::pair<int, bool>, bool> configured;
    {
      std::lock_guard<std::mutex> lock_attr(m);
      if (!configured[{t->device, massdep}]) {
        // just enough shared memory for the resident CTAs: what is left is L1 for the plan arrays
        const int carveout = (int)std::min<size_t>(
            100, (kL11MinBlocks * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
        const void* kernel = massdep ? (const void*)occupation_l11_kernel<true>
                                     : (const void*)occupation_l11_kernel<false>;
        TC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
        TC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     carveout));
        configured[{t->device, massdep}] = true;
      }
    }
    const long long n_blocks = (n_draws + kL11DrawsPerBlock - 1) / kL11DrawsPerBlock;
    const int grid = (int)std::max<long long>(
        1, std::min<long long>(n_blocks, (long long)n_sm * kL11MinBlocks));
    if (massdep)
      occupation_l11_kernel<true><<<grid, kL11Threads, smem, static_cast<cudaStream_t>(stream)>>>(args);
    else
      occupation_l11_kernel<false><<<grid, kL11Threads, smem, static_cast<cudaStream_t>(stream)>>>(args);
    TC_CUDA(cudaGetLastError());
    return TC_OK;
  }
  if (host_mass_dependent(model)) {
    const long long n_pairs = n_draws * args.plan.n_groups;
    const int grid_md = (int)std::max<long long>(
        1, std::min<long long>((n_pairs + 255) / 256, (long long)n_sm * 8));
    occupation_massdep_kernel<<<grid_md, 256, 0, static_cast<cudaStream_t>(stream)>>>(args);
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

size_t tc_predict_workspace_bytes_for(const tc_table* t, int64_t n_draws, int separate,
                                      int precision) {
  size_t need = tc_predict_workspace_bytes(t, n_draws, separate);
  if (need == 0 || precision != TC_PRECISION_3XTF32 || !tcgen_eligible(t, separate ? 1 : 0))
    return need;
  return std::max(need, plan_tcgen_workspace(t, n_draws).total);
}

}  // extern "C"

namespace {

// The 3xTF32 contraction on tcgen05 (csrc/tcgen05_contract.cuh): weights_image_kernel (FP64
// occupations -> operand images) -> tcgen_contract_kernel -> finalize_kernel.  The caller holds
// the table mutex and has validated the arguments.
int predict_tcgen(tc_table* t, const tc_model* model, int n_gauss, const double* theta,
                  int64_t theta_ld, int64_t n_draws, double* ngal, int64_t ngal_stride, double* xi,
                  int64_t xi_stride, void* workspace, cudaStream_t stream, int n_sm) {
  int rc = build_tcgen(t);
  if (rc != TC_OK) return rc;
  if ((rc = build_plan(t, 0, n_gauss))) return rc;
  Layout& L = t->layouts[0];
  const TcgenWorkspace ws = plan_tcgen_workspace(t, n_draws);
  char* base = static_cast<char*>(workspace);
  base += (256 - reinterpret_cast<uintptr_t>(base) % 256) % 256;
  float* h_img = reinterpret_cast<float*>(base);
  float* c_img = reinterpret_cast<float*>(base + ws.a_bytes);
  double* parts = reinterpret_cast<double*>(base + ws.a_bytes + ws.c_bytes);
  double* ngal_tile = reinterpret_cast<double*>(base + ws.a_bytes + ws.c_bytes + ws.parts_bytes);
  double* ngal_parts = reinterpret_cast<double*>(base + ws.a_bytes + ws.c_bytes + ws.parts_bytes +
                                                 ws.ngal_bytes);
  int* error_flag = reinterpret_cast<int*>(base + ws.a_bytes + ws.c_bytes + ws.parts_bytes +
                                           ws.ngal_bytes + ws.ngal_parts_bytes);
  TC_CUDA(cudaMemsetAsync(error_flag, 0, sizeof(int), stream));
  const bool profile = g_profile.enabled;
  if (profile) TC_CUDA(cudaEventRecord(g_profile.ev[0], stream));

  WeightsImageArgs wa{};
  wa.plan = L.plans[n_gauss].dev;
  wa.model = *model;
  wa.theta = theta;
  wa.theta_ds = theta_ld ? 1 : TC_N_THETA;
  wa.theta_ps = theta_ld ? theta_ld : 1;
  wa.n_draws = n_draws;
  // (never the one-draw item shape: the order in which a draw's weights are summed into its
  // number density follows the items, and results must not depend on the batch size)
  pick_series_ranges(wa.plan, 1, std::max<int64_t>(n_draws, 2),
                     (int)std::min<long long>(16, std::max<long long>(
                         2, (long long)n_sm * kWarps / ((n_draws + 7) / 8))),
                     &wa.n_ranges_cen, &wa.n_ranges_sat, &wa.pieces_cen, &wa.pieces_sat);
  wa.max_groups = std::max(wa.plan.n_cen_groups, wa.plan.n_groups - wa.plan.n_cen_groups);
  if (wa.n_ranges_cen + wa.n_ranges_sat > ws.n_ranges_max)
    return fail(TC_EUNSUPPORTED, "tcgen05 path: too many occupation ranges");
  wa.kp = L.tcgen.kp;
  wa.n_pad = L.dev.n_pad;
  wa.n_rows = t->n_rows;
  wa.h_img = h_img;
  wa.c_img = c_img;
  wa.ngal_parts = ngal_parts;
  wa.ngal_ld = ws.n_tiles * kTcM;
  {
    const long long n_items = (n_draws + 7) / 8 * (wa.n_ranges_cen + wa.n_ranges_sat);
    const int grid = (int)std::max<long long>(
        1, std::min<long long>((n_items + kWarps - 1) / kWarps, (long long)n_sm));
    const size_t wsmem = (size_t)kWarps * 8 * wa.max_groups * sizeof(double);
    if (wsmem + 16 * 1024 > (size_t)kSmemLimit)
      return fail(TC_EUNSUPPORTED, "tcgen05 path: too many mass-bin groups per galaxy type");
    TC_CUDA(cudaFuncSetAttribute(weights_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)wsmem));
    weights_image_kernel<<<grid, kThreads, wsmem, stream>>>(wa);
    TC_CUDA(cudaGetLastError());
  }

  TcgenArgs ta{};
  ta.tc = L.tcgen;
  ta.h_img = h_img;
  ta.c_img = c_img;
  ta.ngal_parts = ngal_parts;
  ta.ngal_ld = wa.ngal_ld;
  ta.n_ranges_cen = wa.n_ranges_cen;
  ta.n_ranges_sat = wa.n_ranges_sat;
  ta.n_pad = L.dev.n_pad;
  ta.n_rows = t->n_rows;
  ta.n_tiles = ws.n_tiles;
  ta.parts = parts;
  ta.ngal_tile = ngal_tile;
  ta.error_flag = error_flag;
  static long long* debug_buf = nullptr;
  if (tune("TCGEN_DEBUG", 0)) {
    if (!debug_buf) TC_CUDA(cudaMalloc(reinterpret_cast<void**>(&debug_buf), 1024 * 8 * sizeof(long long)));
    TC_CUDA(cudaMemsetAsync(debug_buf, 0, 1024 * 8 * sizeof(long long), stream));
    ta.debug = debug_buf;
  }
  const size_t smem = (size_t)kTcStages * kTcCopyBytes + kTcBarriers * 8 + 16;
  {
    static std::mutex m;
    static std::map<int, bool> configured;
    std::lock_guard<std::mutex> lock_attr(m);
    if (!configured[t->device]) {
      TC_CUDA(cudaFuncSetAttribute(tcgen_contract_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
      configured[t->device] = true;
    }
  }
  const int grid = (int)std::min<long long>(ws.n_tiles * L.tcgen.n_ib, n_sm);
  tcgen_contract_kernel<<<grid, kTcThreads, smem, stream>>>(ta);
  TC_CUDA(cudaGetLastError());
  if (ta.debug) {   // developer aid: cycle breakdown of CTA 0, 1 and the last one on stderr
    TC_CUDA(cudaStreamSynchronize(stream));
    std::vector<long long> h((size_t)grid * 8);
    TC_CUDA(cudaMemcpy(h.data(), ta.debug, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int c : {0, 1, grid - 1})
      std::fprintf(stderr, "tcgen cta %d: mma total %lld wait_a %lld wait_acc %lld wait_full %lld | "
                   "epi total %lld a_store %lld c_load %lld wait_acc_full %lld\n", c, h[c * 8], h[c * 8 + 1],
                   h[c * 8 + 2], h[c * 8 + 3], h[c * 8 + 4], h[c * 8 + 5], h[c * 8 + 6], h[c * 8 + 7]);
  }
  if (profile) TC_CUDA(cudaEventRecord(g_profile.ev[1], stream));

  FinalizeArgs fa{};
  fa.lay = L.dev;
  fa.lay.n_parts = L.tcgen.n_parts;
  fa.lay.out_ptr = L.tcgen.out_ptr;
  fa.lay.out_parts = L.tcgen.out_parts;
  fa.parts = parts;
  fa.ngal_tile = ngal_tile;
  fa.n_draws = n_draws;
  fa.bm = kTcM;
  fa.mode = TC_MODE_AUTO;
  fa.separate = 0;
  fa.n_tables = t->n_tables;
  fa.ngal_out = ngal;
  fa.ngal_stride = ngal_stride;
  fa.xi_out = xi;
  fa.xi_stride = xi_stride;
  const int n_out = L.dev.n_out;
  int fy, fsmem;
  finalize_grid(n_out, kTcM, ws.n_tiles, n_sm, &fy, &fsmem);
  finalize_kernel<<<dim3((unsigned)ws.n_tiles, fy), 256, fsmem, stream>>>(fa);
  TC_CUDA(cudaGetLastError());
  tcgen_poison_kernel<<<64, 256, 0, stream>>>(error_flag, xi, n_draws, xi_stride, n_out);
  TC_CUDA(cudaGetLastError());
  if (profile) {
    TC_CUDA(cudaEventRecord(g_profile.ev[2], stream));
    g_profile.recorded = true;
  }
  return TC_OK;
}

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
  if (theta) {
    int rc_model = check_model(model);
    if (rc_model != TC_OK) return rc_model;
  }
  if (theta && host_mass_dependent(model))
    return fail(TC_EUNSUPPORTED, "tc_predict_batch: models with mass-dependent assembly bias are "
                                 "evaluated by tc_occupation_batch (strength and split vary from "
                                 "quadrature node to node); pass its result as occ_dev");
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
  if (precision == TC_PRECISION_3XTF32 && theta && !theta_inline &&
      tcgen_eligible(t, separate) &&
      n_draws >= tune("TCGEN_MIN_DRAWS", 1) && workspace &&
      workspace_bytes >= plan_tcgen_workspace(t, n_draws).total) {
    // Blackwell-native contraction (tcgen05 + TMEM + TMA) for every batch size, so that results do
    // not depend on how a batch is cut; split predictions, precomputed occupations, the one-draw
    // latency call and larger tables take the warp-level TF32 MMA below
    int n_sm_tc = 0;
    if ((rc = device_sms(t->device, &n_sm_tc))) return rc;
    if (ngal_stride < (int64_t)t->n_tables || xi_stride < (int64_t)t->n_tables * t->n_r)
      return fail(TC_EINVAL, "tc_predict_batch: output stride smaller than one draw's outputs");
    return predict_tcgen(t, model, n_gauss, theta, theta_ld, n_draws, ngal, ngal_stride, xi,
                         xi_stride, workspace, stream, n_sm_tc);
  }
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
  // Occupation items: series items (occupation_item_series) where they pay -- cross tables
  // (bound by the occupation arithmetic), many quadrature nodes, tables of 200+ rows -- else the
  // node-by-node items whose code is smaller (instruction-cache footprint of the fused kernel)
  const int series_auto = t->mode == TC_MODE_CROSS || plan_g > 16 || L.dev.n_pad >= 200;
  if (theta != nullptr || theta_inline != nullptr ? tune("SERIES_FUSED", series_auto) != 0 : false) {
    pick_series_ranges(args.plan, ws.nt, n_draws,
                       t->mode == TC_MODE_CROSS ? kOccSeriesItemsPerTileCross : kOccItemsPerTile,
                       &args.n_ranges_cen, &args.n_ranges_sat, &args.pieces_cen, &args.pieces_sat,
                       n_draws <= 64);   // measured: 8 / 32 draws 88 / 84 -> 76 / 78 us, 128+ slower
  } else {
    pick_ranges(args.plan, ws.nt, &args.n_ranges_cen, &args.n_ranges_sat,
                t->mode == TC_MODE_CROSS ? kOccItemsPerTileCross : kOccItemsPerTile);
    args.pieces_cen = args.pieces_sat = 0;
  }
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
  args.stress_ns = std::max(0, tune("STRESS", 0));
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
  int fy, fsmem;
  finalize_grid(n_out, bm, ws.n_tiles, n_sm, &fy, &fsmem);
  dim3 fgrid((unsigned)ws.n_tiles, fy);
  finalize_kernel<<<fgrid, 256, fsmem, stream>>>(fa);
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

int tc_halo_bins(int device, const double* log_prim, const double* sec_pct, const double* prim,
                 int64_t n_halos, const double* prim_edges, int n_prim, const double* sec_edges,
                 int n_sec, double* n_h_out, double* n_members_out, double* mean_out,
                 void* stream_) {
  if (!prim_edges || !sec_edges || !n_h_out || !n_members_out || !mean_out || n_prim <= 0 ||
      n_sec <= 0 || n_halos < 0 || (n_halos > 0 && (!log_prim || !sec_pct || !prim)))
    return fail(TC_EINVAL, "tc_halo_bins: bad argument");
  for (int i = 0; i < n_prim; i++)
    if (!(prim_edges[i] < prim_edges[i + 1]))
      return fail(TC_EINVAL, "tc_halo_bins: prim_edges must increase strictly");
  for (int i = 0; i < n_sec; i++)
    if (!(sec_edges[i] < sec_edges[i + 1]))
      return fail(TC_EINVAL, "tc_halo_bins: sec_edges must increase strictly");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_halo_bins: cannot select CUDA device " +
                                           std::to_string(device) + " (no CUDA device available?)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int n_cells = n_prim * n_sec;
  const size_t smem = (size_t)(6 * n_cells + 2) * sizeof(unsigned) +
                      (size_t)(n_prim + n_sec + 2) * sizeof(double);
  if (smem > (size_t)kSmemLimit)
    return fail(TC_EUNSUPPORTED, "tc_halo_bins: " + std::to_string(n_cells) +
                                     " cells do not fit the per-block table in shared memory");
  // 10**edge like the reference (x_min, x_max of distribution_index, tabcorr.py:219-220)
  std::vector<double> host((size_t)2 * n_cells + n_prim + n_sec + 2);
  double* cell_min = host.data();
  double* cell_inv = cell_min + n_cells;
  for (int s = 0; s < n_sec; s++) {
    for (int p = 0; p < n_prim; p++) {
      const double lo = std::pow(10.0, prim_edges[p]), hi = std::pow(10.0, prim_edges[p + 1]);
      cell_min[s * n_prim + p] = lo;
      cell_inv[s * n_prim + p] = 1.0 / (hi - lo);
    }
  }
  std::copy(prim_edges, prim_edges + n_prim + 1, cell_inv + n_cells);
  std::copy(sec_edges, sec_edges + n_sec + 1, cell_inv + n_cells + n_prim + 1);
  double* d_in = nullptr;
  unsigned long long* d_out = nullptr;
  // stream-ordered allocations: served from the driver's pool after the first call
  TC_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_in), host.size() * sizeof(double), stream));
  if (cudaMallocAsync(reinterpret_cast<void**>(&d_out),
                      (size_t)4 * n_cells * sizeof(unsigned long long), stream) != cudaSuccess) {
    (void)cudaFreeAsync(d_in, stream);
    (void)cudaGetLastError();
    return fail(TC_ENOMEM, "tc_halo_bins: out of device memory");
  }
  auto cleanup = [&]() { (void)cudaFreeAsync(d_in, stream); (void)cudaFreeAsync(d_out, stream); };
#define TC_HB(expr)                                                                        \
  do {                                                                                     \
    cudaError_t err__ = (expr);                                                            \
    if (err__ != cudaSuccess) {                                                            \
      (void)cudaGetLastError();                                                            \
      cleanup();                                                                           \
      return fail(TC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));        \
    }                                                                                      \
  } while (0)
  TC_HB(cudaMemcpyAsync(d_in, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice,
                        stream));
  TC_HB(cudaMemsetAsync(d_out, 0, (size_t)4 * n_cells * sizeof(unsigned long long), stream));
  HaloBinArgs args{};
  args.log_prim = log_prim;
  args.sec_pct = sec_pct;
  args.prim = prim;
  args.n_halos = n_halos;
  args.cell_min = d_in;
  args.cell_inv_width = d_in + n_cells;
  args.prim_edges = d_in + 2 * n_cells;
  args.sec_edges = d_in + 2 * n_cells + n_prim + 1;
  args.n_prim = n_prim;
  args.n_sec = n_sec;
  args.counts = d_out;
  args.counts_open = d_out + n_cells;
  args.sum_hi = d_out + 2 * n_cells;
  args.sum_lo = d_out + 3 * n_cells;
  int n_sm = 0;
  int rc = device_sms(device, &n_sm);
  if (rc) { cleanup(); return rc; }
  if (n_halos > 0) {
    TC_HB(cudaFuncSetAttribute(halo_bins_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
    // blocks of contiguous halo ranges: a few waves for balance, and never more haloes per
    // block than the 32-bit per-block tables hold
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)kSmemLimit / (smem + 1024)));
    const long long for_balance = std::min<long long>((n_halos + 8191) / 8192, 4LL * n_sm * per_sm);
    const long long for_range = (n_halos + kHaloBinsPerBlock - 1) / kHaloBinsPerBlock;
    const int grid = (int)std::max<long long>(1, std::max(for_balance, for_range));
    halo_bins_kernel<<<grid, 256, smem, stream>>>(args);
    TC_HB(cudaGetLastError());
  }
  std::vector<unsigned long long> res((size_t)4 * n_cells);
  TC_HB(cudaMemcpyAsync(res.data(), d_out, res.size() * sizeof(unsigned long long),
                        cudaMemcpyDeviceToHost, stream));
  TC_HB(cudaStreamSynchronize(stream));
#undef TC_HB
  cleanup();
  for (int c = 0; c < n_cells; c++) {
    n_h_out[c] = (double)res[c];
    const unsigned long long members = res[n_cells + c];
    n_members_out[c] = (double)members;
    if (members == 0) {
      mean_out[c] = std::nan("");
    } else {
      // mean offset = (sum_hi 2^26 + sum_lo) / 2^52 / members, in long double (64-bit mantissa)
      const long double total = (long double)res[2 * n_cells + c] * 67108864.0L +
                                (long double)res[3 * n_cells + c];
      const long double frac = total / 4503599627370496.0L / (long double)members;
      mean_out[c] = (double)((long double)cell_min[c] + frac / (long double)cell_inv[c]);
    }
  }
  return TC_OK;
}

int tc_debug_math(int kind, const double* x, const double* y, double* out, int64_t n,
                  void* stream) {
  if (!x || !out || (kind != 0 && !y) || n < 0 || kind < 0 || kind > 3)
    return fail(TC_EINVAL, "tc_debug_math: bad argument");
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

int tc_measure_dfma_peak(int device, double* tflops) {
  if (!tflops) return fail(TC_EINVAL, "tc_measure_dfma_peak: NULL output");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_measure_dfma_peak: cannot select CUDA device");
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
    dfma_peak_kernel<<<n_sm, 512>>>(scratch, iters);
    TC_CUDA(cudaEventRecord(e1));
    TC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    TC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double tf = (double)n_sm * 512 * iters * 8 * 2.0 / (ms * 1e-3) * 1e-12;
    if (rep > 0) best = std::max(best, tf);
  }
  (void)cudaEventDestroy(e0);
  (void)cudaEventDestroy(e1);
  (void)cudaFree(scratch);
  *tflops = best;
  return TC_OK;
}

// ------------------------------------------------------------------------------------------
// peer-visible result slabs (multi-GPU path): CUDA IPC around plain cudaMalloc memory
// ------------------------------------------------------------------------------------------
int tc_peer_alloc(int device, size_t bytes, void** ptr_out, unsigned char* handle_out) {
  if (!ptr_out || !handle_out || bytes == 0) return fail(TC_EINVAL, "tc_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_peer_alloc: cannot select CUDA device");
  void* p = nullptr;
  TC_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t err = cudaIpcGetMemHandle(&h, p);
  if (err != cudaSuccess) {
    (void)cudaFree(p);
    (void)cudaGetLastError();
    return fail(TC_ECUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(err));
  }
  std::memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return TC_OK;
}

int tc_peer_open(int device, const unsigned char* handle, void** ptr_out) {
  if (!ptr_out || !handle) return fail(TC_EINVAL, "tc_peer_open: bad argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_peer_open: cannot select CUDA device");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  TC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return TC_OK;
}

int tc_peer_close(int device, void* ptr) {
  if (!ptr) return TC_OK;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_peer_close: cannot select CUDA device");
  TC_CUDA(cudaIpcCloseMemHandle(ptr));
  return TC_OK;
}

int tc_peer_free(int device, void* ptr) {
  if (!ptr) return TC_OK;
  DeviceGuard guard(device);
  if (!guard.ok) return fail(TC_ECUDA, "tc_peer_free: cannot select CUDA device");
  TC_CUDA(cudaFree(ptr));
  return TC_OK;
}

}  // extern "C"
