/* A plain C client of the C ABI in include/tabcorr_b200.h -- no Python, no torch: what a binding
 * from any host language does.  Used by tests/test_gpu_cabi_client.py, which writes the inputs,
 * runs this program on the GPU box and compares its output with the Python API bit for bit.
 *
 * usage: predict_client <input.bin> <output.bin>
 * input : int32 mode, n_rows, n_r, n_gauss, n_draws, separate; then float64 n_h[n_rows],
 *         log_min[n_rows], log_max[n_rows], sec_pct[n_rows], dist_index[n_rows]; int32
 *         is_sat[n_rows]; float64 matrix[n_r * cols], x01[n_gauss], w[n_gauss],
 *         theta[n_draws * TC_N_THETA]
 * output: float64 ngal[n_draws * n_ng], xi[n_draws * n_r * n_comp]
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tabcorr_b200.h"

#define CHECK_TC(call)                                                          \
  do {                                                                          \
    int rc_ = (call);                                                           \
    if (rc_ != TC_OK) {                                                         \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, tc_last_error());     \
      return 2;                                                                 \
    }                                                                           \
  } while (0)
#define CHECK_CUDA(call)                                                        \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));        \
      return 3;                                                                 \
    }                                                                           \
  } while (0)

static double* read_doubles(FILE* f, size_t n) {
  double* p = (double*)malloc((n ? n : 1) * sizeof(double));
  if (!p || fread(p, sizeof(double), n, f) != n) {
    fprintf(stderr, "short read\n");
    exit(4);
  }
  return p;
}

int main(int argc, char** argv) {
  if (argc != 3) {
    fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]);
    return 1;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t head[6];
  if (fread(head, sizeof(int32_t), 6, f) != 6) return 4;
  const int mode = head[0], n_rows = head[1], n_r = head[2], n_gauss = head[3];
  const int64_t n_draws = head[4];
  const int separate = head[5];
  const size_t cols = mode == TC_MODE_AUTO ? (size_t)n_rows * (n_rows + 1) / 2 : (size_t)n_rows;
  double* n_h = read_doubles(f, n_rows);
  double* log_min = read_doubles(f, n_rows);
  double* log_max = read_doubles(f, n_rows);
  double* pct = read_doubles(f, n_rows);
  double* dist = read_doubles(f, n_rows);
  int32_t* is_sat = (int32_t*)malloc(n_rows * sizeof(int32_t));
  if (fread(is_sat, sizeof(int32_t), n_rows, f) != (size_t)n_rows) return 4;
  double* matrix = read_doubles(f, (size_t)n_r * cols);
  double* x01 = read_doubles(f, n_gauss);
  double* w = read_doubles(f, n_gauss);
  double* theta = read_doubles(f, (size_t)n_draws * TC_N_THETA);
  fclose(f);

  if (tc_version() != TC_VERSION) {
    fprintf(stderr, "header / library version mismatch\n");
    return 5;
  }
  tc_table* table = NULL;
  const double* matrices[1] = {matrix};
  CHECK_TC(tc_table_create(&table, mode, n_rows, n_r, 1, n_h, log_min, log_max, pct, dist, is_sat,
                           matrices, 0));
  CHECK_TC(tc_table_plan(table, n_gauss, x01, w));
  if (tc_table_n_rows(table) != n_rows || tc_table_n_r(table) != n_r) return 5;

  const int n_ng = separate ? 2 : 1;
  const int n_comp = !separate ? 1 : (mode == TC_MODE_AUTO ? 3 : 2);
  const size_t n_ngal = (size_t)n_draws * n_ng, n_xi = (size_t)n_draws * n_r * n_comp;
  const size_t ws_bytes = tc_predict_workspace_bytes(table, n_draws, separate);
  double *theta_dev, *ngal_dev, *xi_dev;
  void* ws_dev;
  CHECK_CUDA(cudaMalloc((void**)&theta_dev, (size_t)n_draws * TC_N_THETA * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&ngal_dev, n_ngal * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&xi_dev, n_xi * sizeof(double)));
  CHECK_CUDA(cudaMalloc(&ws_dev, ws_bytes ? ws_bytes : 8));
  CHECK_CUDA(cudaMemcpy(theta_dev, theta, (size_t)n_draws * TC_N_THETA * sizeof(double),
                        cudaMemcpyHostToDevice));
  tc_model model = {TC_FAMILY_ZHENG07, 1 /* decorated */, 0, 0, 0.5, 0.0, 0.0};
  if (tc_model_n_theta(&model) != TC_N_THETA) return 5;
  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));
  CHECK_TC(tc_predict_batch(table, &model, n_gauss, theta_dev, 0, NULL, n_draws, separate,
                            TC_PRECISION_FP64, ngal_dev, n_ng, xi_dev, (int64_t)n_r * n_comp,
                            ws_dev, ws_bytes, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  double* ngal = (double*)malloc(n_ngal * sizeof(double));
  double* xi = (double*)malloc(n_xi * sizeof(double));
  CHECK_CUDA(cudaMemcpy(ngal, ngal_dev, n_ngal * sizeof(double), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(xi, xi_dev, n_xi * sizeof(double), cudaMemcpyDeviceToHost));

  /* the one-draw entry point takes its parameters from host memory */
  double one_ngal_host[2], *one_ngal_dev, *one_xi_dev;
  CHECK_CUDA(cudaMalloc((void**)&one_ngal_dev, 2 * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&one_xi_dev, (size_t)n_r * n_comp * sizeof(double)));
  CHECK_TC(tc_predict_one(table, &model, n_gauss, theta, separate, TC_PRECISION_FP64, one_ngal_dev,
                          n_ng, one_xi_dev, (int64_t)n_r * n_comp, ws_dev, ws_bytes, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  CHECK_CUDA(cudaMemcpy(one_ngal_host, one_ngal_dev, n_ng * sizeof(double), cudaMemcpyDeviceToHost));
  if (one_ngal_host[0] != ngal[0]) {
    fprintf(stderr, "tc_predict_one disagrees with tc_predict_batch\n");
    return 6;
  }

  f = fopen(argv[2], "wb");
  if (!f) return 1;
  fwrite(ngal, sizeof(double), n_ngal, f);
  fwrite(xi, sizeof(double), n_xi, f);
  fclose(f);
  CHECK_TC(tc_table_destroy(table));
  printf("ok: %lld draws, ngal[0] = %.17g\n", (long long)n_draws, ngal[0]);
  return 0;
}
