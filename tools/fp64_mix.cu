// How much FP64 tensor (DMMA) throughput does an interleaved scalar FP64 instruction cost on B200?
// 12 warps per SM: DW of them run a register-only DMMA loop, the others dependent DFMA chains with
// ILP independent chains each.  Reports DMMA and DFMA rates and the DMMA pipe time lost per DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(512, 1) mix_kernel(double* out, long long* clocks, int dmma_warps_per_smsp,
                                                     int dfma_warps_per_smsp, int dmma_iters, int dfma_iters) {
  const int warp = threadIdx.x >> 5, slot = warp >> 2;  // warp % 4 = SMSP, slot = index on the SMSP
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - 1e-9 * threadIdx.x;
  double s = 0;
  long long t0 = clock64();
  if (slot < dmma_warps_per_smsp) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < dmma_iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma884(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  } else if (slot < dmma_warps_per_smsp + dfma_warps_per_smsp) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = i;
    for (int it = 0; it < dfma_iters; it++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i];
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) clocks[blockIdx.x * 16 + warp] = t1 - t0;
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int dmma_w, int dfma_w, int dmma_iters, int dfma_iters, double* out, long long* clocks, int n_sm) {
  CK(cudaMemset(clocks, 0, n_sm * 16 * sizeof(long long)));
  mix_kernel<ILP><<<n_sm, 512>>>(out, clocks, dmma_w, dfma_w, dmma_iters, dfma_iters);
  CK(cudaDeviceSynchronize());
  mix_kernel<ILP><<<n_sm, 512>>>(out, clocks, dmma_w, dfma_w, dmma_iters, dfma_iters);
  CK(cudaDeviceSynchronize());
  static long long h[148 * 16 + 64];
  CK(cudaMemcpy(h, clocks, n_sm * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  // average over SM 0..n_sm-1 of warps on SMSP 0
  double dm = 0, df = 0;
  int ndm = 0, ndf = 0;
  for (int sm = 0; sm < n_sm; sm++)
    for (int w = 0; w < 16; w++) {
      int slot = w >> 2;
      if (slot < dmma_w) { dm += h[sm * 16 + w]; ndm++; }
      else if (slot < dmma_w + dfma_w) { df += h[sm * 16 + w]; ndf++; }
    }
  dm = ndm ? dm / ndm : 0;
  df = ndf ? df / ndf : 0;
  // per SMSP: cycles per DMMA while the DFMA warps run (if DFMA warps outlast the DMMA warps)
  double cyc_per_dmma = dm / ((double)dmma_iters * 8 * (dmma_w ? dmma_w : 1));
  double cyc_per_dfma = df / ((double)dfma_iters * ILP * (dfma_w ? dfma_w : 1));
  printf("dmma_warps/smsp=%d dfma_warps/smsp=%d ilp=%d : dmma warps %.0f cyc (%.2f cyc/DMMA/smsp), dfma warps %.0f cyc (%.2f cyc/DFMA/smsp)\n",
         dmma_w, dfma_w, ILP, dm, cyc_per_dmma, df, cyc_per_dfma);
}

int main() {
  int n_sm = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
  double* out; long long* clocks;
  CK(cudaMalloc(&out, (size_t)n_sm * 512 * sizeof(double)));
  CK(cudaMalloc(&clocks, (size_t)n_sm * 16 * sizeof(long long)));
  const int N = 20000;
  run<1>(2, 0, N, 0, out, clocks, n_sm);
  run<1>(3, 0, N, 0, out, clocks, n_sm);
  run<1>(0, 1, 0, 16 * N, out, clocks, n_sm);
  run<4>(0, 1, 0, 4 * N, out, clocks, n_sm);
  run<8>(0, 1, 0, 2 * N, out, clocks, n_sm);
  run<8>(0, 3, 0, 2 * N, out, clocks, n_sm);
  // mixes sized so that the DFMA warps run at least as long as the DMMA warps
  run<1>(2, 1, N, 16 * N, out, clocks, n_sm);
  run<2>(2, 1, N, 16 * N, out, clocks, n_sm);
  run<4>(2, 1, N, 16 * N, out, clocks, n_sm);
  run<8>(2, 1, N, 16 * N, out, clocks, n_sm);
  run<1>(3, 1, N, 16 * N, out, clocks, n_sm);
  run<4>(3, 1, N, 16 * N, out, clocks, n_sm);
  run<1>(2, 2, N, 16 * N, out, clocks, n_sm);
  run<4>(2, 2, N, 16 * N, out, clocks, n_sm);
  run<1>(1, 1, N, 16 * N, out, clocks, n_sm);
  run<1>(1, 3, N, 16 * N, out, clocks, n_sm);
  return 0;
}
