// DMMA k-loop microbenchmark for B200 (sm_100a): what fraction of the FP64 tensor peak does the
// contraction loop of predict_kernel reach as a function of its structure?  Same operand paths as
// the product kernel: A fragments streamed from global memory (L2 resident, one coalesced
// 16-byte load per lane and k-step, prefetched PF steps ahead), B fragments from shared memory
// (one LDS.64 per n-tile), 2 NT DMMAs per k-step, accumulators reset every TILE_K k-steps with an
// optional register row-dot epilogue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_loop dmma_loop.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// EPI: 0 none, 1 row-dot with W after every tile.  PF: prefetch distance in k-steps (1 or 2).
template <int NT, int EPI, int PF, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) loop_kernel(const double2* __restrict__ afrag, long long a_steps,
                                                      int n_rows4, int tile_k, int tiles, double* out) {
  extern __shared__ double Ws[];
  for (int i = threadIdx.x; i < n_rows4 * NT * 32; i += blockDim.x) Ws[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  double sums[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) sums[nt][0] = sums[nt][1] = 0.0;
  long long pos = ((long long)blockIdx.x * 16 + warp) * 977 % a_steps;
  for (int t = 0; t < tiles; t++) {
    pos = (pos + 7919) % (a_steps - tile_k - 4);
    const double2* ap = afrag + pos * 32 + lane;
    const double* wk = Ws + lane;
    double acc[2][NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
      acc[0][nt][0] = acc[0][nt][1] = acc[1][nt][0] = acc[1][nt][1] = 0.0;
    double2 a_q[PF];
#pragma unroll
    for (int p = 0; p < PF; p++) a_q[p] = ld_stream(ap + 32 * p);
    for (int ks = 0; ks < tile_k; ks++) {
      const double2 a = a_q[0];
#pragma unroll
      for (int p = 0; p + 1 < PF; p++) a_q[p] = a_q[p + 1];
      a_q[PF - 1] = ld_stream(ap + 32 * PF);
      ap += 32;
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const double b = wk[nt * 32];
        dmma884(acc[0][nt], a.x, b);
        dmma884(acc[1][nt], a.y, b);
      }
      wk += NT * 32;
      if ((ks & 63) == 63) wk = Ws + lane;
    }
    if (EPI) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int row = 16 * (t & 3) + 8 * h + g;
        const double* wr = Ws + (size_t)(row >> 2) * NT * 32 + (row & 3) + tig * 8;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          sums[nt][0] = fma(acc[h][nt][0], wr[nt * 32], sums[nt][0]);
          sums[nt][1] = fma(acc[h][nt][1], wr[nt * 32 + 4], sums[nt][1]);
        }
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        sums[nt][0] += acc[0][nt][0] + acc[1][nt][0];
        sums[nt][1] += acc[0][nt][1] + acc[1][nt][1];
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int nt = 0; nt < NT; nt++) s += sums[nt][0] + sums[nt][1];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT, int EPI, int PF, int WARPS>
void run(const char* label, const double2* afrag, long long a_steps, int tile_k, double* out,
         int n_sm) {
  const int n_rows4 = (tile_k < 64 ? tile_k : 64) + 8;
  const int warps = WARPS;
  size_t smem = (size_t)n_rows4 * NT * 32 * sizeof(double);
  CK(cudaFuncSetAttribute(loop_kernel<NT, EPI, PF, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = 2000 * 32 / tile_k;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    loop_kernel<NT, EPI, PF, WARPS><<<n_sm, warps * 32, smem>>>(afrag, a_steps, n_rows4, tile_k, tiles, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double flops = (double)n_sm * warps * tiles * tile_k * 2.0 * NT * 512.0;
  printf("%-28s NT=%d warps=%2d tile_k=%3d epi=%d pf=%d : %8.3f ms  %6.2f TFLOP/s\n", label, NT, warps,
         tile_k, EPI, PF, best, flops / (best * 1e-3) * 1e-12);
}

int main() {
  int n_sm = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
  const long long a_steps = 20 * 480 + 64;  // the N=240, R=20 table: 4.9 MB
  double2* afrag;
  CK(cudaMalloc(&afrag, (size_t)a_steps * 32 * sizeof(double2)));
  CK(cudaMemset(afrag, 0, (size_t)a_steps * 32 * sizeof(double2)));
  double* out;
  CK(cudaMalloc(&out, (size_t)n_sm * 512 * sizeof(double)));
#define RUN(NT, EPI, PF, W, label, tk) run<NT, EPI, PF, W>(label, afrag, a_steps, tk, out, n_sm)
  RUN(8, 0, 1, 4, "long loop", 128);
  RUN(8, 0, 1, 8, "long loop", 128);
  RUN(8, 0, 1, 12, "long loop", 128);
  RUN(4, 0, 1, 16, "long loop", 128);
  RUN(8, 0, 2, 12, "long loop", 128);
  RUN(7, 0, 1, 12, "long loop", 128);
  RUN(8, 1, 1, 8, "tile 32 + rowdot", 32);
  RUN(8, 1, 1, 12, "tile 32 + rowdot", 32);
  RUN(8, 1, 2, 12, "tile 32 + rowdot", 32);
  RUN(7, 1, 1, 12, "tile 32 + rowdot", 32);
  RUN(7, 1, 2, 12, "tile 32 + rowdot", 32);
  RUN(4, 1, 1, 12, "tile 32 + rowdot", 32);
  RUN(4, 1, 2, 12, "tile 32 + rowdot", 32);
  RUN(4, 1, 2, 16, "tile 32 + rowdot", 32);
  RUN(3, 1, 2, 12, "tile 32 + rowdot", 32);
  RUN(3, 1, 2, 16, "tile 32 + rowdot", 32);
  RUN(2, 1, 2, 16, "tile 32 + rowdot", 32);
  RUN(1, 1, 2, 16, "tile 32 + rowdot", 32);
  RUN(8, 1, 1, 12, "tile 8 + rowdot", 8);
  RUN(8, 1, 1, 12, "tile 16 + rowdot", 16);
  RUN(8, 1, 1, 12, "tile 64 + rowdot", 64);
  return 0;
}
