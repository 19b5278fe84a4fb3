// Legacy warp-level mma.sync rates on B200 (register operands only): TF32 m16n8k8, BF16 m16n8k16,
// next to FP64 m8n8k4.  Decides whether a 3xTF32 variant of the contraction on mma.sync can pay.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_peaks mma_sync_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;} } while (0)

template <int ILP>
__global__ void __launch_bounds__(512, 1) tf32_kernel(float* out, int iters) {
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800100u, 0x3f800200u, 0x3f800300u};
  unsigned b[2] = {0x3f800400u, 0x3f800500u + threadIdx.x};
  float c[ILP][4];
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 1234.5f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(512, 1) bf16_kernel(float* out, int iters) {
  unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f81u, 0x3f803f82u, 0x3f803f83u};
  unsigned b[2] = {0x3f803f84u, 0x3f803f85u + threadIdx.x};
  float c[ILP][4];
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = i + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 1234.5f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
double timeit(K launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  int n_sm = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
  float* out;
  CK(cudaMalloc(&out, (size_t)n_sm * 512 * sizeof(float)));
  const int iters = 1 << 15;
  for (int warps : {4, 8, 16}) {
    double ms = timeit([&] { tf32_kernel<8><<<n_sm, warps * 32>>>(out, iters); });
    printf("tf32 m16n8k8  warps/SM %2d ilp 8 : %8.3f ms  %8.1f TFLOP/s\n", warps, ms,
           (double)n_sm * warps * iters * 8 * 2.0 * 16 * 8 * 8 / (ms * 1e-3) * 1e-12);
    ms = timeit([&] { bf16_kernel<8><<<n_sm, warps * 32>>>(out, iters); });
    printf("bf16 m16n8k16 warps/SM %2d ilp 8 : %8.3f ms  %8.1f TFLOP/s\n", warps, ms,
           (double)n_sm * warps * iters * 8 * 2.0 * 16 * 8 * 16 / (ms * 1e-3) * 1e-12);
  }
  CK(cudaGetLastError());
  return 0;
}
