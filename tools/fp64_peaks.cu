// FP64 pipe microbenchmark for B200 (sm_100a): measures the DMMA (mma.sync f64) and DFMA
// peaks that bound the quadratic-form kernel, whether the two contend for the same pipe,
// and the cost of the double-precision transcendentals used by the occupation kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: all warps DMMA; mode 1: all warps DFMA; mode 2: even warps DMMA, odd warps DFMA
template <int ILP>
__global__ void pipe_kernel(double* out, int iters, int mode, double seed) {
  int warp = threadIdx.x >> 5;
  double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-9 * threadIdx.x;
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = i; c[i][1] = -i; }
  bool do_mma = (mode == 0) || (mode == 2 && (warp & 1) == 0);
  if (do_mma) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) dmma884(c[i][0], c[i][1], a, b);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        c[i][0] = fma(c[i][0], a, b);
        c[i][1] = fma(c[i][1], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// transcendental throughput: op 0 erf, 1 exp, 2 log, 3 pow, 4 exp10, 5 log+exp chain
__global__ void trans_kernel(double* out, int iters, int op, double seed) {
  double x = seed + 1e-6 * (threadIdx.x + blockIdx.x * blockDim.x);
  double acc = 0;
  for (int it = 0; it < iters; it++) {
    double y;
    switch (op) {
      case 0: y = erf(x); break;
      case 1: y = exp(x); break;
      case 2: y = log(x + 1.5); break;
      case 3: y = pow(x + 1.5, 0.83); break;
      case 4: y = exp10(x); break;
      default: y = exp(0.83 * log(x + 1.5)); break;
    }
    acc += y;
    x += 1e-7;
  }
  if (acc == 12345.678) out[0] = acc;
}

// DMMA fed from shared memory like the real kernel: each warp computes an (8*MT) x (8*NT) tile,
// A fragments streamed from global (L2) via 128-bit loads, B fragments from shared memory.
template <int NT>
__global__ void fed_kernel(double* out, const double* __restrict__ A, int ksteps, int reps, size_t a_stride_warp) {
  extern __shared__ double sB[];  // [ksteps][NT][32]
  for (int i = threadIdx.x; i < ksteps * NT * 32; i += blockDim.x) sB[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double c[2][NT][2];
#pragma unroll
  for (int m = 0; m < 2; m++)
#pragma unroll
    for (int n = 0; n < NT; n++) { c[m][n][0] = 0; c[m][n][1] = 0; }
  const double2* Ap = reinterpret_cast<const double2*>(A + (blockIdx.x * (blockDim.x >> 5) + warp) * a_stride_warp) + lane;
  for (int rep = 0; rep < reps; rep++) {
    const double2* ap = Ap;
    for (int ks = 0; ks < ksteps; ks++) {
      double2 a = __ldg(ap); ap += 32;
#pragma unroll
      for (int n = 0; n < NT; n++) {
        double b = sB[(ks * NT + n) * 32 + lane];
        dmma884(c[0][n][0], c[0][n][1], a.x, b);
        dmma884(c[1][n][0], c[1][n][1], a.y, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int m = 0; m < 2; m++)
#pragma unroll
    for (int n = 0; n < NT; n++) s += c[m][n][0] + c[m][n][1];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, 1 << 24));
  const int iters = 1 << 14;
  printf("# DMMA.884 / DFMA peak: mode(0=dmma,1=dfma,2=mixed) warps/SM ILP -> TFLOP/s\n");
  for (int mode = 0; mode < 3; mode++) {
    for (int warps : {4, 8, 16, 32}) {
      for (int ilp : {1, 2, 4, 8}) {
        auto launch = [&]() {
          dim3 g(sms), b(warps * 32);
          if (ilp == 1) pipe_kernel<1><<<g, b>>>(out, iters, mode, 1.0);
          if (ilp == 2) pipe_kernel<2><<<g, b>>>(out, iters, mode, 1.0);
          if (ilp == 4) pipe_kernel<4><<<g, b>>>(out, iters, mode, 1.0);
          if (ilp == 8) pipe_kernel<8><<<g, b>>>(out, iters, mode, 1.0);
        };
        float ms = time_ms(launch);
        double mma_warps = mode == 0 ? warps : (mode == 2 ? warps / 2.0 : 0);
        double fma_warps = mode == 1 ? warps : (mode == 2 ? warps / 2.0 : 0);
        double fl_mma = (double)sms * mma_warps * iters * ilp * 8 * 8 * 4 * 2;
        double fl_fma = (double)sms * fma_warps * iters * ilp * 2 * 32 * 2;
        printf("mode %d warps %2d ilp %d : %8.3f ms  dmma %7.2f TF  dfma %7.2f TF  total %7.2f TF\n", mode, warps, ilp, ms,
               fl_mma / ms * 1e-9, fl_fma / ms * 1e-9, (fl_mma + fl_fma) / ms * 1e-9);
      }
    }
  }
  printf("# transcendental throughput (G evals/s), 148*8 CTAs x 256 threads\n");
  const char* names[] = {"erf", "exp", "log", "pow", "exp10", "exp(a*log)"};
  for (int op = 0; op < 6; op++) {
    int it2 = 2048;
    auto launch = [&]() { trans_kernel<<<sms * 8, 256>>>(out, it2, op, 0.3); };
    float ms = time_ms(launch);
    printf("%-12s %8.3f ms  %8.2f Geval/s\n", names[op], ms, (double)sms * 8 * 256 * it2 / ms * 1e-6);
  }
  printf("# smem/L2-fed DMMA (16 x 8NT warp tiles): NT warps/SM -> TFLOP/s\n");
  {
    const int ksteps = 240 / 4;
    size_t a_stride = (size_t)ksteps * 64;  // doubles per warp stream
    double* A; size_t nA = a_stride * sms * 16;
    CK(cudaMalloc(&A, nA * sizeof(double)));
    CK(cudaMemset(A, 0, nA * sizeof(double)));
    for (int warps : {4, 8, 12, 16}) {
      for (int nt : {2, 4, 8}) {
        int reps = 256;
        size_t smem = (size_t)ksteps * nt * 32 * sizeof(double);
        auto launch = [&]() {
          if (nt == 2) { fed_kernel<2><<<sms, warps * 32, smem>>>(out, A, ksteps, reps, a_stride); }
          if (nt == 4) { fed_kernel<4><<<sms, warps * 32, smem>>>(out, A, ksteps, reps, a_stride); }
          if (nt == 8) { cudaFuncSetAttribute(fed_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                         fed_kernel<8><<<sms, warps * 32, smem>>>(out, A, ksteps, reps, a_stride); }
        };
        float ms = time_ms(launch);
        double fl = (double)sms * warps * reps * ksteps * nt * 2 * 512;
        printf("fed NT %d warps %2d : %8.3f ms  %7.2f TF\n", nt, warps, ms, fl / ms * 1e-9);
      }
    }
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
