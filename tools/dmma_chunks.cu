// The product kernel's contraction code (run_chunk, copied from csrc/tabcorr_b200.cu) driven by a
// static schedule without occupation, flags or work queue: what does the chunk structure itself
// (triangular tiles of 4..60 k-steps, two-part k loop, row-dot, butterfly, scratch store) reach?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_chunks dmma_chunks.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct Chunk { int r, mt0, mt1, k_begin, k_cap, part_row, pad0, pad1; };
struct LayoutDev { long long ks_per_r; const double2* afrag; const Chunk* chunks; int n_chunks; int n_parts; };

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

template <int NT, int VARIANT>
__device__ __forceinline__ void run_chunk(const LayoutDev& lay, const Chunk& ch, const double* __restrict__ Ws,
                                          double* __restrict__ parts, int lane) {
  constexpr int BM = 8 * NT;
  const int g = lane >> 2, tig = lane & 3;
  double sums[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) sums[nt][0] = sums[nt][1] = 0.0;
  for (int mt = ch.mt0; mt < ch.mt1; mt++) {
    const int k_tile = 4 * (mt + 1);
    const int k_end = min(k_tile, ch.k_cap);
    const int k_both = VARIANT == 1 ? k_end : min(k_end, k_tile - 2);
    const double2* ap = lay.afrag + ((size_t)ch.r * lay.ks_per_r + 2 * (size_t)mt * (mt + 1) + ch.k_begin) * 32 + lane;
    const double* wk = Ws + (size_t)ch.k_begin * NT * 32 + lane;
    double acc[2][NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) acc[0][nt][0] = acc[0][nt][1] = acc[1][nt][0] = acc[1][nt][1] = 0.0;
    double2 a_next = ld_stream(ap);
    int ks = ch.k_begin;
    for (; ks < k_both; ks++) {
      const double2 a = a_next;
      ap += 32;
      a_next = ld_stream(ap);
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const double b = wk[nt * 32];
        dmma884(acc[0][nt], a.x, b);
        dmma884(acc[1][nt], a.y, b);
      }
      wk += NT * 32;
    }
    for (; ks < k_end; ks++) {
      const double2 a = a_next;
      ap += 32;
      a_next = ld_stream(ap);
#pragma unroll
      for (int nt = 0; nt < NT; nt++) dmma884(acc[1][nt], a.y, wk[nt * 32]);
      wk += NT * 32;
    }
    if (VARIANT != 2) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int row = 16 * mt + 8 * h + g;
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
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      double v = sums[nt][e];
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      sums[nt][e] = v;
    }
  }
  if (g == 0) {
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
      *reinterpret_cast<double2*>(parts + (size_t)ch.part_row * BM + 8 * nt + 2 * tig) = make_double2(sums[nt][0], sums[nt][1]);
  }
}

// SCHED 0: static round robin (warp w takes chunks w, w + 12, ...); 1: shared-memory counter
template <int NT, int VARIANT, int SCHED>
__global__ void __launch_bounds__(384, 1) chunk_kernel(LayoutDev lay, int n_pad, int tiles, double* parts) {
  extern __shared__ double Ws[];
  __shared__ int counter;
  for (int i = threadIdx.x; i < n_pad * 8 * NT; i += blockDim.x) Ws[i] = 1.0 + 1e-9 * i;
  if (threadIdx.x == 0) counter = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* my_parts = parts + (size_t)blockIdx.x * lay.n_parts * 8 * NT;
  if (SCHED == 0) {
    for (int i = warp; i < tiles * lay.n_chunks; i += 12) {
      const Chunk ch = lay.chunks[i % lay.n_chunks];
      run_chunk<NT, VARIANT>(lay, ch, Ws, my_parts, lane);
    }
  } else {
    for (;;) {
      int i = 0;
      if (lane == 0) i = atomicAdd(&counter, 1);
      i = __shfl_sync(0xffffffffu, i, 0);
      if (i >= tiles * lay.n_chunks) break;
      const Chunk ch = lay.chunks[i % lay.n_chunks];
      run_chunk<NT, VARIANT>(lay, ch, Ws, my_parts, lane);
    }
  }
}

template <int NT, int VARIANT, int SCHED>
void run(const char* label, LayoutDev lay, int n_pad, double dmma_per_tile, double* parts, int n_sm) {
  size_t smem = (size_t)n_pad * 8 * NT * sizeof(double);
  CK(cudaFuncSetAttribute(chunk_kernel<NT, VARIANT, SCHED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = 10;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0));
    chunk_kernel<NT, VARIANT, SCHED><<<n_sm, 384, smem>>>(lay, n_pad, tiles, parts);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double flops = (double)n_sm * tiles * dmma_per_tile * NT * 512.0;
  printf("%-44s NT=%d : %8.3f ms  %6.2f TFLOP/s\n", label, NT, best, flops / (best * 1e-3) * 1e-12);
}

int main() {
  int n_sm = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
  const int n_pad = 240, T16 = 15, R = 20, pieces = 4;
  const long long ks_per_r = 2LL * T16 * (T16 + 1);
  std::vector<Chunk> chunks;
  int n_parts = 0;
  for (int r = 0; r < R; r++) {
    long long total = 0; for (int mt = 0; mt < T16; mt++) total += 4 * (mt + 1);
    long long acc = 0; int start = 0, piece = 0;
    for (int mt = 0; mt < T16; mt++) {
      acc += 4 * (mt + 1);
      if (mt == T16 - 1 || acc * pieces >= total * (piece + 1)) {
        Chunk c{}; c.r = r; c.mt0 = start; c.mt1 = mt + 1; c.k_begin = 0; c.k_cap = 1 << 28; c.part_row = n_parts++;
        chunks.push_back(c); start = mt + 1; piece++;
      }
    }
  }
  auto cost = [](const Chunk& c) { long long s = 0; for (int mt = c.mt0; mt < c.mt1; mt++) s += 4 * (mt + 1); return s; };
  std::stable_sort(chunks.begin(), chunks.end(), [&](const Chunk& a, const Chunk& b) { return cost(a) > cost(b); });
  double dmma_per_tile = 0;  // per n-tile
  for (int mt = 0; mt < T16; mt++) dmma_per_tile += R * (8.0 * mt + 6.0);
  double2* afrag; Chunk* dchunks; double* parts;
  size_t a_elems = (size_t)R * ks_per_r * 32 + 32;
  CK(cudaMalloc(&afrag, a_elems * sizeof(double2)));
  CK(cudaMemset(afrag, 0, a_elems * sizeof(double2)));
  CK(cudaMalloc(&dchunks, chunks.size() * sizeof(Chunk)));
  CK(cudaMemcpy(dchunks, chunks.data(), chunks.size() * sizeof(Chunk), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&parts, (size_t)n_sm * n_parts * 64 * sizeof(double)));
  LayoutDev lay{ks_per_r, afrag, dchunks, (int)chunks.size(), n_parts};
  printf("%d chunks per tile\n", (int)chunks.size());
  run<8, 0, 0>("real chunk code, static schedule", lay, n_pad, dmma_per_tile, parts, n_sm);
  run<8, 0, 1>("real chunk code, smem counter", lay, n_pad, dmma_per_tile, parts, n_sm);
  run<7, 0, 0>("real chunk code, static schedule", lay, n_pad, dmma_per_tile, parts, n_sm);
  run<7, 0, 1>("real chunk code, smem counter", lay, n_pad, dmma_per_tile, parts, n_sm);
  {
    double full = 0; for (int mt = 0; mt < T16; mt++) full += R * (8.0 * mt + 8.0);
    run<7, 1, 0>("no skipped half tiles (8mt+8 DMMA)", lay, n_pad, full, parts, n_sm);
  }
  run<7, 2, 0>("no row-dot", lay, n_pad, dmma_per_tile, parts, n_sm);
  run<4, 0, 0>("real chunk code, static schedule", lay, n_pad, dmma_per_tile, parts, n_sm);
  return 0;
}
