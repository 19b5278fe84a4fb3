// tabcorr_b200 -- error handling, layout structs shared by host and device, MMA / load helpers.
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tabcorr_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------
thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define TC_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t err__ = (expr);                                                            \
    if (err__ != cudaSuccess) {                                                            \
      (void)cudaGetLastError();                                                            \
      return fail(TC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));        \
    }                                                                                      \
  } while (0)

// ------------------------------------------------------------------------------------------
// constants shared by host and device
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 384;        // 12 warps, 3 per SM sub-partition, <= 170 registers each
constexpr int kWarps = kThreads / 32;
constexpr int kOccUnroll = 5;        // quadrature nodes in flight per lane (n_gauss_prim = 10 default)
constexpr int kGroupRows = 2;        // rows (secondary-percentile bins) sharing one mass bin
constexpr int kSmemLimit = 227 * 1024;

// One unit of contraction work.  Auto mode: radial bin `r`, 16-row tiles [mt0, mt1); for tile mt
// the k-steps (4 table rows each) [k_begin, min(4 (mt + 1), k_cap)) are multiplied.  Cross mode:
// 16-radial-bin tile `r`, k-steps [k_begin, k_cap).  `part_row` is where the result goes.
struct Chunk {
  int r, mt0, mt1, k_begin, k_cap, part_row, pad0, pad1;
};

// One mass bin of the leauthaud11 kernel (leauthaud11.cuh): the centrals group and / or the
// satellites group over the same quadrature nodes, with everything the kernel needs about their
// rows in one record (no chains of dependent loads at the start and the end of a bin).
struct alignas(16) L11Bin {
  int cen, sat;          // group indices, -1: absent
  int same_w;            // 1: the satellites rows carry the same normalised weights as the centrals rows
  int pad0;
  int row[4];            // padded rows {centrals 0, centrals 1, satellites 0, satellites 1}, -1: absent
  int dst[4];            // their reference row index (column of the occupation output), -1: none
  double first_logm;     // log10 mass of the first node
  double pad1;
  double pct[4];         // secondary-property percentile of the rows
};

struct OccPlan {       // device pointers, one per (layout, n_gauss)
  int n_groups;
  int n_cen_groups;         // groups are ordered centrals first
  int n_gauss;
  int n_gauss_pad;          // n_gauss rounded up to a multiple of `unroll`; padding nodes have zero weight
  int unroll;               // nodes evaluated per iteration: kOccUnroll when it divides n_gauss, else 2
  int zero_row;             // index of an all-zero row of row_c (second row of 1-row groups)
  const double* node_logm;  // [n_groups, G]  log10 of the node masses
  const double* node_m;     // [n_groups, G]  node masses
  const double* node_inv_m; // [n_groups, G]  1 / node mass
  const int* grp_rows;      // [n_groups, kGroupRows] padded row index or -1
  const int* grp_is_sat;    // [n_groups]
  const double* row_c;      // [n_pad, G] normalised quadrature weights
  const double* row_nh;     // [n_pad]
  const double* row_pct;    // [n_pad] secondary-property percentile of the row
  // series evaluation of the bin averages (occupation.cuh, occupation_item_series)
  const double4* grp_ser;   // [n_groups] centrals {log10 M centre, half range d of the nodes, 0, 0};
                            //            satellites {m_ref, u_max, lowest, highest node mass}
  const double2* grp_mom;   // [kSerMom, n_groups] {row 0, row 1}: k-th scaled node moment / k!
  const unsigned char* cen_terms;  // [kSerBuckets] Hermite-series terms per bucket of h; 255: nodes
  const unsigned char* sat_terms;  // [kSerBuckets] binomial-series terms per bucket of y; 255: nodes
  double cen_d_max;         // largest half range of a centrals group (log10 M)
  // leauthaud11 kernel: mass bins = {centrals group, satellites group} over identical node masses
  // (-1: no such group), evaluated together (leauthaud11.cuh)
  const L11Bin* l11_bins;   // [n_l11_bins]
  int n_l11_bins;
};

struct LayoutDev {
  int n_rows;          // N of the table
  int n_pad;           // padded rows, multiple of 16
  int nc_pad;          // first satellite row in padded order
  int n_parts;         // scratch rows per draw tile
  int n_chunks;
  int n_out;           // outputs per draw: Reff * n_comp
  long long ks_per_r;  // k-steps per radial bin (auto) / per 16-bin tile (cross) in the A stream
  const double2* afrag;
  const float4* afrag32;   // 3xTF32 mode: per k8-step a 32-lane block of high parts, then of low parts
  long long ks8_per_r;     // k8-steps per radial bin in afrag32
  const Chunk* chunks;
  const long long* chunk_cost_prefix;  // [n_chunks + 1] cumulative cost of the sorted chunks
  const int* out_ptr;    // [n_out + 1] CSR: which scratch rows sum to output o
  const int* out_parts;
  const int* pad_to_row;  // [n_pad] reference row index or -1
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// The table stream: read-only, not worth an L1 line (each warp reads its fragment once per chunk),
// but it is THE data to keep in L2 -- every draw tile of every CTA re-reads all of it, while the
// scratch rows and results written beside it are touched once.  Without a hint the no-allocate
// loads are evict-first in L2 and the streamed writes push table lines out (ncu: 7.6 M of the
// 282 M evict-first sector reads missed, 150 MB of DRAM reads per launch for a 4.9 MB table).
__device__ __forceinline__ unsigned long long l2_keep_policy() {
  unsigned long long policy;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
__device__ __forceinline__ double2 ld_stream(const double2* p, unsigned long long policy) {
  double2 v;
#ifdef TC_A_NO_HINT
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#else
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
               : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
#endif
  return v;
}

// internal kernel mode: auto-correlation table contracted in 3xTF32 (tc_predict_batch precision 1)
constexpr int kModeAutoTf32 = 2;

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float4& a, float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
        "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// d = a * b with a fresh (zero) accumulator
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const float4& a, float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
        "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)), "f"(0.0f));
}

__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// 3xTF32 mode: index (in floats) of the HIGH part of (padded row i, draw b) in the shared W tile;
// the low part is two floats further.  m16n8k8 B-fragment order: per k8-step and n-tile the lane
// holding B[k = i % 4 (+ 4)][n = b % 8] reads one float4 {hi(k), hi(k + 4), lo(k), lo(k + 4)}.
// The tile has the same size as the FP64 one (8 bytes per weight).
template <int NT>
__device__ __forceinline__ int widx32(int i, int b) {
  return (((((i >> 3) * NT + (b >> 3)) << 5) + ((b & 7) << 2) + (i & 3)) << 2) + ((i >> 2) & 1);
}

// index of (padded row i, draw b) in the shared W tile: DMMA B-fragment order, so that the lane
// holding B[k = i % 4][n = b % 8] of k-step i / 4 and n-tile b / 8 reads consecutive doubles.
template <int NT>
__device__ __forceinline__ int widx(int i, int b) {
  return (((i >> 2) * NT + (b >> 3)) << 5) + ((b & 7) << 2) + (i & 3);
}

// store / load one tracer weight of the shared W tile in the representation of the mode
template <int NT, int MODE>
__device__ __forceinline__ void store_weight(double* Ws, int row, int b, double w) {
  if (MODE == kModeAutoTf32) {
    float* wf = reinterpret_cast<float*>(Ws) + widx32<NT>(row, b);
    const float hi = to_tf32((float)w);
    wf[0] = hi;
    wf[2] = to_tf32((float)(w - (double)hi));
  } else {
    Ws[widx<NT>(row, b)] = w;
  }
}
template <int NT, int MODE>
__device__ __forceinline__ double load_weight(const double* Ws, int row, int b) {
  if (MODE == kModeAutoTf32) {
    const float* wf = reinterpret_cast<const float*>(Ws) + widx32<NT>(row, b);
    return (double)wf[0] + (double)wf[2];
  }
  return Ws[widx<NT>(row, b)];
}

}  // namespace
