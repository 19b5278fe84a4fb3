// tabcorr_b200 -- halo-bin reductions of the tabulation side (the cheap half):
//   the n_h histogram over (log10 primary property, secondary-property percentile) cells and the
//   per-cell mean of the primary property that defines prim_haloprop_dist_index
//   (tabcorr/tabcorr.py:194-227; sort_into_bins :676-737).
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md).
//
// HBM-bound byte/integer work: 24 bytes per halo in, a table of a few hundred cells out.  Every
// CTA keeps a private copy of the cell table in shared memory (the mass function puts most haloes
// into a handful of low-mass cells: global atomics on those would serialise in L2) and flushes it
// once.  Counts are integers; the per-cell sums are accumulated in FIXED POINT (52 fractional bits
// of (x - x_min) / (x_max - x_min), split into two 26-bit halves summed in 64-bit integers), so
// the result does not depend on the order of the atomics: bit-reproducible for any grid.
#pragma once

#include "common.cuh"

namespace {

struct HaloBinArgs {
  const double* log_prim;   // [n] log10 of the primary halo property: decides the cell
  const double* sec_pct;    // [n] secondary-property percentile
  const double* prim;       // [n] primary halo property: what is averaged
  long long n_halos;
  const double* prim_edges; // [n_prim + 1] ascending
  const double* sec_edges;  // [n_sec + 1] ascending
  int n_prim, n_sec;
  const double* cell_min;   // [n_sec * n_prim] 10**lower edge of the cell's primary bin
  const double* cell_inv_width;   // 1 / (10**upper - 10**lower)
  unsigned long long* counts;     // [n_sec * n_prim] histogram2d semantics (last edge inclusive)
  unsigned long long* counts_open;  // digitize semantics (last edge exclusive): members of the mean
  unsigned long long* sum_hi;     // fixed-point sums of the members' scaled offsets
  unsigned long long* sum_lo;
};

// number of edges <= x (np.searchsorted(edges, x, side='right')); NaN sorts behind everything
__device__ __forceinline__ int edges_at_or_below(const double* edges, int n_edges, double x) {
  if (!(x == x)) return n_edges;
  int lo = 0, hi = n_edges;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (edges[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) halo_bins_kernel(const HaloBinArgs args) {
  extern __shared__ unsigned long long hb_smem[];
  const int n_cells = args.n_prim * args.n_sec;
  unsigned long long* s_count = hb_smem;
  unsigned long long* s_open = s_count + n_cells;
  unsigned long long* s_hi = s_open + n_cells;
  unsigned long long* s_lo = s_hi + n_cells;
  double* s_pe = reinterpret_cast<double*>(s_lo + n_cells);
  double* s_se = s_pe + args.n_prim + 1;
  for (int i = threadIdx.x; i < 4 * n_cells; i += blockDim.x) hb_smem[i] = 0ull;
  for (int i = threadIdx.x; i <= args.n_prim; i += blockDim.x) s_pe[i] = args.prim_edges[i];
  for (int i = threadIdx.x; i <= args.n_sec; i += blockDim.x) s_se[i] = args.sec_edges[i];
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < args.n_halos;
       i += (long long)gridDim.x * blockDim.x) {
    const double lp = args.log_prim[i], sp = args.sec_pct[i];
    // np.histogramdd: searchsorted(side='right'), values on the last edge belong to the last bin
    int ip = edges_at_or_below(s_pe, args.n_prim + 1, lp);
    int is = edges_at_or_below(s_se, args.n_sec + 1, sp);
    const bool p_edge = lp == s_pe[args.n_prim], s_edge = sp == s_se[args.n_sec];
    const int ip_closed = ip - (p_edge ? 1 : 0), is_closed = is - (s_edge ? 1 : 0);
    if (ip_closed >= 1 && ip_closed <= args.n_prim && is_closed >= 1 && is_closed <= args.n_sec)
      atomicAdd(&s_count[(is_closed - 1) * args.n_prim + ip_closed - 1], 1ull);
    // np.digitize(right=False) of sort_into_bins: the last edge is outside
    if (ip >= 1 && ip <= args.n_prim && is >= 1 && is <= args.n_sec) {
      const int cell = (is - 1) * args.n_prim + ip - 1;
      double v = (args.prim[i] - args.cell_min[cell]) * args.cell_inv_width[cell];
      v = fmin(fmax(v, 0.0), 1.0);     // log10 rounding can leave a member a hair outside
      const unsigned long long q = (unsigned long long)(v * 4503599627370496.0);   // 2^52
      atomicAdd(&s_open[cell], 1ull);
      atomicAdd(&s_hi[cell], q >> 26);
      atomicAdd(&s_lo[cell], q & 0x3ffffffull);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_cells; c += blockDim.x) {
    if (s_count[c]) atomicAdd(&args.counts[c], s_count[c]);
    if (s_open[c]) {
      atomicAdd(&args.counts_open[c], s_open[c]);
      atomicAdd(&args.sum_hi[c], s_hi[c]);
      atomicAdd(&args.sum_lo[c], s_lo[c]);
    }
  }
}

}  // namespace
