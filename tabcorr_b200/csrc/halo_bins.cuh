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
// of (x - x_min) / (x_max - x_min), in 13-bit slices summed in 32-bit words per block and in 64-bit
// words across blocks), so the result does not depend on the order of the atomics:
// bit-reproducible for any grid.  Measured: 3.0 TB/s = 46 % of the HBM copy rate.
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

// number of edges <= x (np.searchsorted(edges, x, side='right')); NaN sorts behind everything.
// The reference's bins are np.linspace edges, so the position is guessed from the first and the
// last edge and then corrected against the actual edges (exact for any ascending edges: up to
// three steps either way, else a binary search).
__device__ __forceinline__ int edges_at_or_below(const double* edges, int n_edges, double x,
                                                 double inv_step) {
  if (!(x == x)) return n_edges;
  if (x < edges[0]) return 0;
  if (x >= edges[n_edges - 1]) return n_edges;
  int k = min(max((int)((x - edges[0]) * inv_step) + 1, 1), n_edges - 1);   // edges[k-1] <= x < edges[k]?
#pragma unroll 1
  for (int step = 0; step < 3; step++) {
    if (edges[k - 1] > x) k--;
    else if (edges[k] <= x) k++;
    else return k;
  }
  if (edges[k - 1] <= x && x < edges[k]) return k;
  int lo = 0, hi = n_edges;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (edges[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Haloes per block at most: the per-block tables are 32-bit (native shared-memory atomics; 64-bit
// ones are compare-and-swap loops that collapse under the contention of the low-mass cells), and
// the fixed-point position is accumulated in four 13-bit slices: 2^13 * 2^18 < 2^32.
constexpr long long kHaloBinsPerBlock = 1LL << 18;

__global__ void __launch_bounds__(256) halo_bins_kernel(const HaloBinArgs args) {
  extern __shared__ __align__(16) unsigned hb_smem[];
  const int n_cells = args.n_prim * args.n_sec;
  unsigned* s_count = hb_smem;                 // [n_cells] histogram2d semantics
  unsigned* s_open = s_count + n_cells;        // [n_cells] digitize semantics
  unsigned* s_q = s_open + n_cells;            // [4][n_cells] 13-bit slices of the positions
  double* s_pe = reinterpret_cast<double*>(s_q + 4 * n_cells + ((6 * n_cells) & 1));
  double* s_se = s_pe + args.n_prim + 1;
  for (int i = threadIdx.x; i < 6 * n_cells; i += blockDim.x) hb_smem[i] = 0u;
  for (int i = threadIdx.x; i <= args.n_prim; i += blockDim.x) s_pe[i] = args.prim_edges[i];
  for (int i = threadIdx.x; i <= args.n_sec; i += blockDim.x) s_se[i] = args.sec_edges[i];
  __syncthreads();
  // a contiguous range of at most kHaloBinsPerBlock haloes per block
  const double inv_step_p = args.n_prim / (s_pe[args.n_prim] - s_pe[0]);
  const double inv_step_s = args.n_sec / (s_se[args.n_sec] - s_se[0]);
  const long long per_block = (args.n_halos + gridDim.x - 1) / gridDim.x;
  const long long lo = blockIdx.x * per_block, hi = min(lo + per_block, args.n_halos);
  // One halo: its three doubles are loaded up front (the mean's operand too, although only members
  // of a cell need it: a second, dependent trip to HBM per halo costs more than the bytes), and
  // two haloes per thread are in flight.
  auto bin_one = [&](double lp, double sp, double pr) {
    // np.histogramdd: searchsorted(side='right'), values on the last edge belong to the last bin
    const int ip = edges_at_or_below(s_pe, args.n_prim + 1, lp, inv_step_p);
    const int is = edges_at_or_below(s_se, args.n_sec + 1, sp, inv_step_s);
    const bool p_edge = lp == s_pe[args.n_prim], s_edge = sp == s_se[args.n_sec];
    const int ip_closed = ip - (p_edge ? 1 : 0), is_closed = is - (s_edge ? 1 : 0);
    if (ip_closed >= 1 && ip_closed <= args.n_prim && is_closed >= 1 && is_closed <= args.n_sec)
      atomicAdd(&s_count[(is_closed - 1) * args.n_prim + ip_closed - 1], 1u);
    // np.digitize(right=False) of sort_into_bins: the last edge is outside
    if (ip >= 1 && ip <= args.n_prim && is >= 1 && is <= args.n_sec) {
      const int cell = (is - 1) * args.n_prim + ip - 1;
      double v = (pr - args.cell_min[cell]) * args.cell_inv_width[cell];
      v = fmin(fmax(v, 0.0), 1.0);     // log10 rounding can leave a member a hair outside
      const unsigned long long q = (unsigned long long)(v * 4503599627370496.0);   // 2^52
      atomicAdd(&s_open[cell], 1u);
      atomicAdd(&s_q[cell], (unsigned)(q & 0x1fffu));
      atomicAdd(&s_q[n_cells + cell], (unsigned)((q >> 13) & 0x1fffu));
      atomicAdd(&s_q[2 * n_cells + cell], (unsigned)((q >> 26) & 0x1fffu));
      atomicAdd(&s_q[3 * n_cells + cell], (unsigned)(q >> 39));   // 14 bits: v = 1 gives 2^13
    }
  };
  long long i = lo + threadIdx.x;
  for (; i + blockDim.x < hi; i += 2 * blockDim.x) {
    const long long j = i + blockDim.x;
    const double lp0 = args.log_prim[i], sp0 = args.sec_pct[i], pr0 = args.prim[i];
    const double lp1 = args.log_prim[j], sp1 = args.sec_pct[j], pr1 = args.prim[j];
    bin_one(lp0, sp0, pr0);
    bin_one(lp1, sp1, pr1);
  }
  if (i < hi) bin_one(args.log_prim[i], args.sec_pct[i], args.prim[i]);
  __syncthreads();
  for (int c = threadIdx.x; c < n_cells; c += blockDim.x) {
    if (s_count[c]) atomicAdd(&args.counts[c], (unsigned long long)s_count[c]);
    if (s_open[c]) {
      atomicAdd(&args.counts_open[c], (unsigned long long)s_open[c]);
      // sum of q = slices recombined; kept as the two 26-bit halves the host expects
      const unsigned long long low = (unsigned long long)s_q[c] +
                                     ((unsigned long long)s_q[n_cells + c] << 13);
      const unsigned long long high = (unsigned long long)s_q[2 * n_cells + c] +
                                      ((unsigned long long)s_q[3 * n_cells + c] << 13);
      atomicAdd(&args.sum_lo[c], low);
      atomicAdd(&args.sum_hi[c], high);
    }
  }
}

}  // namespace
