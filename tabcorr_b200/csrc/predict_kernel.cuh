// tabcorr_b200 -- the fused occupation + DMMA quadratic-form kernel and its finalize kernel.
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md).
#pragma once

#include "common.cuh"
#include "device_math.cuh"
#include "occupation.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// pipeline flags in shared memory: monotonically increasing counters, so a waiter can never be
// lapped (a parity-based mbarrier can: a warp that only ran occupation items of a tile may meet
// that tile's barrier one or two phases later)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void flag_wait(const int* counter, int target) {
  const volatile int* c = counter;
  while (*c < target) __nanosleep(40);
  __threadfence_block();   // acquire: order the W reads / writes that follow after the flag read
}
// all lanes call it after their last shared-memory access of the item
__device__ __forceinline__ void flag_signal(int* counter, int lane) {
  __threadfence_block();   // release: this lane's W accesses before the flag update
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicAdd(counter, 1);
  }
}

// ------------------------------------------------------------------------------------------
// fused occupation + contraction kernel
// ------------------------------------------------------------------------------------------
struct PredictArgs {
  LayoutDev lay;
  OccPlan plan;
  tc_model model;
  const double* theta;   // parameter draws or nullptr: theta[draw * theta_ds + k * theta_ps]
  long long theta_ds, theta_ps;
  const double* occ;     // [B, n_rows] or nullptr
  int theta_is_inline;   // one draw whose parameters travel in the launch arguments
  double theta_inline[TC_N_THETA];
  long long n_draws;
  long long n_tiles;
  double* parts;         // [n_tiles, n_parts, BM]
  double* ngal_tile;     // [n_tiles, 2, BM]  centrals / satellites number density
  int n_buf;             // W tiles in shared memory: 2 = occupation of tile t + 1 overlaps tile t
  int n_ranges_cen;      // occupation items per n-tile: group ranges of centrals ...
  int n_ranges_sat;      // ... and of satellites
  int pieces_cen, pieces_sat;   // draw pieces per type (series_item); 0: node-by-node items
  int occ_stride;        // every occ_stride-th slot of a tile's work list is an occupation item
  int tf32_segment;      // 3xTF32 mode: k8-steps per FP32 accumulation chain
  int stress_ns;         // test hook (TC_TUNE_STRESS): pseudo-random delays of up to this many
                         // nanoseconds before every occupation item and contraction chunk, to
                         // perturb the schedule the full / empty counters have to order
};

constexpr int kMaxWBuffers = 4;

struct PredictCtrl {
  int full[kMaxWBuffers];       // occupation items finished, per W buffer (n_occ per tile)
  int empty[kMaxWBuffers];      // warps that left a tile's work list, per W buffer (kWarps per tile)
  int next;                     // work-list cursor
  int first_lo, last_hi;        // chunk range of the CTA's first / last tile
  int n_local;                  // tiles this CTA works on
  long long tile_first;
  double theta_inline[TC_N_THETA];   // shared-memory copy of the inline parameters
  int ser_queue[kWarps][kSerQueue];  // per warp: (draw, group) pairs waiting for the node path
};

// One contraction chunk by one warp.  W is the draw tile in B-fragment order.
template <int NT, int MODE>
__device__ __forceinline__ void run_chunk(const LayoutDev& lay, const Chunk& ch,
                                          const double* __restrict__ Ws,
                                          double* __restrict__ parts, int lane) {
  constexpr int BM = 8 * NT;
  const int g = lane >> 2, tig = lane & 3;
  const unsigned long long keep = l2_keep_policy();
  if (MODE == TC_MODE_AUTO) {
    double sums[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) sums[nt][0] = sums[nt][1] = 0.0;
    for (int mt = ch.mt0; mt < ch.mt1; mt++) {
      const int k_tile = 4 * (mt + 1);                 // k-steps of the full lower-triangular tile
      const int k_end = min(k_tile, ch.k_cap);
      // the upper 8 rows of the tile are zero in its last two k-steps: skip their DMMAs
      const int k_both = min(k_end, k_tile - 2);
      const double2* ap = lay.afrag +
          ((size_t)ch.r * lay.ks_per_r + 2 * (size_t)mt * (mt + 1) + ch.k_begin) * 32 + lane;
      const double* wk = Ws + (size_t)ch.k_begin * NT * 32 + lane;
      double acc[2][NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; nt++)
        acc[0][nt][0] = acc[0][nt][1] = acc[1][nt][0] = acc[1][nt][1] = 0.0;
      double2 a_next = ld_stream(ap, keep);
      int ks = ch.k_begin;
      for (; ks < k_both; ks++) {
        const double2 a = a_next;
        ap += 32;
        a_next = ld_stream(ap, keep);  // the stream is padded by one k-step, always safe
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
        a_next = ld_stream(ap, keep);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(acc[1][nt], a.y, wk[nt * 32]);
        wk += NT * 32;
      }
      // row-dot: acc[h][nt][e] = (M' W)[row = 16 mt + 8 h + g][draw = 8 nt + 2 tig + e]
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
    }
    // fixed-order butterfly over the 8 row groups of the warp (lane xor 16, 8, 4): every lane ends
    // with the full sums; the lanes of row group 0 store them
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
        *reinterpret_cast<double2*>(parts + (size_t)ch.part_row * BM + 8 * nt + 2 * tig) =
            make_double2(sums[nt][0], sums[nt][1]);
    }
  } else {
    // cross mode: a 16-radial-bin tile times a k-range of W; the product is the output
    const double2* ap = lay.afrag + ((size_t)ch.r * lay.ks_per_r + ch.k_begin) * 32 + lane;
    const double* wk = Ws + (size_t)ch.k_begin * NT * 32 + lane;
    double acc[2][NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
      acc[0][nt][0] = acc[0][nt][1] = acc[1][nt][0] = acc[1][nt][1] = 0.0;
    double2 a_next = ld_stream(ap, keep);
    for (int ks = ch.k_begin; ks < ch.k_cap; ks++) {
      const double2 a = a_next;
      ap += 32;
      a_next = ld_stream(ap, keep);
#pragma unroll
      for (int nt = 0; nt < NT; nt++) {
        const double b = wk[nt * 32];
        dmma884(acc[0][nt], a.x, b);
        dmma884(acc[1][nt], a.y, b);
      }
      wk += NT * 32;
    }
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++)
        *reinterpret_cast<double2*>(parts + (size_t)(ch.part_row + 8 * h + g) * BM + 8 * nt +
                                    2 * tig) = make_double2(acc[h][nt][0], acc[h][nt][1]);
  }
}

// One contraction chunk in 3xTF32: every table entry and every weight is split into a TF32 high
// part and a TF32 low part (22 significant bits together); hi*hi + lo*hi + hi*lo are accumulated in
// FP32 by the warp-level m16n8k8 MMA (k8-steps of 8 table columns, 16-row tiles as in the FP64
// path), the row-dot and everything after it stay in FP64.  Relative error ~1e-7 of the sum of the
// term magnitudes (tests: 1e-6).
constexpr int kTf32Segment = 1 << 20;   // k8-steps per FP32 running sum (default: the whole tile)

template <int NT>
__device__ __forceinline__ void run_chunk_tf32(const LayoutDev& lay, const Chunk& ch,
                                               const double* __restrict__ Ws,
                                               double* __restrict__ parts, int lane,
                                               int segment) {
  constexpr int BM = 8 * NT;
  const int g = lane >> 2, tig = lane & 3;
  const float* Wf = reinterpret_cast<const float*>(Ws);
  const int k_begin = ch.k_begin >> 1, k_cap = ch.k_cap >> 1;   // k4-steps -> k8-steps
  double sums[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) sums[nt][0] = sums[nt][1] = 0.0;
  for (int mt = ch.mt0; mt < ch.mt1; mt++) {
    const int k_end = min(2 * (mt + 1), k_cap);
    const float4* ap = lay.afrag32 +
        (((size_t)ch.r * lay.ks8_per_r + (size_t)mt * (mt + 1) + k_begin) * 2) * 32 + lane;
    const float4* wk = reinterpret_cast<const float4*>(Wf) + (size_t)k_begin * NT * 32 + lane;
    float4 hi_next = ld_stream4(ap), lo_next = ld_stream4(ap + 32);
    int ks = k_begin;
    while (ks < k_end) {
      // optional: cut the FP32 running sum every `segment` k8-steps (row-dot into the FP64 sums)
      const int seg_end = min(ks + segment, k_end);
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
      for (; ks < seg_end; ks++) {
        const float4 a_hi = hi_next, a_lo = lo_next;
        ap += 64;
        hi_next = ld_stream4(ap);        // the stream is padded by one k8-step, always safe
        lo_next = ld_stream4(ap + 32);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          const float4 w = wk[nt * 32];  // {hi(k), hi(k + 4), lo(k), lo(k + 4)}
          // The tensor core truncates its FP32 accumulator after every MMA -- up to one ulp of the
          // RUNNING sum, always towards zero: chained over a 16-row tile's 30 k8-steps that is a
          // bias of 1e-6 (measured, tools/tf32_error.py).  So each k8-step starts from a zero
          // accumulator (its truncations are relative to the small increment) and is added to the
          // running sum with round-to-nearest FADDs, which are unbiased and nearly free.
          float d[4];
          mma_tf32_zero(d, a_lo, w.x, w.y);   // small terms first
          mma_tf32(d, a_hi, w.z, w.w);
          mma_tf32(d, a_hi, w.x, w.y);
#pragma unroll
          for (int j = 0; j < 4; j++) acc[nt][j] += d[j];
        }
        wk += NT * 32;
      }
      // row-dot: acc[nt][2 h + e] = (M' W)[row = 16 mt + 8 h + g][draw = 8 nt + 2 tig + e]
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int row = 16 * mt + 8 * h + g;
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const float* wf = Wf + widx32<NT>(row, 8 * nt + 2 * tig + e);
            const float a = acc[nt][2 * h + e];
            sums[nt][e] += (double)fmaf(a, wf[2], a * wf[0]);
          }
        }
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
      *reinterpret_cast<double2*>(parts + (size_t)ch.part_row * BM + 8 * nt + 2 * tig) =
          make_double2(sums[nt][0], sums[nt][1]);
  }
}

// The kernel is a barrier-free software pipeline over the CTA's draw tiles.  Work is a sequence of
// per-tile lists of S slots that the 12 warps take from one shared cursor:
//   slot 0                          number densities of tile j (one warp, sequential row order)
//   every occ_stride-th next slot   occupation item (n-tile, group range) of tile j + n_buf - 1,
//                                   written into the other W buffer
//   the remaining slots             contraction chunks of tile j, longest first
// Dependencies always point backwards in that sequence, so taking slots in order cannot deadlock:
// a chunk waits until full[buf] counts all occupation items of its tile, an occupation item until
// empty[buf] counts every warp having left the list of the tile that used its buffer before.
// SERIES selects the occupation items at compile time (series items / node-by-node items), so that
// a kernel only carries the code of the items it runs: the hot code of the 12 warps (DMMA loops +
// occupation + slot dispatch) has to stay inside the 32 KB instruction cache.
template <int NT, int MODE, bool SERIES>
__global__ void __launch_bounds__(kThreads, 1) predict_kernel(const PredictArgs args) {
  constexpr int BM = 8 * NT;
  extern __shared__ __align__(16) double smem[];
  const LayoutDev& lay = args.lay;
  const int n_buf = args.n_buf;
  const size_t tile_doubles = (size_t)lay.n_pad * BM;
  double* tab = smem + n_buf * tile_doubles;                               // math tables
  PredictCtrl* ctrl = reinterpret_cast<PredictCtrl*>(tab + kTabDoubles);
  const int tid = threadIdx.x, lane = tid & 31;

  for (size_t i = tid; i < n_buf * tile_doubles; i += kThreads) smem[i] = 0.0;  // padding rows stay 0
  load_math_tables(tab);

  const int n_occ = NT * (args.n_ranges_cen + args.n_ranges_sat);
  {
    // The grid cuts the total COST (n_tiles x per-tile chunk cost) into equal contiguous ranges, so
    // that every CTA gets the same amount of DMMA work whatever the number of draws; a CTA
    // recomputes the weights of the (at most two) tiles it shares with its neighbours.
    const long long tile_cost = lay.chunk_cost_prefix[lay.n_chunks];
    const long long total_cost = tile_cost * args.n_tiles;
    const long long cost_lo = total_cost / gridDim.x * blockIdx.x +
                              total_cost % gridDim.x * blockIdx.x / gridDim.x;
    const long long cost_hi = total_cost / gridDim.x * (blockIdx.x + 1) +
                              total_cost % gridDim.x * (blockIdx.x + 1) / gridDim.x;
    long long tile_first = cost_lo / tile_cost;
    long long tile_last = min((cost_hi + tile_cost - 1) / tile_cost, args.n_tiles);  // exclusive
    // first chunk whose start cost is >= the range boundary = number of prefix entries below it
    // (neighbouring CTAs count against the same value, so their chunk ranges meet exactly).  The
    // whole CTA counts in parallel: one round trip to L2 instead of a dependent binary search, which
    // is what a one-draw call waits for.
    const long long bound_lo = cost_lo - tile_first * tile_cost;
    const long long bound_hi = cost_hi - (tile_last - 1) * tile_cost;
    int below_lo = 0, below_hi = 0;
    for (int base = 0; base < lay.n_chunks; base += kThreads) {
      const int i = base + tid;
      const long long start = i < lay.n_chunks ? lay.chunk_cost_prefix[i] : tile_cost;
      below_lo += __syncthreads_count(i < lay.n_chunks && start < bound_lo);
      below_hi += __syncthreads_count(i < lay.n_chunks && start < bound_hi);
    }
    if (tid == 0) {
      int first_lo = bound_lo <= 0 ? 0 : bound_lo >= tile_cost ? lay.n_chunks : below_lo;
      if (first_lo >= lay.n_chunks) { tile_first++; first_lo = 0; }
      int last_hi = lay.n_chunks;
      if (tile_last > tile_first) {
        last_hi = bound_hi <= 0 ? 0 : bound_hi >= tile_cost ? lay.n_chunks : below_hi;
        if (last_hi <= (tile_last - 1 == tile_first ? first_lo : 0)) { tile_last--; last_hi = lay.n_chunks; }
      }
      ctrl->tile_first = tile_first;
      ctrl->n_local = (int)max(tile_last - tile_first, 0LL);
      ctrl->first_lo = first_lo;
      ctrl->last_hi = last_hi;
      ctrl->next = 0;
      for (int b = 0; b < kMaxWBuffers; b++) ctrl->full[b] = ctrl->empty[b] = 0;
    }
    if (args.theta_is_inline && tid < TC_N_THETA) ctrl->theta_inline[tid] = args.theta_inline[tid];
  }
  __syncthreads();
  // parameters of the draws: device (or mapped host) memory, or the launch arguments of a
  // one-draw call (a mapped-host read costs every CTA a PCIe round trip: 14 us per call)
  const double* theta_base = args.theta_is_inline ? ctrl->theta_inline : args.theta;
  const int n_local = ctrl->n_local;
  const long long tile_first = ctrl->tile_first;
  const int first_lo = ctrl->first_lo, last_hi = ctrl->last_hi;
  const int occ_ahead = n_buf - 1;
  const int S = 1 + lay.n_chunks + n_occ;
  const int stride = args.occ_stride;

  int left = -occ_ahead;   // lists [.., left) have been left behind by this warp
  int full_seen = -1;      // newest tile whose W this warp has seen complete
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(&ctrl->next, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    const int list = i / S - occ_ahead;          // tile (local index) whose list the slot is in
    const int s = i - (list + occ_ahead) * S;
    // this warp has finished everything it took from earlier lists: release those tiles
    const int upto = min(list, n_local);
    for (int t = max(left, 0); t < upto; t++) flag_signal(&ctrl->empty[t % n_buf], lane);
    left = max(left, upto);
    if (list >= n_local) break;

    int kind = 0, idx = 0;                       // 0 ngal, 1 occupation, 2 chunk
    if (s > 0) {
      const int u = s - 1, q = u / stride;
      if (u - q * stride == stride - 1 && q < n_occ) { kind = 1; idx = q; }
      else { kind = 2; idx = u - min(n_occ, q); }
    }

    if (args.stress_ns > 0 && kind != 0) {   // warp-uniform pseudo-random delay (test hook)
      const unsigned h = ((unsigned)i * 2654435761u) ^ ((unsigned)blockIdx.x * 40503u);
      __nanosleep((h >> 8) % (unsigned)args.stress_ns);
    }
    if (kind == 1) {
      // ---- occupation item idx of tile list + occ_ahead -> W[(list + occ_ahead) % n_buf] ------
      const int j = list + occ_ahead;
      if (j >= n_local) continue;
      const int buf = j % n_buf;
      if (j >= n_buf) flag_wait(&ctrl->empty[buf], (j / n_buf) * kWarps);
      double* Ws = smem + buf * tile_doubles;
      const int nt = idx % NT, q = idx / NT;
      if (!SERIES && theta_base != nullptr) {
        // node-by-node items (small tables: the series items' code would push the working set of
        // the 12 warps out of the instruction cache); a one-draw call spreads all 32 lanes over
        // the groups of the range (column 0 of the tile)
        const bool one_draw = args.n_draws == 1;
        const int b = one_draw ? 0 : 8 * nt + (lane & 7);
        long long draw = (tile_first + j) * BM + b;
        if (draw >= args.n_draws) draw = args.n_draws - 1;  // tail tile: recompute the last draw
        int g_begin, g_end;
        occupation_range(args.plan, args.n_ranges_cen, args.n_ranges_sat, q, g_begin, g_end);
        occupation_item(args.plan, args.model, theta_base + draw * args.theta_ds, args.theta_ps,
                        g_begin, g_end, tab,
                        [&](int row, double occ, double nh) {
                          store_weight<NT, MODE>(Ws, row, b, occ * nh);
                        },
                        one_draw ? lane : (lane >> 3), one_draw ? 32 : 4);
      } else if (SERIES && theta_base != nullptr) {
        const SeriesItem it = series_item(args.plan, args.n_ranges_cen, args.n_ranges_sat,
                                          args.pieces_cen, args.pieces_sat, q);
        const int col0 = 8 * nt + it.b_begin;
        const long long draw0 = (tile_first + j) * BM + col0;
        // draws beyond the batch are skipped: their columns keep weights of an earlier tile (or the
        // initial zeros), every column is contracted on its own and finalize ignores them
        const int n_b = (int)min((long long)it.n_b, args.n_draws - draw0);
        if (n_b > 0 && it.g_end > it.g_begin) {
          auto store = [&](int b, int, int row, double occ, double nh) {
            store_weight<NT, MODE>(Ws, row, col0 + b, occ * nh);
          };
          const double* theta0 = theta_base + draw0 * args.theta_ds;
          if (it.sat)
            occupation_item_series<true>(args.plan, args.model, theta0, args.theta_ds,
                                         args.theta_ps, n_b, it.g_begin, it.g_end, tab,
                                         ctrl->ser_queue[tid >> 5], store);
          else
            occupation_item_series<false>(args.plan, args.model, theta0, args.theta_ds,
                                          args.theta_ps, n_b, it.g_begin, it.g_end, tab,
                                          ctrl->ser_queue[tid >> 5], store);
        }
      } else {
        // precomputed occupations occ[draw, row]: a warp reads 32 consecutive rows of one draw
        // (coalesced; the padded order keeps the table's row order inside a galaxy type), the 8
        // draws of the n-tile one after the other with all their loads in flight
        const int n_q = args.n_ranges_cen + args.n_ranges_sat;
        const int r_begin = (int)((long long)lay.n_pad * q / n_q);
        const int r_end = (int)((long long)lay.n_pad * (q + 1) / n_q);
        const long long draw0 = (tile_first + j) * BM + 8 * nt;
        for (int row0 = r_begin; row0 < r_end; row0 += 32) {
          const int row = row0 + lane;
          const int src = row < r_end ? lay.pad_to_row[row] : -1;
          const double nh = src >= 0 ? args.plan.row_nh[row] : 0.0;
          double w[8];
#pragma unroll
          for (int b = 0; b < 8; b++) {
            const long long draw = min(draw0 + b, args.n_draws - 1);   // tail tile: repeat the last
            w[b] = src >= 0 ? args.occ[draw * lay.n_rows + src] : 0.0;
          }
          if (src >= 0) {
#pragma unroll
            for (int b = 0; b < 8; b++) store_weight<NT, MODE>(Ws, row, 8 * nt + b, w[b] * nh);
          }
        }
      }
      flag_signal(&ctrl->full[buf], lane);
      continue;
    }

    if (list < 0) continue;                      // prologue lists hold occupation items only
    const int c_lo = list == 0 ? first_lo : 0;
    const int c_hi = list == n_local - 1 ? last_hi : lay.n_chunks;
    if (kind == 0 ? c_lo != 0 : (idx < c_lo || idx >= c_hi)) continue;
    const int buf = list % n_buf;
    if (full_seen < list) {
      flag_wait(&ctrl->full[buf], (list / n_buf + 1) * n_occ);
      full_seen = list;
    }
    const double* Ws = smem + buf * tile_doubles;
    const long long tile = tile_first + list;

    if (kind == 0) {
      // ---- number densities (by the CTA that owns the tile's first chunk) ---------------------
      for (int b = lane; b < BM; b += 32) {
        double nc = 0.0, ns = 0.0;
        for (int r = 0; r < lay.nc_pad; r++) nc += load_weight<NT, MODE>(Ws, r, b);
        for (int r = lay.nc_pad; r < lay.n_pad; r++) ns += load_weight<NT, MODE>(Ws, r, b);
        args.ngal_tile[(tile * 2 + 0) * BM + b] = nc;
        args.ngal_tile[(tile * 2 + 1) * BM + b] = ns;
      }
    } else {
      const Chunk ch = lay.chunks[idx];
      if constexpr (MODE == kModeAutoTf32)
        run_chunk_tf32<NT>(lay, ch, Ws, args.parts + (size_t)tile * lay.n_parts * BM, lane,
                           args.tf32_segment);
      else
        run_chunk<NT, MODE>(lay, ch, Ws, args.parts + (size_t)tile * lay.n_parts * BM, lane);
    }
  }
}

// ------------------------------------------------------------------------------------------
// finalize: sum the scratch rows of every output in fixed order and normalise by ngal
// ------------------------------------------------------------------------------------------
struct FinalizeArgs {
  LayoutDev lay;
  const double* parts;
  const double* ngal_tile;
  long long n_draws;
  int bm;
  int mode;
  int separate;
  int n_tables;
  double* ngal_out;
  long long ngal_stride;
  double* xi_out;
  long long xi_stride;
};

// Block (tile, y) handles the draws of one tile and the outputs [o_lo, o_hi) of output block y.
// The sums are staged in shared memory and written out row by row, so that a warp stores
// consecutive outputs of a draw: coalesced also when the destination is the result slab of
// another GPU (peer memory over NVLink, distributed.PeerSlab), where scattered 8-byte stores
// would cost a 32-byte packet each.
__global__ void __launch_bounds__(256) finalize_kernel(const FinalizeArgs args) {
  extern __shared__ double fin_tile[];
  const int bm = args.bm;
  const long long tile = blockIdx.x;
  const int n_out = args.lay.n_out;
  const int o_per_block = (n_out + gridDim.y - 1) / gridDim.y;
  const int o_lo = blockIdx.y * o_per_block, o_hi = min(o_lo + o_per_block, n_out);
  const int row_len = o_per_block | 1;                     // odd: conflict-free column access
  const int b = threadIdx.x % bm;
  const long long draw = tile * bm + b;
  const int n_live = (int)min((long long)bm, args.n_draws - tile * bm);
  if (threadIdx.x < (blockDim.x / bm) * bm && b < n_live) {
    const double nc = args.ngal_tile[(tile * 2 + 0) * bm + b];
    const double ns = args.ngal_tile[(tile * 2 + 1) * bm + b];
    const double ngal = nc + ns;
    const double norm = args.mode == TC_MODE_AUTO ? ngal * ngal : ngal;
    if (blockIdx.y == 0 && threadIdx.x < bm) {
      for (int t = 0; t < args.n_tables; t++) {
        if (args.separate) {
          args.ngal_out[draw * args.ngal_stride + 2 * t + 0] = nc;
          args.ngal_out[draw * args.ngal_stride + 2 * t + 1] = ns;
        } else {
          args.ngal_out[draw * args.ngal_stride + t] = ngal;
        }
      }
    }
    const double* parts = args.parts + (size_t)tile * args.lay.n_parts * bm + b;
    for (int o = o_lo + threadIdx.x / bm; o < o_hi; o += blockDim.x / bm) {
      double s = 0.0;
      for (int j = args.lay.out_ptr[o]; j < args.lay.out_ptr[o + 1]; j++)
        s += parts[(size_t)args.lay.out_parts[j] * bm];
      fin_tile[b * row_len + (o - o_lo)] = s / norm;
    }
  }
  __syncthreads();
  const int n_o = o_hi - o_lo;
  for (int idx = threadIdx.x; idx < n_live * n_o; idx += blockDim.x) {
    const int bb = idx / n_o, ol = idx - bb * n_o;
    args.xi_out[(tile * bm + bb) * args.xi_stride + o_lo + ol] = fin_tile[bb * row_len + ol];
  }
}

}  // namespace
