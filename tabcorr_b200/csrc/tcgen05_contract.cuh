// tabcorr_b200 -- the 3xTF32 contraction on Blackwell's 5th-generation tensor cores:
//   tcgen05.mma kind::tf32 issued by one thread, accumulators in tensor memory (TMEM), operands
//   staged in shared memory by the TMA engine (cp.async.bulk + mbarrier), epilogue via tcgen05.ld.
//
// Part of the single translation unit tabcorr_b200.cu (see its header comment and DESIGN.md 3.3b).
//
// What is computed (tabcorr/tabcorr.py:641-647): xi_r = w^T M_r w / (sum w)^2 with the symmetric
// M_r.  With w = h + l (h = w rounded to TF32, l the remainder) and M = Mh + Ml (TF32 high and low
// parts; exact for the float32 tables the reference writes, tabcorr.py:418-419,448)
//     w^T M w = sum_i (h_i + 2 l_i) (M h)_i + l^T M l,
// so ONE operand plane per draw tile (h) and two per table tile (Mh, Ml) reach 22 significant bits:
// the GEMM  D[b, (r, i)] = sum_k h[b, k] (Mh + Ml)_r[i, k]  runs on the tensor cores with the 128
// draws of a tile as the MMA M dimension -- one TMEM lane per draw, so the row-dot with
// c_i = h_i + 2 l_i = 2 w_i - h_i is local to the epilogue thread that owns the lane -- and the
// dropped l^T M l is 2^-22 relative.  The occupation arithmetic stays FP64 (weights_image_kernel).
//
// Data flow per CTA (persistent, one per SM, 192 threads):
//   warps 0-3          thread m owns TMEM lane m = draw m of the tile.  Per tile they write the
//                      draw tile h (128 x Kp TF32) into TMEM columns [0, Kp) with tcgen05.st -- the
//                      MMA's A operand lives in tensor memory, so shared memory holds nothing but
//                      the table stream -- and run the epilogue: tcgen05.ld of an accumulator,
//                      FP32 products with c_i, FP64 sums, one scratch row per (column block,
//                      radial bin) for finalize_kernel
//   warp 4 (one lane)  TMA producer: the table stream in 64 KB bulk copies (two K segments of one
//                      plane, Ml or Mh, for 2 radial bins x 64 table rows) through a ring of three
//                      stages (cp.async.bulk + mbarrier).  A bulk copy costs ~470 cycles whatever
//                      its size (tools/tcgen_micro.cu): 64 KB copies stream at 110 B/cycle/SM.
//   warp 5 (one lane)  MMA issuer: tcgen05.mma kind::tf32, M = 128 draws, N = 128, K = 8, A from
//                      TMEM, B from shared memory, into one of two 128-column accumulators.  The
//                      issue loops are fully unrolled with constant operand offsets: descriptor
//                      arithmetic on the uniform datapath costs ~100 cycles per MMA otherwise,
//                      more than the 64 cycles the tensor core needs (tools/tcgen_micro.cu).
// Accumulation chains: the tensor core truncates its FP32 accumulator (and the aligned products)
// towards zero, a bias of ~3e-8 of the running sum per MMA of a same-sign chain (measured).  So the
// low-plane products of all K, 2^-11 of the result, form one chain of their own, and the high
// plane is cut into chains of 8 MMAs (one K segment of 64) that the epilogue adds in FP64.
// The table images are pre-formatted in the canonical K-major no-swizzle UMMA layout (8-row x
// 16-byte core matrices), so every copy is one contiguous bulk transfer.
#pragma once

#include "common.cuh"
#include "occupation.cuh"

namespace {

constexpr int kTcM = 128;                 // draws per tile = UMMA M = TMEM lanes
constexpr int kTcNI = 64;                 // table rows i per column block
constexpr int kTcRB = 2;                  // radial bins per MMA
constexpr int kTcN = kTcNI * kTcRB;       // UMMA N
constexpr int kTcKS = 64;                 // k per segment (8 MMAs of K = 8)
constexpr int kTcPlaneBytes = kTcN * kTcKS * 4;        // 32 KB: one plane (lo or hi) of one segment
constexpr int kTcCopyBytes = 2 * kTcPlaneBytes;        // a bulk copy carries two segments of a plane
constexpr int kTcStages = 3;
constexpr int kTcSbo = (kTcKS / 4) * 128;              // bytes between 8-row groups of a plane
constexpr int kTcChain = 4;               // MMAs per high-plane accumulation chain (K = 32)
constexpr int kTcThreads = 320;           // warps 0-3, 6-9: epilogue; 4: TMA producer; 5: MMA issuer
constexpr int kTcEpiWarps = 8;            // 2 groups (radial bin of the pair) x 4 lane quarters
constexpr int kTcMaxKp = 256;             // the draw tile occupies Kp TMEM columns
constexpr int kTcTmemCols = 512;
constexpr int kTcAccCol = 256;            // accumulator ring: 2 x 128 columns behind the draw tile
constexpr int kTcAccBufs = 2;
constexpr int kTcBarriers = 2 * kTcStages + 1 + 2 * kTcAccBufs;
// cycle counters of the roles (TC_TUNE_TCGEN_DEBUG=1 prints them); compile with -DTC_TCGEN_DEBUG
#ifdef TC_TCGEN_DEBUG
constexpr bool kTcDebug = true;
#else
constexpr bool kTcDebug = false;
#endif

struct TcgenDev {
  int kp;        // table rows padded to a multiple of 64 (K extent of the images)
  int n_seg;     // kp / 64: K segments
  int n_pairs;   // segment pairs = bulk copies per plane
  int n_ib;      // column blocks of 64 table rows
  int n_reff;    // radial bins x tables
  int n_rp;      // radial-bin pairs
  int n_parts;   // scratch rows per tile: n_ib * 2 n_rp
  const uint8_t* b_img;
  const int* out_ptr;
  const int* out_parts;
};

struct TcgenArgs {
  TcgenDev tc;
  const float* h_img;        // [n_tiles][kp / 32][128 draws][32]  TF32-rounded weights
  const float* c_img;        // [n_tiles][n_pad][128]              2 w - h
  const double* ngal_parts;  // [n_ranges][ngal_ld] number densities per occupation range
  long long ngal_ld;
  int n_ranges_cen, n_ranges_sat;
  int n_pad, n_rows;
  long long n_tiles;
  double* parts;             // [n_tiles][n_parts][128]
  double* ngal_tile;         // [n_tiles][2][128]
  int* error_flag;
  long long* debug;          // optional [grid][8] cycle counters (TC_TUNE_TCGEN_DEBUG), else nullptr
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline must not hang the device.  Returns false after ~2 s.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) return false;
  }
  return true;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 bytes;
// lbo = bytes between the two core matrices an MMA reads along K, sbo = bytes between 8-row groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
         ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same with the A operand in tensor memory (lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-uniform issue: the whole warp executes the instruction stream and one elected lane issues.
// Inside a lane-0-only branch the compiler moves every operand to a uniform register per MMA
// (4 R2UR + an ELECT loop, ~100 cycles); with uniform control flow the operands stay in uniform
// registers and MMAs issue back to back.
__device__ __forceinline__ void umma_tf32_ts_elect(uint32_t d_tmem, uint32_t a_tmem,
                                                   uint64_t b_desc, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p, q;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n.reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n"
      ::"r"(bar)
      : "memory");
}
// Bounded wait whose failure is sticky and never changes the control flow of the caller.
__device__ __forceinline__ void mbar_wait_sticky(uint32_t bar, uint32_t parity, bool& failed) {
  if (failed) return;
  if (!mbar_wait(bar, parity)) failed = true;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float4 (&v)[8]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
        "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}

// 64 consecutive columns of the thread's lane: two loads in flight, one wait
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r[64];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t* q = r + 32 * h;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]),
          "=r"(q[7]), "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]),
          "=r"(q[14]), "=r"(q[15]), "=r"(q[16]), "=r"(q[17]), "=r"(q[18]), "=r"(q[19]),
          "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]), "=r"(q[25]),
          "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
        : "r"(taddr + 32 * h)
        : "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 64; j++) v[j] = __uint_as_float(r[j]);
}

// byte offset of element (row, k) in a canonical K-major no-swizzle image whose 8-row groups are
// `sbo` bytes apart (k counted from the start of the image's K extent)
__host__ __device__ __forceinline__ size_t canon_offset(int row, int k, size_t sbo) {
  return (size_t)(row >> 3) * sbo + (size_t)(k >> 2) * 128 + (size_t)(row & 7) * 16 +
         (size_t)(k & 3) * 4;
}

constexpr uint32_t kTcIdesc = (1u << 4)                       // D format F32
                              | (2u << 7) | (2u << 10)        // A, B format TF32; both K-major
                              | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

// ------------------------------------------------------------------------------------------
// the contraction kernel
// ------------------------------------------------------------------------------------------
// One accumulation chain: NKS MMAs over consecutive k-steps of a plane in shared memory.
// b_lo is the low word of the descriptor of the chain's first k-step (address and leading offset),
// b_hi the high word; a k-step advances the address field by 256 bytes = 16 units and the TMEM
// operand by 8 columns.  Executed by the whole MMA warp (see umma_tf32_ts_elect).
template <int NKS>
__device__ __forceinline__ void issue_chain(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo,
                                            uint32_t b_hi, uint32_t accumulate_first) {
#pragma unroll
  for (int ks = 0; ks < NKS; ks++)
    umma_tf32_ts_elect(d_tmem, a_tmem + ks * 8, ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + ks * 16),
                       kTcIdesc, ks == 0 ? accumulate_first : 1u);
}
__device__ __forceinline__ void issue_chain_n(int n_ks, uint32_t d_tmem, uint32_t a_tmem,
                                              uint32_t b_lo, uint32_t b_hi,
                                              uint32_t accumulate_first) {
  if (n_ks == 8) {
    issue_chain<8>(d_tmem, a_tmem, b_lo, b_hi, accumulate_first);
  } else if (n_ks == 4) {
    issue_chain<4>(d_tmem, a_tmem, b_lo, b_hi, accumulate_first);
  } else {
    for (int ks = 0; ks < n_ks; ks++)
      umma_tf32_ts_elect(d_tmem, a_tmem + ks * 8,
                         ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + ks * 16), kTcIdesc,
                         ks == 0 ? accumulate_first : 1u);
  }
}

__global__ void __launch_bounds__(kTcThreads, 1) tcgen_contract_kernel(const TcgenArgs args) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  const TcgenDev& tc = args.tc;
  uint8_t* b_s = tc_smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + kTcStages * kTcCopyBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTcBarriers);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kTcStages);
  const uint32_t bar_a_full = smem_u32(bars + 2 * kTcStages);
  const uint32_t bar_acc_full = smem_u32(bars + 2 * kTcStages + 1);
  const uint32_t bar_acc_empty = smem_u32(bars + 2 * kTcStages + 1 + kTcAccBufs);
  // the shuffle tells the compiler that the warp index is warp-uniform (role branches stay uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; s++) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_a_full, 4);                  // warps 0-3 store the draw tile
    for (int b = 0; b < kTcAccBufs; b++) {
      mbar_init(bar_acc_full + 8 * b, 1);
      mbar_init(bar_acc_empty + 8 * b, kTcEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(kTcTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns all 512 columns of the SM's tensor memory, so the allocation starts at lane 0,
  // column 0; the MMA warp relies on it (its operands stay compile-time offsets).
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0 && threadIdx.x == 0) atomicExch(args.error_flag, 4);

  // Work items are (tile, column block) pairs in tile-major order; every CTA takes an equal
  // contiguous range, so at most two of its tiles are shared with a neighbour (both load the tile).
  const long long n_items = args.n_tiles * tc.n_ib;
  const long long item_lo = n_items * blockIdx.x / gridDim.x;
  const long long item_hi = n_items * (blockIdx.x + 1) / gridDim.x;
  const size_t rp_bytes = (size_t)2 * tc.n_pairs * kTcCopyBytes;   // table stream of one (ib, rp)
  const int n_ksteps = (args.n_pad + 7) / 8;                       // k-steps that hold table rows
  const int n_hi_units = (n_ksteps + kTcChain - 1) / kTcChain;

  if (warp == 4) {
    // ===== TMA producer: the table stream ========================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (long long item = item_lo; item < item_hi && ok; item++) {
        const int ib = (int)(item % tc.n_ib);
        for (int rp = 0; rp < tc.n_rp && ok; rp++) {
          const uint8_t* src = tc.b_img + ((size_t)ib * tc.n_rp + rp) * rp_bytes;
          for (int cp = 0; cp < 2 * tc.n_pairs; cp++) {       // low plane pairs, then high plane pairs
            const int pair = cp % tc.n_pairs;
            const uint32_t bytes = (uint32_t)min(2, tc.n_seg - 2 * pair) * kTcPlaneBytes;
            ok = mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (!ok) break;
            mbar_expect_tx(bar_full + 8 * stage, bytes);
            bulk_g2s(smem_u32(b_s + stage * kTcCopyBytes), src + (size_t)cp * kTcCopyBytes, bytes,
                     bar_full + 8 * stage);
            if (++stage == kTcStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (!ok) atomicExch(args.error_flag, 1);
    }
  } else if (warp == 5) {
    // ===== MMA issuer: warp-uniform control flow, one elected lane issues ==========================
    // The instruction stream of this warp is the critical path (uniform-datapath instructions
    // issue every ~6 cycles): a chain of 4 MMAs is 256 cycles of tensor-core work, so the loops
    // below are unrolled to constant operand offsets and carry no index arithmetic.
    int stage = 0;
    uint32_t phase = 0;
    uint32_t buf = 0, emp_par0 = 1, emp_par1 = 1;   // accumulator ring: next buffer, wait parities
    long long tile_prev = -1;
    int t_local = 0;
    bool failed = false;
    long long t_wait_a = 0, t_wait_acc = 0, t_wait_full = 0;
    const long long t_begin = clock64();
    const bool dbg = kTcDebug && args.debug != nullptr;
    const uint32_t b_hi_word = (uint32_t)(kTcSbo >> 4) | (1u << 14);   // stride offset, version 1
    auto acquire_acc = [&]() -> uint32_t {          // returns the accumulator's first TMEM column
      const uint32_t b = buf;
      const long long t0 = dbg ? clock64() : 0;
      mbar_wait_sticky(bar_acc_empty + 8 * b, b ? emp_par1 : emp_par0, failed);
      if (dbg) t_wait_acc += clock64() - t0;
      tc_fence_after();
      if (b) emp_par1 ^= 1; else emp_par0 ^= 1;
      buf ^= 1;
      return kTcAccCol + b * kTcN;
    };
    for (long long item = item_lo; item < item_hi; item++) {
      const long long tile = item / tc.n_ib;
      if (tile != tile_prev) {                 // wait for the draw tile in tensor memory
        const long long t0 = dbg ? clock64() : 0;
        mbar_wait_sticky(bar_a_full, t_local & 1, failed);
        if (dbg) t_wait_a += clock64() - t0;
        tc_fence_after();
        tile_prev = tile;
        t_local++;
      }
      for (int rp = 0; rp < tc.n_rp; rp++) {
        // ---- low plane: one chain over all K ----------------------------------------------------
        {
          const uint32_t d_tmem = acquire_acc();
          const uint32_t acc_bar = bar_acc_full + 8 * ((d_tmem - kTcAccCol) / kTcN);
          for (int pair = 0; pair < tc.n_pairs; pair++) {
            const long long t0 = dbg ? clock64() : 0;
            mbar_wait_sticky(bar_full + 8 * stage, phase, failed);
            if (dbg) t_wait_full += clock64() - t0;
            tc_fence_after();
            const uint32_t b_lo = ((smem_u32(b_s + stage * kTcCopyBytes) & 0x3FFFFu) >> 4) | (8u << 16);
            const int ks_left = n_ksteps - pair * 16;          // k-steps of this pair with table rows
            if (ks_left >= 16) {
              issue_chain<8>(d_tmem, (uint32_t)(pair * 128), b_lo, b_hi_word, pair == 0 ? 0u : 1u);
              issue_chain<8>(d_tmem, (uint32_t)(pair * 128 + 64), b_lo + (kTcPlaneBytes >> 4),
                             b_hi_word, 1u);
            } else {
              for (int ks = 0; ks < ks_left; ks++)
                umma_tf32_ts_elect(d_tmem, (uint32_t)(pair * 128 + ks * 8),
                                   ((uint64_t)b_hi_word << 32) |
                                       (uint64_t)(b_lo + (ks >> 3) * (kTcPlaneBytes >> 4) + (ks & 7) * 16),
                                   kTcIdesc, (pair == 0 && ks == 0) ? 0u : 1u);
            }
            umma_commit_elect(bar_empty + 8 * stage);
            if (++stage == kTcStages) { stage = 0; phase ^= 1; }
          }
          umma_commit_elect(acc_bar);
        }
        // ---- high plane: chains of kTcChain k-steps, each into its own accumulator ----------------
        for (int pair = 0; pair < tc.n_pairs; pair++) {
          const long long t0 = dbg ? clock64() : 0;
          mbar_wait_sticky(bar_full + 8 * stage, phase, failed);
          if (dbg) t_wait_full += clock64() - t0;
          tc_fence_after();
          const uint32_t b_lo = ((smem_u32(b_s + stage * kTcCopyBytes) & 0x3FFFFu) >> 4) | (8u << 16);
          const int ks_left = n_ksteps - pair * 16;
#pragma unroll
          for (int q = 0; q < 16 / kTcChain; q++) {           // the chains of the pair of planes
            const int n = ks_left - q * kTcChain;             // k-steps left for this chain
            if (n > 0) {
              const uint32_t d_tmem = acquire_acc();
              const uint32_t acc_bar = bar_acc_full + 8 * ((d_tmem - kTcAccCol) / kTcN);
              const uint32_t a_tmem = (uint32_t)(pair * 128 + q * kTcChain * 8);
              const uint32_t b_q = b_lo + (uint32_t)((q * kTcChain) >> 3) * (kTcPlaneBytes >> 4) +
                                   (uint32_t)((q * kTcChain) & 7) * 16;
              if (n >= kTcChain) {
                issue_chain<kTcChain>(d_tmem, a_tmem, b_q, b_hi_word, 0u);
              } else {
                for (int ks = 0; ks < n; ks++)
                  umma_tf32_ts_elect(d_tmem, a_tmem + ks * 8,
                                     ((uint64_t)b_hi_word << 32) | (uint64_t)(b_q + ks * 16), kTcIdesc,
                                     ks == 0 ? 0u : 1u);
              }
              umma_commit_elect(acc_bar);
            }
          }
          umma_commit_elect(bar_empty + 8 * stage);
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (failed && lane == 0) atomicExch(args.error_flag, 2);
    if (dbg && lane == 0) {
      long long* d = args.debug + (size_t)blockIdx.x * 8;
      d[0] = clock64() - t_begin; d[1] = t_wait_a; d[2] = t_wait_acc; d[3] = t_wait_full;
    }
  } else {
    // ===== epilogue warps.  Thread m owns TMEM lane m = draw m of the tile.  Eight warps in two
    // groups: group g takes radial bin g of the pair (64 accumulator columns).  What a short chain
    // waits for is the hand-off of the accumulator (commit -> load -> arrive -> next chain), not the
    // arithmetic (TC_TUNE_TCGEN_DEBUG counters): the accumulator is handed back right after the
    // TMEM load.  Tried and measured slower: 16 warps of 32 columns (arrivals on one mbarrier
    // serialise), one polling lane per warp plus a named barrier and a single arrival. ==============
    const int quarter = warp & 3;                       // TMEM lanes a warp may access: 32 (warp % 4)
    const int rr = warp < 4 ? 0 : 1;                    // warps 0-3, 6-9
    const int m = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t unit_count = 0;
    long long tile_prev = -1;
    bool ok = true;
    const bool dbg = kTcDebug && args.debug != nullptr;
    long long t_store = 0, t_cload = 0, t_wait = 0;
    const long long t_begin = clock64();
    const int n_ranges = args.n_ranges_cen + args.n_ranges_sat;
    const int n_chunks = tc.kp / 32;
    for (long long item = item_lo; item < item_hi && ok; item++) {
      const long long tile = item / tc.n_ib;
      const int ib = (int)(item - tile * tc.n_ib);
      const long long t_s0 = dbg ? clock64() : 0;
      if (tile != tile_prev && rr == 0) {
        // Every MMA that reads the previous tile has completed: this thread has waited for the
        // accumulator of the last chain, whose commit follows all earlier MMAs.
        const float4* src = reinterpret_cast<const float4*>(
            args.h_img + ((size_t)tile * n_chunks * kTcM + m) * 32);
        for (int ch = 0; ch < n_chunks; ch++) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; j++) v[j] = __ldg(src + (size_t)ch * kTcM * 8 + j);
          tmem_st32(lane_base + (uint32_t)(ch * 32), v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a_full);
        {
          // number densities of the tile (both CTAs that share a tile write the same values)
          double nc = 0.0, ns = 0.0;
          for (int q = 0; q < n_ranges; q++) {
            const double v = args.ngal_parts[(size_t)q * args.ngal_ld + tile * kTcM + m];
            if (q < args.n_ranges_cen) nc += v; else ns += v;
          }
          args.ngal_tile[(tile * 2 + 0) * kTcM + m] = nc;
          args.ngal_tile[(tile * 2 + 1) * kTcM + m] = ns;
        }
      }
      tile_prev = tile;
      const long long t_s1 = dbg ? clock64() : 0;
      double* parts = args.parts + ((size_t)tile * tc.n_parts + (size_t)ib * 2 * tc.n_rp) * kTcM + m;
      float c[kTcNI];
#pragma unroll
      for (int j = 0; j < kTcNI; j++) {
        const int i = ib * kTcNI + j;
        c[j] = i < args.n_rows ? args.c_img[((size_t)tile * args.n_pad + i) * kTcM + m] : 0.0f;
      }
      if (dbg) {
        float keep = 0.f;
#pragma unroll
        for (int j = 0; j < kTcNI; j++) keep += c[j];
        if (keep == 1.2345e-30f) t_wait++;       // forces the loads to complete before the clock
        t_store += t_s1 - t_s0;
        t_cload += clock64() - t_s1;
      }
      for (int rp = 0; rp < tc.n_rp && ok; rp++) {
        double sum = 0.0;
        for (int u = 0; u <= n_hi_units; u++, unit_count++) {   // the low chain, then the high chains
          const uint32_t buf = unit_count & 1, use = unit_count >> 1;
          const long long t_w0 = dbg ? clock64() : 0;
          ok = mbar_wait(bar_acc_full + 8 * buf, use & 1);
          ok = __all_sync(0xffffffffu, ok);
          if (dbg) t_wait += clock64() - t_w0;
          if (!ok) break;
          tc_fence_after();
          float v[kTcNI];
          tmem_ld64(lane_base + (uint32_t)(kTcAccCol + buf * kTcN + rr * kTcNI), v);
          // hand the accumulator back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
#pragma unroll
          for (int j = 0; j < kTcNI; j += 8) {
            float p = v[j] * c[j];
#pragma unroll
            for (int e = 1; e < 8; e++) p = fmaf(v[j + e], c[j + e], p);
            sum += (double)p;
          }
        }
        if (!ok) break;
        parts[(size_t)(2 * rp + rr) * kTcM] = sum;
      }
    }
    if (!ok && lane == 0) atomicExch(args.error_flag, 3);
    if (dbg && threadIdx.x == 0) {
      long long* d = args.debug + (size_t)blockIdx.x * 8;
      d[4] = clock64() - t_begin; d[5] = t_store; d[6] = t_cload; d[7] = t_wait;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTcTmemCols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// occupation -> operand images of the draw tiles (FP64 arithmetic, same code as occupation_kernel)
// ------------------------------------------------------------------------------------------
struct WeightsImageArgs {
  OccPlan plan;
  tc_model model;
  const double* theta;
  long long theta_ds, theta_ps;
  long long n_draws;
  int n_ranges_cen, n_ranges_sat;
  int pieces_cen, pieces_sat;       // draw pieces per type (series_item)
  int max_groups;                   // groups of the larger galaxy type
  int kp, n_pad, n_rows;
  float* h_img;
  float* c_img;
  double* ngal_parts;
  long long ngal_ld;
};

__global__ void __launch_bounds__(kThreads, 1) weights_image_kernel(const WeightsImageArgs args) {
  __shared__ double tab[kTabDoubles];
  __shared__ int queue[kWarps][kSerQueue];
  // per warp: the weight sum of every (draw, group) pair of the item, [8][max_groups] -- summed
  // afterwards in group order, so that a draw's number density does not depend on which lane
  // happened to evaluate which pair (bitwise independence of the batch composition)
  extern __shared__ double pair_sums_all[];
  load_math_tables(tab);
  for (int i = threadIdx.x; i < kWarps * 8 * args.max_groups; i += kThreads) pair_sums_all[i] = 0.0;
  __syncthreads();
  // one warp per item = a piece of an 8-draw block x a range of groups of one galaxy type
  // (series items, occupation.cuh); h image: [tile][row / 32][draw][row % 32], so that the thread
  // owning a TMEM lane reads 128 contiguous bytes per 32-column store
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* pair_sums = pair_sums_all + (size_t)warp * 8 * args.max_groups;
  const int n_ranges = args.n_ranges_cen + args.n_ranges_sat;
  const long long n_blocks = (args.n_draws + 7) / 8;
  const long long n_items = n_blocks * n_ranges;
  const long long warp0 = (long long)blockIdx.x * kWarps + warp;
  const int n_chunks = args.kp / 32;
  for (long long item = warp0; item < n_items; item += (long long)gridDim.x * kWarps) {
    const long long block = item / n_ranges;
    const int q = (int)(item - block * n_ranges);
    const SeriesItem it = series_item(args.plan, args.n_ranges_cen, args.n_ranges_sat,
                                      args.pieces_cen, args.pieces_sat, q);
    const long long draw0 = block * 8 + it.b_begin;
    const int n_b = (int)max(0LL, min((long long)it.n_b, args.n_draws - draw0));
    if (q == 0) {   // K padding of the block's image rows (the table stream is zero there too)
      const long long draw = block * 8 + (lane & 7);
      if (draw < args.n_draws) {
        const long long tile = draw / kTcM;
        float* h_tile = args.h_img + ((size_t)tile * n_chunks * kTcM + (draw - tile * kTcM)) * 32;
        for (int row = args.n_rows + (lane >> 3); row < args.kp; row += 4)
          h_tile[(size_t)(row >> 5) * kTcM * 32 + (row & 31)] = 0.0f;
      }
    }
    if (n_b > 0 && it.g_end > it.g_begin) {
      auto store = [&](int b, int grp, int row, double occ, double nh) {
        const double w = occ * nh;
        const float h = to_tf32((float)w);
        const long long draw = draw0 + b;
        const long long tile = draw / kTcM;
        const int m = (int)(draw - tile * kTcM);
        args.h_img[((size_t)tile * n_chunks * kTcM + m) * 32 + (size_t)(row >> 5) * kTcM * 32 +
                   (row & 31)] = h;
        args.c_img[(size_t)tile * args.n_pad * kTcM + (size_t)row * kTcM + m] =
            (float)(2.0 * w - (double)h);
        pair_sums[b * args.max_groups + grp - it.g_begin] += w;   // rows of a pair: one lane
      };
      const double* theta0 = args.theta + draw0 * args.theta_ds;
      if (it.sat)
        occupation_item_series<true>(args.plan, args.model, theta0, args.theta_ds, args.theta_ps,
                                     n_b, it.g_begin, it.g_end, tab, queue[warp], store);
      else
        occupation_item_series<false>(args.plan, args.model, theta0, args.theta_ds, args.theta_ps,
                                      n_b, it.g_begin, it.g_end, tab, queue[warp], store);
    }
    __syncwarp();
    // number density of the item per draw: the lanes' sums in a fixed order; draws of the block
    // outside the item's piece get an explicit zero
    const int n_grp = it.g_end - it.g_begin;
    for (int b = 0; b < it.n_b; b++) {
      double v = 0.0;
      for (int g = lane; g < n_grp; g += 32) {   // ascending groups per lane, then a fixed tree
        v += pair_sums[b * args.max_groups + g];
        pair_sums[b * args.max_groups + g] = 0.0;
      }
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      const long long draw = block * 8 + it.b_begin + b;
      if (lane == 0 && draw < args.n_draws) args.ngal_parts[(size_t)q * args.ngal_ld + draw] = v;
    }
    if (lane < 8 && (lane < it.b_begin || lane >= it.b_begin + it.n_b) &&
        block * 8 + lane < args.n_draws)
      args.ngal_parts[(size_t)q * args.ngal_ld + block * 8 + lane] = 0.0;
    __syncwarp();
  }
}

// finalize for the tcgen05 path: like finalize_kernel, plus poisoning when the pipeline timed out
__global__ void tcgen_poison_kernel(const int* error_flag, double* xi, long long n_draws,
                                    long long xi_stride, int n_out) {
  if (*error_flag == 0) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_draws * n_out;
       i += (long long)gridDim.x * blockDim.x)
    xi[(i / n_out) * xi_stride + i % n_out] = CUDART_NAN;
}

}  // namespace
