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
//   warp 4 (one lane)  TMA producer: the draw tile's h image (128 x Kp TF32, 128 KB at N = 240) and
//                      the table stream, 32 KB stages {Ml, Mh} of 128 table rows x 32 k
//   warp 5 (one lane)  MMA issuer: per accumulator block (64 table rows i x 2 radial bins) and
//                      k-block 8 tcgen05.mma (M = 128 draws, N = 128, K = 8), accumulators in TMEM,
//                      optionally one accumulator per K segment (FP32 accumulation error)
//   warps 0-3          epilogue: tcgen05.ld of the lane's 128 columns, FP32 products with c_i,
//                      FP64 sums, one scratch row per (column block, radial bin) for finalize_kernel
// All operand images are pre-formatted in the canonical K-major no-swizzle UMMA layout (8-row x
// 16-byte core matrices), so every copy is one contiguous bulk transfer.
#pragma once

#include "common.cuh"
#include "occupation.cuh"

namespace {

constexpr int kTcM = 128;                 // draws per tile = UMMA M = TMEM lanes
constexpr int kTcNI = 64;                 // table rows i per accumulator block
constexpr int kTcRB = 2;                  // radial bins per MMA
constexpr int kTcN = kTcNI * kTcRB;       // UMMA N
constexpr int kTcKB = 32;                 // k per table stage (4 MMAs of K = 8 per plane)
constexpr int kTcStages = 3;
constexpr int kTcPlaneBytes = kTcN * kTcKB * 4;        // 16 KB
constexpr int kTcStageBytes = 2 * kTcPlaneBytes;       // {lo, hi}
constexpr int kTcThreads = 192;
constexpr int kTcMaxKp = 256;             // A image of 128 draws x Kp x 4 B must leave 3 stages
constexpr int kTcTmemCols = 512;
constexpr int kTcBarriers = 16;

struct TcgenDev {
  int kp;        // table rows padded to a multiple of 32 (K extent of the images)
  int n_kb;      // kp / 32
  int n_ib;      // column blocks of 64 table rows
  int n_rp;      // radial-bin pairs
  int n_parts;   // scratch rows per tile: n_ib * 2 n_rp
  const uint8_t* b_img;
  const int* out_ptr;
  const int* out_parts;
};

struct TcgenArgs {
  TcgenDev tc;
  const uint8_t* a_img;      // [n_tiles][128 x kp TF32, canonical layout]
  const float* c_img;        // [n_tiles][n_pad][128]
  const double* ngal_parts;  // [n_ranges][ngal_ld] number densities per occupation range
  long long ngal_ld;
  int n_ranges_cen, n_ranges_sat;
  int n_pad, n_rows;
  int seg;                   // K segments with their own TMEM accumulator (1, 2 or 4)
  long long n_tiles;
  double* parts;             // [n_tiles][n_parts][128]
  double* ngal_tile;         // [n_tiles][2][128]
  int* error_flag;
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
__global__ void __launch_bounds__(kTcThreads, 1) tcgen_contract_kernel(const TcgenArgs args) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  const TcgenDev& tc = args.tc;
  const uint32_t a_bytes = (uint32_t)kTcM * tc.kp * 4;
  uint8_t* a_s = tc_smem;
  uint8_t* b_s = tc_smem + a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + kTcStages * kTcStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kTcBarriers);
  // barriers: full[3], empty[3], a_full, a_empty, acc_full[4], acc_empty[4]
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 3);
  const uint32_t bar_a_full = smem_u32(bars + 6), bar_a_empty = smem_u32(bars + 7);
  const uint32_t bar_acc_full = smem_u32(bars + 8), bar_acc_empty = smem_u32(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; s++) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_empty, 1);
    for (int b = 0; b < 4; b++) {
      mbar_init(bar_acc_full + 8 * b, 1);
      mbar_init(bar_acc_empty + 8 * b, 4);   // one arrival per epilogue warp
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
  const uint32_t tmem_base = *tmem_slot;

  const int seg = args.seg;
  const int acc_stride = kTcN * seg;               // TMEM columns per accumulator buffer
  const int n_acc = kTcTmemCols / acc_stride;      // 4, 2 or 1 buffers
  const int n_jobs = tc.n_ib * tc.n_rp;            // accumulator blocks per tile
  bool ok = true;

  if (warp == 4) {
    // ===== TMA producer ==========================================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int t_local = 0;
      for (long long tile = blockIdx.x; tile < args.n_tiles && ok; tile += gridDim.x, t_local++) {
        if (t_local > 0) ok = mbar_wait(bar_a_empty, (t_local - 1) & 1);
        if (!ok) break;
        mbar_expect_tx(bar_a_full, a_bytes);
        const uint8_t* src = args.a_img + (size_t)tile * a_bytes;
        for (uint32_t off = 0; off < a_bytes; off += 32768)
          bulk_g2s(smem_u32(a_s + off), src + off, min(32768u, a_bytes - off), bar_a_full);
        for (int job = 0; job < n_jobs && ok; job++) {
          for (int kb = 0; kb < tc.n_kb; kb++) {
            ok = mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (!ok) break;
            mbar_expect_tx(bar_full + 8 * stage, kTcStageBytes);
            bulk_g2s(smem_u32(b_s + stage * kTcStageBytes),
                     tc.b_img + ((size_t)job * tc.n_kb + kb) * kTcStageBytes, kTcStageBytes,
                     bar_full + 8 * stage);
            if (++stage == kTcStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (!ok) atomicExch(args.error_flag, 1);
    }
  } else if (warp == 5) {
    // ===== MMA issuer ============================================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int t_local = 0;
      long long job_count = 0;
      const uint32_t sbo_a = (uint32_t)tc.kp * 32;   // 8-row groups of the A image
      for (long long tile = blockIdx.x; tile < args.n_tiles && ok; tile += gridDim.x, t_local++) {
        ok = mbar_wait(bar_a_full, t_local & 1);
        if (!ok) break;
        tc_fence_after();
        for (int job = 0; job < n_jobs && ok; job++, job_count++) {
          const int buf = (int)(job_count % n_acc);
          const uint32_t use = (uint32_t)(job_count / n_acc);
          ok = mbar_wait(bar_acc_empty + 8 * buf, (use & 1) ^ 1);
          if (!ok) break;
          tc_fence_after();
          int seg_prev = -1;
          for (int kb = 0; kb < tc.n_kb; kb++) {
            ok = mbar_wait(bar_full + 8 * stage, phase);
            if (!ok) break;
            tc_fence_after();
            const int s = kb * seg / tc.n_kb;
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * acc_stride + s * kTcN);
            const uint32_t a_addr = smem_u32(a_s) + (uint32_t)kb * 8 * 128;
            const uint32_t b_addr = smem_u32(b_s + stage * kTcStageBytes);
#pragma unroll
            for (int plane = 0; plane < 2; plane++) {      // low parts first
#pragma unroll
              for (int ks = 0; ks < 4; ks++) {
                const uint64_t ad = umma_desc(a_addr + ks * 256, 128, sbo_a);
                const uint64_t bd = umma_desc(b_addr + plane * kTcPlaneBytes + ks * 256, 128, 1024);
                umma_tf32(d_tmem, ad, bd, kTcIdesc, (s != seg_prev && plane == 0 && ks == 0) ? 0u : 1u);
              }
            }
            seg_prev = s;
            umma_commit(bar_empty + 8 * stage);     // the stage is free once these MMAs have read it
            if (++stage == kTcStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(bar_acc_full + 8 * buf);      // accumulators complete
        }
        umma_commit(bar_a_empty);                   // the draw tile may be overwritten
      }
      if (!ok) atomicExch(args.error_flag, 2);
    }
  } else {
    // ===== epilogue warps: thread m owns TMEM lane m = draw m of the tile ==========================
    const int m = threadIdx.x;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    long long job_count = 0;
    const int n_ranges = args.n_ranges_cen + args.n_ranges_sat;
    for (long long tile = blockIdx.x; tile < args.n_tiles && ok; tile += gridDim.x) {
      {
        double nc = 0.0, ns = 0.0;
        for (int q = 0; q < n_ranges; q++) {
          const double v = args.ngal_parts[(size_t)q * args.ngal_ld + tile * kTcM + m];
          if (q < args.n_ranges_cen) nc += v; else ns += v;
        }
        args.ngal_tile[(tile * 2 + 0) * kTcM + m] = nc;
        args.ngal_tile[(tile * 2 + 1) * kTcM + m] = ns;
      }
      double* parts = args.parts + (size_t)tile * tc.n_parts * kTcM + m;
      for (int ib = 0; ib < tc.n_ib && ok; ib++) {
        float c[kTcNI];
#pragma unroll
        for (int j = 0; j < kTcNI; j++) {
          const int i = ib * kTcNI + j;
          c[j] = i < args.n_rows ? args.c_img[((size_t)tile * args.n_pad + i) * kTcM + m] : 0.0f;
        }
        for (int rp = 0; rp < tc.n_rp; rp++, job_count++) {
          const int buf = (int)(job_count % n_acc);
          const uint32_t use = (uint32_t)(job_count / n_acc);
          ok = mbar_wait(bar_acc_full + 8 * buf, use & 1);
          ok = __all_sync(0xffffffffu, ok);
          if (!ok) break;
          tc_fence_after();
#pragma unroll
          for (int rr = 0; rr < kTcRB; rr++) {
            double sum = 0.0;
#pragma unroll
            for (int ch = 0; ch < kTcNI / 32; ch++) {
              const uint32_t col = (uint32_t)(buf * acc_stride + rr * kTcNI + ch * 32);
              float v[32];
              tmem_ld32(tmem_base + lane_base + col, v);
              for (int s = 1; s < seg; s++) {
                float u[32];
                tmem_ld32(tmem_base + lane_base + col + s * kTcN, u);
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] += u[j];
              }
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float p = v[j] * c[ch * 32 + j];
                p = fmaf(v[j + 1], c[ch * 32 + j + 1], p);
                p = fmaf(v[j + 2], c[ch * 32 + j + 2], p);
                p = fmaf(v[j + 3], c[ch * 32 + j + 3], p);
                sum += (double)p;
              }
            }
            parts[(size_t)(ib * 2 * tc.n_rp + 2 * rp + rr) * kTcM] = sum;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
        }
      }
    }
    if (!ok && lane == 0) atomicExch(args.error_flag, 3);
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
  int kp, n_pad, n_rows;
  uint8_t* a_img;
  float* c_img;
  double* ngal_parts;
  long long ngal_ld;
};

__global__ void __launch_bounds__(kThreads, 1) weights_image_kernel(const WeightsImageArgs args) {
  __shared__ double tab[kTabDoubles];
  load_math_tables(tab);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int n_ranges = args.n_ranges_cen + args.n_ranges_sat;
  const long long n_blocks = (args.n_draws + 7) / 8;
  const long long n_items = n_blocks * n_ranges;
  const long long warp0 = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const size_t a_bytes = (size_t)kTcM * args.kp * 4, sbo_a = (size_t)args.kp * 32;
  for (long long item = warp0; item < n_items; item += (long long)gridDim.x * kWarps) {
    const long long block = item / n_ranges;
    const int q = (int)(item - block * n_ranges);
    const long long draw = block * 8 + (lane & 7);
    const bool live = draw < args.n_draws;
    const long long tile = draw / kTcM;
    const int m = (int)(draw - tile * kTcM);
    uint8_t* a_tile = args.a_img + (size_t)tile * a_bytes;
    float* c_tile = args.c_img + (size_t)tile * args.n_pad * kTcM + m;
    int g_begin, g_end;
    occupation_range(args.plan, args.n_ranges_cen, args.n_ranges_sat, q, g_begin, g_end);
    if (q == 0 && live) {   // K padding of the draw's image row (the table stream is zero there too)
      for (int row = args.n_rows + (lane >> 3); row < args.kp; row += 4)
        *reinterpret_cast<float*>(a_tile + canon_offset(m, row, sbo_a)) = 0.0f;
    }
    double total = 0.0;
    occupation_item(args.plan, args.model,
                    args.theta + (live ? draw : args.n_draws - 1) * args.theta_ds, args.theta_ps,
                    g_begin, g_end, tab,
                    [&](int row, double occ, double nh) {
                      const double w = occ * nh;
                      const float h = to_tf32((float)w);
                      total += w;
                      if (live) {
                        *reinterpret_cast<float*>(a_tile + canon_offset(m, row, sbo_a)) = h;
                        c_tile[(size_t)row * kTcM] = (float)(2.0 * w - (double)h);
                      }
                    },
                    threadIdx.x >> 3 & 3, 4);
    // number density of the range: the four lanes of a draw hold disjoint groups (fixed order)
    total += __shfl_xor_sync(0xffffffffu, total, 8);
    total += __shfl_xor_sync(0xffffffffu, total, 16);
    if (lane < 8 && live) args.ngal_parts[(size_t)q * args.ngal_ld + draw] = total;
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
