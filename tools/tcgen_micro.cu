// Microbenchmarks behind csrc/tcgen05_contract.cuh (round 2): cycles per tcgen05.mma kind::tf32 for
// A in shared memory (SS) or tensor memory (TS) and N = 64/128/256, and per-SM throughput of
// cp.async.bulk from L2 with a ring of stages.  One CTA per SM; operands are whatever the memory
// holds (zeros).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tcgen_micro tools/tcgen_micro.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if (clock64() - t0 > 2000000000LL) return false;
  }
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

// mode 0: SS, mode 1: TS.  One thread issues `n_mma` MMAs into `n_acc` rotating accumulators.
__global__ void __launch_bounds__(128, 1) mma_rate(int mode, int n, int n_mma, int n_acc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 131072);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i++) {
      const uint32_t d = tm + 256 + (uint32_t)((i % n_acc) * n) % 256;
      const uint64_t bd = umma_desc(b_addr + (i & 7) * 256, 128, 2048);
      if (mode == 0) {
        const uint64_t ad = umma_desc(a_addr + (i & 31) * 256, 128, 8192);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
      } else {
        const uint32_t at = tm + (uint32_t)((i & 31) * 8);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(1) : "memory");
      }
    }
    commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

// `n_issuers` warps (lane 0 each) issue n_mma / n_issuers MMAs with loop-invariant operands into
// their own accumulator: is the ~160 cycles per instruction a per-thread issue cost?
__global__ void __launch_bounds__(128, 1) mma_issuers(int mode, int n, int n_mma, int n_issuers, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ long long t_end[4];
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(smem_u32(&bar[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  const int w = threadIdx.x >> 5;
  const long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && w < n_issuers) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24);
    const uint32_t d = tm + 256 + (uint32_t)(w * n) % 256;
    const uint64_t bd = umma_desc(smem_u32(smem + 131072), 128, 2048);
    const uint64_t ad = umma_desc(smem_u32(smem), 128, 8192);
    const uint32_t at = tm;
    const int count = n_mma / n_issuers;
    if (mode == 0) {
#pragma unroll 8
      for (int i = 0; i < count; i++)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
    } else {
#pragma unroll 8
      for (int i = 0; i < count; i++)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(at), "l"(bd), "r"(idesc), "r"(1) : "memory");
    }
    commit(smem_u32(&bar[w]));
    mbar_wait(smem_u32(&bar[w]), 0);
    t_end[w] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long mx = 0;
    for (int i = 0; i < n_issuers; i++) mx = t_end[i] > mx ? t_end[i] : mx;
    cycles[blockIdx.x] = mx;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

// One warp issues (warp-uniform code, elect.sync) chains of `chain` MMAs with constant operands,
// alternating between two accumulators; after every chain a tcgen05.commit to one of 8 barriers
// (nobody waits for them except at the end): what do short chains and frequent commits cost?
__global__ void __launch_bounds__(128, 1) mma_chains(int chain, int n_chains, int do_commit, int fresh, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[9];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 196608 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { for (int i = 0; i < 9; i++) mbar_init(smem_u32(&bar[i]), 1 << 20); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (w == 1) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (16u << 17) | (8u << 24);
    const uint64_t bd = umma_desc(smem_u32(smem + 131072), 128, 2048);
    const long long t0 = clock64();
    for (int c = 0; c < n_chains; c++) {
      const uint32_t d = 256 + (c & 1) * 128;
      for (int i = 0; i < chain; i++)
        asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"((uint32_t)(i * 8)), "l"(bd), "r"(idesc), "r"((fresh && i == 0) ? 0 : 1) : "memory");
      if (do_commit)
        asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(&bar[c & 7])) : "memory");
    }
    if (threadIdx.x == 32) {
      mbar_init(smem_u32(&bar[8]), 1);
      commit(smem_u32(&bar[8]));
      mbar_wait(smem_u32(&bar[8]), 0);
      cycles[blockIdx.x] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

// Ring of `stages` buffers of `bytes`; one thread keeps them all in flight, `n_copy` copies in total,
// all CTAs read the same `span` bytes of `src` in the same order (like the table stream).
__global__ void __launch_bounds__(128, 1) bulk_rate(const uint8_t* src, size_t span, int stages, int bytes, int n_copy, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; s++) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    size_t off = 0;
    for (int i = 0; i < n_copy + stages; i++) {
      const int s = i % stages;
      if (i >= stages) mbar_wait(smem_u32(&bars[s]), ((i / stages) - 1) & 1);
      if (i < n_copy) {
        mbar_expect_tx(smem_u32(&bars[s]), bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * bytes)), "l"(src + off), "r"(bytes), "r"(smem_u32(&bars[s])) : "memory");
        off += bytes;
        if (off + bytes > span) off = 0;
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  int n_sm = 0;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  long long* d_cycles;
  cudaMalloc(&d_cycles, n_sm * sizeof(long long));
  std::vector<long long> h(n_sm);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(bulk_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n_mma = 20000;
  for (int mode = 0; mode < 2; mode++)
    for (int n : {64, 128, 256})
      for (int n_acc : {1, 2, 4}) {
        if (n * n_acc > 256) continue;
        mma_rate<<<n_sm, 128, 196608>>>(mode, n, n_mma, n_acc, d_cycles);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma_rate error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_cycles, n_sm * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto c : h) mx = c > mx ? c : mx;
        printf("{\"bench\": \"mma\", \"a\": \"%s\", \"n\": %d, \"accumulators\": %d, \"cycles_per_mma\": %.1f}\n", mode ? "tmem" : "smem", n, n_acc, (double)mx / n_mma);
      }
  cudaFuncSetAttribute(mma_issuers, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int mode = 0; mode < 2; mode++)
    for (int n : {64, 128, 256})
      for (int n_issuers : {1, 2, 4}) {
        if (n * n_issuers > 256 && n_issuers > 1 && n == 256) continue;
        mma_issuers<<<n_sm, 128, 196608>>>(mode, n, n_mma, n_issuers, d_cycles);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma_issuers error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_cycles, n_sm * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto c : h) mx = c > mx ? c : mx;
        printf("{\"bench\": \"mma_fixed_operands\", \"a\": \"%s\", \"n\": %d, \"issuers\": %d, \"cycles_per_mma\": %.1f}\n", mode ? "tmem" : "smem", n, n_issuers, (double)mx / n_mma);
      }
  cudaFuncSetAttribute(mma_chains, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int chain : {4, 8, 30})
    for (int do_commit : {0, 1})
      for (int fresh : {0, 1}) {
        const int n_chains = 24000 / chain;
        mma_chains<<<n_sm, 128, 196608>>>(chain, n_chains, do_commit, fresh, d_cycles);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma_chains error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_cycles, n_sm * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto c : h) mx = c > mx ? c : mx;
        printf("{\"bench\": \"mma_chains_ts_n128\", \"chain\": %d, \"commit_per_chain\": %d, \"first_overwrites\": %d, \"cycles_per_mma\": %.1f}\n", chain, do_commit, fresh, (double)mx / (n_chains * chain));
      }
  const size_t span = 10u << 20;
  uint8_t* d_src; cudaMalloc(&d_src, span); cudaMemset(d_src, 0, span);
  for (int bytes : {32768, 65536})
    for (int stages : {2, 3}) {
      if ((size_t)bytes * stages > 196608) continue;
      const int n_copy = (64 << 20) / bytes;
      for (int grid : {1, n_sm}) {
        bulk_rate<<<grid, 128, 196608>>>(d_src, span, stages, bytes, n_copy, d_cycles);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("bulk_rate error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
        printf("{\"bench\": \"bulk\", \"bytes\": %d, \"stages\": %d, \"ctas\": %d, \"bytes_per_cycle_per_sm\": %.1f}\n", bytes, stages, grid, (double)bytes * n_copy / mx);
      }
    }
  return 0;
}
