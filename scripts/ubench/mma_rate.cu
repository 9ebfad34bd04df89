// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 operands from shared memory, SWIZZLE_128B K-major) on sm_100a
// for the shapes / accumulator orders the Gram kernel can choose from.  One CTA per SM, operands are whatever is in shared
// memory (values do not matter for timing).  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
//   scripts/ubench/mma_rate.cu -o scripts/ubench/mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
  return (uint64_t)((a >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                   smem_u32(bar)),
               "r"(parity)
               : "memory");
}

// mode: 0 = N=128, tile-outer (4 k-steps on one accumulator, 4 accumulators per chunk)
//       1 = N=256, two accumulators per chunk (same MAC count as mode 0)
//       2 = N=128, one accumulator only (16 MMAs per chunk on the same tile)
//       3 = N=128, k-step outer (accumulator changes every MMA)
//       4 = N=256, one A for both (A slot shared), B slots distinct  [same as 1]
//       5 = N=64, 8 accumulators...  (not used)
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int mode, int n_chunks, int stages, int slots, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // fill shared memory with small finite bf16 values
  for (int i = threadIdx.x; i < stages * slots * 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem);
    const uint32_t stage_bytes = slots * 16384;
    constexpr uint32_t i128 = make_idesc(128, 128), i256 = make_idesc(128, 256);
    long long t0 = clock64();
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t st = base + (uint32_t)(c % stages) * stage_bytes;
      if (mode == 0) {
        for (int t = 0; t < 4; ++t)
          for (int ks = 0; ks < 4; ++ks)
            mma(tb + t * 128, make_desc(st + (t & 1) * 16384 + ks * 32), make_desc(st + ((t + 1) % slots) * 16384 + ks * 32), i128, 1);
      } else if (mode == 1) {
        for (int t = 0; t < 2; ++t)
          for (int ks = 0; ks < 4; ++ks)
            mma(tb + t * 256, make_desc(st + (t % slots) * 16384 + ks * 32), make_desc(st + ((slots >= 3) ? 16384 : 0) + ks * 32), i256, 1);
      } else if (mode == 2) {
        for (int t = 0; t < 4; ++t)
          for (int ks = 0; ks < 4; ++ks) mma(tb, make_desc(st + ks * 32), make_desc(st + ((t + 1) % slots) * 16384 + ks * 32), i128, 1);
      } else if (mode == 3) {
        for (int ks = 0; ks < 4; ++ks)
          for (int t = 0; t < 4; ++t)
            mma(tb + t * 128, make_desc(st + (t & 1) * 16384 + ks * 32), make_desc(st + ((t + 1) % slots) * 16384 + ks * 32), i128, 1);
      } else if (mode == 6) {  // N=256, a single accumulator
        for (int t = 0; t < 2; ++t)
          for (int ks = 0; ks < 4; ++ks) mma(tb, make_desc(st + ks * 32), make_desc(st + ((slots >= 3) ? 16384 : 0) + ks * 32), i256, 1);
      } else if (mode == 7) {  // N=128, A and B the same block (diagonal tile): one operand fetch?
        for (int t = 0; t < 4; ++t)
          for (int ks = 0; ks < 4; ++ks) mma(tb + t * 128, make_desc(st + ks * 32), make_desc(st + ks * 32), i128, 1);
      }
    }
    commit(&bar);
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}


// ---- cta_group::2: a CTA pair issues M=256 (128 rows per CTA) x N=256 MMAs; each CTA holds its A rows and HALF of B ----
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__global__ void __launch_bounds__(128, 1) mma2_rate_kernel(int n_chunks, int stages, int slots, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < stages * slots * 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tbase;
  long long t0 = clock64();
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem);
    const uint32_t stage_bytes = slots * 16384;
    constexpr uint32_t i256 = make_idesc(256, 256);
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t st = base + (uint32_t)(c % stages) * stage_bytes;
      for (int t = 0; t < 2; ++t)  // two ops per chunk: A in slot 0, this CTA's half of B in slot 1 / 2
        for (int ks = 0; ks < 4; ++ks) mma2(tb + t * 256, make_desc(st + ks * 32), make_desc(st + (1 + t) * 16384 + ks * 32), i256, 1);
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
  if (threadIdx.x == 0) {
    mbar_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  const int stages = 4, slots = 3, n_chunks = 400;
  const int smem = stages * slots * 16384 + 2048;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[8] = {"N=128 tile-outer, 4 accumulators", "N=256, 2 accumulators", "N=128, 1 accumulator", "N=128 k-step-outer",
                          "", "", "N=256, 1 accumulator", "N=128 A==B (diagonal)"};
  for (int grid : {1, 148}) {
    for (int mode : {0, 1, 2, 3, 6, 7}) {
      long long h[296];
      float best = 1e30f;
      long long cyc = 0, issue = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        mma_rate_kernel<<<grid, 128, smem>>>(mode, n_chunks, stages, slots, d);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0, mi = 0;
        for (int i = 0; i < grid; ++i) { if (h[2 * i + 1] > mx) mx = h[2 * i + 1]; if (h[2 * i] > mi) mi = h[2 * i]; }
        if (ms < best) { best = ms; cyc = mx; issue = mi; }
      }
      // MAC-equivalent: one "unit" = a 128x128x16 MMA (N=256 counts as two)
      const double units = n_chunks * 16.0;
      printf("grid %3d  %-34s  %7.1f cycles / 128x128x16-equivalent (issue loop %6.1f), kernel %.1f us -> %.2f GHz\n", grid, names[mode],
             cyc / units, issue / units, best * 1e3, cyc / (best * 1e-3) / 1e9);
    }
  }
  // cta_group::2
  {
    const int smem2 = stages * slots * 16384 + 2048;
    cudaFuncSetAttribute(mma2_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    for (int grid : {2, 148}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem2;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      long long h[148];
      long long mx = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, mma2_rate_kernel, n_chunks, stages, slots, d);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cta_group::2 error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        mx = 0;
        for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
      }
      // per CTA: n_chunks * 2 ops * 4 k-steps of a 256x256x16 pair-MMA = 128x256x16 per SM = two 128x128x16 equivalents
      printf("grid %3d  cta_group::2 M=256 N=256 (half of B per CTA)  %7.1f cycles / 128x128x16-equivalent per SM\n", grid,
             mx / (n_chunks * 16.0));
    }
  }
  return 0;
}
