// Gram matrix G = X X^T of the flattened bf16 frame vectors X [T, K], fp64 output.
//
//  impl 0: tcgen05 / TMA / TMEM kernel (sm_100a).  X is both operands (K-major, 128-byte swizzle),
//          so products are exact bf16 x bf16 and no limb splitting is needed.  Output tiles are
//          128 x 128; only the upper triangle is computed.  A CTA owns up to four tiles (512 TMEM
//          columns) and one K range; the ranges are summed afterwards in fp64 in a fixed order.
//  impl 1: CUDA-core kernel with the same tiling / partial layout (validation oracle on device).
//
// Partial layout: partial[kr][tile][128][128] fp32, tile = upper-triangular (bi <= bj) row-major id.
#include <cuda.h>

#include "csb_common.cuh"

namespace csb {

constexpr int kTile = 128;       // output tile edge
constexpr int kChunk = 64;       // K elements per pipeline stage (128 bytes = one swizzle atom)
constexpr int kMaxSlots = 4;     // distinct 128-row blocks resident per stage
constexpr int kMaxMma = 4;       // output tiles (TMEM 128-column groups) per CTA
constexpr int kMaxStages = 8;
constexpr int kBlockBytes = kTile * kChunk * 2;  // 16 KB
constexpr int kStageBudget = 200 * 1024;         // shared memory spent on pipeline stages
constexpr int kSegChunks = 96;                   // accumulator chain length (chunks) between drains
constexpr int kPrefetch = 8;                     // chunks of L2 prefetch distance (148 CTAs x 8 x 48 KB = 57 MB < L2)
constexpr int kMaxParamUnits = 16;               // unit tables up to this size travel as a kernel parameter

struct GramUnit {
  int n_slots;
  int n_tiles;
  int slot_block[kMaxSlots];  // row block loaded into each smem slot
  int a_slot[kMaxMma];        // per tile: smem slot of the A (row) block
  int b_slot[kMaxMma];        // per tile: smem slot of the B (column) block
  int tile_id[kMaxMma];       // upper-triangular tile id (for the partial buffer)
  int pair[kMaxMma];          // 1: head of an N=256 MMA that also covers tile t+1 (same A, adjacent B slots); 2: covered; 0: N=128
};

struct GramUnitTable {
  GramUnit u[kMaxParamUnits];
};

struct GramPlan {
  int nb;        // 128-row blocks
  int b0;        // first column block to compute (0 = whole matrix)
  int n_tiles;   // nb*(nb+1)/2
  int n_units;
  int n_kr;      // K ranges
  int n_chunks;  // ceil(K / 64)
  int max_slots; // largest number of resident blocks any unit needs
  int n_stages;  // pipeline depth that fits the shared-memory budget
};

// Upper-triangular tiles (bi <= bj) of the column blocks bj >= b0, column-major: id = sum_{c=b0}^{bj-1} (c+1) + bi.
// b0 = 0 is the whole Gram matrix; b0 > 0 is the extension by new frames (old columns are already known).
static inline int tri_id(int bi, int bj, int b0) { return bj * (bj + 1) / 2 - b0 * (b0 + 1) / 2 + bi; }
static inline int tri_count(int nb, int b0) { return nb * (nb + 1) / 2 - b0 * (b0 + 1) / 2; }

// Pack the upper-triangular tiles into units of <= 4 tiles touching <= 4 distinct row blocks (fewer
// resident blocks per stage = deeper TMA pipeline and less L2 traffic).  Small problems: greedy "seed +
// the tile that adds the fewest new blocks"; large ones: 2x2 super-blocks (4 tiles / 4 blocks off the
// diagonal, 3 tiles / 2 blocks on it).  T=600 (5 blocks, 15 tiles) packs into 4 units of 3 blocks each.
static void unit_add(GramUnit& u, int bi, int bj, int b0) {
  auto slot = [&](int blk) {
    for (int s = 0; s < u.n_slots; ++s)
      if (u.slot_block[s] == blk) return s;
    u.slot_block[u.n_slots] = blk;
    return u.n_slots++;
  };
  const int sa = slot(bi), sb = slot(bj);
  u.a_slot[u.n_tiles] = sa;
  u.b_slot[u.n_tiles] = sb;
  u.tile_id[u.n_tiles] = tri_id(bi, bj, b0);
  ++u.n_tiles;
}

static int unit_new_blocks(const GramUnit& u, int bi, int bj) {
  bool ha = false, hb = false;
  for (int s = 0; s < u.n_slots; ++s) {
    ha |= u.slot_block[s] == bi;
    hb |= u.slot_block[s] == bj;
  }
  return (ha ? 0 : 1) + ((hb || bj == bi) ? 0 : 1);
}

static void plan_units(int nb, int b0, GramUnit* units, int* n_units, int* max_slots) {
  int n = 0, ms = 0;
  auto emit = [&](GramUnit u) {
    // (1) slots in ascending block order, so that neighbouring column blocks sit in neighbouring 16 KB slots
    int order[kMaxSlots], inv[kMaxSlots];
    for (int i = 0; i < u.n_slots; ++i) order[i] = i;
    for (int i = 1; i < u.n_slots; ++i)
      for (int j = i; j > 0 && u.slot_block[order[j - 1]] > u.slot_block[order[j]]; --j) { int x = order[j]; order[j] = order[j - 1]; order[j - 1] = x; }
    int nb_sorted[kMaxSlots];
    for (int i = 0; i < u.n_slots; ++i) { nb_sorted[i] = u.slot_block[order[i]]; inv[order[i]] = i; }
    for (int i = 0; i < u.n_slots; ++i) u.slot_block[i] = nb_sorted[i];
    for (int t = 0; t < u.n_tiles; ++t) { u.a_slot[t] = inv[u.a_slot[t]]; u.b_slot[t] = inv[u.b_slot[t]]; }
    // (2) tiles sorted by (A slot, B slot)
    for (int i = 1; i < u.n_tiles; ++i)
      for (int j = i; j > 0 && (u.a_slot[j - 1] > u.a_slot[j] || (u.a_slot[j - 1] == u.a_slot[j] && u.b_slot[j - 1] > u.b_slot[j])); --j) {
        int x;
        x = u.a_slot[j]; u.a_slot[j] = u.a_slot[j - 1]; u.a_slot[j - 1] = x;
        x = u.b_slot[j]; u.b_slot[j] = u.b_slot[j - 1]; u.b_slot[j - 1] = x;
        x = u.tile_id[j]; u.tile_id[j] = u.tile_id[j - 1]; u.tile_id[j - 1] = x;
      }
    // (3) same A + adjacent B slots -> one N=256 MMA (A is read from shared memory once for both tiles)
    for (int t = 0; t < kMaxMma; ++t) u.pair[t] = 0;
    for (int t = 0; t + 1 < u.n_tiles; ++t)
      if (u.pair[t] == 0 && u.a_slot[t] == u.a_slot[t + 1] && u.b_slot[t + 1] == u.b_slot[t] + 1) { u.pair[t] = 1; u.pair[t + 1] = 2; }
    if (units) units[n] = u;
    ++n;
    if (u.n_slots > ms) ms = u.n_slots;
  };
  if (tri_count(nb, b0) <= 136) {
    const int nt = tri_count(nb, b0);
    int tb[136][2];
    bool done[136];
    int c = 0;
    for (int bi = 0; bi < nb; ++bi)
      for (int bj = (bi > b0 ? bi : b0); bj < nb; ++bj) { tb[c][0] = bi; tb[c][1] = bj; done[c] = false; ++c; }
    for (int seed = 0; seed < nt; ++seed) {
      if (done[seed]) continue;
      GramUnit u;
      memset(&u, 0, sizeof(u));
      unit_add(u, tb[seed][0], tb[seed][1], b0);
      done[seed] = true;
      while (u.n_tiles < kMaxMma) {
        int best = -1, best_new = 99;
        for (int t = 0; t < nt; ++t) {
          if (done[t]) continue;
          const int nw = unit_new_blocks(u, tb[t][0], tb[t][1]);
          if (u.n_slots + nw <= kMaxSlots && nw < best_new) { best = t; best_new = nw; }
        }
        if (best < 0) break;
        unit_add(u, tb[best][0], tb[best][1], b0);
        done[best] = true;
      }
      emit(u);
    }
  } else {
    const int ns = (nb + 1) / 2;
    for (int I = 0; I < ns; ++I)
      for (int J = I; J < ns; ++J) {
        GramUnit u;
        memset(&u, 0, sizeof(u));
        for (int bi = 2 * I; bi < 2 * I + 2 && bi < nb; ++bi)
          for (int bj = 2 * J; bj < 2 * J + 2 && bj < nb; ++bj)
            if (bi <= bj && bj >= b0) unit_add(u, bi, bj, b0);
        if (u.n_tiles) emit(u);
      }
  }
  *n_units = n;
  if (max_slots) *max_slots = ms;
}

static GramPlan make_plan(int64_t T, int64_t K, int64_t T_old = 0) {
  GramPlan p;
  p.nb = (int)((T + kTile - 1) / kTile);
  p.b0 = (int)(T_old / kTile);
  p.n_tiles = tri_count(p.nb, p.b0);
  plan_units(p.nb, p.b0, nullptr, &p.n_units, &p.max_slots);
  p.n_chunks = (int)((K + kChunk - 1) / kChunk);
  int kr = kNumSMs / p.n_units;
  if (kr < 1) kr = 1;
  if (kr > p.n_chunks) kr = p.n_chunks;
  p.n_kr = kr;
  int st = kStageBudget / (p.max_slots * kBlockBytes);
  if (st > kMaxStages) st = kMaxStages;
  if (st < 2) st = 2;
  p.n_stages = st;
  return p;
}

// ---------------------------------------------------------------------------------------------------
// impl 1: CUDA cores.  CTA = one 128x128 tile x one K range; 256 threads, 8x8 register tile.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_simt_kernel(const __nv_bfloat16* __restrict__ X, int T, int64_t K, int nb, int b0,
                                                        int n_tiles, int n_kr, int n_chunks, float* __restrict__ partial) {
  __shared__ float sA[32][kTile + 4];
  __shared__ float sB[32][kTile + 4];
  // decode tile id -> (bi, bj)
  int tid = blockIdx.x, bj = b0;
  while (tid >= bj + 1) { tid -= bj + 1; ++bj; }
  const int bi = tid;
  const int kr = blockIdx.y;
  const int c0 = (int)((int64_t)kr * n_chunks / n_kr), c1 = (int)((int64_t)(kr + 1) * n_chunks / n_kr);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int64_t k0 = (int64_t)c0 * kChunk; k0 < (int64_t)c1 * kChunk && k0 < K; k0 += 32) {
    for (int e = threadIdx.x; e < kTile * 32; e += 256) {
      const int r = e >> 5, kk = e & 31;
      const int64_t kcol = k0 + kk;
      const int ra = bi * kTile + r, rb = bj * kTile + r;
      sA[kk][r] = (ra < T && kcol < K) ? __bfloat162float(X[(int64_t)ra * K + kcol]) : 0.f;
      sB[kk][r] = (rb < T && kcol < K) ? __bfloat162float(X[(int64_t)rb * K + kcol]) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < 32; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = sA[kk][ty * 8 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = sB[kk][tx * 8 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + ((int64_t)kr * n_tiles + blockIdx.x) * (kTile * kTile);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(ty * 8 + i) * kTile + tx * 8 + j] = acc[i][j];
}

// ---------------------------------------------------------------------------------------------------
// Sum the K ranges in fp64 (fixed order) and write both triangles of G.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_reduce_kernel(const float* __restrict__ partial, int T, int nb, int b0, int n_tiles,
                                                          int n_kr, double* __restrict__ G, int64_t ldg) {
  // grid (n_tiles, 16): a thread owns 4 consecutive columns of one tile row; the K ranges are streamed with 8
  // independent 128-bit loads in flight and added in ascending range order (fixed -> deterministic).
  int tid = blockIdx.x, bj = b0;
  while (tid >= bj + 1) { tid -= bj + 1; ++bj; }
  const int bi = tid;
  const int e4 = blockIdx.y * 256 + threadIdx.x;  // float4 index inside the tile, 4096 per tile
  const int r = e4 >> 5, c = (e4 & 31) * 4;
  const int gi = bi * kTile + r, gj = bj * kTile + c;
  if (gi >= T || gj >= T) return;
  const float4* base = reinterpret_cast<const float4*>(partial + (int64_t)blockIdx.x * (kTile * kTile)) + e4;
  const int64_t kr_stride4 = (int64_t)n_tiles * kTile * kTile / 4;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int kr = 0;
  for (; kr + 8 <= n_kr; kr += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcs(base + (int64_t)(kr + u) * kr_stride4);
#pragma unroll
    for (int u = 0; u < 8; ++u) { s0 += (double)v[u].x; s1 += (double)v[u].y; s2 += (double)v[u].z; s3 += (double)v[u].w; }
  }
  for (; kr < n_kr; ++kr) {
    const float4 v = __ldcs(base + (int64_t)kr * kr_stride4);
    s0 += (double)v.x; s1 += (double)v.y; s2 += (double)v.z; s3 += (double)v.w;
  }
  const double sv[4] = {s0, s1, s2, s3};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (gj + u < T) {
      G[(int64_t)gi * ldg + gj + u] = sv[u];
      if (bi != bj) G[(int64_t)(gj + u) * ldg + gi] = sv[u];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// impl 0: tcgen05 + TMA.
// ---------------------------------------------------------------------------------------------------
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace ptx

// K-major, 128-byte-swizzled shared-memory operand descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO=1 [16,30) | SBO=1024>>4 [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct __align__(8) GramSmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t accum_done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(160, 1) gram_tc_kernel(const __grid_constant__ CUtensorMap xmap,
                                                         const __grid_constant__ GramUnitTable ptable,
                                                         const GramUnit* __restrict__ units, int n_tiles, int n_kr,
                                                         int n_chunks, int kStages, int stage_bytes,
                                                         float* __restrict__ partial, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // stage buffers first (1024-byte aligned for SWIZZLE_128B), control block after
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GramSmemCtl* ctl = reinterpret_cast<GramSmemCtl*>(smem + (size_t)kStages * stage_bytes);
  __shared__ GramUnit unit;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kr = blockIdx.y;
  const int c0 = (int)((int64_t)kr * n_chunks / n_kr), c1 = (int)((int64_t)(kr + 1) * n_chunks / n_kr);
  const int nchunk = c1 - c0;

  if (threadIdx.x == 0) unit = units ? units[blockIdx.x] : ptable.u[blockIdx.x];
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    ptx::mbar_init(&ctl->accum_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&ctl->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const int n_slots = unit.n_slots, n_ut = unit.n_tiles;

  if (warp == 4) {
    if (lane == 0) {
      // ===== TMA producer (dedicated warp: it may run ahead across accumulator drains) =====
      asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
      // L2 prefetch runs kPrefetch chunks ahead of the shared-memory ring: the ring then only has to cover L2-hit
      // latency, the (longer) DRAM latency is hidden without spending shared memory on it.
      for (int c = 0; c < kPrefetch && c < nchunk && !(dbg & 2); ++c)
        for (int q = 0; q < n_slots; ++q) ptx::tma_prefetch_l2_2d(&xmap, (c0 + c) * kChunk, unit.slot_block[q] * kTile);
      for (int c = 0; c < nchunk; ++c) {
        if (!(dbg & 2) && c + kPrefetch < nchunk)
          for (int q = 0; q < n_slots; ++q)
            ptx::tma_prefetch_l2_2d(&xmap, (c0 + c + kPrefetch) * kChunk, unit.slot_block[q] * kTile);
        const int s = c % kStages;
        const uint32_t ph = (uint32_t)(c / kStages) & 1u;
        ptx::mbar_wait(&ctl->empty[s], ph ^ 1u);
        if (dbg & 2) {  // profiling aid: no loads, the MMAs run on whatever is in shared memory
          ptx::mbar_expect_tx(&ctl->full[s], 0);
          continue;
        }
        ptx::mbar_expect_tx(&ctl->full[s], (uint32_t)(n_slots * kBlockBytes));
        uint8_t* stage = smem + (size_t)s * stage_bytes;
        for (int q = 0; q < n_slots; ++q)
          ptx::tma_load_2d(stage + (size_t)q * kBlockBytes, &xmap, &ctl->full[s], (c0 + c) * kChunk, unit.slot_block[q] * kTile);
      }
    }
  } else {
    // ===== warps 0-3: MMA issue (warp 1, lane 0) + accumulator drain (all four; warp w owns TMEM lanes 32w..32w+31) =====
    // The tensor core accumulates in fp32 with truncation, so a chain of n MMAs drifts by up to ~n * 2^-24 relative.  Chains
    // are cut every kSegChunks chunks: the accumulators are drained and added (round-to-nearest fp32) into the partial
    // tile, then restarted from zero.  Short K ranges (config 2: 88 chunks per CTA) are a single segment.
    constexpr uint32_t idesc = make_idesc(128, 128), idesc256 = make_idesc(128, 256);
    const int nseg = (nchunk + kSegChunks - 1) / kSegChunks;
    for (int seg = 0; seg < nseg; ++seg) {
      const int cs = seg * kSegChunks, ce = min(nchunk, cs + kSegChunks);
      if (warp == 1 && lane == 0) {
        for (int c = cs; c < ce; ++c) {
          const int s = c % kStages;
          const uint32_t ph = (uint32_t)(c / kStages) & 1u;
          ptx::mbar_wait(&ctl->full[s], ph);
          ptx::tc_fence_after();
          const uint32_t stage = ptx::smem_u32(smem + (size_t)s * stage_bytes);
          // tile-outer / k-step-inner: four back-to-back MMAs accumulate into the same TMEM tile.  (Measured: the k-step-outer
          // order, which would let tiles that share A reuse it through the A collector, runs 2x slower -- alternating the
          // accumulator between consecutive MMAs costs far more than the saved shared-memory reads.)
          for (int t = 0; t < n_ut && !(dbg & 1); ++t) {  // dbg&1: profiling aid, loads only
            const int pr = (dbg & 4) ? 0 : unit.pair[t];
            if (pr == 2) continue;  // covered by the N=256 MMA of tile t-1
            const uint32_t a_addr = stage + (uint32_t)unit.a_slot[t] * kBlockBytes;
            const uint32_t b_addr = stage + (uint32_t)unit.b_slot[t] * kBlockBytes;
            const uint32_t id = pr == 1 ? idesc256 : idesc;
#pragma unroll
            for (int ks = 0; ks < kChunk / 16; ++ks) {
              ptx::umma_bf16(tmem_base + (uint32_t)t * 128u, make_smem_desc(a_addr + ks * 32), make_smem_desc(b_addr + ks * 32),
                             id, (c > cs || ks > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit(&ctl->empty[s]);  // frees the stage once these MMAs have read it
        }
        ptx::umma_commit(&ctl->accum_done);
      }
      __syncwarp();
      ptx::mbar_wait(&ctl->accum_done, (uint32_t)seg & 1u);
      ptx::tc_fence_after();
      for (int t = 0; t < n_ut; ++t) {
        float* out = partial + ((int64_t)kr * n_tiles + unit.tile_id[t]) * (kTile * kTile) + (int64_t)(warp * 32 + lane) * kTile;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 128 + cb * 32), r);
          ptx::tmem_ld_wait();
          float4* o4 = reinterpret_cast<float4*>(out + cb * 32);
          if (seg == 0) {
#pragma unroll
            for (int v = 0; v < 8; ++v)
              o4[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                  __uint_as_float(r[4 * v + 3]));
          } else {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              float4 o = o4[v];
              o.x = __fadd_rn(o.x, __uint_as_float(r[4 * v]));
              o.y = __fadd_rn(o.y, __uint_as_float(r[4 * v + 1]));
              o.z = __fadd_rn(o.z, __uint_as_float(r[4 * v + 2]));
              o.w = __fadd_rn(o.w, __uint_as_float(r[4 * v + 3]));
              o4[v] = o;
            }
          }
        }
      }
      if (seg + 1 < nseg) {
        ptx::tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");  // every accumulator has been read before the next segment overwrites TMEM
        ptx::tc_fence_after();
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace csb

using namespace csb;

extern "C" {

static size_t gram_ws_bytes(const GramPlan& p) {
  // the CUDA-core path may use more K ranges than the tensor path; size for the larger
  int kr_simt = (2 * kNumSMs + p.n_tiles - 1) / p.n_tiles;
  if (kr_simt > p.n_chunks) kr_simt = p.n_chunks;
  const int kr = kr_simt > p.n_kr ? kr_simt : p.n_kr;
  return align_up((size_t)kr * p.n_tiles * kTile * kTile * sizeof(float), 256) + align_up(sizeof(GramUnit) * p.n_units, 256);
}

size_t csb_gram_workspace_bytes(int64_t T, int64_t K) {
  if (T <= 0 || K <= 0) return 0;
  return gram_ws_bytes(make_plan(T, K, 0));  // the whole matrix is the largest job for a given T
}

int csb_gram_extend(const void* x, int64_t T_old, int64_t T, int64_t K, double* G, int64_t ldg, void* workspace,
                    size_t workspace_bytes, int impl, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(x && G && workspace && T > 0 && K > 0, "gram: null argument or empty shape");
  CSB_REQUIRE(T_old >= 0 && T_old < T && ldg >= T, "gram: need 0 <= T_old < T <= ldg (T_old=%lld T=%lld ldg=%lld)", (long long)T_old,
              (long long)T, (long long)ldg);
  CSB_REQUIRE(K % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0, "gram: K %% 8 != 0 or x not 16-byte aligned");
  CSB_REQUIRE(T < (1 << 24), "gram: T too large");
  GramPlan p = make_plan(T, K, T_old);
  if (workspace_bytes < gram_ws_bytes(p)) {
    set_error("gram: workspace %zu < %zu", workspace_bytes, gram_ws_bytes(p));
    return CSB_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  float* partial = reinterpret_cast<float*>(workspace);
  int n_kr;
  if ((impl & 15) == 1) {
    n_kr = (2 * kNumSMs + p.n_tiles - 1) / p.n_tiles;
    if (n_kr > p.n_chunks) n_kr = p.n_chunks;
    dim3 grid((unsigned)p.n_tiles, (unsigned)n_kr);
    gram_simt_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (int)T, K, p.nb, p.b0, p.n_tiles, n_kr, p.n_chunks, partial);
    int st = check_launch("gram_simt");
    if (st) return st;
  } else if ((impl & 15) == 0) {
    n_kr = p.n_kr;
    int kr_simt = (2 * kNumSMs + p.n_tiles - 1) / p.n_tiles;
    if (kr_simt > p.n_chunks) kr_simt = p.n_chunks;
    const int kr_max = kr_simt > p.n_kr ? kr_simt : p.n_kr;
    GramUnit* d_units = nullptr;
    static thread_local GramUnitTable table;  // plan for small problems, passed by value
    cudaError_t e = cudaSuccess;
    if (p.n_units <= kMaxParamUnits) {
      static thread_local int table_nb = -1, table_b0 = -1;
      if (table_nb != p.nb || table_b0 != p.b0) {
        int nu = 0;
        plan_units(p.nb, p.b0, table.u, &nu, nullptr);
        table_nb = p.nb;
        table_b0 = p.b0;
      }
    } else {
      d_units = reinterpret_cast<GramUnit*>(reinterpret_cast<char*>(workspace) +
                                            align_up((size_t)kr_max * p.n_tiles * kTile * kTile * sizeof(float), 256));
      GramUnit* h_units = (GramUnit*)malloc(sizeof(GramUnit) * p.n_units);
      int nu = 0;
      plan_units(p.nb, p.b0, h_units, &nu, nullptr);
      e = cudaMemcpyAsync(d_units, h_units, sizeof(GramUnit) * nu, cudaMemcpyHostToDevice, s);
      cudaStreamSynchronize(s);  // pageable source buffer is freed below (large-T path only)
      free(h_units);
      if (e != cudaSuccess) { set_error("gram: unit table copy: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
    }

    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) { set_error("gram: cuTensorMapEncodeTiled unavailable"); return CSB_ERR_CUDA; }
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)T};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kChunk, (cuuint32_t)kTile};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("gram: cuTensorMapEncodeTiled failed (%d)", (int)r); return CSB_ERR_CUDA; }
    const int stage_bytes = p.max_slots * kBlockBytes;
    const size_t smem_bytes = (size_t)p.n_stages * stage_bytes + sizeof(GramSmemCtl) + 1024;
    e = cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) { set_error("gram: smem attr: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
    dim3 grid((unsigned)p.n_units, (unsigned)n_kr);
    gram_tc_kernel<<<grid, 160, smem_bytes, s>>>(map, table, d_units, p.n_tiles, n_kr, p.n_chunks, p.n_stages, stage_bytes, partial,
                                                     impl >> 4);
    int st = check_launch("gram_tc");
    if (st) return st;
  } else {
    set_error("gram: unknown impl %d", impl);
    return CSB_ERR_ARG;
  }
  dim3 rgrid((unsigned)p.n_tiles, 16);
  gram_reduce_kernel<<<rgrid, 256, 0, s>>>(partial, (int)T, p.nb, p.b0, p.n_tiles, n_kr, G, ldg);
  return check_launch("gram_reduce");
}

int csb_gram(const void* x, int64_t T, int64_t K, double* G, void* workspace, size_t workspace_bytes, int impl,
             csb_stream_t stream) {
  return csb_gram_extend(x, 0, T, K, G, T, workspace, workspace_bytes, impl, stream);
}

}  // extern "C"
