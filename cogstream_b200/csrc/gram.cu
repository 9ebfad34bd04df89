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
  int n_ops;
  int slot_block[kMaxSlots];  // row block loaded into each smem slot
  // output tiles = TMEM 128-column groups, in op order
  int tile_id[kMaxMma];       // upper-triangular tile id (for the partial buffer)
  int tile_rows[kMaxMma];     // valid rows / columns of the stored tile (the drain skips the rest)
  int tile_cols[kMaxMma];
  int tile_ablk[kMaxMma];     // row block of the accumulator rows (A) and of its columns (B): used when the drain writes G itself
  int tile_bblk[kMaxMma];
  // MMA ops: D[128 x n] (+)= A(slot a)[128 x 64] * B(slots b, b+1 ...)[n x 64]^T; n <= 256 covers one or two adjacent B slots,
  // i.e. one or two output tiles starting at TMEM tile t0.  An N=256 MMA runs at the tensor-pipe floor (64 cycles per
  // 128x128x16), two N=128 MMAs take 84 each: shared-memory operand bandwidth (scripts/ubench/mma_rate.cu).
  int op_a[kMaxMma], op_b[kMaxMma], op_n[kMaxMma], op_t0[kMaxMma];
};

struct GramUnitTable {
  GramUnit u[kMaxParamUnits];
};

// Tiles whose op had A = the higher block are stored flipped ([bj rows][bi cols]); bit = upper-triangular tile id.
// Only the small-problem planner produces them (tile ids < 136).
struct GramFlipMask {
  unsigned long long w[3];
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

// ---- planner -------------------------------------------------------------------------------------------------------------
// An op is "row block a against one or two column blocks"; G is symmetric, so a tile (bi, bj) may be produced by either
// orientation (A = bi or A = bj).  Ops are packed into units of <= 4 tiles touching <= 4 distinct row blocks (fewer resident
// blocks per stage = deeper TMA ring, less L2 traffic); inside a unit the two B blocks of a pair sit in adjacent 16 KB slots.
struct PlanOp {
  int a, nbk, b[2];
};

static int op_blocks(const PlanOp& o, int* out) {
  int n = 0;
  auto add = [&](int v) {
    for (int i = 0; i < n; ++i)
      if (out[i] == v) return;
    out[n++] = v;
  };
  add(o.a);
  for (int i = 0; i < o.nbk; ++i) add(o.b[i]);
  return n;
}

// Slot order for the blocks of a unit such that every pair's B blocks are adjacent (b[0] then b[1]); false if impossible.
static bool unit_slot_order(const PlanOp* ops, int n_ops, int* blocks, int n_blocks) {
  int perm[kMaxSlots];
  for (int i = 0; i < n_blocks; ++i) perm[i] = i;
  // blocks[] is ascending on entry; enumerate permutations in lexicographic order (deterministic choice)
  for (;;) {
    bool ok = true;
    for (int o = 0; o < n_ops && ok; ++o) {
      if (ops[o].nbk != 2) continue;
      int p0 = -1, p1 = -1;
      for (int i = 0; i < n_blocks; ++i) {
        if (blocks[perm[i]] == ops[o].b[0]) p0 = i;
        if (blocks[perm[i]] == ops[o].b[1]) p1 = i;
      }
      ok = p1 == p0 + 1;
    }
    if (ok) {
      int tmp[kMaxSlots];
      for (int i = 0; i < n_blocks; ++i) tmp[i] = blocks[perm[i]];
      for (int i = 0; i < n_blocks; ++i) blocks[i] = tmp[i];
      return true;
    }
    // next permutation
    int i = n_blocks - 2;
    while (i >= 0 && perm[i] > perm[i + 1]) --i;
    if (i < 0) return false;
    int j = n_blocks - 1;
    while (perm[j] < perm[i]) --j;
    int x = perm[i]; perm[i] = perm[j]; perm[j] = x;
    for (int l = i + 1, r = n_blocks - 1; l < r; ++l, --r) { x = perm[l]; perm[l] = perm[r]; perm[r] = x; }
  }
}

static int union_blocks(const PlanOp* ops, int n_ops, int* out) {
  int n = 0;
  for (int o = 0; o < n_ops; ++o) {
    int b[3];
    const int m = op_blocks(ops[o], b);
    for (int i = 0; i < m; ++i) {
      bool have = false;
      for (int j = 0; j < n; ++j) have |= out[j] == b[i];
      if (!have) out[n++] = b[i];
    }
  }
  for (int i = 1; i < n; ++i)
    for (int j = i; j > 0 && out[j - 1] > out[j]; --j) { int x = out[j]; out[j] = out[j - 1]; out[j - 1] = x; }
  return n;
}

static bool unit_feasible(const PlanOp* ops, int n_ops) {
  int tiles = 0;
  for (int o = 0; o < n_ops; ++o) tiles += ops[o].nbk;
  if (tiles > kMaxMma) return false;
  int blocks[3 * (kMaxMma + 1)];
  const int nbk = union_blocks(ops, n_ops, blocks);
  if (nbk > kMaxSlots) return false;
  return unit_slot_order(ops, n_ops, blocks, nbk);
}

static void plan_units(int nb, int b0, int64_t T, GramUnit* units, int* n_units, int* max_slots, GramFlipMask* flips) {
  int n = 0, ms = 0;
  if (flips) flips->w[0] = flips->w[1] = flips->w[2] = 0;
  const int last_rows = (int)(T - (int64_t)(nb - 1) * kTile);  // valid rows of the last block
  auto block_rows = [&](int blk) { return blk == nb - 1 ? last_rows : kTile; };
  auto emit = [&](const PlanOp* ops_in, int n_ops) {
    // pairs first (their accumulators start at even TMEM tiles), then singles
    PlanOp ops[kMaxMma];
    int m = 0;
    for (int pass = 2; pass >= 1; --pass)
      for (int o = 0; o < n_ops; ++o)
        if (ops_in[o].nbk == pass) ops[m++] = ops_in[o];
    GramUnit u;
    memset(&u, 0, sizeof(u));
    int blocks[3 * (kMaxMma + 1)];
    u.n_slots = union_blocks(ops, n_ops, blocks);
    unit_slot_order(ops, n_ops, blocks, u.n_slots);
    for (int i = 0; i < u.n_slots; ++i) u.slot_block[i] = blocks[i];
    auto slot_of = [&](int blk) {
      for (int i = 0; i < u.n_slots; ++i)
        if (u.slot_block[i] == blk) return i;
      return -1;
    };
    for (int o = 0; o < n_ops; ++o) {
      u.op_a[o] = slot_of(ops[o].a);
      u.op_b[o] = slot_of(ops[o].b[0]);
      u.op_t0[o] = u.n_tiles;
      int ncols = 0;
      for (int i = 0; i < ops[o].nbk; ++i) {
        const int a = ops[o].a, b = ops[o].b[i];
        const int bi = a < b ? a : b, bj = a < b ? b : a;
        const int id = tri_id(bi, bj, b0);
        u.tile_id[u.n_tiles] = id;
        u.tile_rows[u.n_tiles] = block_rows(a);
        u.tile_cols[u.n_tiles] = block_rows(b);
        u.tile_ablk[u.n_tiles] = a;
        u.tile_bblk[u.n_tiles] = b;
        if (a > b && flips && id < 192) flips->w[id >> 6] |= 1ull << (id & 63);
        ++u.n_tiles;
        // MMA N: full 128 for a leading block, the valid rows rounded up to 16 for the trailing one
        ncols += (i + 1 < ops[o].nbk) ? kTile : ((block_rows(b) + 15) / 16) * 16;
      }
      u.op_n[o] = ncols;
    }
    u.n_ops = n_ops;
    if (units) units[n] = u;
    ++n;
    if (u.n_slots > ms) ms = u.n_slots;
  };
  if (nb <= 16) {  // <= 136 tiles
    // (1) ops: per row block a, its not-yet-covered tiles (a, b >= a) plus tiles (x, a), x < a, that star x left over, paired
    //     up in ascending block order.  An odd tile out is handed to the star of its other block when there is one.
    PlanOp ops[136];
    int n_ops = 0;
    int deferred[17][17], n_def[17];
    for (int i = 0; i < 17; ++i) n_def[i] = 0;
    for (int a = 0; a < nb; ++a) {
      int ent[40], ne = 0;
      for (int i = 0; i < n_def[a]; ++i) ent[ne++] = deferred[a][i];
      for (int b = (a > b0 ? a : b0); b < nb; ++b) ent[ne++] = b;
      if (ne & 1) {
        if (ent[ne - 1] > a) {  // largest entry is a block with a star still to come
          const int x = ent[--ne];
          deferred[x][n_def[x]++] = a;
        }
      }
      int single = -1;
      if (ne & 1) {  // nothing to hand on: the diagonal tile (cheapest single: one operand fetch) or the last entry stays alone
        int pos = ne - 1;
        for (int i = 0; i < ne; ++i)
          if (ent[i] == a) pos = i;
        single = ent[pos];
        for (int i = pos; i + 1 < ne; ++i) ent[i] = ent[i + 1];
        --ne;
      }
      for (int i = 0; i + 1 < ne; i += 2) { ops[n_ops].a = a; ops[n_ops].nbk = 2; ops[n_ops].b[0] = ent[i]; ops[n_ops].b[1] = ent[i + 1]; ++n_ops; }
      if (single >= 0) {
        // orientation of a single is free: keep A = the lower block so that a trailing partial block is the (trimmed) B
        const int lo = single < a ? single : a, hi = single < a ? a : single;
        ops[n_ops].a = lo; ops[n_ops].nbk = 1; ops[n_ops].b[0] = hi; ops[n_ops].b[1] = -1; ++n_ops;
      }
    }
    // (2) units: seed + the op that adds the fewest new blocks while the unit stays feasible
    bool done[136];
    for (int i = 0; i < n_ops; ++i) done[i] = false;
    for (int seed = 0; seed < n_ops; ++seed) {
      if (done[seed]) continue;
      PlanOp uo[kMaxMma + 1];
      int nu = 0;
      uo[nu++] = ops[seed];
      done[seed] = true;
      for (;;) {
        int cur[3 * (kMaxMma + 1)];
        const int ncur = union_blocks(uo, nu, cur);
        int best = -1, best_new = 99;
        for (int o = 0; o < n_ops; ++o) {
          if (done[o]) continue;
          uo[nu] = ops[o];
          if (!unit_feasible(uo, nu + 1)) continue;
          int tmp[3 * (kMaxMma + 1)];
          const int nw = union_blocks(uo, nu + 1, tmp) - ncur;
          if (nw < best_new) { best = o; best_new = nw; }
        }
        if (best < 0) break;
        uo[nu++] = ops[best];
        done[best] = true;
      }
      emit(uo, nu);
    }
  } else {
    // 2x2 super-blocks: rows {2I, 2I+1} x columns {2J, 2J+1}; each row block against its (one or two) column blocks.
    // Units are emitted macro-tile by macro-tile (12 x 12 super-blocks = 144 units ~ one wave of CTAs): the CTAs that run
    // concurrently then share 24 + 24 row blocks instead of ~300, so a block's K chunk is fetched from HBM once per wave and
    // served to its 12 users out of L2 (X does not fit in L2 for long streams: 531 MB at config 4).
    const int ns = (nb + 1) / 2;
    constexpr int kMacro = 12;
    for (int MI = 0; MI < ns; MI += kMacro)
      for (int MJ = MI; MJ < ns; MJ += kMacro)
        for (int I = MI; I < MI + kMacro && I < ns; ++I)
          for (int J = (MJ > I ? MJ : I); J < MJ + kMacro && J < ns; ++J) {
            PlanOp uo[kMaxMma];
            int nu = 0;
            for (int bi = 2 * I; bi < 2 * I + 2 && bi < nb; ++bi) {
              PlanOp o;
              o.a = bi; o.nbk = 0; o.b[0] = o.b[1] = -1;
              for (int bj = 2 * J; bj < 2 * J + 2 && bj < nb; ++bj)
                if (bi <= bj && bj >= b0) o.b[o.nbk++] = bj;
              if (o.nbk) uo[nu++] = o;
            }
            if (nu) emit(uo, nu);
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
  plan_units(p.nb, p.b0, T, nullptr, &p.n_units, &p.max_slots, nullptr);
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
__global__ void __launch_bounds__(512, 2) gram_reduce_kernel(const float* __restrict__ partial, int T, int nb, int b0, int n_tiles,
                                                             int n_kr, double* __restrict__ G, int64_t ldg,
                                                             const __grid_constant__ GramFlipMask flips, int transposed_store) {
  int tid = blockIdx.x, bj = b0;
  while (tid >= bj + 1) { tid -= bj + 1; ++bj; }
  int bi = tid;
  // Stored tile S[x][y]: the CUDA-core kernel writes rows = block bi, columns = block bj.  The tensor kernel stores its
  // accumulators transposed (rows = the op's B block), and an op of the small-problem planner may have used A = the higher
  // block (flip bit): rows = block bj unless both or neither apply.
  const int flipped = (blockIdx.x < 192) ? (int)((flips.w[blockIdx.x >> 6] >> (blockIdx.x & 63)) & 1ull) : 0;
  if (flipped != transposed_store) { const int x = bi; bi = bj; bj = x; }
  // grid (n_tiles, 32) x 512 threads: the CTA owns 4 rows x 128 columns of its tile, a lane pair 2 consecutive columns of one
  // row; the even lane sums the even K ranges, the odd lane the odd ones (each in ascending order), and the two halves are
  // added even + odd: a fixed order, so G is deterministic.  Up to 40 K ranges (148 SMs / 4 units = 37 at config 2) every load
  // of a thread is in flight at once -- one L2 round trip -- at 64 registers, i.e. with the whole grid resident.
  const int half = threadIdx.x & 1;
  const int e2 = blockIdx.y * 256 + (threadIdx.x >> 1);  // float2 index inside the tile, 8192 per tile
  const int r = e2 >> 6, c = (e2 & 63) * 2;
  const int gi = bi * kTile + r, gj = bj * kTile + c;
  const bool valid = gi < T && gj < T;  // no early return: the pair exchange below is a full-warp shuffle
  // launched with programmatic stream serialization: the blocks are resident before the Gram kernel ends; wait for its
  // completion (and the visibility of its partial tiles) here
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // address = (uniform: tile base + 2u K ranges) + (per thread, 32 bits: element + the lane's parity range): one address register
  const unsigned stride_b = (unsigned)n_tiles * (kTile * kTile * 4u);  // bytes between K ranges (< 2^31: n_tiles <= 136 x 64 KB)
  const unsigned toff = (unsigned)e2 * 8u + (unsigned)half * stride_b;
  const char* tbase = reinterpret_cast<const char*>(partial) + (size_t)blockIdx.x * (kTile * kTile * 4);
  double s0 = 0.0, s1 = 0.0;
  constexpr int kBatch = 20;
  for (int k0 = 0; k0 < n_kr; k0 += 2 * kBatch) {
    float2 v[kBatch];
    const char* ub = tbase + (size_t)k0 * stride_b;
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int kr = k0 + 2 * u + half;
      v[u] = (valid && kr < n_kr) ? __ldcs(reinterpret_cast<const float2*>(ub + (size_t)(2 * u) * stride_b + toff)) : make_float2(0.f, 0.f);
    }
    // ascending order; the padding terms are +0.0 (x + 0.0 == x)
#pragma unroll
    for (int u = 0; u < kBatch; ++u) { s0 += (double)v[u].x; s1 += (double)v[u].y; }
  }
  {
    const double o0 = __shfl_xor_sync(0xffffffffu, s0, 1), o1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    // even + odd on both lanes (commutative, so both lanes hold the same bits)
    s0 = half ? o0 + s0 : s0 + o0;
    s1 = half ? o1 + s1 : s1 + o1;
  }
  // Stores.  Even lanes write their 2 consecutive doubles of G[gi][gj..] (a warp covers 256 contiguous bytes); the mirrored
  // triangle is staged through shared memory so that the 4 rows of a column go out as one 32-byte run G[gj][gi0..gi0+3]
  // (column-strided 8-byte stores from the registers touched one sector per element).
  __shared__ double s_t[kTile][5];
  const bool vec_ok = (ldg & 1) == 0 && (reinterpret_cast<uintptr_t>(G) & 15u) == 0;
  if (valid && !half) {
    double* g = G + (int64_t)gi * ldg + gj;
    if (vec_ok && gj + 1 < T) {
      *reinterpret_cast<double2*>(g) = make_double2(s0, s1);
    } else {
      g[0] = s0;
      if (gj + 1 < T) g[1] = s1;
    }
  }
  if (bi != bj) {  // uniform over the CTA
    if (half) { s_t[c][r & 3] = s0; s_t[c + 1][r & 3] = s1; }
    __syncthreads();
    if (threadIdx.x < kTile) {
      const int col = threadIdx.x;
      const int gi0 = bi * kTile + blockIdx.y * 4, gc = bj * kTile + col;
      if (gc < T && gi0 < T) {
        double* g = G + (int64_t)gc * ldg + gi0;
        if (vec_ok && gi0 + 3 < T) {
          *reinterpret_cast<double2*>(g) = make_double2(s_t[col][0], s_t[col][1]);
          *reinterpret_cast<double2*>(g + 2) = make_double2(s_t[col][2], s_t[col][3]);
        } else {
          for (int u = 0; u < 4; ++u)
            if (gi0 + u < T) g[u] = s_t[col][u];
        }
      }
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
// long waits (the drain warps sit out the whole main loop): poll with a sleep in between instead of spinning on the barrier,
// so that the idle warps do not take issue slots / shared-memory cycles from the MMA operand fetch and from co-resident CTAs
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(500);
  }
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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
                                                         float* __restrict__ partial, int dbg,
                                                         unsigned long long* __restrict__ tl, double* __restrict__ Gout,
                                                         int64_t ldg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // stage buffers first (1024-byte aligned for SWIZZLE_128B), control block after
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GramSmemCtl* ctl = reinterpret_cast<GramSmemCtl*>(smem + (size_t)kStages * stage_bytes);
  __shared__ GramUnit unit;

  // warp index through a shuffle: provably warp-uniform, so the role branches and everything derived from them stay on the
  // uniform datapath (no per-MMA "elect + R2UR" waterfall in the issue loop)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int kr = blockIdx.y;
  // profiling aid (dbg & 8): 8 globaltimer marks per CTA -- entry, setup done, first operands landed, last MMA issued,
  // accumulators complete, drain done, last TMA issued, SM id
  unsigned long long* mark = tl ? tl + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
  auto stamp = [&](int i) {
    if (mark) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      mark[i] = t;
    }
  };
  if (threadIdx.x == 0) {
    stamp(0);
    if (mark) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); mark[7] = sm; }
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the reduce kernel may be scheduled; it waits for our end
  const int c0 = (int)((int64_t)kr * n_chunks / n_kr), c1 = (int)((int64_t)(kr + 1) * n_chunks / n_kr);
  const int nchunk = c1 - c0;

  if (threadIdx.x < (int)(sizeof(GramUnit) / sizeof(int))) {
    const int* src = units ? reinterpret_cast<const int*>(units + blockIdx.x) : reinterpret_cast<const int*>(&ptable.u[blockIdx.x]);
    reinterpret_cast<int*>(&unit)[threadIdx.x] = src[threadIdx.x];
  }
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
  if (threadIdx.x == 0) stamp(1);

  if (warp == 4) {
    if (lane == 0) {
      // ===== TMA producer (dedicated warp: it may run ahead across accumulator drains) =====
      asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
      // L2 prefetch runs kPrefetch chunks ahead of the shared-memory ring: the ring then only has to cover L2-hit
      // latency, the (longer) DRAM latency is hidden without spending shared memory on it.  The first ring fill is issued
      // before the first prefetches so that it is not queued behind them.
      for (int c = 0; c < nchunk; ++c) {
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
        if (c == kStages - 1 || (nchunk < kStages && c == nchunk - 1)) {
          // ring filled once: now start the L2 prefetch stream, kPrefetch chunks deep
          for (int f = c + 1; f < c + 1 + kPrefetch && f < nchunk; ++f)
            for (int q = 0; q < n_slots; ++q) ptx::tma_prefetch_l2_2d(&xmap, (c0 + f) * kChunk, unit.slot_block[q] * kTile);
        } else if (c >= kStages && c + kPrefetch < nchunk) {
          for (int q = 0; q < n_slots; ++q) ptx::tma_prefetch_l2_2d(&xmap, (c0 + c + kPrefetch) * kChunk, unit.slot_block[q] * kTile);
        }
      }
      stamp(6);
    }
  } else {
    // ===== warps 0-3: MMA issue (warp 1, one elected lane) + accumulator drain (all four; warp w owns TMEM lanes 32w..32w+31) =====
    // The tensor core accumulates in fp32 with truncation, so a chain of n MMAs drifts by up to ~n * 2^-24 relative.  Chains
    // are cut every kSegChunks chunks: the accumulators are drained and added (round-to-nearest fp32) into the partial
    // tile, then restarted from zero.  Short K ranges (config 2: 88 chunks per CTA) are a single segment.
    const int nseg = (nchunk + kSegChunks - 1) / kSegChunks;
    // per-op constants of the issue loop, made warp-uniform (shuffle) and hoisted: descriptor low words without the stage
    // offset, instruction descriptor, accumulator column
    uint32_t op_alo[kMaxMma], op_blo[kMaxMma], op_idesc[kMaxMma], op_col[kMaxMma];
    const int n_ops = __shfl_sync(0xffffffffu, unit.n_ops, 0);
#pragma unroll
    for (int o = 0; o < kMaxMma; ++o) {
      const int a = __shfl_sync(0xffffffffu, unit.op_a[o], 0), b = __shfl_sync(0xffffffffu, unit.op_b[o], 0);
      const int n = __shfl_sync(0xffffffffu, unit.op_n[o], 0), t0 = __shfl_sync(0xffffffffu, unit.op_t0[o], 0);
      op_alo[o] = ((uint32_t)(a * kBlockBytes) >> 4) | (1u << 16);
      op_blo[o] = ((uint32_t)(b * kBlockBytes) >> 4) | (1u << 16);
      op_idesc[o] = make_idesc(128, n);
      op_col[o] = tmem_base + (uint32_t)t0 * 128u;
    }
    constexpr uint32_t kDescHi = 64u | (1u << 14) | (2u << 29);  // SBO = 1024 B, version 1, SWIZZLE_128B
    const uint32_t smem_lo = ptx::smem_u32(smem) >> 4;
    for (int seg = 0; seg < nseg; ++seg) {
      const int cs = seg * kSegChunks, ce = min(nchunk, cs + kSegChunks);
      if (warp == 1) {
        for (int c = cs; c < ce; ++c) {
          const int s = c % kStages;
          const uint32_t ph = (uint32_t)(c / kStages) & 1u;
          ptx::mbar_wait(&ctl->full[s], ph);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            if (c == 0) stamp(2);
            const uint32_t so = smem_lo + (uint32_t)(s * stage_bytes >> 4);
            // op-outer / k-step-inner: four back-to-back MMAs accumulate into the same TMEM tile(s)
#pragma unroll
            for (int o = 0; o < kMaxMma; ++o) {
              if (o < n_ops && !(dbg & 1)) {  // dbg&1: profiling aid, loads only
#pragma unroll
                for (int ks = 0; ks < kChunk / 16; ++ks) {
                  const uint64_t ad = ((uint64_t)kDescHi << 32) | (uint64_t)(op_alo[o] + so + 2u * ks);
                  const uint64_t bd = ((uint64_t)kDescHi << 32) | (uint64_t)(op_blo[o] + so + 2u * ks);
                  ptx::umma_bf16(op_col[o], ad, bd, op_idesc[o], (c > cs || ks > 0) ? 1u : 0u);
                }
              }
            }
            ptx::umma_commit(&ctl->empty[s]);  // frees the stage once these MMAs have read it
            if (c == ce - 1) {
              ptx::umma_commit(&ctl->accum_done);
              stamp(3);
            }
          }
          __syncwarp();
        }
      }
      ptx::mbar_wait_relaxed(&ctl->accum_done, (uint32_t)seg & 1u);
      ptx::tc_fence_after();
      if (threadIdx.x == 0) stamp(4);
      // Drain: lane = accumulator row.  The tile is stored TRANSPOSED ([column][row]): for a fixed column the 32 lanes of a
      // warp write 32 consecutive floats = one full 128-byte line per store instruction (row-major float4 stores would touch
      // 32 different lines per instruction and serialise in the LSU: measured 10 us for 38 MB).  The reduce kernel undoes it.
      const int row = warp * 32 + lane;
      const bool direct = Gout != nullptr && seg == nseg - 1;
      float* stile = reinterpret_cast<float*>(smem);  // [128][129] fp32 staging tile (the operand ring is idle by now)
      for (int t = 0; t < n_ut; ++t) {
        float* out = partial + ((int64_t)kr * n_tiles + unit.tile_id[t]) * (kTile * kTile) + row;
        const bool row_ok = row < unit.tile_rows[t];
        const int ncols = unit.tile_cols[t];
        const int ncb = (ncols + 31) >> 5;  // 32-column groups that hold valid columns
        const bool diag = unit.tile_ablk[t] == unit.tile_bblk[t];
#pragma unroll 1
        for (int cb = 0; cb < ncb; ++cb) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 128 + cb * 32), r);
          ptx::tmem_ld_wait();
          if (!row_ok) continue;
          float* o = out + (int64_t)(cb * 32) * kTile;
          float val[32];
          if (seg == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) val[j] = __uint_as_float(r[j]);
          } else {
            // all 32 loads first (independent L2 round trips), then the adds: a load -> add -> store chain per element
            // serialises 32 round trips per column group (measured: 230 us per drain at config 4)
            float prev[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) prev[j] = (cb * 32 + j < ncols) ? __ldcg(o + j * kTile) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) val[j] = __fadd_rn(prev[j], __uint_as_float(r[j]));
          }
          if (direct) {
            // single K range (long streams): the last drain writes G itself -- both triangles -- and no reduce kernel runs.
            // Same value the reduce kernel would produce: (double) of the fp32 segment sum.  The mirror triangle goes out
            // straight from registers (for a fixed column the warp's 32 rows are 32 consecutive doubles); the other
            // orientation is staged through shared memory so that it is written in full 256-byte runs as well (one thread
            // per row would touch 32 lines per store instruction).
            const int64_t ga = (int64_t)unit.tile_ablk[t] * kTile + row;
            const int64_t gb0 = (int64_t)unit.tile_bblk[t] * kTile + cb * 32;
            if (!diag) {
              double* gcol = Gout + gb0 * ldg + ga;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cb * 32 + j < ncols) gcol[(int64_t)j * ldg] = (double)val[j];
            }
            float* srow = stile + row * (kTile + 1) + cb * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) srow[j] = val[j];  // bank = (row + column) mod 32: conflict-free
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cb * 32 + j < ncols) o[j * kTile] = val[j];
          }
        }
        if (direct) {
          asm volatile("bar.sync 1, 128;" ::: "memory");  // the tile is staged
          const int nrows = unit.tile_rows[t];
          double* gbase = Gout + ((int64_t)unit.tile_ablk[t] * kTile) * ldg + (int64_t)unit.tile_bblk[t] * kTile;
          for (int rr = warp; rr < nrows; rr += 4) {
            const float* sr = stile + rr * (kTile + 1);
            double* grow = gbase + (int64_t)rr * ldg;
            for (int c = lane; c < ncols; c += 32) grow[c] = (double)sr[c];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");  // before the next tile overwrites the staging buffer
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
  if (threadIdx.x == 0) stamp(5);
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
  GramFlipMask flips;
  flips.w[0] = flips.w[1] = flips.w[2] = 0;
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
    static thread_local GramFlipMask table_flips;
    cudaError_t e = cudaSuccess;
    if (p.n_units <= kMaxParamUnits) {
      static thread_local int64_t table_T = -1;
      static thread_local int table_b0 = -1;
      if (table_T != T || table_b0 != p.b0) {
        int nu = 0;
        plan_units(p.nb, p.b0, T, table.u, &nu, nullptr, &table_flips);
        table_T = T;
        table_b0 = p.b0;
      }
      flips = table_flips;
    } else {
      d_units = reinterpret_cast<GramUnit*>(reinterpret_cast<char*>(workspace) +
                                            align_up((size_t)kr_max * p.n_tiles * kTile * kTile * sizeof(float), 256));
      // The plans of the first kPlanCache long-stream shapes a thread sees are kept on the host for good, so their copy can stay
      // asynchronous: no stream synchronisation on the call path, and the call can be captured into a CUDA graph (the memcpy
      // node re-reads the same, unchanged host table at every replay).  Further shapes (e.g. a cache that grows frame by
      // frame) take the synchronous path: plan, copy, wait, free.
      constexpr int kPlanCache = 16;
      struct PlanCache { int64_t T, T_old; int n; GramUnit* units; GramFlipMask flips; };
      static thread_local PlanCache cache[kPlanCache];
      static thread_local int cache_n = 0;
      PlanCache* pc = nullptr;
      for (int i = 0; i < cache_n; ++i)
        if (cache[i].T == T && cache[i].T_old == T_old) pc = &cache[i];
      if (!pc && cache_n < kPlanCache) {
        pc = &cache[cache_n++];
        pc->T = T; pc->T_old = T_old;
        pc->units = (GramUnit*)malloc(sizeof(GramUnit) * p.n_units);
        pc->n = 0;
        plan_units(p.nb, p.b0, T, pc->units, &pc->n, nullptr, &pc->flips);
      }
      if (pc) {
        flips = pc->flips;
        e = cudaMemcpyAsync(d_units, pc->units, sizeof(GramUnit) * pc->n, cudaMemcpyHostToDevice, s);
      } else {
        GramUnit* h_units = (GramUnit*)malloc(sizeof(GramUnit) * p.n_units);
        int nu = 0;
        plan_units(p.nb, p.b0, T, h_units, &nu, nullptr, &flips);
        e = cudaMemcpyAsync(d_units, h_units, sizeof(GramUnit) * nu, cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);  // pageable source buffer is freed right below
        free(h_units);
      }
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
    // profiling aid: impl bit 7 (dbg & 8) + a workspace with room behind the documented size -> per-CTA time marks there
    unsigned long long* tl = nullptr;
    const size_t tl_bytes = (size_t)p.n_units * n_kr * 8 * sizeof(unsigned long long);
    if (((impl >> 4) & 8) && workspace_bytes >= gram_ws_bytes(p) + tl_bytes)
      tl = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + gram_ws_bytes(p));
    // one K range per tile (more units than SMs): the kernel's last drain writes G directly, no reduce pass
    const bool direct = n_kr == 1 && !((impl >> 4) & 32);  // impl bit 9: force the partial + reduce path (validation)
    if (!((impl >> 4) & 16))  // profiling aid: impl bit 8 = reduce kernel only
      gram_tc_kernel<<<grid, 160, smem_bytes, s>>>(map, table, d_units, p.n_tiles, n_kr, p.n_chunks, p.n_stages, stage_bytes, partial,
                                                     impl >> 4, tl, direct ? G : nullptr, ldg);
    int st = check_launch("gram_tc");
    if (st) return st;
    if (direct) return CSB_OK;
  } else {
    set_error("gram: unknown impl %d", impl);
    return CSB_ERR_ARG;
  }
  dim3 rgrid((unsigned)p.n_tiles, 32);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = rgrid;
  cfg.blockDim = dim3(512);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gram_reduce_kernel, (const float*)partial, (int)T, p.nb, p.b0, p.n_tiles, n_kr, G, ldg,
                                      flips, (impl & 15) == 0 ? 1 : 0);
  if (le != cudaSuccess) { set_error("gram_reduce: launch: %s", cudaGetErrorString(le)); return CSB_ERR_CUDA; }
  return check_launch("gram_reduce");
}

int csb_gram(const void* x, int64_t T, int64_t K, double* G, void* workspace, size_t workspace_bytes, int impl,
             csb_stream_t stream) {
  return csb_gram_extend(x, 0, T, K, G, T, workspace, workspace_bytes, impl, stream);
}

}  // extern "C"
