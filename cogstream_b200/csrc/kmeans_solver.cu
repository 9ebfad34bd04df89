// Temporal-semantic k-means on the Gram matrix: k-means++ init + Lloyd loop in ONE cooperative
// launch (reference: model/kmeans_with_time.py:41-131).
//
// Every feature distance the reference takes from torch.cdist is ||x_i||^2 + ||c_j||^2 - 2<x_i,c_j>
// with c_j a frame (init) or the mean of a set of frames M_j (Lloyd), hence
//     <x_i, c_j> = (1/n_j) * sum_{m in M_j} G[i][m]            =: S[i][j] / n_j
//     ||c_j||^2  = (1/n_j^2) * sum_{i in M_j} S[i][j]          =: q[j]
//     <c_j, c'_j> = (1/(n_j n'_j)) * sum_{i in M_j} S'[i][j]   (centre shift, :123)
// so the whole loop runs on the T x T fp64 Gram matrix (L2-resident) instead of streaming the
// [T, P*D] features 2*I + k - 1 times.  All sums are fp64 in a fixed order: deterministic, and
// unchanged memberships give a centre shift of exactly 0, as in the reference.
//
// Grid: nb co-resident CTAs (cooperative launch) separated by a software grid barrier (2 per Lloyd
// iteration).  k-means++ runs on CTA 0; in the loop every CTA rebuilds the membership CSR in its own
// shared memory (identical by construction), so only the assignments and the S columns cross CTAs.
// Columns of S are recomputed only for clusters whose membership changed (version stamps per buffer).
#include <stdlib.h>
#include <string.h>

#include "csb_common.cuh"

namespace csb {

struct SolverArgs {
  const double* G;
  int64_t ldg;  // leading dimension of G
  const void* ts;
  int T, k;
  int max_iter;
  double alpha, tol;
  int n_init_given, n_reseed_pool;
  int64_t* init_indices;
  const float* exp_noise;
  const int64_t* reseed_pool;
  int64_t* assignments;
  void* ct_out;
  int32_t* member_offsets;
  int32_t* members;
  int32_t* event_order;    // nullable
  int32_t* event_offsets;  // nullable
  csb_kmeans_status* status;
  // workspace
  double* diag;       // [T] Gram diagonal, compact copy
  double* S[2];       // [T,k]
  unsigned* barrier;  // zeroed before launch
  int32_t* used;      // reseed-pool entries consumed by the init (written by CTA 0)
  unsigned long long* tl;  // phase timeline (globaltimer ns), [0] = entries written; profiling aid
  int diag_in_smem;   // 1: keep a shared-memory copy of the diagonal (T small enough)
  int multiwarp_fill; // 1: per-warp cluster counters fit in shared memory
  const unsigned long long* rng;  // device {seed, offset} of the built-in generator, or nullptr (noise / pool supplied by the caller)
  int fused_init;     // 1: CTA 0 of the solver runs the k-means++ init itself (no separate kmeanspp launch)
  float* noise_ws;    // [(k-1), T] or nullptr: the built-in generator's noise, filled by CTAs 1.. of the fused launch while CTA 0
                      // sets up (it waits for them before its first sampling round)
  unsigned* noise_ready;  // arrival counter of those CTAs
  int init_exact;     // validation switch (CSB_INIT_EXACT): every sampling round evaluates the exact IEEE scores
  int32_t* assign32;  // [T] compact copy of the assignments (what every CTA reloads after the barrier)
  int assign_words;   // 4-byte words reserved for the shared-memory assignment array (16-bit entries; T words when the fused init
                      // aliases it as float[T])
  int s_no_pairs;     // validation switch (CSB_S_NO_PAIRS): the row-streaming S pass with one frame per thread
  int assign_exact;   // validation switch (CSB_ASSIGN_EXACT): the wide assignment pass evaluates every cluster exactly (no filter)
};

__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& target, unsigned nblocks) {
  __syncthreads();
  if (nblocks > 1) {
    if (threadIdx.x == 0) {
      target += nblocks;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      } while (v < target);
    }
    __syncthreads();
  }
}

// the same barrier with a sleeping poll: for the long wait while CTA 0 runs the fused k-means++ init (tens of microseconds),
// so that the other CTAs' pollers do not compete with CTA 0 for the L2
__device__ __forceinline__ void grid_sync_backoff(unsigned* counter, unsigned& target, unsigned nblocks) {
  __syncthreads();
  if (nblocks > 1) {
    if (threadIdx.x == 0) {
      target += nblocks;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned v;
      for (;;) {
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) break;
        __nanosleep(256);
      }
    }
    __syncthreads();
  }
}

// ---- built-in generator (csb_kmeans_params.rng_state): Philox4x32-10, counter = (index, stream, offset), key = seed ----------
// The fused pipeline draws its randomness here instead of launching torch's generator kernels in front of the solver: the Exp(1)
// noise of sampling round r for frame i is -log(U(r, i)), the host random.randint draws of kmeans_with_time.py:41 / :57 / :118
// are floor(U * T).  Every thread that needs draw (stream, index) computes the same value; the offset advances once per solve.
__device__ __forceinline__ uint4 philox_draw4(unsigned long long seed, unsigned long long offset, uint32_t stream, uint32_t index) {
  uint32_t c0 = index, c1 = stream, c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t philox_draw(unsigned long long seed, unsigned long long offset, uint32_t stream, uint32_t index) {
  return philox_draw4(seed, offset, stream, index).x;
}
__device__ __forceinline__ float exp1_from_bits(uint32_t x) {
  return -logf(((float)(x >> 8) + 0.5f) * 5.9604644775390625e-8f);  // U in (0, 1): 24 bits + 1/2, never 0 or 1
}
constexpr uint32_t kRngFirst = 0xFFFFFFFFu, kRngReseed = 0xFFFFFFFEu;  // streams; sampling round r uses stream r

struct RngKey {
  unsigned long long seed, offset;
};
__device__ __forceinline__ RngKey load_rng(const SolverArgs& a) {
  RngKey k;
  k.seed = a.rng ? a.rng[0] : 0ull;
  k.offset = a.rng ? a.rng[1] : 0ull;
  return k;
}
// Exp(1) noise of sampling round `row` for frame i from the built-in generator: component i & 3 of the call (i >> 2, row).
// Out of line on purpose: inlined into the k-means++ round loop (three call sites) the Philox rounds cost the torch-noise path
// 16 us per solve (register allocation / code size of the hot loop), although they never run there.
__device__ __noinline__ float noise_generate(unsigned long long seed, unsigned long long offset, int row, int i) {
  const uint4 x = philox_draw4(seed, offset, (uint32_t)row, (uint32_t)(i >> 2));
  const uint32_t w = (i & 2) ? ((i & 1) ? x.w : x.z) : ((i & 1) ? x.y : x.x);
  return exp1_from_bits(w);
}
// Exp(1) noise of sampling round `row` for frame i: the caller's array, the rows pre-generated by the idle CTAs, or the generator
// The k-means++ round loop is instantiated once per noise source, so that the array paths carry none of the generator's code
// (measured: with the generator reachable from the loop a solve takes 16 us longer even when it never runs):
//   kNoiseArray  the caller's array (read-only for the kernel: ld.global.nc)
//   kNoiseWs     rows pre-generated into the workspace by the idle CTAs of the fused launch (ld.global.cg: written in this kernel)
//   kNoiseGen    generated on the spot (separate init kernels)
enum { kNoiseArray = 0, kNoiseWs = 1, kNoiseGen = 2 };
template <int SRC>
__device__ __forceinline__ float noise_at(const SolverArgs& a, const RngKey& rk, int row, int i) {
  if (SRC == kNoiseArray) return __ldg(a.exp_noise + (int64_t)row * a.T + i);
  if (SRC == kNoiseWs) return __ldcg(a.noise_ws + (int64_t)row * a.T + i);
  return noise_generate(rk.seed, rk.offset, row, i);
}
// n-th random.randint(0, T-1) draw: the caller's pool (flags an overflow), or the built-in generator
__device__ __forceinline__ int reseed_at(const SolverArgs& a, const RngKey& rk, int n, bool* overflow) {
  if (a.rng && !a.reseed_pool) return (int)(((unsigned long long)philox_draw(rk.seed, rk.offset, kRngReseed, (uint32_t)n) * (unsigned)a.T) >> 32);
  if (n < a.n_reseed_pool) return (int)a.reseed_pool[n];
  *overflow = true;
  return 0;
}

constexpr int kTimelineMax = 440;
__device__ __forceinline__ void mark(const SolverArgs& a, int tag) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const unsigned long long n = a.tl[0];
    if (n < kTimelineMax) {
      a.tl[1 + n] = (t << 8) | (unsigned long long)(tag & 0xff);
      a.tl[0] = n + 1;
    }
  }
}

template <typename TS>
__device__ __forceinline__ TS load_ts(const void* ts, int i) {
  return reinterpret_cast<const TS*>(ts)[i];
}

constexpr int kSolverThreads = 256;
constexpr int kSolverWarps = kSolverThreads / 32;

// ---- k-means++ on feature distance, one CTA (kmeans_with_time.py:44-62) ------------------------------
// Sampling round: probs = nearest^2 / sum (:53-59), centre = argmax(probs / Exp(1) noise) -- the noise
// rows are consumed one per SAMPLING round, like the torch generator in the reference.  A round whose
// probs sum to 0 (:55-57) takes the next entry of the host-drawn random.randint pool instead.
// A round is one L2 round trip (row c of G, the next noise row alongside) and two block barriers:
//   * the exact score is RN(RN(n2 / sum) / noise): two IEEE divisions per frame.  The argmax is found from the
//     approximate score u = n2 * (1 / noise) (fast division, <= 2 ulp; the common factor 1 / sum is dropped): a frame with
//     u < u_max * (1 - 2^-20) has an exact score strictly below the exact score of the u_max frame (5 ulp of slack), so when the
//     second largest u of the block is below that bound the u_max frame IS the reference's pick -- no IEEE division at all.
//     Otherwise (near tie: rare) the round falls back to the exact scores, first max = lowest index.
//   * per-warp partial sums of nearest^2 are double-buffered in shared memory (added in warp order by every thread -> one fixed
//     order), so the update pass needs no barrier before publishing them.
struct PPScratch {
  float wv[32];
  int wi[32];
  float w2[32];
  float ws[2][32];
};

template <int NT, int RNG>
__device__ void kmeanspp_block_impl(const SolverArgs& a, float* s_near, const double* diag, PPScratch* sc, int* used_io,
                                    int* n_noise_used, int* n_degenerate) {
  constexpr int NW = NT / 32;
  const int T = a.T, k = a.k;
  const double* G = a.G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < T; i += blockDim.x) s_near[i] = __int_as_float(0x7f800000);
  __syncthreads();
  int noise_row = 0, degenerate = 0, used = *used_io;
  const RngKey rk = load_rng(a);
  // T <= 4 * NT (one frame batch per thread): the Exp(1) row of the NEXT sampling round is fetched (or generated) together with
  // the Gram row of the update pass, one L2 round trip per round instead of two dependent ones
  const bool prefetch_noise = T <= 4 * NT;
  float nzp[4] = {1.f, 1.f, 1.f, 1.f};
  bool nzp_valid = false;
  int wsb = 0;  // ws buffer the next sampling round reads
  bool ws_ready = false;
  for (int m = 0; m < k; ++m) {
    int c;
    if (m < a.n_init_given) {
      c = (int)a.init_indices[m];
    } else if (m == 0) {  // :41 the first centre, from the built-in generator (n_init_given == 0)
      c = (int)(((unsigned long long)philox_draw(rk.seed, rk.offset, kRngFirst, 0u) * (unsigned)T) >> 32);
      if (threadIdx.x == 0) a.init_indices[0] = c;
    } else {
      // :53-54 sum of nearest^2: per-warp partials were left in ws[wsb] by the update pass of the previous round
      const float* ws = sc->ws[wsb];
      float sum = ws[0];
#pragma unroll
      for (int w = 1; w < NW; ++w) sum += ws[w];
      if (sum == 0.f) {                           // :55-57
        ++degenerate;
        bool ovf = false;
        c = reseed_at(a, rk, used, &ovf);
        if (ovf && threadIdx.x == 0) a.status->reseed_overflow = 1;
        ++used;
      } else {
        const int this_row = noise_row;
        ++noise_row;
        // ---- approximate scores: this thread's best / second best ----
        float best = 0.f, second = 0.f;
        int besti = 0x7fffffff;
        for (int i0 = threadIdx.x; i0 < T; i0 += 4 * NT) {
          float nd[4], nz[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ii = i0 + u * NT;
            nd[u] = ii < T ? s_near[ii] : 0.f;
            nz[u] = nzp_valid ? nzp[u] : (ii < T ? noise_at<RNG>(a, rk, this_row, ii) : 1.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ii = i0 + u * NT;
            const float n2 = __fmul_rn(nd[u], nd[u]);
            const float uq = __fdividef(n2, nz[u]);  // frames that coincide with a centre: +0, never chosen
            if (ii < T) {
              if (uq > best) { second = best; best = uq; besti = ii; }
              else if (uq > second) second = uq;
            }
          }
        }
        // scores are >= +0, so their bit patterns order like unsigned integers: hardware warp reductions (REDUX)
        {
          const unsigned bits = __float_as_uint(best);
          const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
          const unsigned widx = __reduce_min_sync(0xffffffffu, bits == wmax ? (unsigned)besti : 0x7fffffffu);
          const bool mine = bits == wmax && (unsigned)besti == widx;
          const unsigned wsec = __reduce_max_sync(0xffffffffu, __float_as_uint(mine ? second : best));
          if (lane == 0) { sc->wv[warp] = __uint_as_float(wmax); sc->wi[warp] = (int)widx; sc->w2[warp] = __uint_as_float(wsec); }
        }
        __syncthreads();
        float bmaxf, secf;
        {
          const unsigned bits = lane < NW ? __float_as_uint(sc->wv[lane]) : 0u;
          const unsigned idx = lane < NW ? (unsigned)sc->wi[lane] : 0x7fffffffu;
          const unsigned s2 = lane < NW ? __float_as_uint(sc->w2[lane]) : 0u;
          const unsigned bmax = __reduce_max_sync(0xffffffffu, bits);
          const unsigned bidx = __reduce_min_sync(0xffffffffu, bits == bmax ? idx : 0x7fffffffu);
          const bool mine = bits == bmax && idx == bidx;
          const unsigned bsec = __reduce_max_sync(0xffffffffu, mine ? s2 : bits);
          bmaxf = __uint_as_float(bmax);
          secf = __uint_as_float(bsec);
          c = (int)bidx;
        }
        const bool unique = bmaxf > 0.f && secf < bmaxf - bmaxf * 9.5367431640625e-7f && !a.init_exact;
        if (!unique) {
          // ---- exact scores (near tie between the two best approximate scores, or the validation switch) ----
          __syncthreads();  // every warp has read wv / wi
          float eb = -1.f;
          int ei = 0x7fffffff;
          for (int i0 = threadIdx.x; i0 < T; i0 += 4 * NT) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int ii = i0 + u * NT;
              if (ii < T) {
                const float ndv = s_near[ii];
                const float nzv = nzp_valid ? nzp[u] : noise_at<RNG>(a, rk, this_row, ii);
                const float n2 = __fmul_rn(ndv, ndv);
                // a zero numerator would send the warp down the IEEE division's slow path: substituted, result selected
                const float p = __fdiv_rn(n2 == 0.f ? 1.f : n2, sum);  // :59
                float scv = __fdiv_rn(p, nzv);                           // multinomial: argmax(p / Exp(1)), first max
                scv = n2 == 0.f ? 0.f : scv;
                if (scv > eb) { eb = scv; ei = ii; }
              }
            }
          }
          {
            const unsigned bits = eb > 0.f ? __float_as_uint(eb) : 0u;
            const unsigned wmax = __reduce_max_sync(0xffffffffu, bits);
            const unsigned widx = __reduce_min_sync(0xffffffffu, bits == wmax ? (unsigned)ei : 0x7fffffffu);
            if (lane == 0) { sc->wv[warp] = __uint_as_float(wmax); sc->wi[warp] = (int)widx; }
          }
          __syncthreads();
          {
            const unsigned bits = lane < NW ? __float_as_uint(sc->wv[lane]) : 0u;
            const unsigned idx = lane < NW ? (unsigned)sc->wi[lane] : 0x7fffffffu;
            const unsigned bmax = __reduce_max_sync(0xffffffffu, bits);
            c = (int)__reduce_min_sync(0xffffffffu, bits == bmax ? idx : 0x7fffffffu);
          }
        }
      }
      if (threadIdx.x == 0) a.init_indices[m] = c;
    }
    if (m + 1 < k) {
      const double gcc = diag[c];
      const double* gc = G + (int64_t)c * a.ldg;
      float part = 0.f;
      // the workspace rows are filled by the other CTAs during the first microseconds of the launch: the first fetch waits for
      // them, and is issued after this round's update pass instead of next to its Gram row to give them that much longer
      const bool late_fetch = RNG == kNoiseWs && !ws_ready && m + 1 >= a.n_init_given;
      if (prefetch_noise && !late_fetch && m + 1 >= a.n_init_given) {  // row of the next sampling round (consumed in order)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ii = threadIdx.x + u * NT;
          nzp[u] = ii < T ? noise_at<RNG>(a, rk, noise_row, ii) : 1.f;
        }
        nzp_valid = true;
      }
      for (int i0 = threadIdx.x; i0 < T; i0 += 4 * NT) {
        double g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ii = i0 + u * NT;
          g[u] = ii < T ? __ldg(gc + ii) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ii = i0 + u * NT;
          if (ii < T) {
            const double d2 = diag[ii] + gcc - 2.0 * g[u];
            const float d = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
            const float nd = fminf(s_near[ii], d);
            s_near[ii] = nd;
            part += __fmul_rn(nd, nd);  // :53
          }
        }
      }
      part = warp_sum(part);
      wsb ^= 1;
      if (lane == 0) sc->ws[wsb][warp] = part;
      if (late_fetch && threadIdx.x == 0) {
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(a.noise_ready) : "memory");
        } while (v < gridDim.x - 1);
      }
      __syncthreads();
      if (late_fetch) {
        ws_ready = true;
        if (prefetch_noise) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int ii = threadIdx.x + u * NT;
            nzp[u] = ii < T ? noise_at<RNG>(a, rk, noise_row, ii) : 1.f;
          }
          nzp_valid = true;
        }
      }
    }
  }
  *used_io = used;
  *n_noise_used = noise_row;
  *n_degenerate = degenerate;
}

template <int NT>
__device__ __forceinline__ void kmeanspp_block(const SolverArgs& a, float* s_near, const double* diag, PPScratch* sc, int* used_io,
                                               int* n_noise_used, int* n_degenerate) {
  if (a.exp_noise) kmeanspp_block_impl<NT, kNoiseArray>(a, s_near, diag, sc, used_io, n_noise_used, n_degenerate);
  else if (a.noise_ws) kmeanspp_block_impl<NT, kNoiseWs>(a, s_near, diag, sc, used_io, n_noise_used, n_degenerate);
  else kmeanspp_block_impl<NT, kNoiseGen>(a, s_near, diag, sc, used_io, n_noise_used, n_degenerate);
}

// ---- membership CSR in shared memory from the new assignments (kmeans_with_time.py:110-120) ---------
// Every CTA builds its own copy (identical by construction).  Before the old CSR is overwritten it is used
// to flag the clusters whose membership changed.  Stable fill (members ascending): each warp owns a
// contiguous range of frames and ranks them with match.any on top of per-warp cluster counters; a single
// warp does the whole range when those counters would not fit.  Empty clusters take the next frame of the
// reseed pool in ascending cluster order (:118).
struct MemberSmem {
  const unsigned short* assign;  // [T] new assignments (k <= 4096: 16 bits each)
  int* offs;     // [k+1]
  int* memb;     // [T+k]
  int* cursor;   // [k+1] scratch
  int* cntw;     // [kSolverWarps][k] per-warp counters (multiwarp_fill) or [k]
  int* changed;  // [k]
  double* ct_new;
  const void* ts_smem;  // shared-memory copy of the timestamps (nullptr: read them from global memory)
};

template <typename TS>
__device__ void rebuild_members_smem(const SolverArgs& a, const MemberSmem& m, int* used_io, bool force_all) {
  const int T = a.T, k = a.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = a.multiwarp_fill ? kSolverWarps : 1;
  const RngKey rk = load_rng(a);
  // (1) the changed flags were set by the caller while it loaded the new assignments (a frame that moved flags both its old
  //     and its new cluster; the first iteration flags everything: the initial CSR lists only the k seed frames); zero the counters
  (void)force_all;
  for (int e = threadIdx.x; e < nw * k; e += blockDim.x) m.cntw[e] = 0;
  __syncthreads();
  // (2) per-warp, per-cluster counts over the warp's frame range
  const int range = ((T + nw - 1) / nw + 31) & ~31;
  const int r0 = warp * range, r1 = min(T, r0 + range);
  if (warp < nw) {
    int* cw = m.cntw + warp * k;
    for (int i0 = r0; i0 < r1; i0 += 32) {
      const int i = i0 + lane;
      const int c = i < r1 ? m.assign[i] : -1 - lane;
      const unsigned peers = __match_any_sync(0xffffffffu, c);
      if (i < r1 && (peers & ((1u << lane) - 1u)) == 0) cw[c] += __popc(peers);
      __syncwarp();
    }
  }
  __syncthreads();
  // (3) totals and exclusive per-warp bases
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    int run = 0;
    for (int w = 0; w < nw; ++w) {
      const int c = m.cntw[w * k + j];
      m.cntw[w * k + j] = run;
      run += c;
    }
    m.cursor[j] = run;  // member count of cluster j
  }
  __syncthreads();
  // (4) offsets + re-seeds: warp 0, each lane owns a contiguous slice of clusters (exclusive warp scans of the
  //     slot counts and of the number of empty clusters, so empties take pool entries in ascending cluster order)
  int used = *used_io;
  if (warp == 0) {
    const int cpl = (k + 31) / 32;
    const int j0 = min(k, lane * cpl), j1 = min(k, j0 + cpl);
    int slots = 0, empties = 0;
    for (int j = j0; j < j1; ++j) {
      const int c = m.cursor[j];
      slots += c > 0 ? c : 1;
      empties += c == 0;
    }
    int s_inc = slots, e_inc = empties;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int vs = __shfl_up_sync(0xffffffffu, s_inc, o), ve = __shfl_up_sync(0xffffffffu, e_inc, o);
      if (lane >= o) { s_inc += vs; e_inc += ve; }
    }
    int run = s_inc - slots, u = used + (e_inc - empties);
    for (int j = j0; j < j1; ++j) {
      const int c = m.cursor[j];
      m.offs[j] = run;
      if (c == 0) {
        bool ovf = false;
        const int r = reseed_at(a, rk, u, &ovf);
        if (ovf) a.status->reseed_overflow = 1;
        ++u;
        m.memb[run] = r;
        m.changed[j] = 1;  // the centre jumps to a freshly drawn frame
        run += 1;
      } else {
        run += c;
      }
    }
    if (lane == 31) {
      m.offs[k] = s_inc;
      m.cursor[k] = used + e_inc;
    }
  }
  __syncthreads();
  used = m.cursor[k];
  // (5) stable fill
  if (warp < nw) {
    int* cw = m.cntw + warp * k;
    for (int i0 = r0; i0 < r1; i0 += 32) {
      const int i = i0 + lane;
      const int c = i < r1 ? m.assign[i] : -1 - lane;
      const unsigned peers = __match_any_sync(0xffffffffu, c);
      if (i < r1) {
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int base = cw[c];
        m.memb[m.offs[c] + base + rank] = i;
        __syncwarp(peers);
        if (rank == 0) cw[c] = base + __popc(peers);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // (6) centre times (:115 mean of member timestamps ascending; :120 a seed alone gives ts[r])
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int beg = m.offs[j], end = m.offs[j + 1];
    TS tsum = TS(0);
    const void* tsrc = m.ts_smem ? m.ts_smem : a.ts;  // k threads x ~T/k dependent reads: shared memory when it fits
    for (int e = beg; e < end; e += 8) {  // 8 member timestamps in flight (global memory when the shared copy does not fit), added in member order
      TS v[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) v[w] = e + w < end ? load_ts<TS>(tsrc, m.memb[e + w]) : TS(0);
#pragma unroll
      for (int w = 0; w < 8; ++w)
        if (e + w < end) tsum += v[w];
    }
    m.ct_new[j] = (double)(TS)(tsum / (TS)(end - beg));
  }
  __syncthreads();
  *used_io = used;
}

// S[i][j] = sum_{m in M_j} G[i][m] for the listed columns j (all CTAs; column fastest so the stores coalesce).
// A thread works on kSE entries at once, 4 Gram loads each in flight (the loop is pure L2 / HBM latency: member addresses come
// from shared memory, 8 warps per SM); every entry is still summed in ascending member order.
constexpr int kSE = 4;
// Large problems: the same sums read through the symmetry of G, S[i][j] = sum_{m in M_j} G[m][i].  A thread owns frame i, a
// CTA a slab of 256 consecutive frames and a group of 32 listed columns: every load instruction of a warp reads 256
// contiguous bytes of Gram row m (the entry-per-thread gather below touches one 8-byte value per 32-byte sector once G no
// longer fits in L2: 1.66 GB at T = 14,400).  Same values added in the same ascending member order -> the same bits.
constexpr int kSRowGroup = 32;
constexpr int kSPairBatch = 8;  // member rows in flight per column and thread in the two-frames-per-thread form (x 4 columns x 16 bytes)
constexpr int64_t kSRowMinEntries = 1 << 20;
// NC columns x NM members in flight per thread (32 loads): 4 x 8 for ordinary member lists (T / k ~ 15 frames: two rounds per
// 4 columns), 16 x 2 when every list of the group is short (the initial pass: one seed frame per cluster).
template <int NC, int NM>
__device__ __forceinline__ void s_rows_columns(const SolverArgs& a, const int* s_offs, const int* s_memb, const int* cols, int c0, int c1,
                                               const double* gi, double* srow) {
  for (int cc = c0; cc < c1; cc += NC) {
    int beg[NC], len[NC], col[NC];
    double s[NC];
    int maxlen = 0;
#pragma unroll
    for (int u = 0; u < NC; ++u) {
      s[u] = 0.0; len[u] = 0; beg[u] = 0; col[u] = -1;
      if (cc + u < c1) {
        col[u] = cols[cc + u];
        beg[u] = s_offs[col[u]];
        len[u] = s_offs[col[u] + 1] - beg[u];
        maxlen = max(maxlen, len[u]);
      }
    }
    for (int e = 0; e < maxlen; e += NM) {
      double v[NC][NM];
#pragma unroll
      for (int u = 0; u < NC; ++u)
#pragma unroll
        for (int w = 0; w < NM; ++w) v[u][w] = (e + w < len[u]) ? __ldg(gi + (int64_t)s_memb[beg[u] + e + w] * a.ldg) : 0.0;
#pragma unroll
      for (int u = 0; u < NC; ++u)
#pragma unroll
        for (int w = 0; w < NM; ++w)
          if (e + w < len[u]) s[u] += v[u][w];  // ascending member order, whatever the batch size
    }
#pragma unroll
    for (int u = 0; u < NC; ++u)
      if (col[u] >= 0) srow[col[u]] = s[u];
  }
}

// Two adjacent frames per thread (one 16-byte load per member row) and 32-bit row offsets: the scalar form above spends ~15
// integer instructions per 8-byte load (64-bit row offset, leading dimension re-read from the constant bank) and was
// instruction-issue-bound (ncu: 59 % of the solver's instructions at BASELINE config 4), not memory-bound.  Needs an even
// leading dimension, a 16-byte aligned G, even slice bounds and (T - 1) * ldg < 2^32; same values added in the same order.
template <int NC, int NM>
__device__ __forceinline__ void s_rows_columns2(const int* s_offs, const int* s_memb, const int* cols, int c0, int c1, const double* gi,
                                                unsigned ldg, double* srow, int k) {
  for (int cc = c0; cc < c1; cc += NC) {
    int beg[NC], len[NC], col[NC];
    double2 s[NC];
    int maxlen = 0;
#pragma unroll
    for (int u = 0; u < NC; ++u) {
      s[u] = make_double2(0.0, 0.0); len[u] = 0; beg[u] = 0; col[u] = -1;
      if (cc + u < c1) {
        col[u] = cols[cc + u];
        beg[u] = s_offs[col[u]];
        len[u] = s_offs[col[u] + 1] - beg[u];
        maxlen = max(maxlen, len[u]);
      }
    }
    for (int e = 0; e < maxlen; e += NM) {
      double2 v[NC][NM];
#pragma unroll
      for (int u = 0; u < NC; ++u)
#pragma unroll
        for (int w = 0; w < NM; ++w) {
          v[u][w] = make_double2(0.0, 0.0);
          if (e + w < len[u]) {
            const unsigned off = (unsigned)s_memb[beg[u] + e + w] * ldg;
            v[u][w] = __ldg(reinterpret_cast<const double2*>(gi + off));
          }
        }
#pragma unroll
      for (int u = 0; u < NC; ++u)
#pragma unroll
        for (int w = 0; w < NM; ++w)
          if (e + w < len[u]) { s[u].x += v[u][w].x; s[u].y += v[u][w].y; }
    }
#pragma unroll
    for (int u = 0; u < NC; ++u)
      if (col[u] >= 0) { srow[col[u]] = s[u].x; srow[k + col[u]] = s[u].y; }
  }
}

// Work items (a slab of 32 or 64 frames x a group of listed columns) differ in cost with the cluster sizes, so every WARP draws
// its items from a ticket counter instead of a static round-robin (the slowest CTA used to finish 40-200 us after the first, and
// a CTA-wide item made seven warps wait for the eighth).  The counter is never reset: a pass consumes n_items + (warps in the
// grid) tickets (every warp draws exactly one ticket past the end), so everybody knows the base of the next pass.  The next
// ticket is drawn before the current item is processed.
template <bool DEEP>  // DEEP: the 255-register wide solver; else 16 loads in flight (the 128-register kernels must not spill)
__device__ void compute_S_rows(const SolverArgs& a, const int* s_offs, const int* s_memb, const int* cols, int ncols, double* S, int f_begin,
                               int f_end, unsigned& ticket_base) {
  const int k = a.k;
  const int lane = threadIdx.x & 31;
  const bool pairs = DEEP && !a.s_no_pairs && (a.ldg & 1) == 0 && (reinterpret_cast<uintptr_t>(a.G) & 15u) == 0 && ((f_begin | f_end) & 1) == 0 &&
                     (int64_t)(a.T - 1) * a.ldg < (1ll << 32);
  const int fpt = pairs ? 2 : 1;                           // frames per thread
  const int group = pairs ? kSRowGroup / 2 : kSRowGroup;   // listed columns per work item (items of about the same size either way)
  const int slab_frames = 32 * fpt;
  const int n_slabs = (f_end - f_begin + slab_frames - 1) / slab_frames;
  const int n_groups = (ncols + group - 1) / group;
  const int n_items = n_slabs * n_groups;
  unsigned* ticket = a.barrier + 1;
  unsigned next = 0;
  if (lane == 0) next = atomicAdd(ticket, 1u) - ticket_base;
  for (;;) {
    const int item = (int)__shfl_sync(0xffffffffu, next, 0);
    if (item >= n_items) break;
    if (lane == 0) next = atomicAdd(ticket, 1u) - ticket_base;  // in flight while this item is processed
    // consecutive tickets = consecutive slabs of one column group: the warps working side by side share the group's member rows
    const int slab = item % n_slabs, c0 = (item / n_slabs) * group;
    const int i = f_begin + slab * slab_frames + lane * fpt;
    if (i < f_end) {  // (lanes past the slice end idle; they rejoin at the shuffle above)
    const int c1 = min(ncols, c0 + group);
    if (DEEP) {
      int gmax = 0;
      for (int cc = c0; cc < c1; ++cc) gmax = max(gmax, s_offs[cols[cc] + 1] - s_offs[cols[cc]]);
      if (pairs) {
        if (gmax <= 2) s_rows_columns2<16, 1>(s_offs, s_memb, cols, c0, c1, a.G + i, (unsigned)a.ldg, S + (int64_t)i * k, k);
        else s_rows_columns2<4, kSPairBatch>(s_offs, s_memb, cols, c0, c1, a.G + i, (unsigned)a.ldg, S + (int64_t)i * k, k);
      } else {
        if (gmax <= 2) s_rows_columns<16, 2>(a, s_offs, s_memb, cols, c0, c1, a.G + i, S + (int64_t)i * k);
        else s_rows_columns<4, 8>(a, s_offs, s_memb, cols, c0, c1, a.G + i, S + (int64_t)i * k);
      }
    } else {
      s_rows_columns<4, 4>(a, s_offs, s_memb, cols, c0, c1, a.G + i, S + (int64_t)i * k);
    }
    }
  }
  __syncwarp();
  ticket_base += (unsigned)n_items + gridDim.x * (blockDim.x >> 5);
}

// Frames [f_begin, f_end) only: all of them in the single-GPU solver, a rank's own frames in the frame-sharded loop.
template <bool DEEP>
__device__ void compute_S(const SolverArgs& a, const int* s_offs, const int* s_memb, const int* cols, int ncols, double* S, int f_begin,
                          int f_end, unsigned& ticket_base) {
  const int k = a.k;
  const int64_t total = (int64_t)(f_end - f_begin) * ncols;
  if (total >= kSRowMinEntries) {
    compute_S_rows<DEEP>(a, s_offs, s_memb, cols, ncols, S, f_begin, f_end, ticket_base);
    return;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (total <= 2 * stride) {
    // few entries per thread (short streams, or only a handful of changed clusters): one entry at a time with 16 member loads in
    // flight -- an event of ~15 frames is ONE L2 round trip instead of four.  Same ascending member order, same bits.
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
      const int il = (int)(idx / ncols), i = f_begin + il;
      const int j = cols[(int)(idx - (int64_t)il * ncols)];
      const int beg = s_offs[j], len = s_offs[j + 1] - beg;
      const double* g = a.G + (int64_t)i * a.ldg;
      double sum = 0.0;
      for (int e = 0; e < len; e += 16) {
        double v[16];
#pragma unroll
        for (int w = 0; w < 16; ++w) v[w] = (e + w < len) ? __ldg(g + s_memb[beg + e + w]) : 0.0;
#pragma unroll
        for (int w = 0; w < 16; ++w)
          if (e + w < len) sum += v[w];
      }
      S[(int64_t)i * k + j] = sum;
    }
    return;
  }
  for (int64_t idx0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx0 < total; idx0 += kSE * stride) {
    const double* g[kSE];
    int beg[kSE], len[kSE];
    int64_t out[kSE];
    double s[kSE];
    int maxlen = 0;
#pragma unroll
    for (int u = 0; u < kSE; ++u) {
      const int64_t idx = idx0 + u * stride;
      s[u] = 0.0;
      len[u] = 0; beg[u] = 0; g[u] = a.G; out[u] = -1;
      if (idx < total) {
        const int il = (int)(idx / ncols), i = f_begin + il;
        const int j = cols[(int)(idx - (int64_t)il * ncols)];
        beg[u] = s_offs[j];
        len[u] = s_offs[j + 1] - beg[u];
        g[u] = a.G + (int64_t)i * a.ldg;
        out[u] = (int64_t)i * k + j;
        maxlen = max(maxlen, len[u]);
      }
    }
    for (int e = 0; e < maxlen; e += 4) {
      double v[kSE][4];
#pragma unroll
      for (int u = 0; u < kSE; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) v[u][w] = (e + w < len[u]) ? __ldg(g[u] + s_memb[beg[u] + e + w]) : 0.0;
#pragma unroll
      for (int u = 0; u < kSE; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w)
          if (e + w < len[u]) s[u] += v[u][w];
    }
#pragma unroll
    for (int u = 0; u < kSE; ++u)
      if (out[u] >= 0) S[out[u]] = s[u];
  }
}

// ---- k-means++ as its own single-CTA launch: while the 39 dependent sampling rounds run, the other SMs stay free for
// other streams' kernels (inside the cooperative kernel they would sit spinning at the grid barrier for >100 us) ----
template <typename TS, int kInitThreads>
__global__ void __launch_bounds__(kInitThreads) kmeanspp_kernel(SolverArgs a) {
  extern __shared__ double smem_init[];
  const int T = a.T;
  double* s_diag = smem_init;                                                   // [T] if diag_in_smem
  float* s_near = reinterpret_cast<float*>(smem_init + (a.diag_in_smem ? T : 0));  // [T]
  __shared__ PPScratch s_pp;
  csb_kmeans_status* st = a.status;
  if (threadIdx.x == 0) {
    st->iterations = 0; st->converged = 0; st->n_reseed_used = 0; st->degenerate_rounds = 0;
    st->reseed_overflow = 0; st->n_noise_used = 0; st->min_margin = __int_as_float(0x7f800000); st->last_shift = 0.f;
  }
  mark(a, 1);
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const double g = a.G[(int64_t)i * a.ldg + i];
    a.diag[i] = g;
    if (a.diag_in_smem) s_diag[i] = g;
  }
  __syncthreads();
  int used = 0, n_noise = 0, n_deg = 0;
  if (a.n_init_given < a.k)
    kmeanspp_block<kInitThreads>(a, s_near, a.diag_in_smem ? s_diag : a.diag, &s_pp, &used, &n_noise, &n_deg);
  if (threadIdx.x == 0) {
    st->degenerate_rounds = n_deg;
    st->n_noise_used = n_noise;
    *a.used = used;  // pool entries consumed by degenerate rounds
  }
  mark(a, 2);
}

// ---- k-means++ over many CTAs (long streams: T > kGridInitMinT) ---------------------------------------------------------
// The single-CTA kernel above spends ~17 us per round at T = 14,400 (one SM pulls a 115 KB Gram row and a 58 KB noise row
// and reduces 14,400 candidates), 16.5 ms for the 959 rounds of BASELINE config 4.  Here every thread owns <= kGridInitU
// frames (nearest distance and diagonal in registers) and a round costs ONE grid barrier:
//   update with row c, per-CTA partial of sum nearest^2, per-CTA CANDIDATES for the next centre  -> grid barrier ->
//   total (fixed tree, identical in every CTA), exact scores of the candidates, global argmax (score, then lowest index).
// The exact score is RN(RN(n2 / sum) / noise) and needs the total; candidates are chosen from u = RN(n2 / noise), which
// orders the frames like the exact score up to two roundings: the true winner i* of a CTA satisfies
// u(i*) >= u_max * (1 - 2^-21), so keeping every frame with u >= u_max * (1 - 2^-20) cannot lose it.  Almost always that is
// one frame; more than kGridInitCand of them (or the force flag) falls back to an exact second pass + a second barrier.
// Same arithmetic per frame as kmeanspp_block; only the (deterministic) order of the fp32 sum differs, which is a common
// factor of all scores.  Cooperative launch (co-residency for the software barrier).
constexpr int kGridInitThreads = 512;
constexpr int kGridInitU = 4;
constexpr int kGridInitMinT = 1024;  // above the fused init of short streams: the cluster kernel beats the single-CTA kernel at every size (profiles/r2_init_midsize_probe.jsonl)
constexpr int kGridInitCand = 4;

struct GridInitBufs {
  unsigned* barrier;  // zeroed before launch
  float* part;        // [2][nb]
  int* cnt;           // [2][nb] candidates published by the CTA (> kGridInitCand: overflow)
  float* cand_n2;     // [2][nb][kGridInitCand] nearest^2
  float* cand_nz;     // [2][nb][kGridInitCand] noise
  int* cand_i;        // [2][nb][kGridInitCand] frame
  float* best_v;      // [nb] exact fallback: per-CTA best score / frame
  int* best_i;
  int force_exact;    // profiling / test aid: always take the two-barrier exact path
  int n_cluster;      // CL: CTAs of the sampling cluster (= the first cluster of the grid); the other clusters are helpers
  int push;           // CL: round records pushed into the peers' buffers (st.async + mbarrier) instead of barrier + remote loads
  int n_helpers;      // CL: helper CTAs that pre-generate the built-in generator's noise rows into a.noise_ws (0: none)
  int profile;        // profiling aid (CSB_INIT_PROFILE=1): thread 0 of CTA 0 sums the clock cycles of the five phases of a
                      // round and stores the per-round averages in part[0..5] (+ the round count) when the kernel ends
};

// CL = true (the default for T <= 32,768): the sampling CTAs are ONE thread-block cluster (<= 16 CTAs).  The per-round exchange goes
// through distributed shared memory instead of global memory: by default every CTA pushes its round record into the peers'
// buffers (st.async + mbarrier, see below); with CSB_INIT_PUSH=0 -- and always in the rare exact-score fallback -- the software
// grid barrier (atomic + acquire poll through the L2, ~1.4 us, plus an L2 round trip for the partials afterwards) becomes the
// hardware cluster barrier (barrier.cluster arrive.release / wait.acquire) followed by four 16-byte ld.shared::cluster loads per
// peer.  Same arithmetic, same candidate rule, same fixed summation tree: the chosen centres are those of the grid form.
// Further clusters of the same launch (blockIdx.x >= n_cluster) are helpers: they pre-generate the built-in generator's noise.
constexpr int kClusterInitMax = 16;
constexpr bool kClusterInitDefault = true;   // CSB_INIT_CLUSTER=0: the cooperative grid with the software barrier (validation)
struct __align__(16) ClusterRound {  // what a CTA publishes per sampling round: 64 bytes = four 16-byte loads for a peer
  float part;
  int cnt;
  int pad[2];
  float n2[kGridInitCand];
  float nz[kGridInitCand];
  int idx[kGridInitCand];
};
static_assert(kGridInitCand == 4 && sizeof(ClusterRound) == 64, "ClusterRound is read with four 16-byte loads");

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint2 ld_cluster_v2(const void* local_smem, unsigned rank) {
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(local_smem), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
  uint2 v;
  asm volatile("ld.shared::cluster.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(remote) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_cluster_v4(const void* local_smem, unsigned rank) {
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(local_smem), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
  uint4 v;
  asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(remote) : "memory");
  return v;
}

// Push exchange (the default): instead of publishing its round record locally and meeting the peers at the cluster barrier, a CTA
// writes the record straight into every peer's receive buffer with st.async (16-byte stores from registers that count their
// bytes on the peer's mbarrier); a peer waits on its own mbarrier (hardware sleep, wake-up on the last byte) and reads local
// shared memory.  One one-way trip through the cluster network per round instead of barrier arrive + wait + remote loads.
// Buffers and mbarriers alternate with the round parity: a CTA can be at most one round ahead of a peer.
__device__ __forceinline__ void st_async_v4(const void* local_dst, const void* local_mbar, unsigned rank, uint4 v) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(local_dst), mb = (uint32_t)__cvta_generic_to_shared(local_mbar), rd, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rd) : "r"(d), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(mb), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rd), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w), "r"(rb)
               : "memory");
}
__device__ __forceinline__ void mbar_init_local(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_local(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_local(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(parity)
      : "memory");
}

template <bool CL, int NT, int U>
__global__ void __launch_bounds__(NT, 1) kmeanspp_grid_kernel(SolverArgs a, GridInitBufs b) {
  constexpr int NW = NT / 32;
  __shared__ float s_wv[NW];
  __shared__ int s_wi[NW];
  __shared__ float s_ws[NW];
  __shared__ float s_cn2[kGridInitCand], s_cnz[kGridInitCand];
  __shared__ int s_ci[kGridInitCand];
  __shared__ int s_cnt, s_c[2], s_ovf[2];
  __shared__ float s_sum[2], s_umax;
  __shared__ ClusterRound s_x[2];  // CL: this CTA's published round data, double-buffered by round parity
  __shared__ ClusterRound s_in[2][CL ? kClusterInitMax : 1];  // CL, push exchange: the peers' round records (by round parity)
  __shared__ __align__(8) unsigned long long s_mbar[2];         // CL, push exchange: bytes of the peers' records, by round parity
  __shared__ __align__(8) unsigned s_fb[2];  // CL, exact fallback: this CTA's best {score bits, frame} (written / read between the two barriers of a round)
  const int T = a.T, k = a.k;
  const unsigned nb = CL ? (unsigned)b.n_cluster : gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  csb_kmeans_status* st = a.status;
  if (CL && blockIdx.x >= nb) {
    // ---- helper clusters (built-in generator): all Exp(1) noise rows, 4 frames per Philox call, into the workspace; the 132
    // SMs the sampling cluster leaves idle finish this in a few microseconds, and the sampling rounds then read a noise row like
    // a caller-supplied one instead of running ~300 generator instructions per frame and round on 16 SMs.  Same values as
    // noise_generate().  Arrival is counted on the grid-init barrier word; the sampling cluster waits for all helpers first.
    const RngKey hk = load_rng(a);
    const int q4 = (T + 3) >> 2;  // calls per row
    const int64_t total = (int64_t)(k - 1) * q4;
    for (int64_t e = (int64_t)(blockIdx.x - nb) * NT + threadIdx.x; e < total; e += (int64_t)b.n_helpers * NT) {
      const int row = (int)(e / q4), cc = (int)(e - (int64_t)row * q4);
      const uint4 x = philox_draw4(hk.seed, hk.offset, (uint32_t)row, (uint32_t)cc);
      float* dst = a.noise_ws + (int64_t)row * T + 4 * cc;
      const float v[4] = {exp1_from_bits(x.x), exp1_from_bits(x.y), exp1_from_bits(x.z), exp1_from_bits(x.w)};
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (4 * cc + u < T) __stcg(dst + u, v[u]);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(b.barrier, 1u);
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->iterations = 0; st->converged = 0; st->n_reseed_used = 0; st->degenerate_rounds = 0;
    st->reseed_overflow = 0; st->n_noise_used = 0; st->min_margin = __int_as_float(0x7f800000); st->last_shift = 0.f;
  }
  const bool push = CL && b.push;
  if (push) {
    if (threadIdx.x == 0) {
      mbar_init_local(&s_mbar[0], 1);
      mbar_init_local(&s_mbar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_barrier();  // every CTA's mbarriers exist before any peer's bytes can arrive
  }
  unsigned n_push0 = 0u, n_push1 = 0u;  // completed uses of each mbarrier (its phase parity)
  mark(a, 1);
  // frames of this thread: base + u * NT, u < U, inside the CTA's contiguous slice
  const int per_cta = (T + (int)nb - 1) / (int)nb;
  const int cta0 = blockIdx.x * per_cta, cta1 = min(T, cta0 + per_cta);
  int fi[U];
  float near[U];    // nearest distance squared (fp32, as the reference computes it)
  double near2[U];  // smallest clamped fp64 squared distance so far
  double dg[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = cta0 + threadIdx.x + u * NT;
    fi[u] = i < cta1 ? i : -1;
    near[u] = __int_as_float(0x7f800000);
    near2[u] = __longlong_as_double(0x7ff0000000000000ll);
    dg[u] = 0.0;
    if (fi[u] >= 0) {
      dg[u] = a.G[(int64_t)i * a.ldg + i];
      a.diag[i] = dg[u];
    }
  }
  unsigned bar = 0;
  int noise_row = 0, degenerate = 0, used = 0;
  const RngKey rk = load_rng(a);
  const bool prof = CL && b.profile && blockIdx.x == 0 && threadIdx.x == 0;  // (the grid form has no registers to spare)
  __shared__ long long s_tacc[7];  // [0..4] phase sums, [5] previous mark, [6] rounds (shared memory: the 64-register variant has none to spare)
  if (prof)
    for (int i = 0; i < 7; ++i) s_tacc[i] = 0;
  auto lap = [&](int i) {
    if (prof) { const long long t = clock64(); s_tacc[i] += t - s_tacc[5]; s_tacc[5] = t; }
  };
  int c = 0;
  if (a.n_init_given == 0) {  // :41 the first centre, from the built-in generator
    c = (int)(((unsigned long long)philox_draw(rk.seed, rk.offset, kRngFirst, 0u) * (unsigned)T) >> 32);
    if (blockIdx.x == 0 && threadIdx.x == 0) a.init_indices[0] = c;
  }
  const bool noise_pre = CL && b.n_helpers > 0;  // the helper clusters' rows (else: the caller's array, or generated on the spot)
  if (noise_pre) {
    if (threadIdx.x == 0) {
      unsigned v;
      for (;;) {
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(b.barrier) : "memory");
        if (v >= (unsigned)b.n_helpers) break;
        __nanosleep(200);
      }
    }
    __syncthreads();
  }
  for (int m = 0; m < k; ++m) {
    if (m < a.n_init_given) c = (int)a.init_indices[m];  // else: c was chosen at the end of the previous iteration
    if (m + 1 >= k) break;
    // ---- update the nearest distances with centre c; per-CTA partial of sum nearest^2 (:45-53) ----
    const bool sample_next = m + 1 >= a.n_init_given;
    if (prof) { s_tacc[5] = clock64(); ++s_tacc[6]; }
    if (push && sample_next && threadIdx.x == 0) mbar_expect_tx_local(&s_mbar[m & 1], nb * (unsigned)sizeof(ClusterRound));
    // (noise row `noise_row` = the one the next round consumes if it samples)
    const double gcc = __ldg(a.G + (int64_t)c * a.ldg + c);
    const double* gc = a.G + (int64_t)c * a.ldg;
    double g[U];
    float nz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      g[u] = fi[u] >= 0 ? __ldg(gc + fi[u]) : 0.0;
      nz[u] = (sample_next && fi[u] >= 0) ? (a.exp_noise ? noise_at<kNoiseArray>(a, rk, noise_row, fi[u])
                                                          : (noise_pre ? noise_at<kNoiseWs>(a, rk, noise_row, fi[u]) : noise_at<kNoiseGen>(a, rk, noise_row, fi[u]))) : 1.f;
    }
    float part = 0.f, n2[U], uq[U], umax = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      n2[u] = 0.f; uq[u] = 0.f;
      if (fi[u] >= 0) {
        // nearest distance = min over the centres of (float)sqrt(max(d2, 0)) = the image of the smallest clamped d2 (clamp, square
        // root and rounding are monotone): the fp64 square root (13 cycles per warp on the SM) runs only when a frame gets a
        // nearer centre -- for whole warps of one scene at a time, and ever more rarely as the rounds go on
        const double d2 = dg[u] + gcc - 2.0 * g[u];
        const double d2c = d2 > 0.0 ? d2 : 0.0;
        if (d2c < near2[u]) {
          near2[u] = d2c;
          const float d = (float)sqrt(d2c);
          near[u] = __fmul_rn(d, d);  // nearest^2, :53
        }
        n2[u] = near[u];
        part += n2[u];
        uq[u] = n2[u] == 0.f ? 0.f : __fdiv_rn(n2[u] == 0.f ? 1.f : n2[u], nz[u]);  // (zero numerators: see kmeanspp_block)
        umax = fmaxf(umax, uq[u]);
      }
    }
    if (!sample_next) continue;  // the next centre is prescribed too: no total needed
    lap(0);  // Gram row + noise loaded, distances updated
    part = warp_sum(part);
    // the scores are non-negative floats: their order is the order of their bit patterns, so the maxima are single REDUX ops
    if (CL) umax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(umax)));
    else umax = warp_max(umax);
    if (CL) {
      // One block barrier, then the cluster barrier: the per-warp partials are combined by warp 0 (fixed shuffle tree), every
      // thread takes the block maximum straight from the per-warp values, and the candidates go directly into the round buffer
      // the peers read.  Buffer m & 1 was last read by the peers in round m - 2, i.e. before the barrier of round m - 1 that
      // this CTA has passed; the per-warp arrays are rewritten only after the cluster barrier every reader arrives at.
      ClusterRound* x = &s_x[m & 1];
      if (lane == 0) { s_ws[warp] = part; s_wv[warp] = umax; }
      if (threadIdx.x == 0) x->cnt = 0;
      __syncthreads();
      const float um = __uint_as_float(__reduce_max_sync(0xffffffffu, lane < NW ? __float_as_uint(s_wv[lane]) : 0u));
      if (warp == 0) {
        const float t = warp_sum(lane < NW ? s_ws[lane] : 0.f);
        if (lane == 0) x->part = t;
      }
      const float thr = um - um * 9.5367431640625e-7f;  // u_max * (1 - 2^-20)
      if (um > 0.f) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (fi[u] >= 0 && uq[u] >= thr) {
            const int slot = atomicAdd(&x->cnt, 1);
            if (slot < kGridInitCand) { x->n2[slot] = n2[u]; x->nz[slot] = nz[u]; x->idx[slot] = fi[u]; }
          }
        }
      }
      lap(1);  // block reduction, candidates published
      if (push) {
        __syncthreads();  // the record is complete
        if (warp == 0 && lane < (int)nb) {
          const uint4* rec = reinterpret_cast<const uint4*>(x);
          const uint4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
          char* dst = reinterpret_cast<char*>(&s_in[m & 1][blockIdx.x]);  // this CTA's slot in the peer's buffer (rank = blockIdx.x)
          st_async_v4(dst, &s_mbar[m & 1], (unsigned)lane, r0);
          st_async_v4(dst + 16, &s_mbar[m & 1], (unsigned)lane, r1);
          st_async_v4(dst + 32, &s_mbar[m & 1], (unsigned)lane, r2);
          st_async_v4(dst + 48, &s_mbar[m & 1], (unsigned)lane, r3);
        }
      } else {
        cluster_barrier();  // every thread of every CTA: releases the writes above, acquires the peers'
      }
    } else {
      if (lane == 0) { s_ws[warp] = part; s_wv[warp] = umax; }
      if (threadIdx.x == 0) { s_cnt = 0; }
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = s_ws[0], um = s_wv[0];
        for (int w = 1; w < NW; ++w) { t += s_ws[w]; um = fmaxf(um, s_wv[w]); }
        b.part[(m & 1) * nb + blockIdx.x] = t;
        s_umax = um;
      }
      __syncthreads();
      {
        const float um = s_umax;
        const float thr = um - um * 9.5367431640625e-7f;  // u_max * (1 - 2^-20)
        if (um > 0.f) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (fi[u] >= 0 && uq[u] >= thr) {
              const int slot = atomicAdd(&s_cnt, 1);
              if (slot < kGridInitCand) { s_cn2[slot] = n2[u]; s_cnz[slot] = nz[u]; s_ci[slot] = fi[u]; }
            }
          }
        }
      }
      __syncthreads();
      const int cnt = s_cnt;
      if (threadIdx.x == 0) b.cnt[(m & 1) * nb + blockIdx.x] = b.force_exact ? kGridInitCand + 1 : cnt;
      if (threadIdx.x < min(cnt, kGridInitCand)) {
        const int64_t o = ((int64_t)(m & 1) * nb + blockIdx.x) * kGridInitCand + threadIdx.x;
        b.cand_n2[o] = s_cn2[threadIdx.x]; b.cand_nz[o] = s_cnz[threadIdx.x]; b.cand_i[o] = s_ci[threadIdx.x];
      }
      lap(1);
      grid_sync(b.barrier, bar, nb);
    }
    lap(2);  // barrier
    const int rb = CL ? (m & 1) : 0;
    // ---- total (fixed shuffle tree: the same bits in every CTA), overflow flag, candidates ----
    // lane q takes CTA q (q += 32): its partial, count and all candidate slots are loaded together -- one L2 round trip
    if (warp == 0) {
      float v = 0.f;
      int ovf = 0;
      float gb = -1.f;
      int gi = 0x7fffffff;
      constexpr int kMaxQ = 5;  // ceil(148 / 32) CTAs per lane at most
      float cn2[kMaxQ][kGridInitCand], cnz[kMaxQ][kGridInitCand];
      int cix[kMaxQ][kGridInitCand], ccnt[kMaxQ];
#pragma unroll
      for (int r = 0; r < kMaxQ; ++r) {
        const unsigned q = lane + 32 * r;
        ccnt[r] = 0;
        if (CL) {
          if (r == 0 && push) mbar_wait_local(&s_mbar[m & 1], ((m & 1) ? n_push1 : n_push0) & 1u);  // (warp 0, all lanes: the peers' records have landed)
          if (r == 0 && q < nb) {  // a cluster has <= 16 CTAs: lane q takes peer q's 64 bytes
            uint4 h, v2, vz, vi;
            if (push) {
              const uint4* rec = reinterpret_cast<const uint4*>(&s_in[m & 1][q]);
              h = rec[0]; v2 = rec[1]; vz = rec[2]; vi = rec[3];
            } else {  // remote loads after the cluster barrier (all four in flight together)
              const char* xb = reinterpret_cast<const char*>(&s_x[m & 1]);
              h = ld_cluster_v4(xb, q); v2 = ld_cluster_v4(xb + 16, q); vz = ld_cluster_v4(xb + 32, q); vi = ld_cluster_v4(xb + 48, q);
            }
            v += __uint_as_float(h.x);
            ccnt[r] = (int)h.y;
            cn2[r][0] = __uint_as_float(v2.x); cn2[r][1] = __uint_as_float(v2.y); cn2[r][2] = __uint_as_float(v2.z); cn2[r][3] = __uint_as_float(v2.w);
            cnz[r][0] = __uint_as_float(vz.x); cnz[r][1] = __uint_as_float(vz.y); cnz[r][2] = __uint_as_float(vz.z); cnz[r][3] = __uint_as_float(vz.w);
            cix[r][0] = (int)vi.x; cix[r][1] = (int)vi.y; cix[r][2] = (int)vi.z; cix[r][3] = (int)vi.w;
          }
        } else if (q < nb) {
          v += __ldcg(b.part + (m & 1) * nb + q);
          ccnt[r] = __ldcg(b.cnt + (m & 1) * nb + q);
          const int64_t o = ((int64_t)(m & 1) * nb + q) * kGridInitCand;
#pragma unroll
          for (int sl = 0; sl < kGridInitCand; ++sl) {
            cn2[r][sl] = __ldcg(b.cand_n2 + o + sl); cnz[r][sl] = __ldcg(b.cand_nz + o + sl); cix[r][sl] = __ldcg(b.cand_i + o + sl);
          }
        }
        ovf |= ccnt[r] > kGridInitCand;
      }
      if (CL) ovf |= b.force_exact;
      v = warp_sum(v);
      ovf = __any_sync(0xffffffffu, ovf);
      if (v != 0.f && !ovf) {
#pragma unroll
        for (int r = 0; r < kMaxQ; ++r)
#pragma unroll
          for (int sl = 0; sl < kGridInitCand; ++sl)
            if (sl < ccnt[r]) {
              const float p = __fdiv_rn(cn2[r][sl], v);       // :59
              const float sc = __fdiv_rn(p, cnz[r][sl]);      // multinomial: argmax(p / Exp(1)), first max
              const int oi = cix[r][sl];
              if (sc > gb || (sc == gb && oi < gi)) { gb = sc; gi = oi; }
            }
        if (CL) {
          // (largest score, then lowest frame) with two REDUX ops: scores are >= +0, a lane without a candidate has key 0
          const unsigned key = gb < 0.f ? 0u : __float_as_uint(gb) + 1u;
          const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
          gi = __reduce_min_sync(0xffffffffu, (key == kmax && kmax != 0u) ? gi : 0x7fffffff);
        } else {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, gb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, gi, o);
            if (ov > gb || (ov == gb && oi < gi)) { gb = ov; gi = oi; }
          }
        }
      }
      if (lane == 0) { s_sum[rb] = v; s_ovf[rb] = ovf; s_c[rb] = gi; }
    }
    if (m & 1) ++n_push1; else ++n_push0;
    __syncthreads();
    lap(3);  // peers read, exact scores, winner broadcast
    const float sum = s_sum[rb];
    if (sum == 0.f) {  // :55-57 every frame coincides with a centre: host-drawn index, no noise row consumed
      ++degenerate;
      bool ovf = false;
      c = reseed_at(a, rk, used, &ovf);
      if (ovf && blockIdx.x == 0 && threadIdx.x == 0) st->reseed_overflow = 1;
      ++used;
    } else {
      ++noise_row;
      if (!s_ovf[rb]) {
        c = s_c[rb];
      } else {
        // ---- exact fallback: scores of all frames with the known total, per-CTA argmax, second barrier ----
        float best = -1.f;
        int besti = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (fi[u] >= 0) {
            const float sc = __fdiv_rn(__fdiv_rn(n2[u], sum), nz[u]);
            if (sc > best) { best = sc; besti = fi[u]; }  // fi ascending in u: first max kept
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
          if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        if (lane == 0) { s_wv[warp] = best; s_wi[warp] = besti; }
        __syncthreads();
        if (threadIdx.x == 0) {
          best = s_wv[0]; besti = s_wi[0];
          for (int w = 1; w < NW; ++w) {
            const float ov = s_wv[w];
            const int oi = s_wi[w];
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
          }
          if (CL) { s_fb[0] = __float_as_uint(best); s_fb[1] = (unsigned)besti; }
          else { b.best_v[blockIdx.x] = best; b.best_i[blockIdx.x] = besti; }
        }
        if (CL) cluster_barrier();
        else grid_sync(b.barrier, bar, nb);
        if (warp == 0) {
          float gb = -1.f;
          int gi = 0x7fffffff;
          for (unsigned q = lane; q < nb; q += 32) {
            float ov;
            int oi;
            if (CL) {
              const uint2 h = ld_cluster_v2(s_fb, q);
              ov = __uint_as_float(h.x); oi = (int)h.y;
            } else {
              ov = __ldcg(b.best_v + q); oi = __ldcg(b.best_i + q);
            }
            if (ov > gb || (ov == gb && oi < gi)) { gb = ov; gi = oi; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, gb, o);
            const int oi = __shfl_xor_sync(0xffffffffu, gi, o);
            if (ov > gb || (ov == gb && oi < gi)) { gb = ov; gi = oi; }
          }
          if (lane == 0) s_c[rb] = gi;
        }
        __syncthreads();
        c = s_c[rb];
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.init_indices[m + 1] = c;
    // s_c / s_sum are rewritten in the next round: a block barrier here, or (CL) results double-buffered by round parity --
    // buffer rb is rewritten in round m + 2, behind the block barrier of round m + 1 that every reader of round m has passed
    if (!CL) __syncthreads();
    lap(4);
  }
  if (prof && s_tacc[6] > 0) {
    for (int i = 0; i < 5; ++i) b.part[i] = (float)((double)s_tacc[i] / (double)s_tacc[6]);
    b.part[5] = (float)s_tacc[6];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->degenerate_rounds = degenerate;
    st->n_noise_used = noise_row;
    *a.used = used;
  }
  mark(a, 2);
  if (CL) cluster_barrier();  // no CTA leaves (and frees its shared memory) while a peer may still be reading it
}

// ---- outputs shared by the grid and the cluster solver: CSR, centre times, chronological event order, status (CTA 0) ----
template <typename TS>
__device__ void solver_outputs(const SolverArgs& a, const int* s_offs, const int* s_memb, const double* ct_cur, const int* s_cntw,
                               int* s_cursor, float my_margin, int it, bool converged, int used, float last_shift) {
  const int k = a.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  csb_kmeans_status* st = a.status;
  my_margin = warp_min(my_margin);  // (the first lane of every frame group tracked its frames' margins)
  if (lane == 0 && my_margin < __int_as_float(0x7f800000))
    atomicMin(reinterpret_cast<unsigned*>(&st->min_margin), __float_as_uint(fmaxf(my_margin, 0.f)));
  if (blockIdx.x == 0) {
    for (int j = threadIdx.x; j <= k; j += blockDim.x) a.member_offsets[j] = s_offs[j];
    const int nmem = s_offs[k];
    for (int m = threadIdx.x; m < nmem; m += blockDim.x) a.members[m] = s_memb[m];
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      if (sizeof(TS) == 8) reinterpret_cast<double*>(a.ct_out)[j] = ct_cur[j];
      else reinterpret_cast<float*>(a.ct_out)[j] = (float)ct_cur[j];
    }
    // chronological per-event order: the member lists minus the seed frames of re-seeded (empty) clusters.
    // After rebuild_members_smem, cntw[w][j] holds warp w's running cursor == exclusive base + its count, so the
    // last warp's cursor is the real member count of cluster j.
    if (a.event_offsets && a.event_order && it > 0) {
      const int nw = a.multiwarp_fill ? kSolverWarps : 1;
      if (warp == 0) {
        const int cpl = (k + 31) / 32;
        const int j0 = min(k, lane * cpl), j1 = min(k, j0 + cpl);
        int mine = 0;
        for (int j = j0; j < j1; ++j) mine += s_cntw[(nw - 1) * k + j];
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += v;
        }
        int run = inc - mine;
        for (int j = j0; j < j1; ++j) {
          s_cursor[j] = run;
          a.event_offsets[j] = run;
          run += s_cntw[(nw - 1) * k + j];
        }
        if (lane == 31) a.event_offsets[k] = inc;
      }
      __syncthreads();
      for (int j = warp; j < k; j += kSolverWarps) {
        const int n = s_cntw[(nw - 1) * k + j], src = s_offs[j], dst = s_cursor[j];
        for (int e = lane; e < n; e += 32) a.event_order[dst + e] = s_memb[src + e];
      }
    }
    if (threadIdx.x == 0) {
      if (a.rng) const_cast<unsigned long long*>(a.rng)[1] += 1;  // every CTA loaded the key long before (grid barriers in between)
      st->iterations = it;
      st->converged = converged ? 1 : 0;
      st->n_reseed_used = used;
      st->last_shift = last_shift;
    }
  }
}

// ---- distances + assignment (kmeans_with_time.py:73-104) of the frames [f_begin, f_end): warp `gwarp` of `n_gwarps` ------------
// S = sum-of-Gram columns of the current centres, q = ||c_j||^2, ctc = centre times, s_rn = 1 / n_j (shared memory), diag = Gram
// diagonal, s_ts = shared-memory copy of the timestamps or nullptr.  Used by the cooperative solver (all frames, every CTA) and
// by the frame-sharded Lloyd kernels (a rank's own frames).
template <typename TS, bool WIDE>
__device__ __forceinline__ void assign_pass(const SolverArgs& a, const double* __restrict__ S, const double* q, const double* ctc,
                                            const double* s_rn, const double* diag, const TS* s_ts, TS alpha, int f_begin, int f_end,
                                            int gwarp, int n_gwarps, float& my_margin) {
  const int k = a.k;
  const int lane = threadIdx.x & 31;
  // A frame is handled by a group of L lanes (L = 32, or 16 / 8 when k <= 64 / 32 and the frame's k columns fit 4 per lane):
  // a warp then works on 32 / L frames side by side -- the same per-frame arithmetic, the dependent chain (S row load, fp64
  // square roots, reductions) of several frames overlapped.  Reductions run over the group's lanes only (min / max / argmin
  // with the lowest-index tie-break do not depend on the reduction tree).
  const int L = (!WIDE && k <= 32) ? 8 : ((!WIDE && k <= 64) ? 16 : 32);
  const int gpw = 32 / L;                      // frames per warp pass
  const int sub = lane & (L - 1), grp = lane / L;
  for (int i0 = f_begin + gwarp * gpw; i0 < f_end; i0 += n_gwarps * gpw) {
    const bool frame_ok = i0 + grp < f_end;
    const int i = frame_ok ? i0 + grp : f_end - 1;  // idle groups shadow the last frame (full-warp shuffles below), no store
    const double gii = diag[i];
    const TS ti = s_ts ? s_ts[i] : load_ts<TS>(a.ts, i);
    float fmin_ = __int_as_float(0x7f800000), fmax_ = -fmin_;
    TS tmin_ = (TS)__int_as_float(0x7f800000), tmax_ = -tmin_;
    constexpr int kWB = WIDE ? 8 : 1;  // 128-column blocks held in registers
    double sv[kWB][4];
    float dc[4];
    TS tc[4];
    double d2lo = __longlong_as_double(0x7ff0000000000000ll), d2hi = -d2lo;  // WIDE: extremes of the squared feature distance
    if (WIDE) {
#pragma unroll
      for (int bk = 0; bk < kWB; ++bk)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = bk * 128 + lane + 32 * u;
          sv[bk][u] = j < k ? __ldcg(S + (int64_t)i * k + j) : 0.0;
        }
      // The row extremes of d = (float)sqrt(max(d2, 0)) are the images of the extremes of d2 (clamp, square root and rounding are
      // monotone), so this pass needs no square root: two per frame follow the reduction.
#pragma unroll
      for (int bk = 0; bk < kWB; ++bk)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = bk * 128 + lane + 32 * u;
          if (j < k) {
            const double d2 = gii - 2.0 * (sv[bk][u] * s_rn[j]) + q[j];
            TS dt = ti - (TS)ctc[j];
            dt = dt < TS(0) ? -dt : dt;
            d2lo = d2 < d2lo ? d2 : d2lo; d2hi = d2 > d2hi ? d2 : d2hi;
            tmin_ = dt < tmin_ ? dt : tmin_; tmax_ = dt > tmax_ ? dt : tmax_;
          }
        }
    } else {
    {
      double s0[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = sub + L * u;
        s0[u] = j < k ? __ldcg(S + (int64_t)i * k + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = sub + L * u;
        if (j < k) {
          const double d2 = gii - 2.0 * (s0[u] * s_rn[j]) + q[j];
          dc[u] = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
          TS dt = ti - (TS)ctc[j];
          tc[u] = dt < TS(0) ? -dt : dt;
          fmin_ = fminf(fmin_, dc[u]); fmax_ = fmaxf(fmax_, dc[u]);
          tmin_ = tc[u] < tmin_ ? tc[u] : tmin_; tmax_ = tc[u] > tmax_ ? tc[u] : tmax_;
        }
      }
    }
    // k > 128 (then L == 32): further blocks of 128 columns, 4 independent S loads per lane in flight
    for (int j0 = 128; j0 < k; j0 += 128) {
      double sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + lane + 32 * u;
        sv[u] = j < k ? __ldcg(S + (int64_t)i * k + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + lane + 32 * u;
        if (j < k) {
          const double d2 = gii - 2.0 * (sv[u] * s_rn[j]) + q[j];
          const float d = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
          TS dt = ti - (TS)ctc[j];
          dt = dt < TS(0) ? -dt : dt;
          fmin_ = fminf(fmin_, d); fmax_ = fmaxf(fmax_, d);
          tmin_ = dt < tmin_ ? dt : tmin_; tmax_ = dt > tmax_ ? dt : tmax_;
        }
      }
    }
    }
    if (WIDE) {
      for (int o = 16; o > 0; o >>= 1) {
        const double f0 = __shfl_xor_sync(0xffffffffu, d2lo, o), f1 = __shfl_xor_sync(0xffffffffu, d2hi, o);
        const TS t0 = __shfl_xor_sync(0xffffffffu, tmin_, o), t1 = __shfl_xor_sync(0xffffffffu, tmax_, o);
        d2lo = f0 < d2lo ? f0 : d2lo; d2hi = f1 > d2hi ? f1 : d2hi;
        tmin_ = t0 < tmin_ ? t0 : tmin_; tmax_ = t1 > tmax_ ? t1 : tmax_;
      }
      fmin_ = (float)sqrt(d2lo > 0.0 ? d2lo : 0.0);
      fmax_ = (float)sqrt(d2hi > 0.0 ? d2hi : 0.0);
    } else {
    for (int o = L >> 1; o > 0; o >>= 1) {  // min / max over the group's lanes
      const float f0 = __shfl_xor_sync(0xffffffffu, fmin_, o), f1 = __shfl_xor_sync(0xffffffffu, fmax_, o);
      const TS t0 = __shfl_xor_sync(0xffffffffu, tmin_, o), t1 = __shfl_xor_sync(0xffffffffu, tmax_, o);
      fmin_ = f0 < fmin_ ? f0 : fmin_; fmax_ = f1 > fmax_ ? f1 : fmax_;
      tmin_ = t0 < tmin_ ? t0 : tmin_; tmax_ = t1 > tmax_ ? t1 : tmax_;
    }
    }
    const float fspread = __fsub_rn(fmax_, fmin_);
    const TS tspread = tmax_ - tmin_;
    float best = __int_as_float(0x7f800000), second = best;
    int bestj = 0x7fffffff;
    auto consider = [&](float d, TS dt, int j) {
      // 0 / spread is exactly +0; a zero numerator (the row minimum: one lane of every warp) would send the warp down the
      // IEEE division's slow path, so it is substituted and the quotient selected afterwards
      const float fnum = __fsub_rn(d, fmin_);
      float nf = (fmax_ > fmin_) ? __fdiv_rn(fnum == 0.f ? 1.f : fnum, fspread) : 0.f;  // :86-91
      nf = fnum == 0.f ? 0.f : nf;
      const TS tnum = dt - tmin_;
      TS nt = (tmax_ > tmin_) ? (tnum == TS(0) ? TS(1) : tnum) / tspread : TS(0);          // :94-99
      nt = tnum == TS(0) ? TS(0) : nt;
      float fin;                                                                               // :101
      if (sizeof(TS) == 8) {
        const double v = __dadd_rn((double)__fmul_rn(nf, nf), __dmul_rn((double)alpha, __dmul_rn((double)nt, (double)nt)));
        fin = __double2float_rn(sqrt(v));
      } else {
        const float ntf = (float)nt;
        fin = __fsqrt_rn(__fadd_rn(__fmul_rn(nf, nf), __fmul_rn((float)alpha, __fmul_rn(ntf, ntf))));
      }
      if (fin < best) { second = best; best = fin; bestj = j; }
      else if (fin < second) second = fin;
    };
    if (WIDE) {
      // Filter.  The exact final distance F_j = sqrt(nf_j^2 + alpha nt_j^2) (:101) costs three fp64 square roots and two IEEE
      // divisions per cluster; with k ~ 1000 clusters per frame that arithmetic IS the pass.  An fp32 estimate F~_j (approximate
      // square root, reciprocal multiplies) is within E of F_j:
      //   |nf~ - nf| <= 4e-7 fmax / (fmax - fmin) + 3e-7   (the estimate of d is within 3 ulp of the double-rounded exact d)
      //   sqrt(alpha) |nt~ - nt| <= 3e-7 sqrt(alpha),  fp32 evaluation of F~ and the final roundings of F: < 5e-7
      // so with E = 5e-7 fmax / (fmax - fmin) + 1e-6 (2 + sqrt(alpha)) the reference's pick (smallest F, lowest index on ties)
      // is among the clusters with F~_j <= min F~ + 2 E.  Only those (and the second smallest estimate, for the margin
      // diagnostic) are evaluated exactly -- typically 2 of k.  A NaN / overflowing estimate counts as a candidate.
      const float inv_f = fmax_ > fmin_ ? __frcp_rn(fspread) : 0.f;
      const float inv_t = tmax_ > tmin_ ? __frcp_rn((float)tspread) : 0.f;
      const float alf = (float)alpha;
      float f2[kWB][4];
      float m1 = __int_as_float(0x7f800000), m2 = m1;
      if (a.assign_exact != 1 && a.assign_exact != 3) {
#pragma unroll
        for (int bk = 0; bk < kWB; ++bk)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = bk * 128 + lane + 32 * u;
            f2[bk][u] = __int_as_float(0x7f800000);
            if (j < k) {
              const double d2 = gii - 2.0 * (sv[bk][u] * s_rn[j]) + q[j];
              float da;
              asm("sqrt.approx.f32 %0, %1;" : "=f"(da) : "f"((float)(d2 > 0.0 ? d2 : 0.0)));
              TS dt = ti - (TS)ctc[j];
              dt = dt < TS(0) ? -dt : dt;
              const float nf = (da - fmin_) * inv_f, nt = (float)(dt - tmin_) * inv_t;
              const float v = fmaf(alf * nt, nt, nf * nf);
              f2[bk][u] = v;
              if (v < m1) { m2 = m1; m1 = v; }
              else if (v < m2) m2 = v;
            }
          }
        for (int o = 16; o > 0; o >>= 1) {
          const float o1 = __shfl_xor_sync(0xffffffffu, m1, o), o2 = __shfl_xor_sync(0xffffffffu, m2, o);
          m2 = fminf(fmaxf(m1, o1), fminf(m2, o2));
          m1 = fminf(m1, o1);
        }
      }
      const float E = 5e-7f * (fmax_ > fmin_ ? fmax_ * inv_f : 0.f) + 1e-6f * (2.f + sqrtf(alf));
      const float thrF = sqrtf(m1) + 2.f * E;
      const float thr2 = a.assign_exact == 1 ? __int_as_float(0x7f800000) : (a.assign_exact >= 2 ? -1.f : fmaxf(thrF * thrF * 1.000001f, m2));
      // candidate slots of this lane as a bit mask; the exact evaluation is ONE copy of the code in a loop over the set bits
      // (unrolled per slot it was 32 copies of ~200 instructions: the pass stalled on instruction fetch).  The S entry of a
      // candidate is re-read (an L2 hit: the row was loaded a moment ago); slots ascend, so ties keep the lowest cluster.
      unsigned cand = 0;
#pragma unroll
      for (int bk = 0; bk < kWB; ++bk)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = bk * 128 + lane + 32 * u;
          if (j < k && (a.assign_exact == 1 || !(f2[bk][u] > thr2))) cand |= 1u << (bk * 4 + u);
        }
#pragma unroll 1
      while (cand) {
        const int slot = __ffs(cand) - 1;
        cand &= cand - 1;
        const int j = (slot >> 2) * 128 + lane + 32 * (slot & 3);
        const double svj = __ldcg(S + (int64_t)i * k + j);
        const double d2 = gii - 2.0 * (svj * s_rn[j]) + q[j];  // same arithmetic as the first pass: the same d2, dt bits
        const float d = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
        TS dt = ti - (TS)ctc[j];
        dt = dt < TS(0) ? -dt : dt;
        consider(d, dt, j);
      }
    } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = sub + L * u;
      if (j < k) consider(dc[u], tc[u], j);
    }
    for (int j0 = 128; j0 < k; j0 += 128) {
      double sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + lane + 32 * u;
        sv[u] = j < k ? __ldcg(S + (int64_t)i * k + j) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + lane + 32 * u;
        if (j < k) {
          const double d2 = gii - 2.0 * (sv[u] * s_rn[j]) + q[j];
          const float d = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
          TS dt = ti - (TS)ctc[j];
          dt = dt < TS(0) ? -dt : dt;
          consider(d, dt, j);
        }
      }
    }
    }
    for (int o = L >> 1; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const float os = __shfl_xor_sync(0xffffffffu, second, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
      const float hi = fmaxf(best, ob);
      second = fminf(hi, fminf(second, os));
      if (ob < best || (ob == best && oj < bestj)) { best = ob; bestj = oj; }
    }
    if (WIDE && a.assign_exact >= 2) bestj = 0;  // (timing aids: no cluster was evaluated)
    if (sub == 0 && frame_ok) {
      a.assignments[i] = bestj;  // :104 lowest index on ties
      a.assign32[i] = bestj;
      if (k > 1) my_margin = fminf(my_margin, second - best);
    }
  }
}

// ---- centre shift (:123-125): sum_j ||c'_j - c_j|| + ||ct' - ct||, identical in every thread of every CTA -------------------------
// Only changed clusters move.  Lane e of the warp owning cluster j sums the values of member e, e+32, ... (fixed order => an
// unchanged list would reproduce the same bits: shift exactly 0).  The value of member m for cluster j is S[m][j] in the single-GPU
// solver (vn / vc = the new / old S, stride k, by_column) and the all-gathered per-frame entry S[m][assign(m)] in the frame-sharded
// loop (stride 1; `per` > 0: the [world][2][per] exchange layout, vn / vc = its two planes).  Updates q_nxt, 1/n, n of the changed
// clusters, copies q of the others.
template <typename TS>
__device__ double shift_pass(const SolverArgs& a, const int* s_offs, const int* s_memb, const int* s_changed, const double* q_cur, double* q_nxt,
                             double* s_rn, double* s_nn, const double* ct_cur, const double* ct_nxt, double* s_red, const double* vn,
                             const double* vc, int per, const int* cnt_real) {
  const int k = a.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // single GPU (per == 0): vn / vc = the new / old S, entry (member m, cluster j) = S[m][j].
  // frame-sharded (per > 0): vn = the exchange buffer [world][2 * per + k]; rank r's segment holds S_nxt[i][a_i] and S_cur[i][a_i] of
  // its frames i (planes of `per`) and, for a re-seeded (empty) cluster j whose seed frame it owns, S_cur[seed][j] at 2 * per + j
  // (the last slot, 2 * per + k, carries the rank's running decision margin); the seed's new entry S_nxt[seed][j] is the Gram
  // diagonal.
  const int64_t seg = 2 * (int64_t)per + k + 1;
  auto at_new = [&](int m, int j) -> double {
    if (per == 0) return __ldcg(vn + (int64_t)m * k + j);
    if (cnt_real[j] == 0) return __ldcg(a.G + (int64_t)m * a.ldg + m);
    const int r = m / per;
    return __ldcg(vn + r * seg + (m - r * per));
  };
  auto at_old = [&](int m, int j) -> double {
    if (per == 0) return __ldcg(vc + (int64_t)m * k + j);
    const int r = m / per;
    if (cnt_real[j] == 0) return __ldcg(vn + r * seg + 2 * per + j);
    return __ldcg(vn + r * seg + per + (m - r * per));
  };
  double part = 0.0;
  constexpr int kSB = 4;  // clusters of a warp whose member loads are issued together (one L2 round trip instead of kSB)
  for (int j0 = warp; j0 < k; j0 += kSolverWarps * kSB) {
    int jj[kSB], beg[kSB], end[kSB];
    double sn[kSB], sc[kSB];
#pragma unroll
    for (int u = 0; u < kSB; ++u) {
      const int j = j0 + u * kSolverWarps;
      jj[u] = -1; beg[u] = 0; end[u] = 0; sn[u] = 0.0; sc[u] = 0.0;
      if (j < k) {
        if (!s_changed[j]) {
          if (lane == 0) q_nxt[j] = q_cur[j];
        } else {
          jj[u] = j; beg[u] = s_offs[j]; end[u] = s_offs[j + 1];
          const int e = beg[u] + lane;
          if (e < end[u]) {
            sn[u] = at_new(s_memb[e], j);
            sc[u] = at_old(s_memb[e], j);
          }
        }
      }
    }
    double my_tn = 0.0, my_tc = 0.0;
    int my_j = -1, my_n = 1;
#pragma unroll
    for (int u = 0; u < kSB; ++u) {
      if (jj[u] < 0) continue;  // warp-uniform
      const int j = jj[u];
      for (int e = beg[u] + lane + 32; e < end[u]; e += 32) {  // clusters of more than 32 members (same per-lane order)
        const double v1 = at_new(s_memb[e], j), v2 = at_old(s_memb[e], j);
        sn[u] += v1;
        sc[u] += v2;
      }
      const double tn = warp_sum(sn[u]), tc = warp_sum(sc[u]);
      if (lane == u) { my_tn = tn; my_tc = tc; my_j = j; my_n = end[u] - beg[u]; }
    }
    // lane u finishes cluster u of the batch: the divisions and the square root of the kSB clusters run side by side
    if (my_j >= 0) {
      const int j = my_j;
      const double n = (double)my_n;
      const double qn = my_tn / (n * n);
      const double cross = my_tc / (n * s_nn[j]);
      const double sh2 = (qn + q_cur[j]) - 2.0 * cross;
      q_nxt[j] = qn;
      s_rn[j] = 1.0 / n;  // cluster j's entries are owned by this warp until the next block barrier
      s_nn[j] = n;
      part += sqrt(sh2 > 0.0 ? sh2 : 0.0);
    }
  }
  const double shift_feat = block_sum<double>(part, s_red);
  double tpart = 0.0;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const TS d = (TS)ct_nxt[j] - (TS)ct_cur[j];
    tpart += (double)(d * d);
  }
  const double shift_time = sqrt(block_sum<double>(tpart, s_red));
  return shift_feat + shift_time;
}

// WIDE (128 < k <= 1024): the assignment pass keeps a frame's whole row of S in registers (one L2 / HBM round trip per frame
// instead of one per 128 columns and per pass); its own instantiation so that the small-k kernel keeps its register budget.
template <typename TS, bool WIDE>
__global__ void __launch_bounds__(kSolverThreads, WIDE ? 1 : 2) kmeans_solver_kernel(SolverArgs a) {
  extern __shared__ double smem_d[];
  const int T = a.T, k = a.k;
  // doubles: q[2][k], ct[2][k], rn[k] (1/n), nn[k] (n), red[32], [diag[T]]
  double* s_q[2] = {smem_d, smem_d + k};
  double* s_ctb[2] = {smem_d + 2 * k, smem_d + 3 * k};
  double* s_rn = smem_d + 4 * k;
  double* s_nn = smem_d + 5 * k;
  double* s_red = smem_d + 6 * k;
  double* s_diag = s_red + 32;
  // ints: assign[T] (16-bit entries, assign_words words), offs[k+1], cursor[k+1], memb[T+k], changed[k], ver[k], verbuf[2][k], cols[k], cntw[nw*k]
  TS* s_ts = reinterpret_cast<TS*>(s_diag + T);  // [T] (8-byte slots) next to the diagonal, same condition
  unsigned short* s_assign = reinterpret_cast<unsigned short*>(s_diag + (a.diag_in_smem ? 2 * T : 0));
  int* s_offs = reinterpret_cast<int*>(s_assign) + a.assign_words;
  int* s_cursor = s_offs + (k + 1);
  int* s_memb = s_cursor + (k + 1);
  int* s_changed = s_memb + (T + k);
  int* s_ver = s_changed + k;
  int* s_verbuf[2] = {s_ver + k, s_ver + 2 * k};
  int* s_cols = s_ver + 3 * k;
  int* s_cntw = s_cols + k;
  __shared__ int s_ncols;

  unsigned bar = 0, s_ticket = 0;
  const unsigned nb = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  csb_kmeans_status* st = a.status;
  int used = 0;
  const double* diag = a.diag_in_smem ? s_diag : a.diag;

  __shared__ PPScratch s_pp;
  if (a.fused_init) {
    // ---------------- k-means++ inside this launch: CTA 0 samples the centres while the other CTAs preload and wait ------------
    if (blockIdx.x > 0 && a.noise_ws) {
      // the idle CTAs pre-generate the Exp(1) noise of all sampling rounds (4 frames per Philox call) before anything else:
      // CTA 0 needs the first row a few microseconds into the launch
      const RngKey rk = load_rng(a);
      const int q4 = (T + 3) >> 2;  // calls per row
      const int64_t total = (int64_t)(k - 1) * q4;
      for (int64_t e = (int64_t)(blockIdx.x - 1) * blockDim.x + threadIdx.x; e < total; e += (int64_t)(nb - 1) * blockDim.x) {
        const int row = (int)(e / q4), c = (int)(e - (int64_t)(e / q4) * q4);
        const uint4 x = philox_draw4(rk.seed, rk.offset, (uint32_t)row, (uint32_t)c);
        float* dst = a.noise_ws + (int64_t)row * T + 4 * c;
        const float v[4] = {exp1_from_bits(x.x), exp1_from_bits(x.y), exp1_from_bits(x.z), exp1_from_bits(x.w)};
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (4 * c + u < T) __stcg(dst + u, v[u]);
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(a.noise_ready, 1u);
    }
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
      const double g = a.G[(int64_t)i * a.ldg + i];
      s_diag[i] = g;
      s_ts[i] = load_ts<TS>(a.ts, i);
      if (blockIdx.x == 0) a.diag[i] = g;
    }
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) {
        st->iterations = 0; st->converged = 0; st->n_reseed_used = 0; st->degenerate_rounds = 0;
        st->reseed_overflow = 0; st->n_noise_used = 0; st->min_margin = __int_as_float(0x7f800000); st->last_shift = 0.f;
      }
      mark(a, 1);
      __syncthreads();
      int n_noise = 0, n_deg = 0;
      if (a.n_init_given < k)  // s_near aliases s_assign: the assignments are not needed before the first iteration
        kmeanspp_block<kSolverThreads>(a, reinterpret_cast<float*>(s_assign), s_diag, &s_pp, &used, &n_noise, &n_deg);
      if (threadIdx.x == 0) {
        st->degenerate_rounds = n_deg;
        st->n_noise_used = n_noise;
        *a.used = used;
      }
      mark(a, 2);
    }
    grid_sync_backoff(a.barrier, bar, nb);
    used = __ldcg(a.used);
  } else {
    // ---------------- the k-means++ kernel ran before this launch: centres, diagonal and pool usage are in global memory ------
    used = *a.used;
    if (a.diag_in_smem)
      for (int i = threadIdx.x; i < T; i += blockDim.x) { s_diag[i] = a.diag[i]; s_ts[i] = load_ts<TS>(a.ts, i); }
  }
  // initial centres: cluster j = { init[j] }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int c = (int)__ldcg(a.init_indices + j);
    s_offs[j] = j;
    s_memb[j] = c;
    s_ctb[0][j] = (double)load_ts<TS>(a.ts, c);  // :66
    s_q[0][j] = diag[c];
    s_rn[j] = 1.0;
    s_nn[j] = 1.0;
    s_ver[j] = 1;
    s_verbuf[0][j] = 1;
    s_verbuf[1][j] = 0;
    s_cols[j] = j;
  }
  if (threadIdx.x == 0) s_offs[k] = k;
  __syncthreads();
  compute_S<WIDE>(a, s_offs, s_memb, s_cols, k, a.S[0], 0, T, s_ticket);
  grid_sync(a.barrier, bar, nb);
  mark(a, 3);

  MemberSmem ms;
  ms.assign = s_assign; ms.offs = s_offs; ms.memb = s_memb; ms.cursor = s_cursor; ms.cntw = s_cntw; ms.changed = s_changed;
  ms.ts_smem = a.diag_in_smem ? s_ts : nullptr;

  const TS alpha = (TS)a.alpha;
  int cur = 0, it = 0;
  bool converged = false, had_empty = false;
  float last_shift = 0.f;
  float my_margin = __int_as_float(0x7f800000);
  while (it < a.max_iter) {
    ++it;
    // ---------- distances + assignment (:73-104), a warp per frame ----------
    const double* S = a.S[cur];
    const double* q = s_q[cur];
    const double* ctc = s_ctb[cur];
    assign_pass<TS, WIDE>(a, S, q, ctc, s_rn, diag, a.diag_in_smem ? s_ts : nullptr, alpha, 0, T, blockIdx.x * kSolverWarps + warp,
                          (int)nb * kSolverWarps, my_margin);
    mark(a, 10);
    grid_sync(a.barrier, bar, nb);
    mark(a, 11);

    // ---------- new centres: every CTA rebuilds the membership in its own shared memory ----------
    const int nxt = cur ^ 1;
    for (int j = threadIdx.x; j < k; j += blockDim.x) s_changed[j] = it == 1 ? 1 : 0;
    __syncthreads();
    int moved = 0;
    for (int i0 = threadIdx.x; i0 < T; i0 += 8 * blockDim.x) {  // 8 loads in flight per thread (one L2 round trip per 2048 frames)
      int now[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) now[u] = i0 + u * blockDim.x < T ? __ldcg(a.assign32 + i0 + u * blockDim.x) : 0;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < T) {
          const int before = s_assign[i];  // (s_assign holds garbage before the first iteration)
          if (it > 1 && now[u] != before) { s_changed[before] = 1; s_changed[now[u]] = 1; moved = 1; }
          s_assign[i] = (unsigned short)now[u];
        }
      }
    }
    // No frame moved and no cluster was empty (an empty cluster is re-seeded with a fresh draw every iteration, :117-120): the
    // membership, hence every centre, is what it was -- the centre shift is exactly 0 and the reference stops here (:130).  Every
    // CTA sees the same assignments and takes the same branch; the lists / centre times of the previous rebuild are the outputs.
    if (!__syncthreads_or(moved) && it > 1 && !had_empty) {
      mark(a, 12);
      last_shift = 0.f;
      converged = true;
      break;
    }
    ms.ct_new = s_ctb[nxt];
    const int used_before = used;
    rebuild_members_smem<TS>(a, ms, &used, it == 1);
    had_empty = used != used_before;
    // version stamps: a changed cluster gets a new version; a column of S[nxt] is recomputed iff it is stale
    if (warp == 0) {
      int nc = 0;
      for (int j0 = 0; j0 < k; j0 += 32) {
        const int j = j0 + lane;
        bool stale = false;
        if (j < k) {
          if (s_changed[j]) s_ver[j] += 1;
          stale = s_verbuf[nxt][j] != s_ver[j];
          if (stale) s_verbuf[nxt][j] = s_ver[j];
        }
        const unsigned bal = __ballot_sync(0xffffffffu, stale);
        if (stale) s_cols[nc + __popc(bal & ((1u << lane) - 1u))] = j;
        nc += __popc(bal);
      }
      if (lane == 0) s_ncols = nc;
    }
    __syncthreads();
    mark(a, 12);
    compute_S<WIDE>(a, s_offs, s_memb, s_cols, s_ncols, a.S[nxt], 0, T, s_ticket);
    mark(a, 13);
    grid_sync(a.barrier, bar, nb);
    mark(a, 14);

    // ---------- centre shift (:123-125); every CTA evaluates it identically ----------
    // Only changed clusters move.  Lane e of the warp owning cluster j sums member e, e+32, ... of column j of the
    // new and the old S (fixed order => an unchanged list would reproduce the same bits: shift exactly 0).
    const double shift = shift_pass<TS>(a, s_offs, s_memb, s_changed, s_q[cur], s_q[nxt], s_rn, s_nn, s_ctb[cur], s_ctb[nxt], s_red,
                                        a.S[nxt], a.S[cur], 0, nullptr);
    __syncthreads();
    cur = nxt;
    mark(a, 15);
    last_shift = (float)shift;
    if (shift <= a.tol) { converged = true; break; }  // :130
  }

  // ---------------- outputs ----------------
  solver_outputs<TS>(a, s_offs, s_memb, s_ctb[cur], s_cntw, s_cursor, my_margin, it, converged, used, last_shift);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Frame-sharded Lloyd loop (first half): a single long stream over the GPUs of one box.
// Every rank holds the Gram matrix (the per-stream cache keeps it up to date) and runs the k-means++ init redundantly (a chain of
// k - 1 dependent sampling rounds: nothing to shard); the Lloyd iteration is split by FRAMES: rank r computes the distance /
// assignment pass and the S rows of its own frames [f_begin, f_end) only.  One iteration = three launches with a collective after
// the first two (the host side issues them: NCCL all-gather over NVLink):
//     ASSIGN  own frames -> assign32[f_begin:f_end)                          | all-gather assign32 [T]
//     UPDATE  (every rank, identical) membership CSR, centre times, changed columns; S[nxt] rows of the own frames for the
//             changed clusters; per own frame i the entries S_nxt[i][a_i], S_cur[i][a_i]     | all-gather of those [T] x 2
//     SHIFT   (every rank, identical) ||c_j||^2, 1/n_j and the centre shift from the gathered entries -> converged?
// The state that the single-GPU kernel keeps in shared memory across iterations lives in global memory between launches
// (LloydState, identical on every rank).  The phases call the SAME device functions as the cooperative solver (assign_pass,
// rebuild_members_smem, compute_S, shift_pass) with the same summation orders: assignments, iteration count, re-seeds and
// centre times are bit-identical to the single-GPU solve.
// ---------------------------------------------------------------------------------------------------------------------------------
enum { kLloydInit = 1, kLloydAssign = 2, kLloydUpdate = 3, kLloydShift = 4, kLloydOutput = 5 };
enum { kScCur = 0, kScIt = 1, kScUsed = 2, kScConverged = 3, kScLastShift = 4, kScMargin = 5, kScGlobalMargin = 6 };

struct LloydState {
  double* q[2];      // [k]
  double* ctb[2];    // [k]
  double* rn;        // [k]
  double* nn;        // [k]
  int* ver;          // [k]
  int* verbuf[2];    // [k]
  int* changed;      // [k]
  int* cntw_last;    // [k] assigned frames per cluster (without the seed of a re-seeded one)
  int* offs;         // [k + 1]
  int* memb;         // [T + k]
  int* assign_prev;  // [T]
  int* scal;         // [8]
  double* xch;       // [world][2][per] exchange planes: S_nxt[i][a_i] / S_cur[i][a_i] of every frame
  int phase, f_begin, f_end, per;
};

template <typename TS, bool WIDE>
__global__ void __launch_bounds__(kSolverThreads, 1) lloyd_phase_kernel(SolverArgs a, LloydState L) {
  extern __shared__ double smem_d[];
  const int T = a.T, k = a.k;
  double* s_q[2] = {smem_d, smem_d + k};
  double* s_ctb[2] = {smem_d + 2 * k, smem_d + 3 * k};
  double* s_rn = smem_d + 4 * k;
  double* s_nn = smem_d + 5 * k;
  double* s_red = smem_d + 6 * k;
  double* s_diag = s_red + 32;
  TS* s_ts = reinterpret_cast<TS*>(s_diag + T);
  unsigned short* s_assign = reinterpret_cast<unsigned short*>(s_diag + (a.diag_in_smem ? 2 * T : 0));
  int* s_offs = reinterpret_cast<int*>(s_assign) + a.assign_words;
  int* s_cursor = s_offs + (k + 1);
  int* s_memb = s_cursor + (k + 1);
  int* s_changed = s_memb + (T + k);
  int* s_ver = s_changed + k;
  int* s_verbuf[2] = {s_ver + k, s_ver + 2 * k};
  int* s_cols = s_ver + 3 * k;
  int* s_cntw = s_cols + k;
  __shared__ int s_ncols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned nb = gridDim.x;
  unsigned s_ticket = 0;  // (the launcher zeroes the ticket counter in front of every launch that runs an S pass)
  const int nw = a.multiwarp_fill ? kSolverWarps : 1;
  const int cur = L.scal[kScCur], nxt = cur ^ 1;
  if (a.diag_in_smem)
    for (int i = threadIdx.x; i < T; i += blockDim.x) { s_diag[i] = a.diag[i]; s_ts[i] = load_ts<TS>(a.ts, i); }
  const double* diag = a.diag_in_smem ? s_diag : a.diag;

  if (L.phase == kLloydInit) {
    // centres = the k seed frames (kmeans_with_time.py:65-66); S[0] rows of the own frames
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      s_offs[j] = j;
      s_memb[j] = (int)__ldcg(a.init_indices + j);
      s_cols[j] = j;
    }
    if (threadIdx.x == 0) s_offs[k] = k;
    __syncthreads();
    compute_S<WIDE>(a, s_offs, s_memb, s_cols, k, a.S[0], L.f_begin, L.f_end, s_ticket);
    if (blockIdx.x == 0) {
      for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const int c = s_memb[j];
        L.q[0][j] = a.diag[c];
        L.ctb[0][j] = (double)load_ts<TS>(a.ts, c);
        L.rn[j] = 1.0; L.nn[j] = 1.0;
        L.ver[j] = 1; L.verbuf[0][j] = 1; L.verbuf[1][j] = 0;
        L.cntw_last[j] = 0;
        L.offs[j] = j; L.memb[j] = c;
      }
      if (threadIdx.x == 0) {
        L.offs[k] = k;
        L.scal[kScCur] = 0; L.scal[kScIt] = 0; L.scal[kScUsed] = *a.used; L.scal[kScConverged] = 0;
        L.scal[kScLastShift] = 0; L.scal[kScMargin] = 0x7f800000;
      }
    }
    return;
  }

  if (L.phase == kLloydAssign) {
    for (int j = threadIdx.x; j < k; j += blockDim.x) { s_q[cur][j] = L.q[cur][j]; s_ctb[cur][j] = L.ctb[cur][j]; s_rn[j] = L.rn[j]; }
    __syncthreads();
    float my_margin = __int_as_float(0x7f800000);
    assign_pass<TS, WIDE>(a, a.S[cur], s_q[cur], s_ctb[cur], s_rn, diag, a.diag_in_smem ? s_ts : nullptr, (TS)a.alpha, L.f_begin, L.f_end,
                          blockIdx.x * kSolverWarps + warp, (int)nb * kSolverWarps, my_margin);
    my_margin = warp_min(my_margin);
    if (lane == 0 && my_margin < __int_as_float(0x7f800000))
      atomicMin(reinterpret_cast<unsigned*>(L.scal + kScMargin), __float_as_uint(fmaxf(my_margin, 0.f)));
    return;
  }

  if (L.phase == kLloydUpdate) {  // cooperative launch (one grid barrier)
    const int it = L.scal[kScIt] + 1;
    int used = L.scal[kScUsed];
    unsigned bar = 0;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      s_changed[j] = it == 1 ? 1 : 0;
      s_ver[j] = L.ver[j]; s_verbuf[nxt][j] = L.verbuf[nxt][j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
      const int now = __ldcg(a.assign32 + i), before = L.assign_prev[i];
      if (it > 1 && now != before) { s_changed[before] = 1; s_changed[now] = 1; }
      s_assign[i] = (unsigned short)now;
    }
    __syncthreads();
    MemberSmem ms;
    ms.assign = s_assign; ms.offs = s_offs; ms.memb = s_memb; ms.cursor = s_cursor; ms.cntw = s_cntw; ms.changed = s_changed;
    ms.ts_smem = a.diag_in_smem ? s_ts : nullptr;
    ms.ct_new = s_ctb[nxt];
    rebuild_members_smem<TS>(a, ms, &used, it == 1);
    if (warp == 0) {
      int nc = 0;
      for (int j0 = 0; j0 < k; j0 += 32) {
        const int j = j0 + lane;
        bool stale = false;
        if (j < k) {
          if (s_changed[j]) s_ver[j] += 1;
          stale = s_verbuf[nxt][j] != s_ver[j];
          if (stale) s_verbuf[nxt][j] = s_ver[j];
        }
        const unsigned bal = __ballot_sync(0xffffffffu, stale);
        if (stale) s_cols[nc + __popc(bal & ((1u << lane) - 1u))] = j;
        nc += __popc(bal);
      }
      if (lane == 0) s_ncols = nc;
    }
    __syncthreads();
    compute_S<WIDE>(a, s_offs, s_memb, s_cols, s_ncols, a.S[nxt], L.f_begin, L.f_end, s_ticket);
    grid_sync(a.barrier, bar, nb);
    // exchange planes: what the (replicated) shift pass needs of this rank's rows
    const int64_t seg = 2 * (int64_t)L.per + k + 1;
    for (int i = L.f_begin + blockIdx.x * blockDim.x + threadIdx.x; i < L.f_end; i += nb * blockDim.x) {
      const int ai = s_assign[i];
      const int r = i / L.per, o = i - r * L.per;
      // an unchanged cluster's column was not recomputed in S[nxt] (its shift term is skipped): send the value it would have
      const double v_cur = __ldcg(a.S[cur] + (int64_t)i * k + ai);
      L.xch[r * seg + o] = s_changed[ai] ? __ldcg(a.S[nxt] + (int64_t)i * k + ai) : v_cur;
      L.xch[r * seg + L.per + o] = v_cur;
    }
    // re-seeded (empty) clusters: the owner of the seed frame exports the seed's entry of the OLD column
    if (blockIdx.x == 0 && L.f_end > L.f_begin) {
      const int rme = L.f_begin / L.per;
      for (int j = threadIdx.x; j < k; j += blockDim.x) {
        double v = 0.0;
        if (s_cntw[(nw - 1) * k + j] == 0) {
          const int seed = s_memb[s_offs[j]];
          if (seed >= L.f_begin && seed < L.f_end) v = __ldcg(a.S[cur] + (int64_t)seed * k + j);
        }
        L.xch[rme * seg + 2 * L.per + j] = v;
      }
      if (threadIdx.x == 0) L.xch[rme * seg + 2 * L.per + k] = (double)__int_as_float(L.scal[kScMargin]);  // own frames' margin so far
    }
    if (blockIdx.x == 0) {
      for (int j = threadIdx.x; j < k; j += blockDim.x) {
        L.changed[j] = s_changed[j]; L.ver[j] = s_ver[j]; L.verbuf[nxt][j] = s_verbuf[nxt][j];
        L.ctb[nxt][j] = s_ctb[nxt][j];
        L.cntw_last[j] = s_cntw[(nw - 1) * k + j];
        L.offs[j] = s_offs[j];
      }
      if (threadIdx.x == 0) { L.offs[k] = s_offs[k]; L.scal[kScUsed] = used; }
      const int nmem = s_offs[k];
      for (int m = threadIdx.x; m < nmem; m += blockDim.x) L.memb[m] = s_memb[m];
      for (int i = threadIdx.x; i < T; i += blockDim.x) L.assign_prev[i] = s_assign[i];
    }
    return;
  }

  if (L.phase == kLloydShift) {  // one CTA
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      s_q[cur][j] = L.q[cur][j]; s_ctb[cur][j] = L.ctb[cur][j]; s_ctb[nxt][j] = L.ctb[nxt][j];
      s_rn[j] = L.rn[j]; s_nn[j] = L.nn[j]; s_changed[j] = L.changed[j]; s_offs[j] = L.offs[j];
      s_cols[j] = L.cntw_last[j];
    }
    if (threadIdx.x == 0) s_offs[k] = L.offs[k];
    __syncthreads();
    const int nmem = s_offs[k];
    for (int m = threadIdx.x; m < nmem; m += blockDim.x) s_memb[m] = L.memb[m];
    __syncthreads();
    const double shift = shift_pass<TS>(a, s_offs, s_memb, s_changed, s_q[cur], s_q[nxt], s_rn, s_nn, s_ctb[cur], s_ctb[nxt], s_red,
                                        L.xch, nullptr, L.per, s_cols);
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) { L.q[nxt][j] = s_q[nxt][j]; L.rn[j] = s_rn[j]; L.nn[j] = s_nn[j]; }
    if (threadIdx.x == 0) {
      float mg = __int_as_float(0x7f800000);  // smallest decision margin over all ranks' frames (diagnostic of csb_kmeans_status)
      const int64_t seg = 2 * (int64_t)L.per + k + 1;
      for (int r = 0; r * L.per < T; ++r) mg = fminf(mg, (float)__ldcg(L.xch + r * seg + 2 * L.per + k));
      L.scal[kScGlobalMargin] = __float_as_int(mg);
      L.scal[kScCur] = nxt;
      L.scal[kScIt] = L.scal[kScIt] + 1;
      L.scal[kScConverged] = shift <= a.tol ? 1 : 0;  // :130
      L.scal[kScLastShift] = __float_as_int((float)shift);
    }
    return;
  }

  if (L.phase == kLloydOutput) {  // one CTA
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      s_offs[j] = L.offs[j]; s_ctb[cur][j] = L.ctb[cur][j]; s_cntw[(nw - 1) * k + j] = L.cntw_last[j];
    }
    if (threadIdx.x == 0) s_offs[k] = L.offs[k];
    __syncthreads();
    const int nmem = s_offs[k];
    for (int m = threadIdx.x; m < nmem; m += blockDim.x) s_memb[m] = L.memb[m];
    if (L.scal[kScIt] > 0)
      for (int i = threadIdx.x; i < T; i += blockDim.x) a.assignments[i] = L.assign_prev[i];  // (:104 of the last iteration)
    __syncthreads();
    solver_outputs<TS>(a, s_offs, s_memb, s_ctb[cur], s_cntw, s_cursor, __int_as_float(L.scal[kScIt] > 0 ? L.scal[kScGlobalMargin] : 0x7f800000), L.scal[kScIt],
                       L.scal[kScConverged] != 0, L.scal[kScUsed], __int_as_float(L.scal[kScLastShift]));
  }
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SolverSmemPlan {
  int diag_in_smem, multiwarp_fill, assign_words;
  size_t bytes;
};

static SolverSmemPlan plan_smem(int64_t T, int64_t k) {
  SolverSmemPlan p;
  // assignments are 16-bit in shared memory (k <= 4096); short streams keep T words, the fused init aliases them as float[T].
  // At BASELINE config 4 (T = 14,400, k = 960) that is what lets the per-warp counters of the membership rebuild fit under the
  // 196 KB carve-out (190 KB in all): the rebuild runs 8 warps wide (143 -> 55 us per iteration) and the L1 keeps 60 KB for the
  // Gram-row streaming of the S pass (with the 228 KB carve-out that pass ran 1.5x slower).
  p.assign_words = (int)(T <= 1024 ? T : (T + 1) / 2);
  const size_t base = sizeof(double) * (6 * k + 32) + sizeof(int) * ((size_t)p.assign_words + T + 8 * k + 4) + 16;
  const size_t budget = 200 * 1024;
  p.multiwarp_fill = (base + sizeof(int) * kSolverWarps * k <= budget && !getenv("CSB_NO_MULTIWARP")) ? 1 : 0;  // (env: tuning aid)
  size_t b = base + sizeof(int) * (p.multiwarp_fill ? kSolverWarps : 1) * k;
  p.diag_in_smem = (b + 2 * sizeof(double) * T <= budget) ? 1 : 0;  // Gram diagonal + timestamps
  if (p.diag_in_smem) b += 2 * sizeof(double) * T;
  p.bytes = b;
  return p;
}

}  // namespace csb

using namespace csb;

extern "C" {

size_t csb_kmeans_workspace_bytes(int64_t T, int64_t k) {
  if (T <= 0 || k <= 0) return 0;
  size_t b = 0;
  b += align_up(sizeof(double) * T, 256);              // diag
  b += 2 * align_up(sizeof(double) * T * k, 256);      // S
  b += align_up(sizeof(int32_t) * T, 256);             // compact assignments
  if (T <= 1024) b += align_up(sizeof(float) * T * k, 256);  // pre-generated noise of the built-in generator (fused init)
  b += 4096;                                           // barrier + used + phase timeline
  return b;
}

int csb_kmeans_solve(const double* G, const void* timestamps, const csb_kmeans_params* p, int64_t* init_indices,
                     const float* exp_noise, const int64_t* reseed_pool, int64_t* assignments, void* centroid_times,
                     int32_t* member_offsets, int32_t* members, int32_t* event_order, int32_t* event_offsets,
                     csb_kmeans_status* status, void* workspace, size_t workspace_bytes, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(p, "kmeans_solve: null params");
  CSB_REQUIRE(G && timestamps && init_indices && assignments && centroid_times && member_offsets && members && status && workspace,
              "kmeans_solve: null pointer");
  CSB_REQUIRE(p->k >= 1 && p->T > p->k, "kmeans_solve: need T > k >= 1 (T=%lld k=%lld); T <= k is the caller's early return",
              (long long)p->T, (long long)p->k);
  CSB_REQUIRE(p->T < (1ll << 30) && p->k <= 4096, "kmeans_solve: T or k too large (k <= 4096)");
  CSB_REQUIRE(p->ts_dtype == CSB_F32 || p->ts_dtype == CSB_F64, "kmeans_solve: timestamps must be f32 or f64");
  CSB_REQUIRE(p->n_init_given >= (p->rng_state ? 0 : 1) && p->n_init_given <= p->k,
              "kmeans_solve: n_init_given must be in [1,k] ([0,k] with the built-in generator)");
  CSB_REQUIRE(p->n_init_given == p->k || exp_noise || p->rng_state, "kmeans_solve: exp_noise or rng_state required when centres are sampled");
  CSB_REQUIRE(p->n_reseed_pool == 0 || reseed_pool, "kmeans_solve: reseed_pool is null");
  // max_iteration < 0: the k-means++ init only (init_indices, Gram diagonal, pool usage): the frame-sharded Lloyd phases follow
  const bool init_only = p->max_iteration < 0;
  CSB_REQUIRE(p->ldg == 0 || p->ldg >= p->T, "kmeans_solve: ldg < T");
  const int64_t T = p->T, k = p->k;
  const SolverSmemPlan sp = plan_smem(T, k);
  CSB_REQUIRE(sp.bytes <= 225 * 1024, "kmeans_solve: T=%lld k=%lld do not fit the shared-memory membership (T <= ~25k)", (long long)T,
              (long long)k);
  if (workspace_bytes < csb_kmeans_workspace_bytes(T, k)) {
    set_error("kmeans_solve: workspace %zu < %zu", workspace_bytes, csb_kmeans_workspace_bytes(T, k));
    return CSB_ERR_WORKSPACE;
  }
  SolverArgs a;
  memset(&a, 0, sizeof(a));
  a.G = G; a.ldg = p->ldg > 0 ? p->ldg : T; a.ts = timestamps; a.T = (int)T; a.k = (int)k; a.max_iter = p->max_iteration; a.alpha = p->alpha; a.tol = p->tol;
  a.n_init_given = (int)p->n_init_given; a.n_reseed_pool = (int)p->n_reseed_pool;
  a.init_indices = init_indices; a.exp_noise = exp_noise; a.reseed_pool = reseed_pool; a.assignments = assignments;
  a.ct_out = centroid_times; a.member_offsets = member_offsets; a.members = members; a.status = status;
  a.event_order = event_order; a.event_offsets = event_offsets;
  a.diag_in_smem = sp.diag_in_smem; a.multiwarp_fill = sp.multiwarp_fill; a.assign_words = sp.assign_words;
  a.init_exact = getenv("CSB_INIT_EXACT") ? 1 : 0;
  a.assign_exact = getenv("CSB_ASSIGN_EXACT") ? atoi(getenv("CSB_ASSIGN_EXACT")) : 0;  // (2 / 3: timing aids, wrong results)
  a.s_no_pairs = getenv("CSB_S_NO_PAIRS") ? 1 : 0;
  a.rng = reinterpret_cast<const unsigned long long*>(p->rng_state);
  char* w = reinterpret_cast<char*>(workspace);
  a.diag = reinterpret_cast<double*>(w); w += align_up(sizeof(double) * T, 256);
  for (int b = 0; b < 2; ++b) { a.S[b] = reinterpret_cast<double*>(w); w += align_up(sizeof(double) * T * k, 256); }
  a.assign32 = reinterpret_cast<int32_t*>(w); w += align_up(sizeof(int32_t) * T, 256);
  float* noise_ws = nullptr;
  if (T <= 1024) { noise_ws = reinterpret_cast<float*>(w); w += align_up(sizeof(float) * T * k, 256); }
  a.barrier = reinterpret_cast<unsigned*>(w);
  a.noise_ready = reinterpret_cast<unsigned*>(w + 192);
  a.noise_ws = nullptr;  // set below, for the fused launch only (the separate init kernels generate their noise inline)
  a.used = reinterpret_cast<int32_t*>(w + 128);
  a.tl = reinterpret_cast<unsigned long long*>(w + 256);
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(w, 0, 264, s);
  if (e != cudaSuccess) { set_error("kmeans_solve: memset: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }

  const size_t smem = sp.bytes;
  const bool wide = k > 128 && k <= 1024;
  auto kern = (p->ts_dtype == CSB_F64) ? (wide ? kmeans_solver_kernel<double, true> : kmeans_solver_kernel<double, false>)
                                       : (wide ? kmeans_solver_kernel<float, true> : kmeans_solver_kernel<float, false>);
  // (the 48 KB default covers static + dynamic shared memory: opt in from 40 KB of dynamic, the kernels have < 1 KB of static)
  if (smem > 40 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("kmeans_solve: smem attr: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSolverThreads, smem);
  if (e != cudaSuccess || per_sm < 1) { set_error("kmeans_solve: occupancy query failed"); return CSB_ERR_CUDA; }
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // k-means++ (or just the diagonal / status reset when every centre is prescribed): one CTA, or a cooperative grid for
  // long streams (CSB_INIT_GRID=0/1 forces either: profiling aid)
  bool grid_init = T > kGridInitMinT && p->n_init_given < k;
  if (const char* o = getenv("CSB_INIT_GRID")) grid_init = atoi(o) != 0 && T <= (int64_t)sms * kGridInitThreads * kGridInitU;
  // short streams (one init thread per <= 4 frames at the solver's block size): the init runs on CTA 0 of the solver launch
  a.fused_init = (!grid_init && T <= 1024 && sp.diag_in_smem && !getenv("CSB_INIT_THREADS") && !init_only) ? 1 : 0;
  if (const char* o = getenv("CSB_FUSED_INIT")) a.fused_init = a.fused_init && atoi(o) != 0;
  if (a.fused_init) {
    // nothing to launch here
  } else if (grid_init) {
    int64_t nbi = (T + kGridInitThreads - 1) / kGridInitThreads;
    if (nbi > sms) nbi = sms;
    GridInitBufs gb;
    gb.barrier = reinterpret_cast<unsigned*>(w + 64);
    // scratch inside S[0]: the Lloyd kernel (re)builds S[0] from scratch after the init
    char* sc0 = reinterpret_cast<char*>(a.S[0]);
    auto carve = [&](size_t bytes) { char* r = sc0; sc0 += align_up(bytes, 256); return r; };
    gb.part = reinterpret_cast<float*>(carve(sizeof(float) * 2 * nbi));
    gb.cnt = reinterpret_cast<int*>(carve(sizeof(int) * 2 * nbi));
    gb.cand_n2 = reinterpret_cast<float*>(carve(sizeof(float) * 2 * nbi * kGridInitCand));
    gb.cand_nz = reinterpret_cast<float*>(carve(sizeof(float) * 2 * nbi * kGridInitCand));
    gb.cand_i = reinterpret_cast<int*>(carve(sizeof(int) * 2 * nbi * kGridInitCand));
    gb.best_v = reinterpret_cast<float*>(carve(sizeof(float) * nbi));
    gb.best_i = reinterpret_cast<int*>(carve(sizeof(int) * nbi));
    gb.force_exact = getenv("CSB_INIT_EXACT") ? 1 : 0;
    gb.profile = getenv("CSB_INIT_PROFILE") ? 1 : 0;
    gb.n_cluster = 0; gb.n_helpers = 0;
    gb.push = (getenv("CSB_INIT_PUSH") && atoi(getenv("CSB_INIT_PUSH")) == 0) ? 0 : 1;  // (0: cluster barrier + DSMEM loads, validation)
    // One thread-block cluster when the stream fits 16 CTAs (T <= 32,768): hardware cluster barrier + DSMEM exchange per round
    // instead of the software grid barrier + L2 round trip.  The cluster size is a power of two (surplus CTAs own an empty
    // slice); sizes above 8 are non-portable and taken only if the device can place such a cluster.  CSB_INIT_CLUSTER=0: off.
    bool clustered = false;
    const char* cl_env = getenv("CSB_INIT_CLUSTER");
    const bool cl_want = cl_env ? atoi(cl_env) != 0 : kClusterInitDefault;
    if (cl_want && T <= (int64_t)kClusterInitMax * kGridInitThreads * kGridInitU) {
      int csz = 2;
      while (csz < nbi && csz < kClusterInitMax) csz *= 2;
      if ((int64_t)csz * kGridInitThreads * kGridInitU < T) csz = kClusterInitMax;
      if (const char* o = getenv("CSB_INIT_CLUSTER_SIZE")) {  // tuning aid: 2 / 4 / 8 / 16, if the stream fits
        const int v = atoi(o);
        if ((v == 2 || v == 4 || v == 8 || v == 16) && (int64_t)v * kGridInitThreads * kGridInitU >= T) csz = v;
      }
      // frames per CTA decide the block: one frame per thread keeps the per-round dependent chain (fp64 distance, square root,
      // two divisions) at one frame's latency -- 1024 threads x 1 frame when the slice fits, else 512 threads x <= 4 frames
      // (CSB_INIT_CLUSTER_BLOCK=512 forces the latter: tuning aid)
      const int64_t per_cta = (T + csz - 1) / csz;
      const char* blk_env = getenv("CSB_INIT_CLUSTER_BLOCK");
      const int variant = (per_cta <= 512) ? 0 : ((per_cta <= 1024 && !(blk_env && atoi(blk_env) == 512)) ? 1 : 2);
      void (*ckern)(SolverArgs, GridInitBufs) = variant == 0 ? kmeanspp_grid_kernel<true, 512, 1>
                                               : (variant == 1 ? kmeanspp_grid_kernel<true, 1024, 1> : kmeanspp_grid_kernel<true, kGridInitThreads, kGridInitU>);
      // built-in generator: further clusters of the same launch pre-generate the noise rows (into S[1], unused before the first
      // Lloyd iteration) on the SMs the sampling cluster leaves idle (CSB_INIT_HELPERS=0: generate inside the rounds)
      int n_hclusters = 0;
      if (a.rng && !exp_noise && k > 1 && p->n_init_given < k && !(getenv("CSB_INIT_HELPERS") && atoi(getenv("CSB_INIT_HELPERS")) == 0)) {
        n_hclusters = (sms - csz) / csz;
        if (n_hclusters > 8) n_hclusters = 8;
        if (n_hclusters < 0) n_hclusters = 0;
      }
      gb.n_cluster = csz;
      gb.n_helpers = n_hclusters * csz;
      if (gb.n_helpers > 0) a.noise_ws = reinterpret_cast<float*>(a.S[1]);
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)(csz * (1 + n_hclusters)));
      cfg.blockDim = dim3(variant == 1 ? 1024 : 512);
      cfg.stream = s;
      cudaLaunchAttribute cattr[1];
      cattr[0].id = cudaLaunchAttributeClusterDimension;
      cattr[0].val.clusterDim.x = (unsigned)csz;
      cattr[0].val.clusterDim.y = 1;
      cattr[0].val.clusterDim.z = 1;
      cfg.attrs = cattr;
      cfg.numAttrs = 1;
      bool ok = true;
      if (csz > 8) ok = cudaFuncSetAttribute(ckern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
      int n_clusters = 0;
      if (ok) ok = cudaOccupancyMaxActiveClusters(&n_clusters, ckern, &cfg) == cudaSuccess && n_clusters >= 1;
      if (!ok) cudaGetLastError();  // not placeable here: clear the error, take the cooperative grid below
      if (ok) {
        e = cudaLaunchKernelEx(&cfg, ckern, a, gb);
        if (e != cudaSuccess) { set_error("kmeans_solve: cluster init launch: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
        clustered = true;
      }
      a.noise_ws = nullptr;  // (the solver launch below sets its own, for the fused init only)
    }
    if (!clustered) {
      void* iargs[] = {&a, &gb};
      e = cudaLaunchCooperativeKernel((const void*)kmeanspp_grid_kernel<false, kGridInitThreads, kGridInitU>, dim3((unsigned)nbi), dim3(kGridInitThreads), iargs, 0, s);
      if (e != cudaSuccess) { set_error("kmeans_solve: init launch: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
    }
    count_launch(1);
  } else {
    // few warps = cheap block barriers: the rounds are a dependent chain, each about (frames / threads) elements of work
    int nt = T <= 1024 ? 256 : (T <= 4096 ? 512 : 1024);
    if (const char* o = getenv("CSB_INIT_THREADS")) nt = atoi(o);
    const bool f64 = p->ts_dtype == CSB_F64;
    void (*ikern)(SolverArgs) = nullptr;
    if (nt <= 128) { nt = 128; ikern = f64 ? kmeanspp_kernel<double, 128> : kmeanspp_kernel<float, 128>; }
    else if (nt <= 256) { nt = 256; ikern = f64 ? kmeanspp_kernel<double, 256> : kmeanspp_kernel<float, 256>; }
    else if (nt <= 512) { nt = 512; ikern = f64 ? kmeanspp_kernel<double, 512> : kmeanspp_kernel<float, 512>; }
    else { nt = 1024; ikern = f64 ? kmeanspp_kernel<double, 1024> : kmeanspp_kernel<float, 1024>; }
    const size_t ismem = sizeof(float) * T + (sp.diag_in_smem ? sizeof(double) * T : 0) + 16;
    if (ismem > 40 * 1024) {  // (static + dynamic must fit the 48 KB default otherwise)
      e = cudaFuncSetAttribute(ikern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ismem);
      if (e != cudaSuccess) { set_error("kmeans_solve: init smem attr: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
    }
    ikern<<<1, nt, ismem, s>>>(a);
    int st = check_launch("kmeanspp");
    if (st) return st;
  }
  // Lloyd loop: frames_per_warp frames per warp in the assignment phase (default 2; T = 600: 38 CTAs).  The loop is
  // latency-bound: one frame per warp (75 CTAs) saves ~2 us per iteration for a stream that runs alone, but with six streams in
  // flight the extra co-resident CTAs cost the bandwidth-bound kernels of the other streams 7 % (measured: 2.15 vs 2.00 M frames/s)
  if (init_only) return CSB_OK;
  int fpw = p->frames_per_warp > 0 ? (int)p->frames_per_warp : 2;  // (CSB_SOLVER_FPW overrides: tuning runs)
  if (const char* o = getenv("CSB_SOLVER_FPW")) fpw = atoi(o) > 0 ? atoi(o) : fpw;
  // a warp pass covers 32 / L frames (L lanes per frame: 8 / 16 for k <= 32 / 64 in the narrow kernel, see the assignment pass)
  const int gpw = wide ? 1 : (k <= 32 ? 4 : (k <= 64 ? 2 : 1));
  if (!(p->frames_per_warp > 0) && !getenv("CSB_SOLVER_FPW") && gpw > 1) fpw = 1;  // the side-by-side frames already halve the CTA count
  int64_t nb = (T + (int64_t)fpw * gpw * kSolverWarps - 1) / ((int64_t)fpw * gpw * kSolverWarps);
  if (nb > sms) nb = sms;
  if (nb < 1) nb = 1;
  // built-in generator + fused init + helpers to spare: CTAs 1.. pre-generate the noise rows while CTA 0 samples
  a.noise_ws = (a.fused_init && a.rng && !exp_noise && nb > 1 && k > 1 && p->n_init_given < k &&
                !getenv("CSB_NO_NOISE_WS")) ? noise_ws : nullptr;
  void* args[] = {&a};
  e = cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)nb), dim3(kSolverThreads), args, smem, s);
  if (e != cudaSuccess) { set_error("kmeans_solve: launch: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  count_launch(1);
  return CSB_OK;
}

size_t csb_kmeans_lloyd_state_bytes(int64_t T, int64_t k) {
  if (T <= 0 || k <= 0) return 0;
  return 6 * align_up(sizeof(double) * k, 256) + 6 * align_up(sizeof(int) * k, 256) + align_up(sizeof(int) * (k + 1), 256) +
         align_up(sizeof(int) * (T + k), 256) + align_up(sizeof(int) * T, 256) + 256;
}

int csb_kmeans_lloyd_phase(const double* G, const void* timestamps, const csb_kmeans_params* p, const csb_lloyd_shard* sh,
                           int64_t* init_indices, const int64_t* reseed_pool, int32_t* assign32, int64_t* assignments, void* centroid_times,
                           int32_t* member_offsets, int32_t* members, int32_t* event_order, int32_t* event_offsets, csb_kmeans_status* status,
                           void* workspace, size_t workspace_bytes, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(p && sh, "kmeans_lloyd_phase: null params");
  CSB_REQUIRE(G && timestamps && init_indices && assign32 && assignments && centroid_times && member_offsets && members && status &&
                  workspace && sh->state && sh->exchange,
              "kmeans_lloyd_phase: null pointer");
  CSB_REQUIRE(p->k >= 1 && p->T > p->k && p->T < (1ll << 30) && p->k <= 4096, "kmeans_lloyd_phase: need T > k >= 1, k <= 4096");
  CSB_REQUIRE(p->ts_dtype == CSB_F32 || p->ts_dtype == CSB_F64, "kmeans_lloyd_phase: timestamps must be f32 or f64");
  CSB_REQUIRE(sh->phase >= kLloydInit && sh->phase <= kLloydOutput, "kmeans_lloyd_phase: unknown phase %d", sh->phase);
  CSB_REQUIRE(sh->per_rank > 0 && sh->f_begin >= 0 && sh->f_begin <= sh->f_end && sh->f_end <= p->T && sh->f_end - sh->f_begin <= sh->per_rank &&
                  (sh->f_begin % sh->per_rank == 0 || sh->f_begin == sh->f_end),
              "kmeans_lloyd_phase: bad frame range");
  CSB_REQUIRE(p->ldg == 0 || p->ldg >= p->T, "kmeans_lloyd_phase: ldg < T");
  const int64_t T = p->T, k = p->k;
  const SolverSmemPlan sp = plan_smem(T, k);
  CSB_REQUIRE(sp.bytes <= 225 * 1024, "kmeans_lloyd_phase: T=%lld k=%lld do not fit the shared-memory membership", (long long)T, (long long)k);
  if (workspace_bytes < csb_kmeans_workspace_bytes(T, k)) { set_error("kmeans_lloyd_phase: workspace too small"); return CSB_ERR_WORKSPACE; }
  if (sh->state_bytes < csb_kmeans_lloyd_state_bytes(T, k)) { set_error("kmeans_lloyd_phase: state too small"); return CSB_ERR_WORKSPACE; }
  SolverArgs a;
  memset(&a, 0, sizeof(a));
  a.G = G; a.ldg = p->ldg > 0 ? p->ldg : T; a.ts = timestamps; a.T = (int)T; a.k = (int)k; a.max_iter = p->max_iteration; a.alpha = p->alpha; a.tol = p->tol;
  a.n_init_given = (int)p->n_init_given; a.n_reseed_pool = (int)p->n_reseed_pool;
  a.init_indices = init_indices; a.reseed_pool = reseed_pool; a.assignments = assignments;
  a.ct_out = centroid_times; a.member_offsets = member_offsets; a.members = members; a.status = status;
  a.event_order = event_order; a.event_offsets = event_offsets;
  a.diag_in_smem = sp.diag_in_smem; a.multiwarp_fill = sp.multiwarp_fill; a.assign_words = sp.assign_words;
  a.rng = reinterpret_cast<const unsigned long long*>(p->rng_state);
  char* w = reinterpret_cast<char*>(workspace);
  a.diag = reinterpret_cast<double*>(w); w += align_up(sizeof(double) * T, 256);
  for (int b = 0; b < 2; ++b) { a.S[b] = reinterpret_cast<double*>(w); w += align_up(sizeof(double) * T * k, 256); }
  w += align_up(sizeof(int32_t) * T, 256);
  if (T <= 1024) w += align_up(sizeof(float) * T * k, 256);
  a.assign32 = assign32;  // the caller's [world * per_rank] buffer (all-gathered between ASSIGN and UPDATE)
  a.barrier = reinterpret_cast<unsigned*>(w);
  a.used = reinterpret_cast<int32_t*>(w + 128);
  a.tl = reinterpret_cast<unsigned long long*>(w + 256);
  LloydState L;
  char* sw = reinterpret_cast<char*>(sh->state);
  auto carve = [&](size_t bytes) { char* r = sw; sw += align_up(bytes, 256); return r; };
  for (int b = 0; b < 2; ++b) L.q[b] = reinterpret_cast<double*>(carve(sizeof(double) * k));
  for (int b = 0; b < 2; ++b) L.ctb[b] = reinterpret_cast<double*>(carve(sizeof(double) * k));
  L.rn = reinterpret_cast<double*>(carve(sizeof(double) * k));
  L.nn = reinterpret_cast<double*>(carve(sizeof(double) * k));
  L.ver = reinterpret_cast<int*>(carve(sizeof(int) * k));
  for (int b = 0; b < 2; ++b) L.verbuf[b] = reinterpret_cast<int*>(carve(sizeof(int) * k));
  L.changed = reinterpret_cast<int*>(carve(sizeof(int) * k));
  L.cntw_last = reinterpret_cast<int*>(carve(sizeof(int) * k));
  (void)carve(sizeof(int) * k);
  L.offs = reinterpret_cast<int*>(carve(sizeof(int) * (k + 1)));
  L.memb = reinterpret_cast<int*>(carve(sizeof(int) * (T + k)));
  L.assign_prev = reinterpret_cast<int*>(carve(sizeof(int) * T));
  L.scal = reinterpret_cast<int*>(carve(32));
  L.xch = reinterpret_cast<double*>(sh->exchange);
  L.phase = sh->phase; L.f_begin = (int)sh->f_begin; L.f_end = (int)sh->f_end; L.per = (int)sh->per_rank;
  cudaStream_t s = as_stream(stream);
  const bool wide = k > 128 && k <= 1024;
  auto kern = (p->ts_dtype == CSB_F64) ? (wide ? lloyd_phase_kernel<double, true> : lloyd_phase_kernel<double, false>)
                                       : (wide ? lloyd_phase_kernel<float, true> : lloyd_phase_kernel<float, false>);
  const size_t smem = sp.bytes;
  cudaError_t e;
  // (the 48 KB default covers static + dynamic shared memory: opt in from 40 KB of dynamic, the kernels have < 1 KB of static)
  if (smem > 40 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("kmeans_lloyd_phase: smem attr: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  }
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t nf = sh->f_end - sh->f_begin;
  int64_t nb = 1;
  if (sh->phase == kLloydAssign) {
    const int gpw = wide ? 1 : (k <= 32 ? 4 : (k <= 64 ? 2 : 1));
    nb = (nf + (int64_t)gpw * kSolverWarps - 1) / ((int64_t)gpw * kSolverWarps);
  } else if (sh->phase == kLloydInit || sh->phase == kLloydUpdate) {
    nb = (nf * (k < 64 ? k : 64) + kSolverThreads * 4 - 1) / (kSolverThreads * 4);  // S rows: ~4 entries per thread and pass
  }
  if (nb > sms) nb = sms;
  if (nb < 1) nb = 1;
  void* args[] = {&a, &L};
  if (sh->phase == kLloydInit || sh->phase == kLloydUpdate) {  // grid barrier word + the ticket counter of the S pass
    e = cudaMemsetAsync(a.barrier, 0, 8, s);
    if (e != cudaSuccess) { set_error("kmeans_lloyd_phase: memset: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  }
  if (sh->phase == kLloydUpdate) {
    e = cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)nb), dim3(kSolverThreads), args, smem, s);
  } else {
    e = cudaLaunchKernel((const void*)kern, dim3((unsigned)nb), dim3(kSolverThreads), args, smem, s);
  }
  if (e != cudaSuccess) { set_error("kmeans_lloyd_phase: launch: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  count_launch(1);
  return CSB_OK;
}

}  // extern "C"
