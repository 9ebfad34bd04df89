// Selection / masking / compaction kernels (reference: model/cogreasoner_chat.py:50-64, :285-296,
// :325-333, :383-432, :449-476).  Integer and byte work is exact; bf16 arithmetic reproduces ATen's
// op-by-op rounding (fp32 math, one round-to-nearest-even per reference op).
#include <stdlib.h>

#include "csb_common.cuh"

namespace csb {

// ---- block-wide exclusive scan of per-thread counts (blockDim <= 1024) -----------------------------
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot /* smem[32] */, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? warp_tot[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    warp_tot[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  return warp_tot[warp] + inc - v;
}

// ------------------------------------------------------------------------------------------------
// distance to own centroid from the Gram matrix, a warp per frame, two passes:
//   rowsum_i = sum_{m in M} G_im                     (M = members of assign[i])
//   dist_i^2 = G_ii - (2/n) rowsum_i + (1/n^2) sum_{m in M} rowsum_m
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gram_rowsum_kernel(const double* __restrict__ G, int64_t ldg, int T, const int64_t* __restrict__ assign,
                                                          const int32_t* __restrict__ offs, const int32_t* __restrict__ memb,
                                                          double* __restrict__ rowsum) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= T) return;
  const int j = (int)assign[i];
  const int beg = offs[j], end = offs[j + 1];
  const double* gi = G + (int64_t)i * ldg;
  double s = 0.0;
  for (int m = beg + lane; m < end; m += 32) s += gi[memb[m]];
  s = warp_sum(s);
  if (lane == 0) rowsum[i] = s;
}

__global__ void __launch_bounds__(256) dist_gram_kernel(const double* __restrict__ G, int64_t ldg, int T, const int64_t* __restrict__ assign,
                                                        const int32_t* __restrict__ offs, const int32_t* __restrict__ memb,
                                                        const double* __restrict__ rowsum, float* __restrict__ dist) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= T) return;
  const int j = (int)assign[i];
  const int beg = offs[j], end = offs[j + 1];
  double q = 0.0;
  for (int m = beg + lane; m < end; m += 32) q += rowsum[memb[m]];
  q = warp_sum(q);
  if (lane == 0) {
    const double nn = (double)(end - beg);
    const double d2 = G[(int64_t)i * ldg + i] - 2.0 * (rowsum[i] / nn) + q / (nn * nn);
    dist[i] = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
  }
}

// ------------------------------------------------------------------------------------------------
// representatives: a warp per event; n_rep rounds of "smallest (dist, index) greater than the last".
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clear_flags_kernel(uint8_t* f, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f[i] = 0;
}

__global__ void __launch_bounds__(256) select_reps_kernel(const float* __restrict__ dist, const int64_t* __restrict__ assign,
                                                          int T, int k, int n_rep, int64_t* __restrict__ reps,
                                                          int32_t* __restrict__ rep_counts, uint8_t* __restrict__ rep_flags) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= k) return;
  int cnt = 0;
  for (int i0 = 0; i0 < T; i0 += 32) {
    const int i = i0 + lane;
    cnt += __popc(__ballot_sync(0xffffffffu, i < T && (int)assign[i] == j));
  }
  int64_t* out = reps + (int64_t)j * n_rep;
  if (cnt <= n_rep) {
    // all members, ascending (torch.nonzero order), cogreasoner_chat.py:58-59
    int pos = 0;
    for (int i0 = 0; i0 < T; i0 += 32) {
      const int i = i0 + lane;
      const bool hit = i < T && (int)assign[i] == j;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        out[pos + __popc(bal & ((1u << lane) - 1u))] = i;
        rep_flags[i] = 1;
      }
      pos += __popc(bal);
    }
    for (int r = cnt + lane; r < n_rep; r += 32) out[r] = -1;
    if (lane == 0) rep_counts[j] = cnt;
    return;
  }
  float last_d = -1.f;
  int last_i = -1;
  for (int r = 0; r < n_rep; ++r) {
    float bd = __int_as_float(0x7f800000);
    int bi = 0x7fffffff;
    for (int i = lane; i < T; i += 32) {
      if ((int)assign[i] != j) continue;
      const float d = dist[i];
      const bool after = d > last_d || (d == last_d && i > last_i);
      if (after && (d < bd || (d == bd && i < bi))) { bd = d; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (lane == 0) {
      out[r] = bi == 0x7fffffff ? -1 : bi;
      if (bi != 0x7fffffff) rep_flags[bi] = 1;
    }
    last_d = bd;
    last_i = bi;
  }
  if (lane == 0) rep_counts[j] = n_rep;
}

// ------------------------------------------------------------------------------------------------
// Fused: distance of every member to its own centroid (from the Gram matrix) + the n_rep nearest members.
// A warp per event; the event's members are event_order[off[j] .. off[j+1]) (ascending frame index).
//   rowsum_m = sum_{m'} G[m][m'],  q = sum_m rowsum_m / n^2,  dist_m^2 = G_mm - 2 rowsum_m / n + q
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void reps_of_event(const double* __restrict__ G, int64_t ldg, int j, int n_rep,
                                              const int32_t* __restrict__ ev_off, const int32_t* __restrict__ order,
                                              float* __restrict__ dist, int64_t* __restrict__ reps,
                                              int32_t* __restrict__ rep_counts, uint8_t* __restrict__ rep_flags) {
  const int lane = threadIdx.x & 31;
  const int beg = ev_off[j], n = ev_off[j + 1] - beg;
  const int32_t* mem = order + beg;
  int64_t* out = reps + (int64_t)j * n_rep;
  // distances (also for small events: the dist output is complete).  The row sums of the first two rows of a lane (events of up
  // to 64 frames: all of them) stay in registers for the second pass; longer events recompute the rest.
  double qpart = 0.0, rs_keep[2] = {0.0, 0.0};
  for (int a = lane, it = 0; a < n; a += 32, ++it) {
    const double* ga = G + (int64_t)mem[a] * ldg;
    double rs = 0.0;
    int b = 0;
    for (; b + 4 <= n; b += 4) {  // 4 independent loads in flight, added in member order
      const double g0 = ga[mem[b]], g1 = ga[mem[b + 1]], g2 = ga[mem[b + 2]], g3 = ga[mem[b + 3]];
      rs += g0; rs += g1; rs += g2; rs += g3;
    }
    for (; b < n; ++b) rs += ga[mem[b]];
    if (it < 2) rs_keep[it] = rs;
    qpart += rs;
  }
  const double nn = (double)n;
  const double q = warp_sum(qpart) / (nn * nn);
  for (int a = lane, it = 0; a < n; a += 32, ++it) {
    const int m = mem[a];
    const double* ga = G + (int64_t)m * ldg;
    double rs;
    if (it < 2) {
      rs = it == 0 ? rs_keep[0] : rs_keep[1];
    } else {
      rs = 0.0;
      for (int b = 0; b < n; ++b) rs += ga[mem[b]];
    }
    const double d2 = ga[m] - 2.0 * (rs / nn) + q;
    dist[m] = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
    rep_flags[m] = 0;
  }
  __syncwarp();
  if (n <= n_rep) {  // all members, ascending (cogreasoner_chat.py:58-59)
    for (int a = lane; a < n; a += 32) { out[a] = mem[a]; rep_flags[mem[a]] = 1; }
    for (int r = n + lane; r < n_rep; r += 32) out[r] = -1;
    if (lane == 0) rep_counts[j] = n;
    return;
  }
  float last_d = -1.f;
  int last_i = -1;
  for (int r = 0; r < n_rep; ++r) {
    float bd = __int_as_float(0x7f800000);
    int bi = 0x7fffffff;
    for (int a = lane; a < n; a += 32) {
      const int i = mem[a];
      const float d = dist[i];
      const bool after = d > last_d || (d == last_d && i > last_i);
      if (after && (d < bd || (d == bd && i < bi))) { bd = d; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (lane == 0) {
      out[r] = bi == 0x7fffffff ? -1 : bi;
      if (bi != 0x7fffffff) rep_flags[bi] = 1;
    }
    last_d = bd;
    last_i = bi;
  }
  if (lane == 0) rep_counts[j] = n_rep;
}

__global__ void __launch_bounds__(256) reps_gram_kernel(const double* __restrict__ G, int64_t ldg, int k, int n_rep,
                                                        const int32_t* __restrict__ ev_off, const int32_t* __restrict__ order,
                                                        float* __restrict__ dist, int64_t* __restrict__ reps,
                                                        int32_t* __restrict__ rep_counts, uint8_t* __restrict__ rep_flags) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= k) return;
  reps_of_event(G, ldg, j, n_rep, ev_off, order, dist, reps, rep_counts, rep_flags);
}

// ------------------------------------------------------------------------------------------------
// chronological per-event order: a warp per event sweeps the assignments (stable, ascending).
// Single CTA (offsets need a scan over k counts).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) event_order_kernel(const int64_t* __restrict__ assign, int T, int k,
                                                           int32_t* __restrict__ order, int32_t* __restrict__ offsets) {
  extern __shared__ int s_cnt[];  // k + 1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int j = warp; j < k; j += nwarp) {
    int cnt = 0;
    for (int i0 = 0; i0 < T; i0 += 32) {
      const int i = i0 + lane;
      cnt += __popc(__ballot_sync(0xffffffffu, i < T && (int)assign[i] == j));
    }
    if (lane == 0) s_cnt[j] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int j = 0; j < k; ++j) {
      const int c = s_cnt[j];
      s_cnt[j] = run;
      offsets[j] = run;
      run += c;
    }
    s_cnt[k] = run;
    offsets[k] = run;
  }
  __syncthreads();
  for (int j = warp; j < k; j += nwarp) {
    int pos = s_cnt[j];
    for (int i0 = 0; i0 < T; i0 += 32) {
      const int i = i0 + lane;
      const bool hit = i < T && (int)assign[i] == j;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) order[pos + __popc(bal & ((1u << lane) - 1u))] = i;
      pos += __popc(bal);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// (c) cosine flags: one CTA per event.  bf16: every ATen op of F.cosine_similarity rounds to bf16:
//   n = bf16(sqrt(sum x^2)); xn = bf16(x / max(n, bf16(1e-8))); cos = bf16(sum bf16(qn * en)).
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ float ld_elem(const void* p, int i) {
  if (BF16) return bf16_bits_to_float(reinterpret_cast<const uint16_t*>(p)[i]);
  return reinterpret_cast<const float*>(p)[i];
}
template <bool BF16>
__device__ __forceinline__ float rnd(float v) { return BF16 ? round_bf16(v) : v; }

template <bool BF16>
__global__ void __launch_bounds__(256) cosine_flags_kernel(const void* __restrict__ q, const void* __restrict__ ev, int Dl,
                                                           float threshold, void* __restrict__ cos_out,
                                                           uint8_t* __restrict__ low) {
  __shared__ float red[32];
  const int j = blockIdx.x;
  const void* e = BF16 ? (const void*)(reinterpret_cast<const uint16_t*>(ev) + (int64_t)j * Dl)
                       : (const void*)(reinterpret_cast<const float*>(ev) + (int64_t)j * Dl);
  float sq = 0.f, se = 0.f;
  for (int i = threadIdx.x; i < Dl; i += blockDim.x) {
    const float a = ld_elem<BF16>(q, i), b = ld_elem<BF16>(e, i);
    sq = fmaf(a, a, sq);
    se = fmaf(b, b, se);
  }
  sq = block_sum<float>(sq, red);
  se = block_sum<float>(se, red);
  const float eps = rnd<BF16>(1e-8f);
  const float qn = fmaxf(rnd<BF16>(__fsqrt_rn(sq)), eps);
  const float en = fmaxf(rnd<BF16>(__fsqrt_rn(se)), eps);
  float dot = 0.f;
  for (int i = threadIdx.x; i < Dl; i += blockDim.x) {
    const float a = rnd<BF16>(__fdiv_rn(ld_elem<BF16>(q, i), qn));
    const float b = rnd<BF16>(__fdiv_rn(ld_elem<BF16>(e, i), en));
    dot += rnd<BF16>(__fmul_rn(a, b));
  }
  dot = block_sum<float>(dot, red);
  if (threadIdx.x == 0) {
    const float c = rnd<BF16>(dot);
    if (BF16) reinterpret_cast<uint16_t*>(cos_out)[j] = (uint16_t)float_to_bf16_bits(c);
    else reinterpret_cast<float*>(cos_out)[j] = c;
    low[j] = c < threshold ? 1 : 0;  // strict, threshold already rounded to the dtype (:329-330)
  }
}

// Fused relevance step: the k cosine CTAs, and the last one to finish derives the minor-frame list
// (cogreasoner_chat.py:325-333 in one launch).  `ticket` is a zero-initialised counter that the last CTA resets.
template <bool BF16>
__global__ void __launch_bounds__(256) relevance_minor_kernel(const void* __restrict__ q, const void* __restrict__ ev, int Dl,
                                                              float threshold, void* __restrict__ cos_out, uint8_t* __restrict__ low,
                                                              const int64_t* __restrict__ assign, const uint8_t* __restrict__ rep,
                                                              int T, uint8_t* __restrict__ flags, int32_t* __restrict__ list,
                                                              int64_t* __restrict__ n_minor, unsigned* ticket) {
  __shared__ float red[32];
  __shared__ int wt[32];
  __shared__ int total;
  __shared__ int is_last;
  const int j = blockIdx.x;
  const void* e = BF16 ? (const void*)(reinterpret_cast<const uint16_t*>(ev) + (int64_t)j * Dl)
                       : (const void*)(reinterpret_cast<const float*>(ev) + (int64_t)j * Dl);
  float sq = 0.f, se = 0.f;
  for (int i = threadIdx.x; i < Dl; i += blockDim.x) {
    const float a = ld_elem<BF16>(q, i), b = ld_elem<BF16>(e, i);
    sq = fmaf(a, a, sq);
    se = fmaf(b, b, se);
  }
  sq = block_sum<float>(sq, red);
  se = block_sum<float>(se, red);
  const float eps = rnd<BF16>(1e-8f);
  const float qn = fmaxf(rnd<BF16>(__fsqrt_rn(sq)), eps);
  const float en = fmaxf(rnd<BF16>(__fsqrt_rn(se)), eps);
  float dot = 0.f;
  for (int i = threadIdx.x; i < Dl; i += blockDim.x) {
    const float a = rnd<BF16>(__fdiv_rn(ld_elem<BF16>(q, i), qn));
    const float b = rnd<BF16>(__fdiv_rn(ld_elem<BF16>(e, i), en));
    dot += rnd<BF16>(__fmul_rn(a, b));
  }
  dot = block_sum<float>(dot, red);
  if (threadIdx.x == 0) {
    const float c = rnd<BF16>(dot);
    if (BF16) reinterpret_cast<uint16_t*>(cos_out)[j] = (uint16_t)float_to_bf16_bits(c);
    else reinterpret_cast<float*>(cos_out)[j] = c;
    low[j] = c < threshold ? 1 : 0;
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int per = (T + blockDim.x - 1) / blockDim.x;
  const int beg = threadIdx.x * per, end = min(T, beg + per);
  int cnt = 0;
  for (int i = beg; i < end; ++i) {
    const uint8_t f = (__ldcg(low + assign[i]) && !rep[i]) ? 1 : 0;
    flags[i] = f;
    cnt += f;
  }
  int pos = block_exclusive_scan(cnt, wt, &total);
  for (int i = beg; i < end; ++i)
    if (flags[i]) list[pos++] = i;
  if (threadIdx.x == 0) {
    *n_minor = total;
    *ticket = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// The whole selection step of phase 2 in ONE launch (cogreasoner_chat.py:284 + :317-318 + :325-333):
//   CTAs [0, k)            cosine of event j against the question; with `partials` the event embedding is first finished from
//                          the split partial sums of the var-len mean pool (sum in split order, one IEEE division, one rounding:
//                          the bits of csb_seq_mean_pool) -- the pooled row never makes a round trip through HBM
//   CTAs [k, k + ceil(k/8)) a warp per event: the n_rep members nearest to the centroid, from the Gram matrix
//   the last CTA to finish derives the minor-frame flags and the ascending list.
// ------------------------------------------------------------------------------------------------
constexpr int kRelSplits = 8;  // == kPoolSplits of embed_pool.cu

struct RelevanceArgs {
  const void* q;
  const void* ev;          // [k, Dl] or nullptr
  const float* partials;   // [k][kRelSplits][Dl] or nullptr
  const int32_t* cu;       // [k + 1] with partials
  int k, Dl;
  float threshold;
  const double* G;
  int64_t ldg;
  const int32_t* ev_off;
  const int32_t* order;
  int n_rep;
  const int64_t* assign;
  int T;
  void* ev_out;  // [k, Dl] with partials (nullable)
  void* cos_out;
  uint8_t* low;
  float* dist;
  int64_t* reps;
  int32_t* rep_counts;
  uint8_t* rep_flags;
  uint8_t* minor_flags;
  int32_t* minor_list;
  int64_t* n_minor;
  unsigned* ticket;
};

template <bool BF16>
__global__ void __launch_bounds__(256) relevance_select_kernel(RelevanceArgs a) {
  extern __shared__ float s_row[];  // [Dl] with partials
  __shared__ float red[32];
  __shared__ int wt[32];
  __shared__ int total;
  __shared__ int is_last;
  const int b = blockIdx.x;
  if (b < a.k) {
    const int Dl = a.Dl;
    const void* e = nullptr;
    if (a.partials) {
      const int len = a.cu[b + 1] - a.cu[b];
      const float* pp = a.partials + (int64_t)b * kRelSplits * Dl;
      for (int d = threadIdx.x; d < Dl; d += blockDim.x) {
        float v[kRelSplits];
#pragma unroll
        for (int sp = 0; sp < kRelSplits; ++sp) v[sp] = pp[(int64_t)sp * Dl + d];
        float sum = 0.f;
#pragma unroll
        for (int sp = 0; sp < kRelSplits; ++sp) sum += v[sp];
        const float m = rnd<BF16>(__fdiv_rn(sum, (float)len));  // torch.mean: fp32 sum, one division, one rounding
        s_row[d] = m;
        if (a.ev_out) {
          if (BF16) reinterpret_cast<uint16_t*>(a.ev_out)[(int64_t)b * Dl + d] = (uint16_t)float_to_bf16_bits(m);
          else reinterpret_cast<float*>(a.ev_out)[(int64_t)b * Dl + d] = m;
        }
      }
      __syncthreads();
    } else {
      e = BF16 ? (const void*)(reinterpret_cast<const uint16_t*>(a.ev) + (int64_t)b * Dl)
               : (const void*)(reinterpret_cast<const float*>(a.ev) + (int64_t)b * Dl);
    }
    auto ev_at = [&](int i) -> float { return e ? ld_elem<BF16>(e, i) : s_row[i]; };
    float sq = 0.f, se = 0.f;
    for (int i = threadIdx.x; i < Dl; i += blockDim.x) {
      const float x = ld_elem<BF16>(a.q, i), y = ev_at(i);
      sq = fmaf(x, x, sq);
      se = fmaf(y, y, se);
    }
    sq = block_sum<float>(sq, red);
    se = block_sum<float>(se, red);
    const float eps = rnd<BF16>(1e-8f);
    const float qn = fmaxf(rnd<BF16>(__fsqrt_rn(sq)), eps);
    const float en = fmaxf(rnd<BF16>(__fsqrt_rn(se)), eps);
    float dot = 0.f;
    for (int i = threadIdx.x; i < Dl; i += blockDim.x) {
      const float x = rnd<BF16>(__fdiv_rn(ld_elem<BF16>(a.q, i), qn));
      const float y = rnd<BF16>(__fdiv_rn(ev_at(i), en));
      dot += rnd<BF16>(__fmul_rn(x, y));
    }
    dot = block_sum<float>(dot, red);
    if (threadIdx.x == 0) {
      const float c = rnd<BF16>(dot);
      if (BF16) reinterpret_cast<uint16_t*>(a.cos_out)[b] = (uint16_t)float_to_bf16_bits(c);
      else reinterpret_cast<float*>(a.cos_out)[b] = c;
      a.low[b] = c < a.threshold ? 1 : 0;
    }
  } else {
    const int j = (b - a.k) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j < a.k) reps_of_event(a.G, a.ldg, j, a.n_rep, a.ev_off, a.order, a.dist, a.reps, a.rep_counts, a.rep_flags);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int T = a.T;
  const int per = (T + blockDim.x - 1) / blockDim.x;
  const int beg = threadIdx.x * per, end = min(T, beg + per);
  int cnt = 0;
  for (int i = beg; i < end; ++i) {
    const uint8_t f = (__ldcg(a.low + a.assign[i]) && !__ldcg(a.rep_flags + i)) ? 1 : 0;
    a.minor_flags[i] = f;
    cnt += f;
  }
  int pos = block_exclusive_scan(cnt, wt, &total);
  for (int i = beg; i < end; ++i)
    if (a.minor_flags[i]) a.minor_list[pos++] = i;
  if (threadIdx.x == 0) {
    *a.n_minor = total;
    *a.ticket = 0;
  }
}

// minor frames: flags + ascending list (single CTA scan)
__global__ void __launch_bounds__(1024) minor_frames_kernel(const int64_t* __restrict__ assign, const uint8_t* __restrict__ low,
                                                            const uint8_t* __restrict__ rep, int T,
                                                            uint8_t* __restrict__ flags, int32_t* __restrict__ list,
                                                            int64_t* __restrict__ n_minor) {
  __shared__ int wt[32];
  __shared__ int total;
  const int per = (T + blockDim.x - 1) / blockDim.x;
  const int beg = threadIdx.x * per, end = min(T, beg + per);
  int cnt = 0;
  for (int i = beg; i < end; ++i) {
    const uint8_t f = (low[assign[i]] && !rep[i]) ? 1 : 0;
    flags[i] = f;
    cnt += f;
  }
  int pos = block_exclusive_scan(cnt, wt, &total);
  for (int i = beg; i < end; ++i)
    if (flags[i]) list[pos++] = i;
  if (threadIdx.x == 0) *n_minor = total;
}

// ------------------------------------------------------------------------------------------------
// (d) pixel-difference mask.  A warp owns one (frame, token): W contiguous elements of the current and
// the previous frame, 128-bit loads; the previous frame's copy is an L2 hit (frames are swept in order).
// ------------------------------------------------------------------------------------------------
// sum over 8 bf16 lanes of | bf16(a - b) |, accumulated in fp32 in lane order
__device__ __forceinline__ float absdiff8_bf16(const uint4& a, const uint4& b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 x, y;
    memcpy(&x, &aw[i], 4);
    memcpy(&y, &bw[i], 4);
    const __nv_bfloat162 d = __hsub2(x, y);
    uint32_t bits;
    memcpy(&bits, &d, 4);
    s += fabsf(__uint_as_float(bits << 16));        // |.| folds into the FADD operand modifier
    s += fabsf(__uint_as_float(bits & 0xffff0000u));
  }
  return s;
}

template <bool BF16>
__global__ void __launch_bounds__(256) pixel_diff_kernel(const void* __restrict__ pix, int t, int P, int W, float thr_cmp,
                                                         float first_val, uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t tok = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= (int64_t)t * P) return;
  if (tok < P) {  // first frame: torch.full_like(threshold + 1) > threshold  (:411-412)
    if (lane == 0) mask[tok] = first_val > thr_cmp ? 1 : 0;
    return;
  }
  float s = 0.f;
  if (BF16) {
    const uint16_t* cur = reinterpret_cast<const uint16_t*>(pix) + tok * W;
    const uint16_t* prv = cur - (int64_t)P * W;
    if ((W & 7) == 0) {
      const uint4* c4 = reinterpret_cast<const uint4*>(cur);
      const uint4* p4 = reinterpret_cast<const uint4*>(prv);
      const int Wv = W >> 3;
      // packed bf16x2 subtract (one RN to bf16, == ATen's fp32 subtract + cast for bf16 operands), |.| by masking the
      // sign bits, fp32 accumulation; two vectors per lane in flight
      int v = lane;
      for (; v + 32 < Wv; v += 64) {
        const uint4 a0 = ldg_stream(c4 + v), a1 = ldg_stream(c4 + v + 32);
        const uint4 b0 = __ldg(p4 + v), b1 = __ldg(p4 + v + 32);
        s += absdiff8_bf16(a0, b0);
        s += absdiff8_bf16(a1, b1);
      }
      if (v < Wv) s += absdiff8_bf16(ldg_stream(c4 + v), __ldg(p4 + v));
    } else {
      for (int v = lane; v < W; v += 32) s += fabsf(round_bf16(__fsub_rn(bf16_bits_to_float(cur[v]), bf16_bits_to_float(prv[v]))));
    }
  } else {
    const float* cur = reinterpret_cast<const float*>(pix) + tok * W;
    const float* prv = cur - (int64_t)P * W;
    for (int v = lane; v < W; v += 32) s += fabsf(__fsub_rn(cur[v], prv[v]));
  }
  s = warp_sum(s);
  if (lane == 0) {
    // mean = sum / W (ATen CPU mean for reduced types: fp32 sum, div, one cast), then * 255, each op rounded
    float m = rnd<BF16>(__fdiv_rn(s, (float)W));
    m = rnd<BF16>(__fmul_rn(m, 255.f));
    mask[tok] = m > thr_cmp ? 1 : 0;
  }
}

// Fast bf16 path (TMA staging): a warp owns token p over a run of F consecutive frames.  Lane 0 streams the rows
// [f0-1, f1) of that token into a private 3-deep shared-memory ring (12 warps per CTA, one CTA per SM) with bulk async copies (cp.async.bulk ->
// mbarrier complete_tx), every lane pulls its 16-byte slices of the landed row into registers and the previous
// frame's row stays in registers, so each pixel row crosses L2->SM once (plus 1/F for the run boundary) instead
// of twice, with up to 3 rows per warp in flight independent of register pressure.
constexpr int kPixWarpsDefault = 12;
constexpr int kPixStagesDefault = 3;

__device__ __forceinline__ uint32_t pix_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NV, int kPixWarps, int kPixStages>
__global__ void __launch_bounds__(kPixWarps * 32, 1) pixel_diff_tma_kernel(const uint8_t* __restrict__ pix, int t, int P, int Wv,
                                                                           int W, int F, float thr_cmp, float first_val,
                                                                           uint8_t* __restrict__ mask, int evict_first) {
  extern __shared__ __align__(128) uint8_t pix_smem[];
  // the pixels are read exactly once: with `evict_first` their L2 lines are the first to go, so a latency-bound kernel of the same
  // stream that runs concurrently (the k-means solver on its 2.9 MB Gram matrix) keeps its working set in L2
  uint64_t l2_policy = 0;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_policy));
  __shared__ __align__(8) uint64_t bars[kPixWarps][kPixStages];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)Wv * 16u;
  uint8_t* ring = pix_smem + (size_t)warp * kPixStages * row_bytes;
  if (lane == 0) {
    for (int s = 0; s < kPixStages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pix_smem_u32(&bars[warp][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int nrun = (t + F - 1) / F;
  const int64_t nitems = (int64_t)nrun * P;
  const int64_t fstride = (int64_t)P * row_bytes;
  uint32_t it_base = 0;  // rows consumed by this warp so far (ring position / parity bookkeeping)
  for (int64_t item = (int64_t)blockIdx.x * kPixWarps + warp; item < nitems; item += (int64_t)gridDim.x * kPixWarps) {
    const int run = (int)(item / P), p = (int)(item - (int64_t)run * P);
    const int f0 = run * F, f1 = min(t, f0 + F);
    const int fa = f0 == 0 ? 0 : f0 - 1;
    const int nrows = f1 - fa;
    const uint8_t* src0 = pix + (int64_t)p * row_bytes + (int64_t)fa * fstride;
    auto issue = [&](int i) {  // lane 0 only
      const uint32_t slot = (it_base + (uint32_t)i) % kPixStages;
      const uint32_t bar = pix_smem_u32(&bars[warp][slot]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
      if (evict_first)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                         pix_smem_u32(ring + (size_t)slot * row_bytes)),
                     "l"(src0 + (int64_t)i * fstride), "r"(row_bytes), "r"(bar), "l"(l2_policy)
                     : "memory");
      else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         pix_smem_u32(ring + (size_t)slot * row_bytes)),
                     "l"(src0 + (int64_t)i * fstride), "r"(row_bytes), "r"(bar)
                     : "memory");
    };
    if (lane == 0)
      for (int i = 0; i < kPixStages && i < nrows; ++i) issue(i);
    uint4 prev[NV];
    for (int i = 0; i < nrows; ++i) {
      const uint32_t pos = it_base + (uint32_t)i;
      const uint32_t slot = pos % kPixStages, parity = (pos / kPixStages) & 1u;
      const uint32_t bar = pix_smem_u32(&bars[warp][slot]);
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "PW_%=:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
          "@p bra PD_%=;\n\t"
          "bra PW_%=;\n\t"
          "PD_%=:\n\t}" ::"r"(bar),
          "r"(parity)
          : "memory");
      const uint4* row = reinterpret_cast<const uint4*>(ring + (size_t)slot * row_bytes);
      uint4 cur[NV];
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int v = lane + 32 * u;
        cur[u] = v < Wv ? row[v] : make_uint4(0, 0, 0, 0);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // our reads of the slot precede the next bulk write to it
      __syncwarp();
      if (lane == 0 && i + kPixStages < nrows) issue(i + kPixStages);
      const int f = fa + i;
      if (i == 0 && f0 == 0) {
        if (lane == 0) mask[p] = first_val > thr_cmp ? 1 : 0;  // :411-412
      } else if (i > 0) {
        float sacc = 0.f;
#pragma unroll
        for (int u = 0; u < NV; ++u) sacc += absdiff8_bf16(cur[u], prev[u]);
        sacc = warp_sum(sacc);
        if (lane == 0) {
          float m = round_bf16(__fdiv_rn(sacc, (float)W));
          m = round_bf16(__fmul_rn(m, 255.f));
          mask[(int64_t)f * P + p] = m > thr_cmp ? 1 : 0;
        }
      }
#pragma unroll
      for (int u = 0; u < NV; ++u) prev[u] = cur[u];
    }
    it_base += (uint32_t)nrows;
  }
}

// per-frame rules: min_tokens padding (:413-414) then the minor-frame override (:416-422). A warp per frame.
__global__ void __launch_bounds__(256) mask_finalize_kernel(uint8_t* __restrict__ mask, int t, int P, int min_tokens,
                                                            const uint8_t* __restrict__ minor_flags, int64_t frame0,
                                                            const uint8_t* __restrict__ kind) {
  const int lane = threadIdx.x & 31;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= t) return;
  uint8_t* m = mask + (int64_t)f * P;
  if (kind && kind[frame0 + f]) {  // image item / single-frame video: all tokens kept
    for (int p = lane; p < P; p += 32) m[p] = 1;
    return;
  }
  if (minor_flags && minor_flags[frame0 + f]) {
    for (int p = lane; p < P; p += 32) m[p] = (p == 0) ? 1 : 0;
    return;
  }
  int cnt = 0;
  for (int p = lane; p < P; p += 32) cnt += m[p];
  cnt = warp_sum(cnt);
  if (cnt < min_tokens)
    for (int p = lane; p < min_tokens && p < P; p += 32) m[p] = 1;
}

// ------------------------------------------------------------------------------------------------
// compaction: scan (single CTA), gather kept rows, emit pooled rows for pooled frames.
// ------------------------------------------------------------------------------------------------
// Two-level scan: a CTA owns a chunk of 1024 elements (256 threads x 4); the last CTA to finish turns the
// chunk counts into exclusive prefixes ("last block done"), so one launch yields counts + prefixes + total.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ void finish_chunk_scan(int* counts, int nchunks, unsigned* ticket, int64_t* total_out, int* wt, int* stotal) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  int carry = 0;
  for (int base = 0; base < nchunks; base += blockDim.x) {
    const int b = base + threadIdx.x;
    const int v = b < nchunks ? __ldcg(counts + b) : 0;
    const int ex = block_exclusive_scan(v, wt, stotal);
    if (b < nchunks) counts[b] = carry + ex;
    carry += *stotal;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[nchunks] = carry;
    if (total_out) *total_out = carry;
    *ticket = 0;  // ready for the next scan on this workspace
  }
}

__global__ void __launch_bounds__(kScanThreads) mask_chunk_count_kernel(const uint8_t* __restrict__ mask, int64_t n,
                                                                         int* __restrict__ counts, int nchunks,
                                                                         unsigned* ticket, int64_t* __restrict__ n_kept) {
  __shared__ int wt[32];
  __shared__ int stotal;
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  int cnt = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u)
    if (base + u < n) cnt += mask[base + u] ? 1 : 0;
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) wt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += wt[w];
    counts[blockIdx.x] = t;
  }
  finish_chunk_scan(counts, nchunks, ticket, n_kept, wt, &stotal);
}

// positions + kept_index of this chunk (the row copy itself is csb_gather_rows: many CTAs, 4 loads in flight)
__global__ void __launch_bounds__(kScanThreads) compact_scatter_kernel(const uint8_t* __restrict__ mask, int64_t n,
                                                                        const int* __restrict__ prefix, int32_t* __restrict__ pos,
                                                                        int32_t* __restrict__ kept_index) {
  __shared__ int wt[32];
  __shared__ int stotal;
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  int f[kScanItems], cnt = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    f[u] = (base + u < n && mask[base + u]) ? 1 : 0;
    cnt += f[u];
  }
  int lp = prefix[blockIdx.x] + block_exclusive_scan(cnt, wt, &stotal);
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    if (base + u < n) {
      pos[base + u] = lp;
      if (f[u]) kept_index[lp++] = (int32_t)(base + u);
    }
  }
}

// Pooled token of every pooled frame, written straight into its slot of the compacted output.  A CTA owns a frame;
// kPoolGroups thread groups each add a contiguous quarter of the frame's tokens (ascending) so that enough loads are
// in flight, then group 0 adds the four partial sums in order: fp32 accumulation, one round-to-nearest to bf16.
constexpr int kPoolGroups = 4;

// Blocks [0, T) pool one flagged frame each; with `kept_index` given, blocks [T, gridDim.x) copy the kept rows (the row gather
// of the compaction) in the same launch -- the small gather rides under the HBM-bound pooling pass.  The gather blocks skip
// token 0 of a pooled frame: that output row belongs to the pooling block.
__global__ void __launch_bounds__(1024) pool_emit_kernel(const uint4* __restrict__ x, int P, int Dv,
                                                         const uint8_t* __restrict__ pool_flags,
                                                         const uint8_t* __restrict__ mask, const int32_t* __restrict__ pos,
                                                         uint4* __restrict__ out, int T, const int32_t* __restrict__ kept_index,
                                                         const int64_t* __restrict__ n_kept) {
  extern __shared__ float pool_part[];  // [kPoolGroups - 1][blockDim.x][8]
  if ((int)blockIdx.x >= T) {
    const int64_t nthreads = (int64_t)blockDim.x * blockDim.y;
    const int64_t nblk = gridDim.x - T;
    const int64_t tid = ((int64_t)blockIdx.x - T) * nthreads + threadIdx.y * blockDim.x + threadIdx.x;
    const int64_t total = *n_kept * Dv;
    const int64_t stride = nblk * nthreads;
    for (int64_t g0 = tid; g0 < total; g0 += 4 * stride) {
      uint4 v[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t g = g0 + u * stride;
        ok[u] = false;
        if (g < total) {
          const int64_t r = g / Dv, c = g - r * Dv;
          const int tok = kept_index[r];
          const int f = tok / P;
          ok[u] = !(pool_flags[f] && tok == f * P);
          if (ok[u]) v[u] = ldg_stream(x + (int64_t)tok * Dv + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (ok[u]) out[g0 + u * stride] = v[u];
    }
    return;
  }
  const int64_t t = blockIdx.x;
  if (!pool_flags[t] || !mask[t * P]) return;
  const int g = threadIdx.y;
  const int per = (P + kPoolGroups - 1) / kPoolGroups;
  const int p0 = min(P, g * per), p1 = min(P, p0 + per);
  const uint4* xf = x + t * (int64_t)P * Dv;
  for (int c0 = 0; c0 < Dv; c0 += blockDim.x) {  // uniform trip count: barriers inside are safe
    const int c = c0 + threadIdx.x;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < Dv) {
      int p = p0;
      for (; p + 8 <= p1; p += 8) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = ldg_stream(xf + (int64_t)(p + u) * Dv + c);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += f[e];
        }
      }
      for (; p < p1; ++p) {
        float f[8];
        unpack8(ldg_stream(xf + (int64_t)p * Dv + c), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
      if (g > 0) {
        float* dst = pool_part + ((size_t)(g - 1) * blockDim.x + threadIdx.x) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) dst[e] = acc[e];
      }
    }
    __syncthreads();
    if (g == 0 && c < Dv) {
#pragma unroll
      for (int gg = 1; gg < kPoolGroups; ++gg) {
        const float* src = pool_part + ((size_t)(gg - 1) * blockDim.x + threadIdx.x) * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += src[e];
      }
      float m[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = __fdiv_rn(acc[e], (float)P);
      uint4 o;
      o.x = float_to_bf16_bits(m[0]) | (float_to_bf16_bits(m[1]) << 16);
      o.y = float_to_bf16_bits(m[2]) | (float_to_bf16_bits(m[3]) << 16);
      o.z = float_to_bf16_bits(m[4]) | (float_to_bf16_bits(m[5]) << 16);
      o.w = float_to_bf16_bits(m[6]) | (float_to_bf16_bits(m[7]) << 16);
      out[(int64_t)pos[t * P] * Dv + c] = o;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// input-id filtering (:459-474), chunked: (1) image-token prefix, (2) keep prefix, (3) scatter.
// keep[i] = !is_image[i] || mask[rank_of_image(i)]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) ids_image_count_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t image_tok,
                                                                        int* __restrict__ counts, int nchunks, unsigned* ticket,
                                                                        int64_t* __restrict__ n_image) {
  __shared__ int wt[32];
  __shared__ int stotal;
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  int cnt = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u)
    if (base + u < n) cnt += ids[base + u] == image_tok;
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) wt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += wt[w];
    counts[blockIdx.x] = t;
  }
  finish_chunk_scan(counts, nchunks, ticket, n_image, wt, &stotal);
}

template <bool SCATTER>
__global__ void __launch_bounds__(kScanThreads) ids_keep_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t image_tok,
                                                                 const uint8_t* __restrict__ mask, int64_t n_mask,
                                                                 const int* __restrict__ img_prefix, int* __restrict__ keep_counts,
                                                                 int nchunks, unsigned* ticket, int64_t* __restrict__ n_out,
                                                                 const int64_t* __restrict__ attn, const int64_t* __restrict__ labels,
                                                                 int64_t* __restrict__ o_ids, int64_t* __restrict__ o_attn,
                                                                 int64_t* __restrict__ o_labels) {
  __shared__ int wt[32];
  __shared__ int stotal;
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  int64_t v[kScanItems];
  int img[kScanItems], icnt = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    v[u] = base + u < n ? ids[base + u] : 0;
    img[u] = (base + u < n && v[u] == image_tok) ? 1 : 0;
    icnt += img[u];
  }
  int r = img_prefix[blockIdx.x] + block_exclusive_scan(icnt, wt, &stotal);
  int keep[kScanItems], kcnt = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    keep[u] = base + u < n ? 1 : 0;
    if (img[u]) {
      keep[u] = (r < n_mask && mask[r]) ? 1 : 0;
      ++r;
    }
    kcnt += keep[u];
  }
  if (!SCATTER) {
    kcnt = warp_sum(kcnt);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) wt[threadIdx.x >> 5] = kcnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < kScanThreads / 32; ++w) t += wt[w];
      keep_counts[blockIdx.x] = t;
    }
    finish_chunk_scan(keep_counts, nchunks, ticket, n_out, wt, &stotal);
  } else {
    int p = keep_counts[blockIdx.x] + block_exclusive_scan(kcnt, wt, &stotal);
#pragma unroll
    for (int u = 0; u < kScanItems; ++u) {
      if (keep[u]) {
        o_ids[p] = v[u];
        if (attn) o_attn[p] = attn[base + u];
        if (labels) o_labels[p] = labels[base + u];
        ++p;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused variants (one cooperative launch instead of 2-3 dependent launches): the chunk counts are published, the CTAs meet at
// a software grid barrier (co-residency guaranteed by the cooperative launch; the grid is ceil(n / 1024) CTAs) and every
// CTA sums the counts of its predecessors itself.  Same outputs as the kernels above, bit for bit (integer work).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void scan_grid_sync(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (gridDim.x > 1) {
    if (threadIdx.x == 0) {
      target += gridDim.x;
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

// sum of counts[0 .. blockIdx.x) by the whole CTA (fixed tree; integers, so the order is irrelevant)
__device__ __forceinline__ int chunk_prefix(const int* counts, int* wt, int* stotal) {
  int v = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += blockDim.x) v += __ldcg(counts + b);
  block_exclusive_scan(v, wt, stotal);
  const int t = *stotal;
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kScanThreads) ids_filter_coop_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t image_tok,
                                                                        const uint8_t* __restrict__ mask, int64_t n_mask,
                                                                        int* __restrict__ img_counts, int* __restrict__ keep_counts,
                                                                        unsigned* barrier, int64_t* __restrict__ n_out,
                                                                        const int64_t* __restrict__ attn,
                                                                        const int64_t* __restrict__ labels,
                                                                        int64_t* __restrict__ o_ids, int64_t* __restrict__ o_attn,
                                                                        int64_t* __restrict__ o_labels) {
  __shared__ int wt[32];
  __shared__ int stotal;
  unsigned bar = 0;
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  int64_t v[kScanItems];
  int img[kScanItems], icnt = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    v[u] = base + u < n ? ids[base + u] : 0;
    img[u] = (base + u < n && v[u] == image_tok) ? 1 : 0;
    icnt += img[u];
  }
  const int iex = block_exclusive_scan(icnt, wt, &stotal);
  const int itot = stotal;
  if (threadIdx.x == 0) img_counts[blockIdx.x] = itot;
  scan_grid_sync(barrier, bar);
  const int ipre = chunk_prefix(img_counts, wt, &stotal);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) n_out[1] = (int64_t)ipre + itot;  // <image> ids seen
  int r = ipre + iex;  // rank of this thread's first <image> id
  int keep[kScanItems], kcnt = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    keep[u] = base + u < n ? 1 : 0;
    if (img[u]) {
      keep[u] = (r < n_mask && mask[r]) ? 1 : 0;
      ++r;
    }
    kcnt += keep[u];
  }
  const int kex = block_exclusive_scan(kcnt, wt, &stotal);
  const int ktot = stotal;
  if (threadIdx.x == 0) keep_counts[blockIdx.x] = ktot;
  scan_grid_sync(barrier, bar);
  const int kpre = chunk_prefix(keep_counts, wt, &stotal);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_out = (int64_t)kpre + ktot;
  int p = kpre + kex;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    if (keep[u]) {
      o_ids[p] = v[u];
      if (attn) o_attn[p] = attn[base + u];
      if (labels) o_labels[p] = labels[base + u];
      ++p;
    }
  }
}

__global__ void __launch_bounds__(kScanThreads) compact_scan_coop_kernel(const uint8_t* __restrict__ mask, int64_t n,
                                                                          int* __restrict__ counts, unsigned* barrier,
                                                                          int64_t* __restrict__ n_kept, int32_t* __restrict__ pos,
                                                                          int32_t* __restrict__ kept_index) {
  __shared__ int wt[32];
  __shared__ int stotal;
  unsigned bar = 0;
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + threadIdx.x * kScanItems;
  int f[kScanItems], cnt = 0;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    f[u] = (base + u < n && mask[base + u]) ? 1 : 0;
    cnt += f[u];
  }
  const int ex = block_exclusive_scan(cnt, wt, &stotal);
  const int tot = stotal;
  if (threadIdx.x == 0) counts[blockIdx.x] = tot;
  scan_grid_sync(barrier, bar);
  const int pre = chunk_prefix(counts, wt, &stotal);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_kept = (int64_t)pre + tot;
  int lp = pre + ex;
#pragma unroll
  for (int u = 0; u < kScanItems; ++u) {
    if (base + u < n) {
      pos[base + u] = lp;
      if (f[u]) kept_index[lp++] = (int32_t)(base + u);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Phase-2 tail in ONE cooperative launch (cogreasoner_chat.py:413-422 + :458-474): per-frame mask rules (min_tokens padding,
// minor-frame override), the token scan (positions, kept_index, n_kept) and the id filter (image rank -> keep -> scatter).
// Token chunks are whole frames (fpc = 1024 / P frames per CTA) so that the per-frame rules are CTA-local; id chunks are 1024
// ids.  Two grid barriers; the barrier word is never reset: every launch starts from the epoch the previous one left behind
// (block 0 publishes it after the last barrier), so no memset node precedes the kernel.
// ------------------------------------------------------------------------------------------------
struct CompactStreamArgs {
  uint8_t* mask;  // [T*P] in: raw difference mask, out: final mask
  int T, P, fpc, n_tok_chunks, min_tokens, finalize;
  const uint8_t* minor;  // [T] or nullptr
  const uint8_t* kind;   // [T] or nullptr: 1 = image-like frame (all tokens kept)
  int* counts;
  int32_t* pos;
  int32_t* kept_index;
  int64_t* n_kept;
  const int64_t* ids;
  int64_t n_ids;
  int64_t image_tok;
  int n_id_chunks;
  const int64_t* attn;
  const int64_t* labels;
  int* img_counts;
  int* keep_counts;
  int64_t* o_ids;
  int64_t* o_attn;
  int64_t* o_labels;
  int64_t* n_out;  // [2]: kept ids, <image> ids seen
  unsigned* barrier;  // [0] arrival counter, [1] epoch (value of the counter when the launch started)
};

__device__ __forceinline__ void epoch_grid_sync(unsigned* counter, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while ((int)(v - target) < 0);
  }
  __syncthreads();
}

// sum of counts[0 .. c) by the whole CTA
__device__ __forceinline__ int chunk_prefix_to(const int* counts, int c, int* wt, int* stotal) {
  int v = 0;
  for (int b = threadIdx.x; b < c; b += blockDim.x) v += __ldcg(counts + b);
  block_exclusive_scan(v, wt, stotal);
  const int t = *stotal;
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kScanThreads) compact_stream_coop_kernel(CompactStreamArgs a) {
  __shared__ int wt[32];
  __shared__ int stotal;
  __shared__ uint8_t s_m[kScanChunk];
  const int c = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned bar = *reinterpret_cast<volatile unsigned*>(a.barrier + 1);
  const bool tok = c < a.n_tok_chunks, idc = c < a.n_id_chunks;
  const int P = a.P;
  // ---- A. per-frame rules + chunk counts ----
  const int f0 = c * a.fpc, f1 = min(a.T, f0 + a.fpc);
  const int64_t tbase = (int64_t)f0 * P;
  const int ntok = tok ? (f1 - f0) * P : 0;
  int f[kScanItems], cnt = 0, ex = 0, tot = 0;
  if (tok) {
    for (int fr = f0 + warp; fr < f1; fr += kScanThreads / 32) {
      uint8_t* m = a.mask + (int64_t)fr * P;
      uint8_t* sm = s_m + (fr - f0) * P;
      if (!a.finalize) {
        for (int p = lane; p < P; p += 32) sm[p] = m[p] ? 1 : 0;
      } else if (a.kind && a.kind[fr]) {  // :401-403 image / single-frame video: every token kept, no per-frame rule
        for (int p = lane; p < P; p += 32) { sm[p] = 1; m[p] = 1; }
      } else if (a.minor && a.minor[fr]) {  // :416-422 minor frames keep only token 0
        for (int p = lane; p < P; p += 32) { const uint8_t v = p == 0 ? 1 : 0; sm[p] = v; m[p] = v; }
      } else {
        int kept = 0;
        for (int p = lane; p < P; p += 32) { const uint8_t v = m[p] ? 1 : 0; sm[p] = v; kept += v; }
        kept = warp_sum(kept);
        if (kept < a.min_tokens)  // :413-414
          for (int p = lane; p < a.min_tokens && p < P; p += 32) { sm[p] = 1; m[p] = 1; }
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kScanItems; ++u) {
      const int sl = threadIdx.x * kScanItems + u;
      f[u] = (sl < ntok && s_m[sl]) ? 1 : 0;
      cnt += f[u];
    }
    ex = block_exclusive_scan(cnt, wt, &stotal);
    tot = stotal;
    if (threadIdx.x == 0) a.counts[c] = tot;
    __syncthreads();
  }
  const int64_t ibase = (int64_t)c * kScanChunk + threadIdx.x * kScanItems;
  int64_t v[kScanItems];
  int img[kScanItems], icnt = 0, iex = 0, itot = 0;
  if (idc) {
#pragma unroll
    for (int u = 0; u < kScanItems; ++u) {
      v[u] = ibase + u < a.n_ids ? a.ids[ibase + u] : 0;
      img[u] = (ibase + u < a.n_ids && v[u] == a.image_tok) ? 1 : 0;
      icnt += img[u];
    }
    iex = block_exclusive_scan(icnt, wt, &stotal);
    itot = stotal;
    if (threadIdx.x == 0) a.img_counts[c] = itot;
  }
  epoch_grid_sync(a.barrier, bar);
  // ---- B. token positions; id keep flags ----
  if (tok) {
    const int pre = chunk_prefix_to(a.counts, c, wt, &stotal);
    if (c == a.n_tok_chunks - 1 && threadIdx.x == 0) *a.n_kept = (int64_t)pre + tot;
    int lp = pre + ex;
#pragma unroll
    for (int u = 0; u < kScanItems; ++u) {
      const int sl = threadIdx.x * kScanItems + u;
      if (sl < ntok) {
        a.pos[tbase + sl] = lp;
        if (f[u]) a.kept_index[lp++] = (int32_t)(tbase + sl);
      }
    }
  }
  int keep[kScanItems], kcnt = 0, kex = 0, ktot = 0;
  if (idc) {
    const int ipre = chunk_prefix_to(a.img_counts, c, wt, &stotal);
    if (c == a.n_id_chunks - 1 && threadIdx.x == 0) a.n_out[1] = (int64_t)ipre + itot;
    int r = ipre + iex;
    const int64_t n_mask = (int64_t)a.T * P;
#pragma unroll
    for (int u = 0; u < kScanItems; ++u) {
      keep[u] = ibase + u < a.n_ids ? 1 : 0;
      if (img[u]) {
        keep[u] = (r < n_mask && __ldcg(a.mask + r)) ? 1 : 0;  // the final mask was written by other CTAs: read it from L2
        ++r;
      }
      kcnt += keep[u];
    }
    kex = block_exclusive_scan(kcnt, wt, &stotal);
    ktot = stotal;
    if (threadIdx.x == 0) a.keep_counts[c] = ktot;
  }
  epoch_grid_sync(a.barrier, bar);
  if (c == 0 && threadIdx.x == 0) a.barrier[1] = bar;  // every CTA has read the old epoch long ago: publish the next one
  // ---- C. id scatter ----
  if (idc) {
    const int kpre = chunk_prefix_to(a.keep_counts, c, wt, &stotal);
    if (c == a.n_id_chunks - 1 && threadIdx.x == 0) a.n_out[0] = (int64_t)kpre + ktot;
    int p = kpre + kex;
#pragma unroll
    for (int u = 0; u < kScanItems; ++u) {
      if (keep[u]) {
        a.o_ids[p] = v[u];
        if (a.attn) a.o_attn[p] = a.attn[ibase + u];
        if (a.labels) a.o_labels[p] = a.labels[ibase + u];
        ++p;
      }
    }
  }
}

// number of CTAs of `kernel` (kScanThreads threads, no dynamic shared memory) that can be co-resident on the current device
template <typename K>
static int coop_capacity(K kernel) {
  int per_sm = 0, dev = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kScanThreads, 0) != cudaSuccess) return 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return per_sm * sms;
}

// ------------------------------------------------------------------------------------------------
// input-id filtering with position-id rebuild (training-side path, rare): single CTA.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) filter_ids_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t image_tok,
                                                          const uint8_t* __restrict__ mask, int64_t n_mask,
                                                          const int64_t* __restrict__ attn, const int64_t* __restrict__ labels,
                                                          const int64_t* __restrict__ posids, int64_t* __restrict__ o_ids,
                                                          int64_t* __restrict__ o_attn, int64_t* __restrict__ o_labels,
                                                          int64_t* __restrict__ o_pos, int64_t* __restrict__ n_out,
                                                          int32_t* __restrict__ seg_start /* ws [n] */) {
  __shared__ int wt[32];
  __shared__ int total;
  const int64_t per = (n + blockDim.x - 1) / blockDim.x;
  const int64_t beg = threadIdx.x * per, end = min(n, beg + per);
  // pass 1: rank among image tokens
  int cnt = 0;
  for (int64_t i = beg; i < end; ++i) cnt += ids[i] == image_tok;
  int irank = block_exclusive_scan(cnt, wt, &total);
  const int img_total = total;
  // pass 2: kept count
  int kept = 0, r = irank;
  for (int64_t i = beg; i < end; ++i) {
    bool keep = true;
    if (ids[i] == image_tok) { keep = r < n_mask ? mask[r] != 0 : false; ++r; }
    kept += keep;
  }
  int opos = block_exclusive_scan(kept, wt, &total);
  const int out_total = total;
  r = irank;
  for (int64_t i = beg; i < end; ++i) {
    bool keep = true;
    if (ids[i] == image_tok) { keep = r < n_mask ? mask[r] != 0 : false; ++r; }
    if (keep) {
      o_ids[opos] = ids[i];
      if (attn) o_attn[opos] = attn[i];
      if (labels) o_labels[opos] = labels[i];
      if (posids) seg_start[opos] = (opos == 0 || posids[i] == 0) ? opos : -1;
      ++opos;
    }
  }
  if (threadIdx.x == 0) { n_out[0] = out_total; n_out[1] = img_total; }
  if (!posids) return;
  // position ids restart at every kept element whose old position id was 0 (:472-474): running max of starts
  __syncthreads();
  const int64_t oper = ((int64_t)out_total + blockDim.x - 1) / blockDim.x;
  const int64_t ob = threadIdx.x * oper, oe = min((int64_t)out_total, ob + oper);
  int last = -1;
  for (int64_t i = ob; i < oe; ++i) last = max(last, seg_start[i]);
  // inclusive max-scan across threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  int inc = last;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc = max(inc, v);
  }
  if (lane == 31) wt[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? wt[lane] : -1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w = max(w, v);
    }
    wt[lane] = w;
  }
  __syncthreads();
  int carry = warp > 0 ? wt[warp - 1] : -1;
  int prev = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane > 0) carry = max(carry, prev);
  for (int64_t i = ob; i < oe; ++i) {
    carry = max(carry, seg_start[i]);
    o_pos[i] = i - carry;
  }
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace csb

using namespace csb;

extern "C" {

int csb_dist_to_assigned_gram(const double* G, int64_t ldg, int64_t T, int64_t k, const int64_t* assignments, const int32_t* offs,
                              const int32_t* memb, float* dist, double* rowsum_ws, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(G && assignments && offs && memb && dist && rowsum_ws && T > 0 && k > 0 && ldg >= T, "dist_to_assigned_gram: null argument or ldg < T");
  cudaStream_t s = as_stream(stream);
  gram_rowsum_kernel<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(G, ldg, (int)T, assignments, offs, memb, rowsum_ws);
  count_launch(1);
  dist_gram_kernel<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(G, ldg, (int)T, assignments, offs, memb, rowsum_ws, dist);
  return check_launch("dist_to_assigned_gram");
}

int csb_representatives_gram(const double* G, int64_t ldg, int64_t T, int64_t k, const int32_t* event_offsets,
                             const int32_t* order, int64_t n_rep, float* dist, int64_t* reps, int32_t* rep_counts,
                             uint8_t* rep_flags, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(G && event_offsets && order && dist && reps && rep_counts && rep_flags && T > 0 && k > 0 && n_rep > 0 && ldg >= T,
              "representatives_gram: null argument or ldg < T");
  reps_gram_kernel<<<(unsigned)((k + 7) / 8), 256, 0, as_stream(stream)>>>(G, ldg, (int)k, (int)n_rep, event_offsets, order, dist, reps,
                                                                            rep_counts, rep_flags);
  return check_launch("representatives_gram");
}

int csb_relevance_minor(const void* question, const void* events, int64_t k, int64_t Dl, int32_t dtype, double threshold,
                        const int64_t* assignments, const uint8_t* rep_flags, int64_t T, void* cos_out, uint8_t* low_flags,
                        uint8_t* minor_flags, int32_t* minor_list, int64_t* n_minor, uint32_t* ticket, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(question && events && cos_out && low_flags && assignments && rep_flags && minor_flags && minor_list && n_minor && ticket &&
                  k > 0 && Dl > 0 && T > 0,
              "relevance_minor: null argument");
  cudaStream_t s = as_stream(stream);
  if (dtype == CSB_BF16) {
    const float thr = __bfloat162float(__float2bfloat16_rn((float)threshold));
    relevance_minor_kernel<true><<<(unsigned)k, 256, 0, s>>>(question, events, (int)Dl, thr, cos_out, low_flags, assignments, rep_flags,
                                                            (int)T, minor_flags, minor_list, n_minor, ticket);
  } else if (dtype == CSB_F32) {
    relevance_minor_kernel<false><<<(unsigned)k, 256, 0, s>>>(question, events, (int)Dl, (float)threshold, cos_out, low_flags, assignments,
                                                             rep_flags, (int)T, minor_flags, minor_list, n_minor, ticket);
  } else {
    set_error("relevance_minor: dtype must be bf16 or f32");
    return CSB_ERR_UNSUPPORTED;
  }
  return check_launch("relevance_minor");
}

int csb_select_representatives(const float* dist, const int64_t* assignments, int64_t T, int64_t k, int64_t n_rep, int64_t* reps,
                               int32_t* rep_counts, uint8_t* rep_flags, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(dist && assignments && reps && rep_counts && rep_flags && T > 0 && k > 0 && n_rep > 0, "select_representatives: null argument");
  cudaStream_t s = as_stream(stream);
  clear_flags_kernel<<<(unsigned)((T + 255) / 256), 256, 0, s>>>(rep_flags, T);
  count_launch(1);
  select_reps_kernel<<<(unsigned)((k + 7) / 8), 256, 0, s>>>(dist, assignments, (int)T, (int)k, (int)n_rep, reps, rep_counts, rep_flags);
  return check_launch("select_representatives");
}

int csb_event_order(const int64_t* assignments, int64_t T, int64_t k, int32_t* order, int32_t* event_offsets, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(assignments && order && event_offsets && T > 0 && k > 0, "event_order: null argument");
  CSB_REQUIRE(k <= 10000, "event_order: k too large");
  event_order_kernel<<<1, 1024, sizeof(int) * (k + 1), as_stream(stream)>>>(assignments, (int)T, (int)k, order, event_offsets);
  return check_launch("event_order");
}

int csb_cosine_flags(const void* question, const void* events, int64_t k, int64_t Dl, int32_t dtype, double threshold,
                     void* cos_out, uint8_t* low_flags, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(question && events && cos_out && low_flags && k > 0 && Dl > 0, "cosine_flags: null argument");
  cudaStream_t s = as_stream(stream);
  if (dtype == CSB_BF16) {
    // `tensor_bf16 < python_float` compares against the scalar rounded to bf16
    const float thr = __bfloat162float(__float2bfloat16_rn((float)threshold));
    cosine_flags_kernel<true><<<(unsigned)k, 256, 0, s>>>(question, events, (int)Dl, thr, cos_out, low_flags);
  } else if (dtype == CSB_F32) {
    cosine_flags_kernel<false><<<(unsigned)k, 256, 0, s>>>(question, events, (int)Dl, (float)threshold, cos_out, low_flags);
  } else {
    set_error("cosine_flags: dtype must be bf16 or f32");
    return CSB_ERR_UNSUPPORTED;
  }
  return check_launch("cosine_flags");
}

int csb_minor_frames(const int64_t* assignments, const uint8_t* low_flags, const uint8_t* rep_flags, int64_t T,
                     uint8_t* minor_flags, int32_t* minor_list, int64_t* n_minor, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(assignments && low_flags && rep_flags && minor_flags && minor_list && n_minor && T > 0, "minor_frames: null argument");
  minor_frames_kernel<<<1, 1024, 0, as_stream(stream)>>>(assignments, low_flags, rep_flags, (int)T, minor_flags, minor_list, n_minor);
  return check_launch("minor_frames");
}

int csb_pixel_diff(const void* pixels, int64_t t, int64_t P, int64_t W, int32_t dtype, double threshold, uint8_t* mask,
                   csb_stream_t stream) {
  return csb_pixel_diff_ex(pixels, t, P, W, dtype, threshold, mask, 0, CSB_PIXEL_EVICT_FIRST, stream);
}

int csb_pixel_diff_ex(const void* pixels, int64_t t, int64_t P, int64_t W, int32_t dtype, double threshold, uint8_t* mask,
                      int64_t max_ctas, int32_t flags, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(pixels && mask && t > 0 && P > 0 && W > 0, "pixel_diff: null argument or empty shape");
  CSB_REQUIRE(W < (1ll << 30) && P < (1ll << 30) && t < (1ll << 30), "pixel_diff: shape too large");
  cudaStream_t s = as_stream(stream);
  const int64_t ntok = t * P;
  const unsigned blocks = (unsigned)((ntok + 7) / 8);
  if (dtype == CSB_BF16) {
    CSB_REQUIRE((W % 8 != 0) || (reinterpret_cast<uintptr_t>(pixels) & 15u) == 0, "pixel_diff: pixels not 16-byte aligned");
    const float thr = __bfloat162float(__float2bfloat16_rn((float)threshold));
    const float first = __bfloat162float(__float2bfloat16_rn((float)(threshold + 1.0)));
    const int64_t Wv = W / 8;
    // (warps, stages) per CTA: CSB_PIX_CFG=<warps>x<stages> overrides the default for tuning runs
    int pw = kPixWarpsDefault, ps = kPixStagesDefault;
    if (const char* cfg = getenv("CSB_PIX_CFG")) sscanf(cfg, "%dx%d", &pw, &ps);
    const size_t ring_bytes = (size_t)pw * ps * Wv * 16;
    if (W % 8 == 0 && Wv <= 320 && t >= 8 && ring_bytes <= 200 * 1024) {
      const int F = t >= 96 ? 12 : 4;
      const int64_t nitems = ((t + F - 1) / F) * P;
      int64_t gb = (nitems + pw - 1) / pw;
      // up to 4 CTAs per SM's worth of grid (one resident at a time: the ring takes the shared memory): CTAs that find their SM
      // taken by a long-running kernel of another stream are simply picked up by the SMs that finish first
      // (CSB_PIX_WAVES overrides for tuning runs; 1 = one persistent CTA per SM)
      int waves = 4;
      if (const char* o = getenv("CSB_PIX_WAVES")) waves = atoi(o) > 0 ? atoi(o) : 4;
      if (gb > (int64_t)kNumSMs * waves) gb = (int64_t)kNumSMs * waves;
      if (max_ctas > 0 && gb > max_ctas) gb = max_ctas;  // throttled grid: the caller overlaps the kernel with a latency-bound one
      if (const char* o = getenv("CSB_PIX_CTAS")) { if (atoi(o) > 0 && gb > atoi(o)) gb = atoi(o); }  // (tuning runs)
      int evict_first = (flags & CSB_PIXEL_EVICT_FIRST) ? 1 : 0;
      if (const char* o = getenv("CSB_PIX_EVICT")) evict_first = atoi(o);
      auto launch = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes);
        if (e != cudaSuccess) { set_error("pixel_diff: smem attr: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
        kern<<<(unsigned)gb, pw * 32, ring_bytes, s>>>((const uint8_t*)pixels, (int)t, (int)P, (int)Wv, (int)W, F, thr, first, mask, evict_first);
        return check_launch("pixel_diff_tma");
      };
#define CSB_PIX_DISPATCH(NVV)                                                       \
      if (pw == 8 && ps == 4) return launch(pixel_diff_tma_kernel<NVV, 8, 4>);      \
      if (pw == 12 && ps == 3) return launch(pixel_diff_tma_kernel<NVV, 12, 3>);    \
      if (pw == 16 && ps == 2) return launch(pixel_diff_tma_kernel<NVV, 16, 2>);    \
      if (pw == 16 && ps == 3) return launch(pixel_diff_tma_kernel<NVV, 16, 3>);    \
      if (pw == 4 && ps == 8) return launch(pixel_diff_tma_kernel<NVV, 4, 8>);
      if (Wv <= 64) { CSB_PIX_DISPATCH(2) }
      else if (Wv <= 192) { CSB_PIX_DISPATCH(6) }
      else { CSB_PIX_DISPATCH(10) }
#undef CSB_PIX_DISPATCH
    }
    pixel_diff_kernel<true><<<blocks, 256, 0, s>>>(pixels, (int)t, (int)P, (int)W, thr, first, mask);
  } else if (dtype == CSB_F32) {
    pixel_diff_kernel<false><<<blocks, 256, 0, s>>>(pixels, (int)t, (int)P, (int)W, (float)threshold, (float)(threshold + 1.0), mask);
  } else {
    set_error("pixel_diff: dtype must be bf16 or f32");
    return CSB_ERR_UNSUPPORTED;
  }
  return check_launch("pixel_diff");
}

int csb_mask_finalize(uint8_t* mask, int64_t t, int64_t P, int64_t min_tokens, const uint8_t* minor_flags, int64_t frame0,
                      csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(mask && t > 0 && P > 0, "mask_finalize: null argument or empty shape");
  mask_finalize_kernel<<<(unsigned)((t + 7) / 8), 256, 0, as_stream(stream)>>>(mask, (int)t, (int)P, (int)min_tokens, minor_flags, frame0,
                                                                               nullptr);
  return check_launch("mask_finalize");
}

int csb_pixel_mask(const void* pixels, int64_t t, int64_t P, int64_t W, int32_t dtype, double threshold, int64_t min_tokens,
                   const uint8_t* minor_flags, int64_t frame0, uint8_t* mask, csb_stream_t stream) {
  int st = csb_pixel_diff(pixels, t, P, W, dtype, threshold, mask, stream);
  if (st) return st;
  return csb_mask_finalize(mask, t, P, min_tokens, minor_flags, frame0, stream);
}

// layout: [pos / seg_start : n int32][chunk counts A : nchunks+1][chunk counts B : nchunks+1][ticket : 256 B]
static inline int64_t ws_n(int64_t n_tokens, int64_t n_ids) { const int64_t n = n_tokens > n_ids ? n_tokens : n_ids; return n > 0 ? n : 1; }
static inline int64_t ws_chunks(int64_t n) { return (n + kScanChunk - 1) / kScanChunk; }

size_t csb_compact_workspace_bytes(int64_t n_tokens, int64_t n_ids) {
  const int64_t n = ws_n(n_tokens, n_ids);
  return align_up(sizeof(int32_t) * (size_t)n, 256) + 2 * align_up(sizeof(int32_t) * (size_t)(ws_chunks(n) + 1), 256) + 256;
}

int csb_compact_tokens(const void* x, int64_t T, int64_t P, int64_t D, const uint8_t* mask, const uint8_t* pool_flags, void* out,
                       int32_t* kept_index, int64_t* n_kept, void* workspace, size_t workspace_bytes, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(x && mask && out && kept_index && n_kept && workspace && T > 0 && P > 0 && D > 0, "compact_tokens: null argument");
  CSB_REQUIRE(D % 8 == 0, "compact_tokens: D %% 8 != 0");
  const int64_t n = T * P;
  CSB_REQUIRE(n < (1ll << 31), "compact_tokens: too many tokens");
  if (workspace_bytes < csb_compact_workspace_bytes(n, 0)) { set_error("compact_tokens: workspace too small"); return CSB_ERR_WORKSPACE; }
  cudaStream_t s = as_stream(stream);
  char* w = reinterpret_cast<char*>(workspace);
  int32_t* pos = reinterpret_cast<int32_t*>(w);
  int* counts = reinterpret_cast<int*>(w + align_up(sizeof(int32_t) * (size_t)n, 256));
  const int nchunks = (int)ws_chunks(n);
  unsigned* ticket = reinterpret_cast<unsigned*>(w + align_up(sizeof(int32_t) * (size_t)n, 256) +
                                                 2 * align_up(sizeof(int32_t) * (size_t)(nchunks + 1), 256));
  cudaError_t e = cudaMemsetAsync(ticket, 0, 4, s);
  if (e != cudaSuccess) { set_error("compact_tokens: memset: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  int st;
  static thread_local int cap_compact = -1;
  if (cap_compact < 0) cap_compact = coop_capacity(compact_scan_coop_kernel);
  if (nchunks <= cap_compact && !getenv("CSB_NO_COOP_SCAN")) {  // count + barrier + scatter in one cooperative launch
    int64_t n_arg = n;
    void* args[] = {(void*)&mask, (void*)&n_arg, (void*)&counts, (void*)&ticket, (void*)&n_kept, (void*)&pos, (void*)&kept_index};
    e = cudaLaunchCooperativeKernel((const void*)compact_scan_coop_kernel, dim3((unsigned)nchunks), dim3(kScanThreads), args, 0, s);
    if (e != cudaSuccess) { set_error("compact_tokens: cooperative launch: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
    count_launch(1);
  } else {
    mask_chunk_count_kernel<<<nchunks, kScanThreads, 0, s>>>(mask, n, counts, nchunks, ticket, n_kept);
    st = check_launch("mask_chunk_count");
    if (st) return st;
    compact_scatter_kernel<<<nchunks, kScanThreads, 0, s>>>(mask, n, counts, pos, kept_index);
    st = check_launch("compact_scatter");
    if (st) return st;
  }
  if (pool_flags) {
    // pooled tokens + the row gather of the kept tokens in ONE launch (the gather blocks follow the T pooling blocks)
    const int Dv = (int)(D / 8);
    const int tx = Dv >= 256 ? 256 : ((Dv + 31) / 32) * 32;
    const size_t smem = (size_t)(kPoolGroups - 1) * tx * 8 * sizeof(float);
    int64_t gblocks = (n * Dv + (int64_t)tx * kPoolGroups * 4 - 1) / ((int64_t)tx * kPoolGroups * 4);
    if (gblocks > kNumSMs * 4) gblocks = kNumSMs * 4;
    if (gblocks < 1) gblocks = 1;
    pool_emit_kernel<<<(unsigned)(T + gblocks), dim3(tx, kPoolGroups), smem, s>>>((const uint4*)x, (int)P, Dv, pool_flags, mask, pos,
                                                                                   (uint4*)out, (int)T, kept_index, n_kept);
    return check_launch("pool_emit");
  }
  return csb_gather_rows(x, D * 2, kept_index, 0, n_kept, n, out, stream);
}

size_t csb_compact_stream_workspace_bytes(int64_t n_tokens, int64_t n_ids) {
  const int64_t n = ws_n(n_tokens, n_ids);
  const size_t own = align_up(sizeof(int32_t) * (size_t)n, 256) + 3 * align_up(sizeof(int32_t) * (size_t)(n + 1), 256) + 256;
  const size_t fallback = csb_compact_workspace_bytes(n_tokens, n_ids);
  return own > fallback ? own : fallback;
}

int csb_compact_stream(const csb_compact_args* p, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(p, "compact_stream: null args");
  CSB_REQUIRE(p->x_bf16 && p->mask && p->out_bf16 && p->kept_index && p->n_kept && p->workspace && p->T > 0 && p->P > 0 && p->D > 0,
              "compact_stream: null argument or empty shape");
  CSB_REQUIRE(p->D % 8 == 0, "compact_stream: D %% 8 != 0");
  CSB_REQUIRE(!p->input_ids || (p->out_ids && p->n_out && p->n_ids > 0), "compact_stream: input_ids given without outputs");
  CSB_REQUIRE(!p->attention_mask || p->out_attention_mask, "compact_stream: attention_mask given without output");
  CSB_REQUIRE(!p->labels || p->out_labels, "compact_stream: labels given without output");
  CSB_REQUIRE(!p->pool_minor || p->minor_flags, "compact_stream: pool_minor needs minor_flags");
  const int64_t T = p->T, P = p->P, n = T * P, n_ids = p->input_ids ? p->n_ids : 0;
  CSB_REQUIRE(n < (1ll << 31) && n_ids < (1ll << 31), "compact_stream: too many tokens / ids");
  if (p->workspace_bytes < csb_compact_stream_workspace_bytes(n, n_ids)) { set_error("compact_stream: workspace too small"); return CSB_ERR_WORKSPACE; }
  cudaStream_t s = as_stream(stream);
  const int fpc = P <= kScanChunk ? (int)(kScanChunk / P) : 0;
  const int n_tok_chunks = fpc ? (int)((T + fpc - 1) / fpc) : 0;
  const int n_id_chunks = (int)ws_chunks(n_ids);
  const int grid = n_tok_chunks > n_id_chunks ? n_tok_chunks : n_id_chunks;
  static thread_local int cap = -1;
  if (cap < 0) cap = coop_capacity(compact_stream_coop_kernel);
  const uint8_t* pool_flags = p->pool_minor ? p->minor_flags : nullptr;
  if (!fpc || grid > cap || getenv("CSB_NO_COOP_SCAN")) {  // frames longer than a chunk, or more chunks than co-resident CTAs
    int st = CSB_OK;
    if (p->finalize) {
      mask_finalize_kernel<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(p->mask, (int)T, (int)P, (int)p->min_tokens, p->minor_flags, 0, p->frame_kind);
      st = check_launch("mask_finalize");
    }
    if (st) return st;
    st = csb_compact_tokens(p->x_bf16, T, P, p->D, p->mask, pool_flags, p->out_bf16, p->kept_index, p->n_kept, p->workspace,
                            p->workspace_bytes, stream);
    if (st || !p->input_ids) return st;
    // (csb_compact_tokens has consumed its part of the workspace by the time the id filter runs: same stream)
    return csb_filter_ids(p->input_ids, n_ids, p->image_token_index, p->mask, n, p->attention_mask, p->labels, nullptr, p->out_ids,
                          p->out_attention_mask, p->out_labels, nullptr, p->n_out, p->workspace, p->workspace_bytes, stream);
  }
  const int64_t nn = ws_n(n, n_ids);
  char* w = reinterpret_cast<char*>(p->workspace);
  const size_t cbytes = align_up(sizeof(int32_t) * (size_t)(nn + 1), 256);
  CompactStreamArgs a;
  a.mask = p->mask; a.T = (int)T; a.P = (int)P; a.fpc = fpc; a.n_tok_chunks = n_tok_chunks; a.min_tokens = (int)p->min_tokens;
  a.finalize = p->finalize; a.minor = p->minor_flags; a.kind = p->frame_kind;
  a.pos = reinterpret_cast<int32_t*>(w);
  w += align_up(sizeof(int32_t) * (size_t)nn, 256);
  a.counts = reinterpret_cast<int*>(w); a.img_counts = reinterpret_cast<int*>(w + cbytes); a.keep_counts = reinterpret_cast<int*>(w + 2 * cbytes);
  a.barrier = reinterpret_cast<unsigned*>(w + 3 * cbytes);
  a.kept_index = p->kept_index; a.n_kept = p->n_kept;
  a.ids = p->input_ids; a.n_ids = n_ids; a.image_tok = p->image_token_index; a.n_id_chunks = n_id_chunks;
  a.attn = p->attention_mask; a.labels = p->labels; a.o_ids = p->out_ids; a.o_attn = p->out_attention_mask; a.o_labels = p->out_labels;
  a.n_out = p->n_out;
  void* args[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)compact_stream_coop_kernel, dim3((unsigned)grid), dim3(kScanThreads), args, 0, s);
  if (e != cudaSuccess) { set_error("compact_stream: cooperative launch: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  count_launch(1);
  if (pool_flags) {
    const int Dv = (int)(p->D / 8);
    const int tx = Dv >= 256 ? 256 : ((Dv + 31) / 32) * 32;
    const size_t smem = (size_t)(kPoolGroups - 1) * tx * 8 * sizeof(float);
    int64_t gblocks = (n * Dv + (int64_t)tx * kPoolGroups * 4 - 1) / ((int64_t)tx * kPoolGroups * 4);
    if (gblocks > kNumSMs * 4) gblocks = kNumSMs * 4;
    if (gblocks < 1) gblocks = 1;
    pool_emit_kernel<<<(unsigned)(T + gblocks), dim3(tx, kPoolGroups), smem, s>>>((const uint4*)p->x_bf16, (int)P, Dv, pool_flags, p->mask, a.pos,
                                                                                   (uint4*)p->out_bf16, (int)T, p->kept_index, p->n_kept);
    return check_launch("pool_emit");
  }
  return csb_gather_rows(p->x_bf16, p->D * 2, p->kept_index, 0, p->n_kept, n, p->out_bf16, stream);
}

size_t csb_relevance_select_workspace_bytes(int64_t k, int64_t Dl) {
  if (k <= 0 || Dl <= 0) return 256;
  return align_up(sizeof(float) * (size_t)k * kRelSplits * (size_t)Dl, 256);
}

int csb_relevance_select(const csb_relevance_args* p, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(p, "relevance_select: null args");
  CSB_REQUIRE(p->question && p->cos_out && p->low_flags && p->assignments && p->minor_flags && p->minor_list && p->n_minor && p->ticket &&
                  p->k > 0 && p->Dl > 0 && p->T > 0,
              "relevance_select: null argument");
  CSB_REQUIRE(p->gram && p->event_offsets && p->event_order && p->dist && p->reps && p->rep_counts && p->rep_flags && p->n_rep > 0 &&
                  p->ldg >= p->T,
              "relevance_select: null selection argument or ldg < T");
  CSB_REQUIRE((p->events != nullptr) != (p->hidden != nullptr), "relevance_select: give either events [k,Dl] or the packed hidden states");
  CSB_REQUIRE(p->dtype == CSB_BF16 || p->dtype == CSB_F32, "relevance_select: dtype must be bf16 or f32");
  cudaStream_t s = as_stream(stream);
  RelevanceArgs a;
  a.q = p->question; a.ev = p->events; a.partials = nullptr; a.cu = p->cu_seqlens; a.k = (int)p->k; a.Dl = (int)p->Dl;
  a.threshold = p->dtype == CSB_BF16 ? __bfloat162float(__float2bfloat16_rn((float)p->threshold)) : (float)p->threshold;
  a.G = p->gram; a.ldg = p->ldg; a.ev_off = p->event_offsets; a.order = p->event_order; a.n_rep = (int)p->n_rep;
  a.assign = p->assignments; a.T = (int)p->T; a.ev_out = p->events_out; a.cos_out = p->cos_out; a.low = p->low_flags; a.dist = p->dist;
  a.reps = p->reps; a.rep_counts = p->rep_counts; a.rep_flags = p->rep_flags; a.minor_flags = p->minor_flags; a.minor_list = p->minor_list;
  a.n_minor = p->n_minor; a.ticket = p->ticket;
  size_t smem = 0;
  if (p->hidden) {
    CSB_REQUIRE(p->cu_seqlens && p->workspace, "relevance_select: hidden states need cu_seqlens and a workspace");
    if (p->workspace_bytes < csb_relevance_select_workspace_bytes(p->k, p->Dl)) { set_error("relevance_select: workspace too small"); return CSB_ERR_WORKSPACE; }
    int st = launch_seq_pool_partial(p->hidden, p->cu_seqlens, p->k, p->Dl, p->dtype, reinterpret_cast<float*>(p->workspace), s);
    if (st) return st;
    a.partials = reinterpret_cast<const float*>(p->workspace);
    smem = sizeof(float) * (size_t)p->Dl;
    CSB_REQUIRE(smem <= 200 * 1024, "relevance_select: Dl too large for the fused pool (Dl <= 51200)");
  }
  const unsigned grid = (unsigned)(p->k + (p->k + 7) / 8);
  auto launch = [&](auto kern) -> int {
    if (smem > 40 * 1024) {  // (static + dynamic must fit the 48 KB default otherwise)
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("relevance_select: smem attr: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
    }
    kern<<<grid, 256, smem, s>>>(a);
    return check_launch("relevance_select");
  };
  return p->dtype == CSB_BF16 ? launch(relevance_select_kernel<true>) : launch(relevance_select_kernel<false>);
}

int csb_filter_ids(const int64_t* input_ids, int64_t n_ids, int64_t image_token_index, const uint8_t* mask, int64_t n_mask,
                   const int64_t* attention_mask, const int64_t* labels, const int64_t* position_ids, int64_t* out_ids,
                   int64_t* out_attention_mask, int64_t* out_labels, int64_t* out_position_ids, int64_t* n_out, void* workspace,
                   size_t workspace_bytes, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(input_ids && mask && out_ids && n_out && n_ids > 0, "filter_ids: null argument");
  CSB_REQUIRE(!attention_mask || out_attention_mask, "filter_ids: attention_mask given without output");
  CSB_REQUIRE(!labels || out_labels, "filter_ids: labels given without output");
  CSB_REQUIRE(!position_ids || (out_position_ids && workspace), "filter_ids: position_ids given without output/workspace");
  CSB_REQUIRE(n_ids < (1ll << 31), "filter_ids: too many ids");
  CSB_REQUIRE(workspace, "filter_ids: null workspace");
  if (workspace_bytes < csb_compact_workspace_bytes(0, n_ids)) { set_error("filter_ids: workspace too small"); return CSB_ERR_WORKSPACE; }
  cudaStream_t s = as_stream(stream);
  if (!position_ids) {
    const int64_t n = ws_n(0, n_ids);
    char* w = reinterpret_cast<char*>(workspace);
    const int nchunks = (int)ws_chunks(n_ids);
    const size_t cbytes = align_up(sizeof(int32_t) * (size_t)(ws_chunks(n) + 1), 256);
    int* img_prefix = reinterpret_cast<int*>(w + align_up(sizeof(int32_t) * (size_t)n, 256));
    int* keep_prefix = reinterpret_cast<int*>(reinterpret_cast<char*>(img_prefix) + cbytes);
    unsigned* ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(keep_prefix) + cbytes);
    cudaError_t e = cudaMemsetAsync(ticket, 0, 4, s);
    if (e != cudaSuccess) { set_error("filter_ids: memset: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
    static thread_local int cap_ids = -1;
    if (cap_ids < 0) cap_ids = coop_capacity(ids_filter_coop_kernel);
    if (nchunks <= cap_ids && !getenv("CSB_NO_COOP_SCAN")) {  // the three dependent passes in one cooperative launch
      void* args[] = {(void*)&input_ids, (void*)&n_ids, (void*)&image_token_index, (void*)&mask, (void*)&n_mask, (void*)&img_prefix,
                      (void*)&keep_prefix, (void*)&ticket, (void*)&n_out, (void*)&attention_mask, (void*)&labels, (void*)&out_ids,
                      (void*)&out_attention_mask, (void*)&out_labels};
      e = cudaLaunchCooperativeKernel((const void*)ids_filter_coop_kernel, dim3((unsigned)nchunks), dim3(kScanThreads), args, 0, s);
      if (e != cudaSuccess) { set_error("filter_ids: cooperative launch: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
      count_launch(1);
      return CSB_OK;
    }
    ids_image_count_kernel<<<nchunks, kScanThreads, 0, s>>>(input_ids, n_ids, image_token_index, img_prefix, nchunks, ticket, n_out + 1);
    int st = check_launch("ids_image_count");
    if (st) return st;
    ids_keep_kernel<false><<<nchunks, kScanThreads, 0, s>>>(input_ids, n_ids, image_token_index, mask, n_mask, img_prefix, keep_prefix,
                                                            nchunks, ticket, n_out, attention_mask, labels, out_ids,
                                                            out_attention_mask, out_labels);
    st = check_launch("ids_keep_count");
    if (st) return st;
    ids_keep_kernel<true><<<nchunks, kScanThreads, 0, s>>>(input_ids, n_ids, image_token_index, mask, n_mask, img_prefix, keep_prefix,
                                                           nchunks, ticket, n_out, attention_mask, labels, out_ids,
                                                           out_attention_mask, out_labels);
    return check_launch("ids_keep_scatter");
  }
  filter_ids_kernel<<<1, 1024, 0, s>>>(input_ids, n_ids, image_token_index, mask, n_mask, attention_mask, labels,
                                       position_ids, out_ids, out_attention_mask, out_labels, out_position_ids,
                                       n_out, reinterpret_cast<int32_t*>(workspace));
  return check_launch("filter_ids");
}

}  // extern "C"
