// The two tensor ops that sit between the clustering and the relevance test of select_events_based_on_summary, and the
// last step of the call site:
//
//  embed_scatter  -- model/cogreasoner_chat.py:303-305 (per-event summary prompt) and :567-572 (final prompt):
//                      inputs_embeds = embed_tokens(input_ids).clone()
//                      inputs_embeds[input_ids == image_token] = inputs_embeds[...] * 0.0 + features
//                    as ONE pass: every output row is read from exactly one source (the embedding table row of a text id, or
//                    the next feature row for an <image> id), no intermediate [n, D] tensors.  The ids may be the
//                    concatenation of several sequences (the k event prompts packed back to back, cu_seqlens style): the
//                    j-th <image> id overall takes feature row j, which is exactly the event-major `packed` buffer phase 1
//                    writes.
//  seq_mean_pool  -- :317-318, :322: torch.mean(last_hidden_state, dim=1) of every sequence of a packed var-len batch
//                    (fp32 accumulation in row order, one division, one rounding), [k, Dl] in one launch instead of k.
//
// Both are HBM-bound streaming kernels: 128-bit loads/stores, >= 4 independent loads in flight per thread.
#include "csb_common.cuh"

namespace csb {

constexpr int kRankThreads = 1024;
constexpr int kRowsPerCta = 32;  // one ballot

// chunk_prefix[c] = number of <image> ids before id 32*c (c = 0 .. ceil(n/32)); status[0] = total, status[1] = number of ids
// outside [0, vocab).  Single CTA: n is a prompt length (1e4 .. 1e6), the scan is a few microseconds.
__global__ void __launch_bounds__(kRankThreads) ids_rank_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t image_token,
                                                                int64_t vocab, int32_t* __restrict__ chunk_prefix,
                                                                int32_t* __restrict__ status) {
  __shared__ int s_warp[kRankThreads / 32];
  __shared__ int s_run, s_bad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_run = 0; s_bad = 0; }
  __syncthreads();
  int bad = 0;
  for (int64_t base = 0; base < n; base += kRankThreads) {
    const int64_t i = base + threadIdx.x;
    const int64_t id = i < n ? ids[i] : -1;
    const bool img = i < n && id == image_token;
    if (i < n && (id < 0 || id >= vocab)) ++bad;
    const unsigned bal = __ballot_sync(0xffffffffu, img);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    // exclusive prefix over the 32 warp counts (every warp computes it redundantly from shared memory)
    int v = s_warp[lane];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const int warp_excl = __shfl_sync(0xffffffffu, inc - v, warp);
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    const int run = s_run;
    if (lane == 0) {
      const int64_t c = (base >> 5) + warp;  // chunk of 32 ids this warp covers; c == ceil(n/32) is the sentinel (= total)
      if (c <= (n + 31) / 32) chunk_prefix[c] = run + warp_excl;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_run = run + total;
    __syncthreads();
  }
  bad = warp_sum(bad);
  if (lane == 0 && bad) atomicAdd(&s_bad, bad);
  __syncthreads();
  if (threadIdx.x == 0) {
    chunk_prefix[(n + 31) / 32] = s_run;  // sentinel = total (also written by the loop when n % 1024 != 0; same value)
    status[0] = s_run;
    status[1] = s_bad;
  }
}

// Multi-CTA form for long id sequences (the k packed event prompts: 1e5+ ids): a CTA ranks kRankChunk ids (relative prefixes per
// 32-id chunk + its total); the last CTA to finish turns the totals into exclusive bases.  The scatter kernel adds the two.
constexpr int kRankChunk = 4096;

__global__ void __launch_bounds__(kRankThreads) ids_rank_multi_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t image_token, int64_t vocab,
                                                                      int32_t* __restrict__ chunk_prefix, int32_t* __restrict__ cta_base,
                                                                      int32_t* __restrict__ cta_bad, unsigned* ticket, int32_t* __restrict__ status) {
  __shared__ int s_warp[kRankThreads / 32];
  __shared__ int s_run, s_bad, s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_run = 0; s_bad = 0; }
  __syncthreads();
  const int64_t begin = (int64_t)blockIdx.x * kRankChunk, end = min(n, begin + kRankChunk);
  int bad = 0;
  for (int64_t base = begin; base < end; base += kRankThreads) {
    const int64_t i = base + threadIdx.x;
    const int64_t id = i < end ? ids[i] : -1;
    const bool img = i < end && id == image_token;
    if (i < end && (id < 0 || id >= vocab)) ++bad;
    const unsigned bal = __ballot_sync(0xffffffffu, img);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int v = s_warp[lane];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const int warp_excl = __shfl_sync(0xffffffffu, inc - v, warp);
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    const int run = s_run;
    if (lane == 0 && base + (int64_t)warp * 32 < end) chunk_prefix[(base >> 5) + warp] = run + warp_excl;  // relative to this CTA
    __syncthreads();
    if (threadIdx.x == 0) s_run = run + total;
    __syncthreads();
  }
  bad = warp_sum(bad);
  if (lane == 0 && bad) atomicAdd(&s_bad, bad);
  __syncthreads();
  if (threadIdx.x == 0) {
    cta_base[blockIdx.x] = s_run;  // totals for now
    cta_bad[blockIdx.x] = s_bad;
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // exclusive scan of the CTA totals (gridDim.x <= 1024), one value per thread
  const int nbk = gridDim.x;
  int v = (int)threadIdx.x < nbk ? __ldcg(cta_base + threadIdx.x) : 0;
  int b = (int)threadIdx.x < nbk ? __ldcg(cta_bad + threadIdx.x) : 0;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  b = warp_sum(b);
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  if (threadIdx.x == 0) s_bad = 0;  // reused as the grand total of out-of-range ids
  __syncthreads();
  if (lane == 0 && b) atomicAdd(&s_bad, b);
  int wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += s_warp[w];
  if ((int)threadIdx.x < nbk) cta_base[threadIdx.x] = wbase + inc - v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < kRankThreads / 32; ++w) tot += s_warp[w];
    status[0] = tot;
    status[1] = s_bad;
    *ticket = 0;
  }
}

// x * 0.0 + f in the arithmetic of the tensor dtype (ATen evaluates the two ops separately, rounding after each)
__device__ __forceinline__ uint32_t zero_mul_add_bf16x2(uint32_t e2, uint32_t f2) {
  const float e_lo = __uint_as_float(e2 << 16), e_hi = __uint_as_float(e2 & 0xffff0000u);
  const float f_lo = __uint_as_float(f2 << 16), f_hi = __uint_as_float(f2 & 0xffff0000u);
  const float z_lo = round_bf16(__fmul_rn(e_lo, 0.0f)), z_hi = round_bf16(__fmul_rn(e_hi, 0.0f));  // +-0, or NaN for inf / NaN
  const uint32_t r_lo = float_to_bf16_bits(__fadd_rn(z_lo, f_lo)), r_hi = float_to_bf16_bits(__fadd_rn(z_hi, f_hi));
  return r_lo | (r_hi << 16);
}

template <bool BF16>
__global__ void __launch_bounds__(256) embed_scatter_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t image_token,
                                                            int64_t vocab, const uint4* __restrict__ table,
                                                            const uint4* __restrict__ feats, int64_t n_feats, int Dv,
                                                            const int32_t* __restrict__ chunk_prefix, uint4* __restrict__ out,
                                                            const int32_t* __restrict__ frame_order, int tokens_per_frame,
                                                            const int32_t* __restrict__ cta_base, int32_t* __restrict__ status) {
  __shared__ const uint4* s_src[kRowsPerCta];
  __shared__ int s_img[kRowsPerCta];
  __shared__ int s_cnt[8], s_badc[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerCta;
  int64_t prefix;
  if (chunk_prefix) {
    prefix = chunk_prefix[blockIdx.x] + (cta_base ? cta_base[row0 / kRankChunk] : 0);
  } else {
    // short id sequences (a single prompt): no rank kernel -- every CTA counts the <image> ids in front of its rows itself
    // (the ids are L2-resident: n <= 16k ids = 128 KB); the last CTA thereby sees every id and writes the status words
    const bool last = blockIdx.x == gridDim.x - 1;
    int cnt = 0, bad = 0;
    for (int64_t i = threadIdx.x; i < row0; i += blockDim.x) {
      const int64_t id = __ldg(ids + i);
      cnt += id == image_token;
      bad += (id < 0 || id >= vocab);
    }
    cnt = warp_sum(cnt); bad = warp_sum(bad);
    if (lane == 0) { s_cnt[warp] = cnt; s_badc[warp] = bad; }
    __syncthreads();
    int tot = 0, btot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { tot += s_cnt[w]; btot += s_badc[w]; }
    prefix = tot;
    if (last && warp == 0) {
      const int64_t i = row0 + lane;
      const int64_t id = i < n ? ids[i] : 0;
      const int own = __popc(__ballot_sync(0xffffffffu, i < n && id == image_token));
      const int ownbad = __popc(__ballot_sync(0xffffffffu, i < n && (id < 0 || id >= vocab)));
      if (lane == 0) { status[0] = tot + own; status[1] = btot + ownbad; }
    }
  }
  if (warp == 0) {
    const int64_t i = row0 + lane;
    const int64_t id = i < n ? ids[i] : -1;
    const bool img = i < n && id == image_token;
    const unsigned bal = __ballot_sync(0xffffffffu, img);
    const int64_t rank = prefix + __popc(bal & ((1u << lane) - 1u));
    const uint4* src = nullptr;  // nullptr -> zeros (id out of range, or more <image> ids than feature rows: flagged in status)
    if (img) {
      if (rank < n_feats) {
        int64_t row = rank;
        if (frame_order) {  // chronological per-event gather folded into the read (cogreasoner_chat.py:291-296)
          const int64_t fr = rank / tokens_per_frame;
          row = (int64_t)frame_order[fr] * tokens_per_frame + (rank - fr * tokens_per_frame);
        }
        src = feats + row * Dv;
      }
    } else if (id >= 0 && id < vocab) {
      src = table + id * Dv;
    }
    s_src[lane] = src;
    s_img[lane] = img ? 1 : 0;
  }
  __syncthreads();
  const uint4* img_row = (image_token >= 0 && image_token < vocab) ? table + image_token * Dv : nullptr;
  for (int r = warp; r < kRowsPerCta; r += 8) {
    const int64_t i = row0 + r;
    if (i >= n) break;
    const uint4* src = s_src[r];
    const bool img = s_img[r] != 0;
    uint4* dst = out + i * Dv;
    for (int c0 = lane; c0 < Dv; c0 += 128) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        v[u] = (src && c < Dv) ? ldg_stream(src + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        if (c >= Dv) continue;
        uint4 o = v[u];
        if (img && src) {
          const uint4 e = img_row ? __ldg(img_row + c) : make_uint4(0, 0, 0, 0);
          if (BF16) {
            o.x = zero_mul_add_bf16x2(e.x, o.x); o.y = zero_mul_add_bf16x2(e.y, o.y);
            o.z = zero_mul_add_bf16x2(e.z, o.z); o.w = zero_mul_add_bf16x2(e.w, o.w);
          } else {
            o.x = __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(e.x), 0.0f), __uint_as_float(o.x)));
            o.y = __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(e.y), 0.0f), __uint_as_float(o.y)));
            o.z = __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(e.z), 0.0f), __uint_as_float(o.z)));
            o.w = __float_as_uint(__fadd_rn(__fmul_rn(__uint_as_float(e.w), 0.0f), __uint_as_float(o.w)));
          }
        }
        dst[c] = o;
      }
    }
  }
}

// ---- var-len mean pool ------------------------------------------------------------------------------------------------
constexpr int kPoolSplits = 8;  // row splits per sequence (partial sums are added in split order: deterministic)

// grid (ceil(Dv / 128), n_seq, kPoolSplits) x 128 threads: thread = one 128-bit column group, rows of its split in order
template <bool BF16>
__global__ void __launch_bounds__(128) seq_pool_partial_kernel(const uint4* __restrict__ h, const int32_t* __restrict__ cu, int Dv,
                                                               float* __restrict__ partial) {
  constexpr int E = BF16 ? 8 : 4;
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int j = blockIdx.y, sp = blockIdx.z;
  if (c >= Dv) return;
  const int beg = cu[j], end = cu[j + 1], len = end - beg;
  const int per = (len + kPoolSplits - 1) / kPoolSplits;
  const int r0 = beg + min(len, sp * per), r1 = beg + min(len, (sp + 1) * per);
  float acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldg_stream(h + (int64_t)(r + u) * Dv + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (BF16) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      } else {
        acc[0] += __uint_as_float(v[u].x); acc[1] += __uint_as_float(v[u].y);
        acc[2] += __uint_as_float(v[u].z); acc[3] += __uint_as_float(v[u].w);
      }
    }
  }
  for (; r < r1; ++r) {
    const uint4 v = ldg_stream(h + (int64_t)r * Dv + c);
    if (BF16) {
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    } else {
      acc[0] += __uint_as_float(v.x); acc[1] += __uint_as_float(v.y);
      acc[2] += __uint_as_float(v.z); acc[3] += __uint_as_float(v.w);
    }
  }
  float* o = partial + ((int64_t)(j * kPoolSplits + sp) * Dv + c) * E;
#pragma unroll
  for (int e = 0; e < E; ++e) o[e] = acc[e];
}

// grid (ceil(D / 256), n_seq): split partials added in order, one IEEE division by the length, one rounding
template <bool BF16>
__global__ void __launch_bounds__(256) seq_pool_final_kernel(const float* __restrict__ partial, const int32_t* __restrict__ cu, int D,
                                                             void* __restrict__ out) {
  const int d = blockIdx.x * 256 + threadIdx.x;
  const int j = blockIdx.y;
  if (d >= D) return;
  const int len = cu[j + 1] - cu[j];
  float s = 0.f;
#pragma unroll
  for (int sp = 0; sp < kPoolSplits; ++sp) s += partial[(int64_t)(j * kPoolSplits + sp) * D + d];
  const float m = __fdiv_rn(s, (float)len);  // len == 0 -> NaN, as torch.mean of an empty dim
  if (BF16) reinterpret_cast<__nv_bfloat16*>(out)[(int64_t)j * D + d] = __float2bfloat16_rn(m);
  else reinterpret_cast<float*>(out)[(int64_t)j * D + d] = m;
}

static inline size_t align_up256(size_t v) { return (v + 255) / 256 * 256; }

int launch_seq_pool_partial(const void* hidden, const int32_t* cu_seqlens, int64_t n_seq, int64_t Dl, int dtype, float* partial,
                            cudaStream_t s) {
  const int per = dtype == CSB_BF16 ? 8 : 4;
  if (Dl % per != 0 || (reinterpret_cast<uintptr_t>(hidden) & 15u) != 0) {
    set_error("seq_mean_pool: rows must be 128-bit multiples, 16-byte aligned");
    return CSB_ERR_ARG;
  }
  if (n_seq > 65535) { set_error("seq_mean_pool: more than 65535 sequences"); return CSB_ERR_ARG; }
  const int Dv = (int)(Dl / per);
  dim3 g1((unsigned)((Dv + 127) / 128), (unsigned)n_seq, kPoolSplits);
  if (dtype == CSB_BF16) seq_pool_partial_kernel<true><<<g1, 128, 0, s>>>((const uint4*)hidden, cu_seqlens, Dv, partial);
  else seq_pool_partial_kernel<false><<<g1, 128, 0, s>>>((const uint4*)hidden, cu_seqlens, Dv, partial);
  return check_launch("seq_pool_partial");
}

}  // namespace csb

using namespace csb;

extern "C" {

size_t csb_embed_scatter_workspace_bytes(int64_t n_ids) {
  if (n_ids <= 0) return 256;
  const int64_t nbk = (n_ids + kRankChunk - 1) / kRankChunk;
  return align_up256(sizeof(int32_t) * (size_t)((n_ids + 31) / 32 + 2)) + align_up256(sizeof(int32_t) * (size_t)(2 * nbk + 2)) + 256;
}

static int embed_scatter_impl(const int64_t* input_ids, int64_t n_ids, int64_t image_token_index, const void* embed_table, int64_t vocab,
                              const void* features, int64_t n_features, const int32_t* frame_order, int64_t tokens_per_frame, int64_t D,
                              int dtype, void* out, int32_t* status, void* workspace, size_t workspace_bytes, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(input_ids && embed_table && out && status && workspace && n_ids > 0 && vocab > 0 && D > 0,
              "embed_scatter: null argument or empty shape");
  CSB_REQUIRE(features || n_features == 0, "embed_scatter: features is null but n_features > 0");
  CSB_REQUIRE(dtype == CSB_BF16 || dtype == CSB_F32, "embed_scatter: dtype must be bf16 or f32");
  const int per = dtype == CSB_BF16 ? 8 : 4;
  CSB_REQUIRE(D % per == 0, "embed_scatter: D must be a multiple of %d (128-bit rows)", per);
  CSB_REQUIRE(((reinterpret_cast<uintptr_t>(embed_table) | reinterpret_cast<uintptr_t>(features) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
              "embed_scatter: table / features / out must be 16-byte aligned");
  CSB_REQUIRE(n_ids < (int64_t)1 << 31, "embed_scatter: too many ids");
  if (workspace_bytes < csb_embed_scatter_workspace_bytes(n_ids)) {
    set_error("embed_scatter: workspace %zu < %zu", workspace_bytes, csb_embed_scatter_workspace_bytes(n_ids));
    return CSB_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  const unsigned grid = (unsigned)((n_ids + kRowsPerCta - 1) / kRowsPerCta);
  const int Dv = (int)(D / per);
  // ranks of the <image> ids: computed inside the scatter kernel for a single prompt, by a multi-CTA rank kernel for long packed
  // sequences (CSB_EMBED_RANK=1: the single-CTA rank kernel, kept as the validation path)
  const int32_t* prefix = nullptr;
  const int32_t* cta_base = nullptr;
  const char* force = getenv("CSB_EMBED_RANK");
  if (force && atoi(force) == 1) {
    int32_t* pf = reinterpret_cast<int32_t*>(workspace);
    ids_rank_kernel<<<1, kRankThreads, 0, s>>>(input_ids, n_ids, image_token_index, vocab, pf, status);
    int st = check_launch("ids_rank");
    if (st) return st;
    prefix = pf;
  } else if (n_ids > 16384 || (force && atoi(force) == 2)) {
    char* w = reinterpret_cast<char*>(workspace);
    int32_t* pf = reinterpret_cast<int32_t*>(w);
    w += align_up256(sizeof(int32_t) * (size_t)((n_ids + 31) / 32 + 2));
    const int64_t nbk = (n_ids + kRankChunk - 1) / kRankChunk;
    CSB_REQUIRE(nbk <= kRankThreads, "embed_scatter: more than 4M ids");
    int32_t* cb = reinterpret_cast<int32_t*>(w);
    int32_t* cbad = cb + nbk;
    unsigned* ticket = reinterpret_cast<unsigned*>(w + align_up256(sizeof(int32_t) * (size_t)(2 * nbk + 2)));
    cudaError_t e = cudaMemsetAsync(ticket, 0, 4, s);
    if (e != cudaSuccess) { set_error("embed_scatter: memset: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
    ids_rank_multi_kernel<<<(unsigned)nbk, kRankThreads, 0, s>>>(input_ids, n_ids, image_token_index, vocab, pf, cb, cbad, ticket, status);
    int st = check_launch("ids_rank_multi");
    if (st) return st;
    prefix = pf;
    cta_base = cb;
  }
  if (dtype == CSB_BF16)
    embed_scatter_kernel<true><<<grid, 256, 0, s>>>(input_ids, n_ids, image_token_index, vocab, (const uint4*)embed_table,
                                                     (const uint4*)features, n_features, Dv, prefix, (uint4*)out, frame_order,
                                                     (int)tokens_per_frame, cta_base, status);
  else
    embed_scatter_kernel<false><<<grid, 256, 0, s>>>(input_ids, n_ids, image_token_index, vocab, (const uint4*)embed_table,
                                                      (const uint4*)features, n_features, Dv, prefix, (uint4*)out, frame_order,
                                                      (int)tokens_per_frame, cta_base, status);
  return check_launch("embed_scatter");
}

int csb_embed_scatter(const int64_t* input_ids, int64_t n_ids, int64_t image_token_index, const void* embed_table, int64_t vocab,
                      const void* features, int64_t n_features, int64_t D, int dtype, void* out, int32_t* status, void* workspace,
                      size_t workspace_bytes, csb_stream_t stream) {
  return embed_scatter_impl(input_ids, n_ids, image_token_index, embed_table, vocab, features, n_features, nullptr, 1, D, dtype, out,
                            status, workspace, workspace_bytes, stream);
}

int csb_embed_scatter_ordered(const int64_t* input_ids, int64_t n_ids, int64_t image_token_index, const void* embed_table, int64_t vocab,
                              const void* features, const int32_t* frame_order, int64_t n_frames, int64_t tokens_per_frame, int64_t D,
                              int dtype, void* out, int32_t* status, void* workspace, size_t workspace_bytes, csb_stream_t stream) {
  CSB_REQUIRE(frame_order && n_frames > 0 && tokens_per_frame > 0 && tokens_per_frame < (1ll << 31),
              "embed_scatter_ordered: null frame order or empty shape");
  return embed_scatter_impl(input_ids, n_ids, image_token_index, embed_table, vocab, features, n_frames * tokens_per_frame, frame_order,
                            tokens_per_frame, D, dtype, out, status, workspace, workspace_bytes, stream);
}

size_t csb_seq_mean_pool_workspace_bytes(int64_t n_seq, int64_t Dl) {
  if (n_seq <= 0 || Dl <= 0) return 256;
  return align_up256(sizeof(float) * (size_t)n_seq * kPoolSplits * (size_t)Dl);
}

int csb_seq_mean_pool(const void* hidden, const int32_t* cu_seqlens, int64_t n_seq, int64_t Dl, int dtype, void* out, void* workspace,
                      size_t workspace_bytes, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(hidden && cu_seqlens && out && workspace && n_seq > 0 && Dl > 0, "seq_mean_pool: null argument or empty shape");
  CSB_REQUIRE(dtype == CSB_BF16 || dtype == CSB_F32, "seq_mean_pool: dtype must be bf16 or f32");
  const int per = dtype == CSB_BF16 ? 8 : 4;
  CSB_REQUIRE(Dl % per == 0 && (reinterpret_cast<uintptr_t>(hidden) & 15u) == 0, "seq_mean_pool: rows must be 128-bit multiples, 16-byte aligned");
  CSB_REQUIRE(n_seq <= 65535, "seq_mean_pool: more than 65535 sequences");
  if (workspace_bytes < csb_seq_mean_pool_workspace_bytes(n_seq, Dl)) {
    set_error("seq_mean_pool: workspace %zu < %zu", workspace_bytes, csb_seq_mean_pool_workspace_bytes(n_seq, Dl));
    return CSB_ERR_WORKSPACE;
  }
  cudaStream_t s = as_stream(stream);
  const int Dv = (int)(Dl / per);
  float* partial = reinterpret_cast<float*>(workspace);
  dim3 g1((unsigned)((Dv + 127) / 128), (unsigned)n_seq, kPoolSplits);
  dim3 g2((unsigned)((Dl + 255) / 256), (unsigned)n_seq);
  if (dtype == CSB_BF16) {
    seq_pool_partial_kernel<true><<<g1, 128, 0, s>>>((const uint4*)hidden, cu_seqlens, Dv, partial);
    int st = check_launch("seq_pool_partial");
    if (st) return st;
    seq_pool_final_kernel<true><<<g2, 256, 0, s>>>(partial, cu_seqlens, (int)Dl, out);
  } else {
    seq_pool_partial_kernel<false><<<g1, 128, 0, s>>>((const uint4*)hidden, cu_seqlens, Dv, partial);
    int st = check_launch("seq_pool_partial");
    if (st) return st;
    seq_pool_final_kernel<false><<<g2, 256, 0, s>>>(partial, cu_seqlens, (int)Dl, out);
  }
  return check_launch("seq_pool_final");
}

}  // extern "C"
