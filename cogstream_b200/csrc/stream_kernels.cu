// HBM-bound streaming kernels over the bf16 feature tensor x [T, P*D]:
//   pool_norm, centroid_update, dist_to_assigned, gather_rows, pool_frames_inplace.
// All use 128-bit coalesced loads (8 bf16 per lane), fp32 accumulation in a fixed order
// (deterministic, atomic-free) and size their grids from the data, not from a fixed tile.
#include "csb_common.cuh"

namespace csb {

// ------------------------------------------------------------------------------------------------
// (a) pool + squared norm: one CTA per frame; a thread owns 8 consecutive channels and walks the P
// tokens with 8 loads in flight.  Algorithmic bytes: 2*T*P*D read, (2|4)*T*D + 4*T written.
// ------------------------------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256) pool_norm_kernel(const uint4* __restrict__ x, int P, int Dv /* D/8 */,
                                                        uint4* __restrict__ pooled_bf16, float4* __restrict__ pooled_f32,
                                                        float* __restrict__ sqnorm) {
  __shared__ double red[32];
  const int64_t t = blockIdx.x;
  const uint4* xf = x + t * (int64_t)P * Dv;
  const float invP = 1.0f / (float)P;  // only used when P is a power of two (exact); else true division below
  (void)invP;
  float ss = 0.f;
  for (int c = threadIdx.x; c < Dv; c += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int p = 0;
    for (; p + UNROLL <= P; p += UNROLL) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(xf + (int64_t)(p + u) * Dv + c);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[e] += f[e];
          ss = fmaf(f[e], f[e], ss);
        }
      }
    }
    for (; p < P; ++p) {
      float f[8];
      unpack8(ldg_stream(xf + (int64_t)p * Dv + c), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e] += f[e];
        ss = fmaf(f[e], f[e], ss);
      }
    }
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = __fdiv_rn(acc[e], (float)P);
    if (pooled_f32) {
      pooled_f32[(t * Dv + c) * 2 + 0] = make_float4(m[0], m[1], m[2], m[3]);
      pooled_f32[(t * Dv + c) * 2 + 1] = make_float4(m[4], m[5], m[6], m[7]);
    }
    if (pooled_bf16) {
      uint4 o;
      o.x = float_to_bf16_bits(m[0]) | (float_to_bf16_bits(m[1]) << 16);
      o.y = float_to_bf16_bits(m[2]) | (float_to_bf16_bits(m[3]) << 16);
      o.z = float_to_bf16_bits(m[4]) | (float_to_bf16_bits(m[5]) << 16);
      o.w = float_to_bf16_bits(m[6]) | (float_to_bf16_bits(m[7]) << 16);
      pooled_bf16[t * Dv + c] = o;
    }
  }
  if (sqnorm) {
    double tot = block_sum<double>((double)ss, red);
    if (threadIdx.x == 0) sqnorm[t] = (float)tot;
  }
}

// ------------------------------------------------------------------------------------------------
// centroid update: grid (column tiles, clusters); a thread owns 8 consecutive columns of one centre
// and adds the member rows in ascending frame order.  Bytes: 2*K*(sum n_j) read + 4*k*K written.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) centroid_update_kernel(const uint4* __restrict__ x, int64_t Kv,
                                                              const int32_t* __restrict__ offsets,
                                                              const int32_t* __restrict__ members,
                                                              float4* __restrict__ C) {
  const int j = blockIdx.y;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Kv) return;
  const int beg = offsets[j], end = offsets[j + 1];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int m = beg;
  for (; m + 4 <= end; m += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldg_stream(x + (int64_t)members[m + u] * Kv + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  }
  for (; m < end; ++m) {
    float f[8];
    unpack8(ldg_stream(x + (int64_t)members[m] * Kv + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  }
  const float n = (float)(end - beg);
  float r[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) r[e] = (end > beg) ? __fdiv_rn(acc[e], n) : 0.f;
  float4* out = C + ((int64_t)j * Kv + c) * 2;
  out[0] = make_float4(r[0], r[1], r[2], r[3]);
  out[1] = make_float4(r[4], r[5], r[6], r[7]);
}

// ------------------------------------------------------------------------------------------------
// centroid update fused with the per-event token gather: the member rows of cluster j are read once, added in
// ascending order AND stored to their event-major position.  Bytes: 2*T*K read + 2*T*K + 4*k*K written.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) centroid_gather_kernel(const uint4* __restrict__ x, int64_t Kv,
                                                              const int32_t* __restrict__ offsets,
                                                              const int32_t* __restrict__ members,
                                                              const int32_t* __restrict__ ev_offsets, float4* __restrict__ C,
                                                              uint4* __restrict__ packed) {
  const int j = blockIdx.y;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Kv) return;
  const int beg = offsets[j], end = offsets[j + 1];
  const int ev0 = ev_offsets[j], nev = ev_offsets[j + 1] - ev0;  // rows that belong to the event (0 for a re-seeded centre)
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int m = beg;
  for (; m + 4 <= end; m += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ldg_stream(x + (int64_t)members[m + u] * Kv + c);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
      if (packed && m + u - beg < nev) packed[(int64_t)(ev0 + m + u - beg) * Kv + c] = v[u];
    }
  }
  for (; m < end; ++m) {
    const uint4 v = ldg_stream(x + (int64_t)members[m] * Kv + c);
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
    if (packed && m - beg < nev) packed[(int64_t)(ev0 + m - beg) * Kv + c] = v;
  }
  if (C) {
    const float n = (float)(end - beg);
    float r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = (end > beg) ? __fdiv_rn(acc[e], n) : 0.f;
    float4* out = C + ((int64_t)j * Kv + c) * 2;
    out[0] = make_float4(r[0], r[1], r[2], r[3]);
    out[1] = make_float4(r[4], r[5], r[6], r[7]);
  }
}

// ------------------------------------------------------------------------------------------------
// distance of every frame to its own centroid: one CTA per frame, fp32 lanes, fp64 block reduce.
// Bytes: 2*T*K (x) + 4*K per frame of centroid (L2-resident across the ~T/k frames of an event).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dist_to_assigned_kernel(const uint4* __restrict__ x, int64_t Kv,
                                                               const float4* __restrict__ C,
                                                               const int64_t* __restrict__ assign,
                                                               float* __restrict__ dist) {
  __shared__ double red[32];
  const int64_t i = blockIdx.x;
  const int64_t a = assign[i];
  const uint4* xr = x + i * Kv;
  const float4* cr = C + a * Kv * 2;
  float s0 = 0.f, s1 = 0.f;
  for (int64_t c = threadIdx.x; c < Kv; c += blockDim.x) {
    float f[8];
    unpack8(ldg_stream(xr + c), f);
    const float4 c0 = __ldg(cr + c * 2), c1 = __ldg(cr + c * 2 + 1);
    float d;
    d = f[0] - c0.x; s0 = fmaf(d, d, s0);
    d = f[1] - c0.y; s1 = fmaf(d, d, s1);
    d = f[2] - c0.z; s0 = fmaf(d, d, s0);
    d = f[3] - c0.w; s1 = fmaf(d, d, s1);
    d = f[4] - c1.x; s0 = fmaf(d, d, s0);
    d = f[5] - c1.y; s1 = fmaf(d, d, s1);
    d = f[6] - c1.z; s0 = fmaf(d, d, s0);
    d = f[7] - c1.w; s1 = fmaf(d, d, s1);
  }
  double tot = block_sum<double>((double)s0 + (double)s1, red);
  if (threadIdx.x == 0) dist[i] = (float)sqrt(tot);
}

// ------------------------------------------------------------------------------------------------
// row gather: dst[r] = src[index[r]], 16-byte vectors, grid-stride over (row, vector) pairs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ src, int64_t vpr,
                                                          const int32_t* __restrict__ index, int64_t n_host,
                                                          const int64_t* __restrict__ n_dev, uint4* __restrict__ dst) {
  const int64_t n = n_dev ? *n_dev : n_host;
  const int64_t total = n * vpr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); g0 < total; g0 += stride) {
    uint4 v[4];
    int64_t g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      g[u] = g0 + (int64_t)u * gridDim.x * blockDim.x;
      if (g[u] < total) {
        const int64_t r = g[u] / vpr, c = g[u] - r * vpr;
        v[u] = ldg_stream(src + (int64_t)index[r] * vpr + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (g[u] < total) dst[g[u]] = v[u];
  }
}

// ------------------------------------------------------------------------------------------------
// token-0 pooling of listed frames, in place (the caller cloned the tensor).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_frames_inplace_kernel(uint4* __restrict__ x, int P, int Dv,
                                                                  const int32_t* __restrict__ frames, int64_t n_host,
                                                                  const int64_t* __restrict__ n_dev) {
  const int64_t n = n_dev ? *n_dev : n_host;
  if ((int64_t)blockIdx.x >= n) return;
  const int64_t t = frames[blockIdx.x];
  uint4* xf = x + t * (int64_t)P * Dv;
  for (int c = threadIdx.x; c < Dv; c += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int p = 0;
    for (; p + 4 <= P; p += 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = xf[(int64_t)(p + u) * Dv + c];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    }
    for (; p < P; ++p) {
      float f[8];
      unpack8(xf[(int64_t)p * Dv + c], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = __fdiv_rn(acc[e], (float)P);
    uint4 o;
    o.x = float_to_bf16_bits(m[0]) | (float_to_bf16_bits(m[1]) << 16);
    o.y = float_to_bf16_bits(m[2]) | (float_to_bf16_bits(m[3]) << 16);
    o.z = float_to_bf16_bits(m[4]) | (float_to_bf16_bits(m[5]) << 16);
    o.w = float_to_bf16_bits(m[6]) | (float_to_bf16_bits(m[7]) << 16);
    xf[c] = o;
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace csb

using namespace csb;

extern "C" {

int csb_pool_norm(const void* x, int64_t T, int64_t P, int64_t D, void* pooled_bf16, float* pooled_f32, float* sqnorm,
                  csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(x && T > 0 && P > 0 && D > 0, "pool_norm: null input or empty shape");
  CSB_REQUIRE(D % 8 == 0 && aligned16(x), "pool_norm: D must be a multiple of 8 and x 16-byte aligned (D=%lld)", (long long)D);
  CSB_REQUIRE(P < (1 << 24) && D < (1ll << 31), "pool_norm: shape too large");
  const int Dv = (int)(D / 8);
  const int threads = Dv >= 256 ? 256 : ((Dv + 31) / 32) * 32;
  pool_norm_kernel<8><<<(unsigned)T, threads, 0, as_stream(stream)>>>((const uint4*)x, (int)P, Dv, (uint4*)pooled_bf16,
                                                                       (float4*)pooled_f32, sqnorm);
  return check_launch("pool_norm");
}

int csb_centroid_update(const void* x, int64_t T, int64_t K, const int32_t* offsets, const int32_t* members, int64_t k,
                        float* centroids, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(x && offsets && members && centroids && T > 0 && k > 0, "centroid_update: null argument");
  CSB_REQUIRE(K % 8 == 0 && aligned16(x) && aligned16(centroids), "centroid_update: K %% 8 != 0 or misaligned");
  CSB_REQUIRE(k <= 65535, "centroid_update: k > 65535");
  const int64_t Kv = K / 8;
  dim3 grid((unsigned)((Kv + 255) / 256), (unsigned)k);
  centroid_update_kernel<<<grid, 256, 0, as_stream(stream)>>>((const uint4*)x, Kv, offsets, members, (float4*)centroids);
  return check_launch("centroid_update");
}

int csb_centroid_gather(const void* x, int64_t T, int64_t K, const int32_t* offsets, const int32_t* members,
                        const int32_t* event_offsets, int64_t k, float* centroids, void* packed, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(x && offsets && members && event_offsets && (centroids || packed) && T > 0 && k > 0, "centroid_gather: null argument");
  CSB_REQUIRE(K % 8 == 0 && aligned16(x) && (!centroids || aligned16(centroids)) && (!packed || aligned16(packed)),
              "centroid_gather: K %% 8 != 0 or misaligned");
  CSB_REQUIRE(k <= 65535, "centroid_gather: k > 65535");
  const int64_t Kv = K / 8;
  dim3 grid((unsigned)((Kv + 255) / 256), (unsigned)k);
  centroid_gather_kernel<<<grid, 256, 0, as_stream(stream)>>>((const uint4*)x, Kv, offsets, members, event_offsets,
                                                               (float4*)centroids, (uint4*)packed);
  return check_launch("centroid_gather");
}

int csb_dist_to_assigned(const void* x, int64_t T, int64_t K, const float* centroids, int64_t k, const int64_t* assignments,
                         float* dist, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(x && centroids && assignments && dist && T > 0 && k > 0, "dist_to_assigned: null argument");
  CSB_REQUIRE(K % 8 == 0 && aligned16(x) && aligned16(centroids), "dist_to_assigned: K %% 8 != 0 or misaligned");
  dist_to_assigned_kernel<<<(unsigned)T, 256, 0, as_stream(stream)>>>((const uint4*)x, K / 8, (const float4*)centroids,
                                                                      assignments, dist);
  return check_launch("dist_to_assigned");
}

int csb_gather_rows(const void* src, int64_t row_bytes, const int32_t* index, int64_t n_host, const int64_t* n_dev,
                    int64_t n_max, void* dst, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(src && index && dst, "gather_rows: null argument");
  CSB_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0 && aligned16(src) && aligned16(dst), "gather_rows: rows must be 16-byte multiples");
  if (n_max <= 0) return CSB_OK;
  const int64_t vpr = row_bytes / 16;
  int64_t blocks = (n_max * vpr + 256 * 4 - 1) / (256 * 4);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const uint4*)src, vpr, index, n_host, n_dev, (uint4*)dst);
  return check_launch("gather_rows");
}

int csb_pool_frames_inplace(void* x, int64_t T, int64_t P, int64_t D, const int32_t* frame_list, int64_t n_host,
                            const int64_t* n_dev, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(x && T > 0 && P > 0 && D > 0, "pool_frames_inplace: null input or empty shape");
  CSB_REQUIRE(D % 8 == 0 && aligned16(x), "pool_frames_inplace: D %% 8 != 0 or misaligned");
  const int64_t n_max = n_dev ? T : n_host;
  if (n_max <= 0) return CSB_OK;
  CSB_REQUIRE(frame_list, "pool_frames_inplace: null frame list");
  const int Dv = (int)(D / 8);
  const int threads = Dv >= 256 ? 256 : ((Dv + 31) / 32) * 32;
  pool_frames_inplace_kernel<<<(unsigned)n_max, threads, 0, as_stream(stream)>>>((uint4*)x, (int)P, Dv, frame_list, n_host, n_dev);
  return check_launch("pool_frames_inplace");
}

}  // extern "C"
