// Encoder tail: the vision tower's post_layernorm + the m x m bilinear token merge in ONE pass over the
// encoder's last hidden states (reference: model/modeling_videollama3_encoder.py:479 post_layernorm, :485-502 view / permute /
// reshape / permute / F.interpolate(mode='bilinear') / permute / view).
//
// The reference materialises the LayerNorm output [n, C], a spatially re-ordered copy of it, the interpolated [n / m^2, C] tensor
// and a last permuted copy: ~4.3x the bytes of one read of the hidden states.  The hidden states arrive merge-block-contiguous
// (the m*m patches of a merged token are consecutive rows, image_processing_videollama3.py:326-345), so the whole chain is
// "normalise the rows of a block, blend them": a warp per output token, rows held in registers, one read, one (1 / m^2 sized)
// write.  F.interpolate(size = grid / m, bilinear, align_corners=False) samples block-local position (m-1)/2 in both directions:
// for even m the four central patches with weights 1/4, for odd m the central patch alone.  Arithmetic follows ATen's CUDA
// kernels op by op: LayerNorm in fp32 (biased variance), rounded to bf16; the blend h0*(w0*a + w1*b) + h1*(w0*c + w1*d) in fp32,
// rounded to bf16.  HBM-bound: 2*n*C bytes read + 2*n*C/m^2 written.
#include "csb_common.cuh"

namespace csb {

// fp32 mean and 1/std of one row held in registers (raw = the lane's 16-byte slices; columns beyond C were loaded as zeros)
template <int NV>
__device__ __forceinline__ void row_stats(const uint4 (&raw)[NV], int Cv, int lane, float eps, float inv_c, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    float f[8];
    unpack8(raw[u], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += f[e];
  }
  mean = warp_sum(s) * inv_c;
  float v = 0.f;
#pragma unroll
  for (int u = 0; u < NV; ++u) {
    if (lane + 32 * u < Cv) {
      float f[8];
      unpack8(raw[u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = f[e] - mean;
        v = fmaf(d, d, v);
      }
    }
  }
  rstd = rsqrtf(warp_sum(v) * inv_c + eps);
}

// LayerNorm output (a bf16 tensor in the reference) of 8 channels of one row
__device__ __forceinline__ void ln8(const uint4& raw, float mean, float rstd, const float (&g)[8], const float (&b)[8], float (&y)[8]) {
  float f[8];
  unpack8(raw, f);
#pragma unroll
  for (int e = 0; e < 8; ++e) y[e] = round_bf16(fmaf(g[e], rstd * (f[e] - mean), b[e]));
}

// A warp per output token.  All rows of the block that carry weight (4 for even m, 1 for odd m) are requested before any is
// processed (4 rows x NV x 16 B per lane in flight, 10 KB per warp at C = 1152); the statistics are taken from the raw registers
// and the normalise + blend + store step walks the columns 8 at a time, so the live state stays below 170 registers: 12 warps / SM.
template <int NV>
__global__ void __launch_bounds__(192, 2) ln_merge_kernel(const uint4* __restrict__ x, int64_t n_out, int Cv, int m, const uint4* __restrict__ gamma,
                                                          const uint4* __restrict__ beta, float eps, uint4* __restrict__ out) {
  extern __shared__ float s_gb[];  // gamma [C], beta [C] as fp32
  float* s_g = s_gb;
  float* s_b = s_gb + Cv * 8;
  const int lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < Cv; c += blockDim.x) {
    float f[8];
    unpack8(__ldg(gamma + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) s_g[c * 8 + e] = f[e];
    unpack8(__ldg(beta + c), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) s_b[c * 8 + e] = f[e];
  }
  __syncthreads();
  const float inv_c = 1.0f / (float)(Cv * 8);
  const int p0 = (m - 1) >> 1;            // block-local sample position floor((m-1)/2) ...
  const bool blend = ((m - 1) & 1) != 0;  // ... + 1/2 for even m: the two central rows / columns with weights 1/2
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n_out; t += warps) {
    const uint4* blk = x + t * (int64_t)m * m * Cv;
    const uint4* rows[4] = {blk + (int64_t)(p0 * m + p0) * Cv, blk + (int64_t)(p0 * m + p0 + 1) * Cv, blk + (int64_t)((p0 + 1) * m + p0) * Cv,
                            blk + (int64_t)((p0 + 1) * m + p0 + 1) * Cv};
    uint4 raw[4][NV];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int c = lane + 32 * u;
        raw[r][u] = (c < Cv && (r == 0 || blend)) ? ldg_stream(rows[r] + c) : make_uint4(0, 0, 0, 0);
      }
    float mean[4], rstd[4];
    row_stats<NV>(raw[0], Cv, lane, eps, inv_c, mean[0], rstd[0]);
    if (blend) {
#pragma unroll
      for (int r = 1; r < 4; ++r) row_stats<NV>(raw[r], Cv, lane, eps, inv_c, mean[r], rstd[r]);
    }
    uint4* o = out + t * Cv;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int c = lane + 32 * u;
      if (c >= Cv) continue;
      const float4 g0 = *reinterpret_cast<const float4*>(s_g + c * 8), g1 = *reinterpret_cast<const float4*>(s_g + c * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(s_b + c * 8), b1 = *reinterpret_cast<const float4*>(s_b + c * 8 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float acc[8];
      ln8(raw[0][u], mean[0], rstd[0], gg, bb, acc);
      if (blend) {
        // h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11), all weights 1/2 (exact scalings: one rounding per addition)
        float y1[8], y2[8], y3[8];
        ln8(raw[1][u], mean[1], rstd[1], gg, bb, y1);
        ln8(raw[2][u], mean[2], rstd[2], gg, bb, y2);
        ln8(raw[3][u], mean[3], rstd[3], gg, bb, y3);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float hi = __fadd_rn(0.5f * acc[e], 0.5f * y1[e]);
          const float lo = __fadd_rn(0.5f * y2[e], 0.5f * y3[e]);
          acc[e] = __fadd_rn(0.5f * hi, 0.5f * lo);
        }
      }
      uint4 w;
      w.x = float_to_bf16_bits(acc[0]) | (float_to_bf16_bits(acc[1]) << 16);
      w.y = float_to_bf16_bits(acc[2]) | (float_to_bf16_bits(acc[3]) << 16);
      w.z = float_to_bf16_bits(acc[4]) | (float_to_bf16_bits(acc[5]) << 16);
      w.w = float_to_bf16_bits(acc[6]) | (float_to_bf16_bits(acc[7]) << 16);
      o[c] = w;
    }
  }
}

}  // namespace csb

using namespace csb;

extern "C" {

int csb_ln_merge(const void* hidden, int64_t n_rows, int64_t C, int64_t merge, const void* gamma, const void* beta, double eps, void* out,
                 csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(hidden && gamma && beta && out && n_rows > 0 && C > 0 && merge >= 1, "ln_merge: null argument or empty shape");
  CSB_REQUIRE(C % 8 == 0 && C <= 2048, "ln_merge: C must be a multiple of 8, at most 2048 (C=%lld)", (long long)C);
  CSB_REQUIRE(n_rows % (merge * merge) == 0, "ln_merge: %lld rows are not whole %lld x %lld blocks", (long long)n_rows, (long long)merge,
              (long long)merge);
  CSB_REQUIRE(((reinterpret_cast<uintptr_t>(hidden) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
              "ln_merge: buffers must be 16-byte aligned");
  const int64_t n_out = n_rows / (merge * merge);
  const int Cv = (int)(C / 8);
  int64_t blocks = (n_out + 5) / 6;
  if (blocks > (int64_t)kNumSMs * 2) blocks = (int64_t)kNumSMs * 2;  // persistent, two CTAs per SM: warps stride over the output tokens
  cudaStream_t s = as_stream(stream);
  const size_t smem = sizeof(float) * 2 * (size_t)C;
#define CSB_LN_LAUNCH(NVV) \
  ln_merge_kernel<NVV><<<(unsigned)blocks, 192, smem, s>>>((const uint4*)hidden, n_out, Cv, (int)merge, (const uint4*)gamma, (const uint4*)beta, (float)eps, (uint4*)out)
  if (Cv <= 64) CSB_LN_LAUNCH(2);
  else if (Cv <= 160) CSB_LN_LAUNCH(5);
  else CSB_LN_LAUNCH(8);
#undef CSB_LN_LAUNCH
  return check_launch("ln_merge");
}

}  // extern "C"
