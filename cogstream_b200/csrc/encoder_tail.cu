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
#include <cstdlib>

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


// Two fp32 values rounded to bf16 (round-to-nearest-even): one packing instruction for the pair
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}

// Bulk-copy form for C <= 1280 (NV <= 5), bit-identical to ln_merge_kernel.  The column-wise kernel above holds a token's rows in
// registers, so a warp has nothing in flight while it computes (12 warps / SM: 0.53 of the HBM peak).  Here every warp owns a
// private ring of shared-memory stages: lane 0 streams the next token's rows into it with bulk async copies (cp.async.bulk ->
// mbarrier complete_tx) while the warp works on the current one straight out of shared memory, row by row -- a row is unpacked
// ONCE, its fp32 values stay in registers from the mean through the variance (kept as x - mean) to the normalised output, which is
// folded into the running blend (rows 0, 1 -> hi, rows 2, 3 -> lo).  One CTA per SM, 12 warps x 2 stages x 9 KB at C = 1152.
constexpr int kLnMaxWarps = 12;
constexpr int kLnMaxStages = 4;

__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NV>
__global__ void __launch_bounds__(kLnMaxWarps * 32, 1) ln_merge_bulk_kernel(const uint8_t* __restrict__ x, int64_t n_out, int Cv, int m,
                                                                            const uint4* __restrict__ gamma, const uint4* __restrict__ beta,
                                                                            float eps, uint4* __restrict__ out, int stages, int gb_bytes) {
  extern __shared__ __align__(128) uint8_t ln_smem[];
  __shared__ __align__(8) uint64_t bars[kLnMaxWarps][kLnMaxStages];
  // gamma / beta as fp32, split so that consecutive lanes read consecutive 16-byte words: [g 0-3][g 4-7][b 0-3][b 4-7] x [Cv]
  float4* s_gb4 = reinterpret_cast<float4*>(ln_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int p0 = (m - 1) >> 1;            // block-local sample position floor((m-1)/2) ...
  const bool blend = ((m - 1) & 1) != 0;  // ... + 1/2 for even m: the two central rows / columns with weights 1/2
  const uint32_t row_bytes = (uint32_t)Cv * 16u;
  const uint32_t stage_bytes = (blend ? 4u : 1u) * row_bytes;
  uint8_t* ring = ln_smem + gb_bytes + (size_t)warp * stages * stage_bytes;
  for (int c = threadIdx.x; c < Cv; c += blockDim.x) {
    float f[8];
    unpack8(__ldg(gamma + c), f);
    s_gb4[c] = make_float4(f[0], f[1], f[2], f[3]);
    s_gb4[Cv + c] = make_float4(f[4], f[5], f[6], f[7]);
    unpack8(__ldg(beta + c), f);
    s_gb4[2 * Cv + c] = make_float4(f[0], f[1], f[2], f[3]);
    s_gb4[3 * Cv + c] = make_float4(f[4], f[5], f[6], f[7]);
  }
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ln_smem_u32(&bars[warp][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float inv_c = 1.0f / (float)(Cv * 8);
  const int nfull = Cv >> 5, rem = Cv & 31;  // column slots (32 lanes x 16 bytes) that are full / lanes of the partial one
  const int64_t tstride = (int64_t)gridDim.x * nwarps;
  const int64_t t0 = (int64_t)blockIdx.x * nwarps + warp;
  const int64_t blk_bytes = (int64_t)m * m * row_bytes;
  const int64_t off_a = (int64_t)(p0 * m + p0) * row_bytes, off_b = (int64_t)((p0 + 1) * m + p0) * row_bytes;
  auto issue = [&](uint32_t i) {  // lane 0 only: the rows of this warp's i-th token -> ring slot i % stages
    const int64_t t = t0 + (int64_t)i * tstride;
    if (t >= n_out) return;
    const uint32_t slot = i % (uint32_t)stages;
    const uint32_t bar = ln_smem_u32(&bars[warp][slot]);
    const uint32_t dst = ln_smem_u32(ring + (size_t)slot * stage_bytes);
    const uint8_t* src = x + t * blk_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
    if (blend) {  // rows (p0, p0), (p0, p0+1) are adjacent in memory, and so are (p0+1, p0), (p0+1, p0+1)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src + off_a),
                   "r"(2u * row_bytes), "r"(bar)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 2u * row_bytes),
                   "l"(src + off_b), "r"(2u * row_bytes), "r"(bar)
                   : "memory");
    } else {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src + off_a),
                   "r"(row_bytes), "r"(bar)
                   : "memory");
    }
  };
  if (lane == 0)
    for (int s = 0; s < stages; ++s) issue((uint32_t)s);
  uint32_t i = 0;
  for (int64_t t = t0; t < n_out; t += tstride, ++i) {
    const uint32_t slot = i % (uint32_t)stages, parity = (i / (uint32_t)stages) & 1u;
    const uint32_t bar = ln_smem_u32(&bars[warp][slot]);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LD_%=;\n\t"
        "bra LW_%=;\n\t"
        "LD_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
    const uint8_t* stage = ring + (size_t)slot * stage_bytes;
    // hi = y(row 0) + y(row 1), lo = y(row 2) + y(row 3); the blend h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11) with
    // all weights 1/2 is 0.25 * RN(RN(v00 + v01) + RN(v10 + v11)): scaling by a power of two commutes with rounding
    float hi[NV][8], lo[NV][8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r == 0 || blend) {  // uniform
        const uint4* row = reinterpret_cast<const uint4*>(stage + (size_t)r * row_bytes);
        float f[NV][8];
        float s = 0.f;  // same summation order as row_stats()
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          if (u < nfull || (u == nfull && lane < rem)) unpack8(row[lane + 32 * u], f[u]);
          else {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[u][e] = 0.f;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) s += f[u][e];
        }
        const float mean = warp_sum(s) * inv_c;
        float v = 0.f;
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          if (u < nfull) {  // uniform: every lane holds a valid column
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float d = f[u][e] - mean;
              f[u][e] = d;
              v = fmaf(d, d, v);
            }
          } else if (u == nfull) {  // the partial slot: columns beyond C contribute nothing
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float d = lane < rem ? f[u][e] - mean : 0.f;
              f[u][e] = d;
              v = fmaf(d, d, v);
            }
          }
        }
        const float rstd = rsqrtf(warp_sum(v) * inv_c + eps);
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          if (u < nfull || (u == nfull && rem > 0)) {  // uniform; lanes beyond C in the partial slot compute on zeros, never stored
            const int c = min(lane + 32 * u, Cv - 1);
            const float4 g0 = s_gb4[c], g1 = s_gb4[Cv + c], b0 = s_gb4[2 * Cv + c], b1 = s_gb4[3 * Cv + c];
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              // LayerNorm output rounded to bf16 (a bf16 tensor in the reference), back in fp32 for the blend
              const uint32_t pk = pack_bf16x2(fmaf(gg[e], __fmul_rn(rstd, f[u][e]), bb[e]), fmaf(gg[e + 1], __fmul_rn(rstd, f[u][e + 1]), bb[e + 1]));
              const float y0 = __uint_as_float(pk << 16), y1 = __uint_as_float(pk & 0xffff0000u);
              if (r == 0) { hi[u][e] = y0; hi[u][e + 1] = y1; }
              else if (r == 1) { hi[u][e] = __fadd_rn(hi[u][e], y0); hi[u][e + 1] = __fadd_rn(hi[u][e + 1], y1); }
              else if (r == 2) { lo[u][e] = y0; lo[u][e + 1] = y1; }
              else { lo[u][e] = __fadd_rn(lo[u][e], y0); lo[u][e + 1] = __fadd_rn(lo[u][e + 1], y1); }
            }
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // our reads of the slot precede the next bulk write to it
    __syncwarp();
    if (lane == 0) issue(i + (uint32_t)stages);
    uint4* o = out + t * Cv;
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      if (u < nfull || (u == nfull && rem > 0)) {  // uniform
        float a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = blend ? __fmul_rn(0.25f, __fadd_rn(hi[u][e], lo[u][e])) : hi[u][e];
        uint4 w;
        w.x = pack_bf16x2(a[0], a[1]);
        w.y = pack_bf16x2(a[2], a[3]);
        w.z = pack_bf16x2(a[4], a[5]);
        w.w = pack_bf16x2(a[6], a[7]);
        if (u < nfull || lane < rem) o[lane + 32 * u] = w;
      }
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
  // bulk-copy kernel (C <= 1280): as many warps (<= 12) and ring stages (2-4) as fit next to the fp32 gamma / beta
  const bool columnwise = getenv("CSB_LN_COLUMNWISE") != nullptr;  // validation: the column-wise kernel for every width
  if (Cv <= 160 && !columnwise) {
    const size_t gb = (smem + 127) & ~(size_t)127;
    const size_t stage_bytes = (size_t)((merge & 1) ? 1 : 4) * (size_t)C * 2;
    const size_t budget = 227 * 1024 - 1024 - gb;
    int nw = kLnMaxWarps, st = 0;
    for (; nw >= 4; nw -= 2) {
      st = (int)(budget / ((size_t)nw * stage_bytes));
      if (st >= 2) break;
    }
    if (st >= 2 && nw >= 4) {
      if (st > kLnMaxStages) st = kLnMaxStages;
      const size_t dyn = gb + (size_t)nw * st * stage_bytes;
      int64_t nblk = (n_out + nw - 1) / nw;
      if (nblk > (int64_t)kNumSMs) nblk = kNumSMs;  // persistent, one CTA per SM: warps stride over the output tokens
#define CSB_LN_BULK_LAUNCH(NVV)                                                                                                        \
  do {                                                                                                                                 \
    cudaError_t e = cudaFuncSetAttribute(ln_merge_bulk_kernel<NVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);            \
    if (e != cudaSuccess) { set_error("ln_merge: smem attr: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }                        \
    ln_merge_bulk_kernel<NVV><<<(unsigned)nblk, nw * 32, dyn, s>>>((const uint8_t*)hidden, n_out, Cv, (int)merge, (const uint4*)gamma, \
                                                                   (const uint4*)beta, (float)eps, (uint4*)out, st, (int)gb);          \
  } while (0)
      if (Cv <= 64) CSB_LN_BULK_LAUNCH(2);
      else CSB_LN_BULK_LAUNCH(5);
#undef CSB_LN_BULK_LAUNCH
      return check_launch("ln_merge");
    }
  }
  if (Cv <= 64) CSB_LN_LAUNCH(2);
  else if (Cv <= 160) CSB_LN_LAUNCH(5);
  else CSB_LN_LAUNCH(8);  // wide rows: three fp32 copies of a row do not fit the register file
#undef CSB_LN_LAUNCH
  return check_launch("ln_merge");
}

}  // extern "C"
