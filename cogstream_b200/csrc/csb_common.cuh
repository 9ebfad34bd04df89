// Shared helpers for the cogstream_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cogstream_b200.h"

namespace csb {

// ---- error plumbing (no exceptions cross the C ABI) --------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // counts one launch; cudaGetLastError -> CSB_OK / CSB_ERR_CUDA
void count_launch(int n);            // launches that do not go through check_launch
int device_ok();                     // cached compute-capability check
// split partial sums of the var-len mean pool (embed_pool.cu): partial [n_seq][8][Dl] fp32; shared with the fused relevance step
int launch_seq_pool_partial(const void* hidden, const int32_t* cu_seqlens, int64_t n_seq, int64_t Dl, int dtype, float* partial,
                            cudaStream_t s);

#define CSB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::csb::set_error(__VA_ARGS__);  \
      return CSB_ERR_ARG;             \
    }                                 \
  } while (0)

#define CSB_DEVICE_OR_RETURN()                 \
  do {                                         \
    int _s = ::csb::device_ok();               \
    if (_s != CSB_OK) return _s;               \
  } while (0)

inline cudaStream_t as_stream(csb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;

// ---- device helpers -------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }

// unpack 8 bf16 (one 128-bit load) to fp32
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16);
  f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16);
  f[7] = __uint_as_float(v.w & 0xffff0000u);
}

// round-to-nearest-even fp32 -> bf16 bits (matches ATen's c10::BFloat16 conversion, NaN aside)
__device__ __forceinline__ uint32_t float_to_bf16_bits(float f) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float round_bf16(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

// Deterministic block-wide sum (fixed tree), result broadcast to all threads. `scratch` >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T r = (lane < nwarp) ? scratch[lane] : T(0);
  r = warp_sum(r);
  return r;
}

}  // namespace csb
