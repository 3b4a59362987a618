// 8-element (16-byte) vector helpers for fp16 / bf16 NHWC tensors.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace dyk {

template <bool kBf16>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t;
    if constexpr (kBf16) t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    else t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <bool kBf16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kBf16) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// WeightedFeatureFusion arithmetic (build_utils/layers.py:82-84: x * w[0] + a * w[1]) with one fixed evaluation order, so
// that the stand-alone kernel and the convolution that forms the sum while staging its operand give the same bits.
__device__ __forceinline__ float fuse2(float x, float a, float w0, float w1) { return __fmaf_rn(x, w0, __fmul_rn(a, w1)); }

template <bool kBf16>
__device__ __forceinline__ float load1(const void* p, long long idx) {
  if constexpr (kBf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
  else return __half2float(reinterpret_cast<const __half*>(p)[idx]);
}
template <bool kBf16>
__device__ __forceinline__ void store1(void* p, long long idx, float v) {
  if constexpr (kBf16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[idx] = __float2half_rn(v);
}

// Dispatch a callable templated on <bool kBf16> from the runtime dtype id.
#define DYK_DISPATCH_DTYPE(dtype, ...)                       \
  do {                                                       \
    if ((dtype) == DYK_BF16) {                               \
      constexpr bool kBf16 = true;                           \
      __VA_ARGS__;                                           \
    } else {                                                 \
      constexpr bool kBf16 = false;                          \
      __VA_ARGS__;                                           \
    }                                                        \
  } while (0)

}  // namespace dyk
