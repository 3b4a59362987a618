// Device / host helpers shared by the tcgen05 convolution kernels (conv_tc.cu, conv_halo.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace dyk {

// Division by a runtime constant without the ~30-instruction integer-division sequence (the per-tile coordinate
// math used to be most of the epilogue's instruction count on small tiles): q = (umulhi(n, mul) + n) >> shr,
// valid for 0 <= n < 2^31 (Granlund-Montgomery round-up method).
struct FastDiv {
  unsigned mul, shr, div;
};
static inline FastDiv make_fastdiv(unsigned d) {
  FastDiv f;
  f.div = d;
  unsigned l = 0;
  while ((1ull << l) < d) ++l;
  f.shr = l;
  f.mul = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  return f;
}
__device__ __forceinline__ unsigned fd_div(unsigned n, const FastDiv& f) { return (__umulhi(n, f.mul) + n) >> f.shr; }

template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (kBf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool kBf16>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (kBf16) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&v));
  }
}


}  // namespace dyk
