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

// Residual rows of the tiles a CTA will reach a few iterations from now, requested into L2 ahead of time.  The epilogue
// loads the residual of a tile when it gets there; on layers with a short main loop (few input channels, large maps) that
// load's DRAM latency (2 - 3 us under load) used to be exposed once per tile or even per column chunk and set the whole
// tile period (4.3 us per 128 x 128 tile on 64->128 3x3 @128x160, three times the HBM time).  bytes % 16 == 0, 16-byte aligned.
constexpr int kResPrefetchTiles = 4;
// mode (DYK_RES_PF, read once on the host): 0 = off, 1 = one prefetch.global.L2 per 128-byte line, 2 = one bulk prefetch per row
__device__ __forceinline__ void l2_prefetch_row(const void* gptr, unsigned bytes, int mode) {
  if (mode == 2) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
  } else {
    const char* c = reinterpret_cast<const char*>(gptr);
    for (unsigned o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + o));
  }
}
int res_prefetch_mode();   // conv_tc.cu

}  // namespace dyk
