// Activation functions of the reference's [convolutional] blocks (models.py:51-64), evaluated in fp32.
#pragma once
#include "../../include/dyk_b200.h"

namespace dyk {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mish_f(float x) {
  // x * tanh(softplus(x)) = x * n / (n + 2),  n = e^x (e^x + 2); for x >= 20 the ratio rounds to 1 in fp32, so
  // clamping the exponent (no overflow to inf / NaN) replaces a divergent branch.
  // The bare MUFU instructions are used: with x <= 20 the operands stay far inside the range where __expf / __fdividef
  // add their scaling fix-ups (8 FMUL + 2 FSETP per element in the SASS of the conv epilogue, now 4 FMUL), and a flushed
  // e^x for x < -87 gives the correct limit 0.  9 instructions per element instead of 16.
  const float e = ex2_approx(fminf(x, 20.f) * 1.4426950408889634f);
  const float n = fmaf(e, e, e + e);
  return x * (n * rcp_approx(n + 2.f));
}

// compile-time activation (conv epilogue)
template <int kAct>
__device__ __forceinline__ float act_apply(float x) {
  if constexpr (kAct == DYK_ACT_LEAKY) return fmaxf(x, 0.1f * x);
  else if constexpr (kAct == DYK_ACT_MISH) return mish_f(x);
  else if constexpr (kAct == DYK_ACT_RELU) return fmaxf(x, 0.f);
  else if constexpr (kAct == DYK_ACT_RELU6) return fminf(fmaxf(x, 0.f), 6.f);
  else if constexpr (kAct == DYK_ACT_HARDSWISH) return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
  else if constexpr (kAct == DYK_ACT_HARDSIGMOID) return fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
  else return x;
}

// Run a statement with the activation id as the compile-time constant kAct.  The per-element kernels (train.cu, conv_direct.cu) are
// specialised on it: with a run-time `switch` inside the element loop the compiler emitted an indexed branch (BRX) per
// element, and the BN apply kernels ran at 3.0 - 4.2 TB/s where a plain 16-bit axpby reaches 6.1 TB/s.
#define DYK_DISPATCH_ACT(act, ...)                                                          \
  do {                                                                                      \
    switch (act) {                                                                          \
      case DYK_ACT_LEAKY: { constexpr int kAct = DYK_ACT_LEAKY; __VA_ARGS__; } break;        \
      case DYK_ACT_MISH: { constexpr int kAct = DYK_ACT_MISH; __VA_ARGS__; } break;          \
      case DYK_ACT_RELU: { constexpr int kAct = DYK_ACT_RELU; __VA_ARGS__; } break;          \
      case DYK_ACT_RELU6: { constexpr int kAct = DYK_ACT_RELU6; __VA_ARGS__; } break;        \
      case DYK_ACT_HARDSWISH: { constexpr int kAct = DYK_ACT_HARDSWISH; __VA_ARGS__; } break; \
      case DYK_ACT_HARDSIGMOID: { constexpr int kAct = DYK_ACT_HARDSIGMOID; __VA_ARGS__; } break; \
      default: { constexpr int kAct = DYK_ACT_LINEAR; __VA_ARGS__; } break;                  \
    }                                                                                       \
  } while (0)

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case DYK_ACT_LEAKY: return x > 0.f ? x : 0.1f * x;
    case DYK_ACT_MISH: return mish_f(x);
    case DYK_ACT_RELU: return fmaxf(x, 0.f);
    case DYK_ACT_RELU6: return fminf(fmaxf(x, 0.f), 6.f);
    case DYK_ACT_HARDSWISH: return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    case DYK_ACT_HARDSIGMOID: return fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    default: return x;
  }
}

}  // namespace dyk
