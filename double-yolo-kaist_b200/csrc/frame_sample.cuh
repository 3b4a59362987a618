// The caller's NCHW frames (fp32 in [0, 1], or uint8 divided by 255) seen through an optional bilinear resize:
// F.interpolate(imgs, size=(H, W), mode='bilinear', align_corners=False) of the multi-scale training loop
// (train_utils/kaist_train_eval_utils.py:59-71) evaluated on the fly by the kernels that read the frames (stem convolution,
// stem weight-gradient operand), so the resized batch is never written to memory.
#pragma once
#include <cstdint>

namespace dyk {

// PyTorch's area_pixel_compute_source_index (align_corners = False, not cubic) and upsample_bilinear2d arithmetic:
//   src = scale * (dst + 0.5) - 0.5, clamped at 0;  i0 = (int)src, i1 = i0 + (i0 < size - 1), lambda1 = src - i0, lambda0 = 1 - lambda1
//   val = h0 * (w0 * p[i0][j0] + w1 * p[i0][j1]) + h1 * (w0 * p[i1][j0] + w1 * p[i1][j1]),   scale = (float)src_size / dst_size
struct ResizeAxis {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ ResizeAxis resize_axis(int dst, float scale, int src_size) {
  ResizeAxis a;
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  a.i0 = static_cast<int>(s);
  if (a.i0 > src_size - 1) a.i0 = src_size - 1;
  a.i1 = a.i0 + (a.i0 < src_size - 1 ? 1 : 0);
  a.l1 = s - static_cast<float>(a.i0);
  a.l0 = 1.f - a.l1;
  return a;
}

template <typename TIn>
__device__ __forceinline__ float frame_value(TIn raw) {
  if constexpr (sizeof(TIn) == 1) return static_cast<float>(raw) * (1.0f / 255.0f);
  else return static_cast<float>(raw);
}

// one resized pixel of plane `p` (Hs x Ws); for uint8 the interpolation runs on the raw bytes and the result is scaled by
// 1/255 (the reference divides first: the same value up to fp32 rounding, far below the 16-bit rounding that follows)
template <typename TIn>
__device__ __forceinline__ float frame_bilinear(const TIn* __restrict__ p, int Ws, const ResizeAxis& h, const ResizeAxis& w) {
  const TIn* r0 = p + static_cast<long long>(h.i0) * Ws;
  const TIn* r1 = p + static_cast<long long>(h.i1) * Ws;
  const float v00 = static_cast<float>(__ldg(r0 + w.i0)), v01 = static_cast<float>(__ldg(r0 + w.i1));
  const float v10 = static_cast<float>(__ldg(r1 + w.i0)), v11 = static_cast<float>(__ldg(r1 + w.i1));
  const float v = h.l0 * (w.l0 * v00 + w.l1 * v01) + h.l1 * (w.l0 * v10 + w.l1 * v11);
  if constexpr (sizeof(TIn) == 1) return v * (1.0f / 255.0f);
  else return v;
}

}  // namespace dyk
