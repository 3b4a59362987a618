// CUDA-core convolutions for the two shapes that do not map onto the tensor-core implicit GEMM:
//  * the stem (Cin = 3, models.py:35-36), reading the caller's NCHW fp32 frames directly so that the
//    layout/precision conversion costs no extra pass over the images;
//  * depthwise convolutions (groups == C; models.py:41, build_utils/layers.py:224), which are HBM-bound.
#include "common.h"
#include <cstdlib>
#include "act.cuh"
#include "vec.cuh"

namespace dyk {

// ------------------------------------------------------------------ stem: NCHW fp32 -> NHWC dtype
// One thread = one output pixel x COUT_T output channels; consecutive threads walk along W, so the
// per-plane input reads are coalesced and each thread stores COUT_T*2 contiguous bytes.
template <typename TIn>
__device__ __forceinline__ float load_frame(const TIn* p);
template <>
__device__ __forceinline__ float load_frame<float>(const float* p) { return __ldg(p); }
// uint8 frames are normalised exactly as the reference's callers do: img.float() / 255.0
// (train_utils/kaist_train_eval_utils.py:54-55, evaluate.py:67-68)
template <>
__device__ __forceinline__ float load_frame<uint8_t>(const uint8_t* p) { return __fdiv_rn((float)__ldg(p), 255.f); }

template <int COUT_T, bool kBf16, typename TIn, int kAct>
__global__ void __launch_bounds__(128)
stem_conv_kernel(const TIn* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                 const float* __restrict__ bias, uint8_t* __restrict__ y, long long ys, int N, int H, int W,
                 int Cin, int Cout, int k, int stride, int pad, int Ho, int Wo, int act) {
  extern __shared__ float wsm[];  // [k*k*Cin][COUT_T] for this block's channel group
  const int co0 = blockIdx.y * COUT_T;
  const int taps = k * k * Cin;
  for (int i = threadIdx.x; i < taps * COUT_T; i += blockDim.x) {
    const int t = i / COUT_T, co = i - t * COUT_T;
    // w is [Cout][k][k][Cin]; t = (r*k + s)*Cin + ci
    wsm[i] = (co0 + co < Cout) ? __ldg(&w[(long long)(co0 + co) * taps + t]) : 0.f;
  }
  __syncthreads();
  const long long total = (long long)N * Ho * Wo;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(pix % Wo);
    const int ho = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((long long)Wo * Ho));
    float acc[COUT_T];
#pragma unroll
    for (int c = 0; c < COUT_T; ++c) acc[c] = 0.f;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    for (int r = 0; r < k; ++r) {
      const int h = h0 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int ww = w0 + s;
        if (ww < 0 || ww >= W) continue;
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = load_frame<TIn>(&x[(((long long)n * Cin + ci) * H + h) * W + ww]);
          const float* wp = &wsm[((r * k + s) * Cin + ci) * COUT_T];
#pragma unroll
          for (int c = 0; c < COUT_T; ++c) acc[c] = fmaf(v, wp[c], acc[c]);
        }
      }
    }
    uint8_t* yp = y + (pix * ys + co0) * 2;
#pragma unroll
    for (int c8 = 0; c8 < COUT_T / 8; ++c8) {
      if (co0 + c8 * 8 >= Cout) break;
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int co = co0 + c8 * 8 + q;
        const float sc = scale ? __ldg(&scale[co]) : 1.f;
        const float bi = bias ? __ldg(&bias[co]) : 0.f;
        o[q] = act_apply<kAct>(fmaf(acc[c8 * 8 + q], sc, bi));
      }
      *(reinterpret_cast<uint4*>(yp) + c8) = pack8<kBf16>(o);
    }
  }
}

// ------------------------------------------------------------------ depthwise k x k
// One thread = one output pixel x 8 channels (16-byte vectors, channel fastest => coalesced).
template <bool kBf16>
__global__ void __launch_bounds__(256)
dwconv_kernel(const uint8_t* __restrict__ x, long long xs, const float* __restrict__ w,
              const float* __restrict__ scale, const float* __restrict__ bias, uint8_t* __restrict__ y,
              long long ys, int N, int H, int W, int cv, int k, int stride, int pad, int Ho, int Wo, int act) {
  const int C = cv * 8;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long t = i / cv;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    for (int r = 0; r < k; ++r) {
      const int h = h0 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int ww = w0 + s;
        if (ww < 0 || ww >= W) continue;
        const long long pix = ((long long)n * H + h) * W + ww;
        float f[8];
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(x + pix * xs * 2) + c), f);
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w + (long long)(r * k + s) * C + c * 8));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w + (long long)(r * k + s) * C + c * 8) + 1);
        acc[0] = fmaf(f[0], wa.x, acc[0]); acc[1] = fmaf(f[1], wa.y, acc[1]);
        acc[2] = fmaf(f[2], wa.z, acc[2]); acc[3] = fmaf(f[3], wa.w, acc[3]);
        acc[4] = fmaf(f[4], wb.x, acc[4]); acc[5] = fmaf(f[5], wb.y, acc[5]);
        acc[6] = fmaf(f[6], wb.z, acc[6]); acc[7] = fmaf(f[7], wb.w, acc[7]);
      }
    }
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float sc = scale ? __ldg(&scale[c * 8 + q]) : 1.f;
      const float bi = bias ? __ldg(&bias[c * 8 + q]) : 0.f;
      o[q] = apply_act(fmaf(acc[q], sc, bi), act);
    }
    const long long opix = ((long long)n * Ho + ho) * Wo + wo;
    *(reinterpret_cast<uint4*>(y + opix * ys * 2) + c) = pack8<kBf16>(o);
  }
}

// ------------------------------------------------------------------ depthwise 3x3 / 5x5, strip version
// The kernel above issues k*k 16-byte activation loads + 2*k*k weight loads per output vector and decomposes a 64-bit
// linear index with three divisions per output: instruction / LSU bound at ~5x the HBM time (49 % of the MobileNetV3
// step).  Here a thread owns one 8-channel vector and a horizontal strip of kDwStrip outputs: the input window slides
// along W (k*(strip*stride + k - stride) loads per strip instead of strip*k*k), the weights of one filter row are loaded
// once per input row, and the index is decomposed once per strip in 32 bits.  The accumulation order per output is
// unchanged (r outer, s inner).  A vertical-strip variant with the whole filter in registers was measured slower
// (MobileNetV3 bs64 depthwise total 7.5 ms vs 5.7 ms: 252 registers for 5x5).
constexpr int kDwStrip = 4;
template <bool kBf16, int K, int STRIDE, int kAct>
__global__ void __launch_bounds__(256)
dwconv_strip_kernel(const uint8_t* __restrict__ x, long long xs, const float* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ bias, uint8_t* __restrict__ y,
                     long long ys, int N, int H, int W, int cv, int pad, int Ho, int Wo, int strips_w, int act,
                     unsigned total) {
  const int C = cv * 8;
  constexpr int kWin = (kDwStrip - 1) * STRIDE + K;      // input columns a strip touches
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % (unsigned)cv;
    unsigned t = i / (unsigned)cv;
    const unsigned sw = t % (unsigned)strips_w; t /= (unsigned)strips_w;
    const unsigned ho = t % (unsigned)Ho;
    const unsigned n = t / (unsigned)Ho;
    const int wo0 = sw * kDwStrip;
    const int h0 = (int)ho * STRIDE - pad, w0 = wo0 * STRIDE - pad;
    float acc[kDwStrip][8];
#pragma unroll
    for (int j = 0; j < kDwStrip; ++j)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[j][q] = 0.f;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int h = h0 + r;
      if (h < 0 || h >= H) continue;
      const uint8_t* xrow = x + (((long long)n * H + h) * W) * xs * 2;
      float wr[K][8];
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w + (long long)(r * K + s) * C + c * 8));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w + (long long)(r * K + s) * C + c * 8) + 1);
        wr[s][0] = wa.x; wr[s][1] = wa.y; wr[s][2] = wa.z; wr[s][3] = wa.w;
        wr[s][4] = wb.x; wr[s][5] = wb.y; wr[s][6] = wb.z; wr[s][7] = wb.w;
      }
#pragma unroll
      for (int col = 0; col < kWin; ++col) {
        const int ww = w0 + col;
        if (ww < 0 || ww >= W) continue;
        float f[8];
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(xrow + (long long)ww * xs * 2) + c), f);
#pragma unroll
        for (int j = 0; j < kDwStrip; ++j) {
          const int s = col - j * STRIDE;
          if (s >= 0 && s < K) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[j][q] = fmaf(f[q], wr[s][q], acc[j][q]);
          }
        }
      }
    }
    float sc[8], bi[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sc[q] = scale ? __ldg(&scale[c * 8 + q]) : 1.f;
      bi[q] = bias ? __ldg(&bias[c * 8 + q]) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kDwStrip; ++j) {
      const int wo = wo0 + j;
      if (wo >= Wo) break;
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = act_apply<kAct>(fmaf(acc[j][q], sc[q], bi[q]));
      const long long opix = ((long long)n * Ho + ho) * Wo + wo;
      *(reinterpret_cast<uint4*>(y + opix * ys * 2) + c) = pack8<kBf16>(o);
    }
  }
}

}  // namespace dyk

namespace dyk {
int stem_tc_try(const void* x, const float* w, const float* scale, const float* bias, void* y, int64_t ys, int N, int H,
                int W, int Cin, int Cout, int k, int stride, int pad, int act, int dtype, int x_kind, cudaStream_t stream,
                int Hs, int Ws);
int stem3x3_try(const void* x, const float* w, const float* scale, const float* bias, void* y, int64_t ys, int N, int H, int W,
                int Cin, int Cout, int k, int stride, int pad, int act, int dtype, int x_kind, cudaStream_t stream);
int dwconv_tile_try(const void* x, int64_t xs, const float* w, const float* scale, const float* bias, void* y, int64_t ys,
                    int N, int H, int W, int C, int k, int stride, int pad, int act, int dtype, cudaStream_t stream);
}
using namespace dyk;

extern "C" __attribute__((visibility("default"))) int dyk_conv2d_stem_nchw_fwd(const void* x, const float* w, const float* scale, const float* bias,
                                        void* y, int64_t ys, int32_t N, int32_t H, int32_t W, int32_t Cin,
                                        int32_t Cout, int32_t k, int32_t stride, int32_t pad, int32_t act,
                                        int32_t dtype, int32_t x_kind, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x && w && y, "dyk_conv2d_stem_nchw_fwd: null pointer");
  DYK_REQUIRE(x_kind == 0 || x_kind == 1, "dyk_conv2d_stem_nchw_fwd: x_kind=%d (0 = fp32, 1 = uint8)", x_kind);
  DYK_REQUIRE(Cin >= 1 && Cin <= 4, "dyk_conv2d_stem_nchw_fwd: Cin=%d (expects 1..4)", Cin);
  DYK_REQUIRE(Cout > 0 && Cout % 8 == 0 && ys % 8 == 0 && ys >= Cout, "dyk_conv2d_stem_nchw_fwd: Cout=%d ys=%lld",
              Cout, (long long)ys);
  DYK_REQUIRE(k >= 1 && k <= 7 && stride >= 1 && pad >= 0, "dyk_conv2d_stem_nchw_fwd: k=%d stride=%d", k, stride);
  DYK_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "dyk_conv2d_stem_nchw_fwd: y must be 16-byte aligned");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_conv2d_stem_nchw_fwd: empty output");
  // 3x3 / stride 1 / 3 -> 32 channels (every shipped cfg): tensor-core kernel (conv_stem_tc.cu); DYK_STEM_TC=0 keeps the
  // CUDA-core kernel below, which also serves all other stem shapes
  static const bool stem_tc_off = getenv("DYK_STEM_TC") != nullptr && getenv("DYK_STEM_TC")[0] == '0';
  if (!stem_tc_off && (dtype == DYK_F16 || dtype == DYK_BF16)) {
    const int rc = stem_tc_try(x, w, scale, bias, y, ys, N, H, W, Cin, Cout, k, stride, pad, act, dtype, x_kind, stream, 0, 0);
    if (rc <= 0) return rc;
  }
  const long long total = (long long)N * Ho * Wo;
  {   // 3x3 / Cin 3 / pad 1 stems with 16 or 32 output channels: unrolled kernel, conv_stem3.cu (DYK_STEM_FAST=0: the generic one)
    const int rc = stem3x3_try(x, w, scale, bias, y, ys, N, H, W, Cin, Cout, k, stride, pad, act, dtype, x_kind, stream);
    if (rc <= 0) return rc;
  }
  const int ct = (Cout % 32 == 0) ? 32 : 16;
  long long gx = (total + 127) / 128;
  const long long cap = (long long)num_sms() * 32;
  if (gx > cap) gx = cap;
  const dim3 grid((unsigned)gx, (Cout + ct - 1) / ct);
  const size_t smem = (size_t)k * k * Cin * ct * sizeof(float);
#define DYK_STEM_LAUNCH(CT, TIN)                                                                              \
  DYK_DISPATCH_ACT(act, DYK_DISPATCH_DTYPE(dtype, (stem_conv_kernel<CT, kBf16, TIN, kAct><<<grid, 128, smem, stream>>>( \
                                static_cast<const TIN*>(x), w, scale, bias, (uint8_t*)y, ys, N, H, W, Cin, Cout, \
                                k, stride, pad, Ho, Wo, act))))
  if (ct == 32) {
    if (x_kind == 0) DYK_STEM_LAUNCH(32, float); else DYK_STEM_LAUNCH(32, uint8_t);
  } else {
    if (x_kind == 0) DYK_STEM_LAUNCH(16, float); else DYK_STEM_LAUNCH(16, uint8_t);
  }
#undef DYK_STEM_LAUNCH
  DYK_LAUNCH_OK("stem_conv_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_conv2d_stem_nchw_resize_fwd(const void* x, const float* w, const float* scale,
                                        const float* bias, void* y, int64_t ys, int32_t N, int32_t Hs, int32_t Ws, int32_t H,
                                        int32_t W, int32_t Cin, int32_t Cout, int32_t k, int32_t stride, int32_t pad,
                                        int32_t act, int32_t dtype, int32_t x_kind, void* stream_) {
  DYK_REQUIRE(x && w && y, "dyk_conv2d_stem_nchw_resize_fwd: null pointer");
  DYK_REQUIRE(x_kind == 0 || x_kind == 1, "dyk_conv2d_stem_nchw_resize_fwd: x_kind=%d (0 = fp32, 1 = uint8)", x_kind);
  DYK_REQUIRE(Hs > 0 && Ws > 0 && H > 0 && W > 0 && N > 0, "dyk_conv2d_stem_nchw_resize_fwd: bad sizes");
  DYK_REQUIRE(ys % 8 == 0 && ys >= Cout && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "dyk_conv2d_stem_nchw_resize_fwd: y layout");
  DYK_REQUIRE((dtype == DYK_F16 || dtype == DYK_BF16) && k == 3 && stride == 1 && pad == 1 && Cin == 3 && Cout == 32,
              "dyk_conv2d_stem_nchw_resize_fwd: only the 3x3 / stride 1 / 3 -> 32 channel stem of the shipped cfgs has the fused resize");
  return stem_tc_try(x, w, scale, bias, y, ys, N, H, W, Cin, Cout, k, stride, pad, act, dtype, x_kind,
                     static_cast<cudaStream_t>(stream_), Hs, Ws);
}

extern "C" __attribute__((visibility("default"))) int dyk_dwconv2d_fwd(const void* x, int64_t xs, const float* w, const float* scale, const float* bias,
                                void* y, int64_t ys, int32_t N, int32_t H, int32_t W, int32_t C, int32_t k,
                                int32_t stride, int32_t pad, int32_t act, int32_t dtype, void* stream_) {
  DYK_REQUIRE(x && w && y, "dyk_dwconv2d_fwd: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && ys % 8 == 0, "dyk_dwconv2d_fwd: C=%d / strides must be multiples of 8", C);
  DYK_REQUIRE(k >= 1 && k <= 9 && stride >= 1 && pad >= 0, "dyk_dwconv2d_fwd: k=%d stride=%d", k, stride);
  DYK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w) & 15) == 0,
              "dyk_dwconv2d_fwd: 16-byte alignment");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_dwconv2d_fwd: empty output");
  const int cv = C / 8;
  {  // TMA-staged tile kernel (dwconv_tile.cu) for the 3x3 / 5x5 shapes; DYK_DW_TILE=0 keeps the strip kernel below
    const int rc = dwconv_tile_try(x, xs, w, scale, bias, y, ys, N, H, W, C, k, stride, pad, act, dtype,
                                   static_cast<cudaStream_t>(stream_));
    if (rc <= 0) return rc;
  }
  if ((k == 3 || k == 5) && (stride == 1 || stride == 2)) {
    const int strips = (Wo + kDwStrip - 1) / kDwStrip;
    const long long tot = (long long)N * Ho * strips * cv;
    if (tot < (1ll << 31)) {
      long long gs = (tot + 255) / 256;
      if (gs > (long long)num_sms() * 16) gs = (long long)num_sms() * 16;
      cudaStream_t st = static_cast<cudaStream_t>(stream_);
#define DYK_DW(KERN, KK, SS)                                                                                        \
  DYK_DISPATCH_ACT(act, DYK_DISPATCH_DTYPE(dtype, (KERN<kBf16, KK, SS, kAct><<<(unsigned)gs, 256, 0, st>>>(              \
                                (const uint8_t*)x, xs, w, scale, bias, (uint8_t*)y, ys, N, H, W, cv, pad, Ho, Wo, strips, act, \
                                (unsigned)tot))))
      if (k == 3 && stride == 1) DYK_DW(dwconv_strip_kernel, 3, 1);
      else if (k == 3) DYK_DW(dwconv_strip_kernel, 3, 2);
      else if (stride == 1) DYK_DW(dwconv_strip_kernel, 5, 1);
      else DYK_DW(dwconv_strip_kernel, 5, 2);
#undef DYK_DW
      DYK_LAUNCH_OK("dwconv strip kernel");
      return DYK_OK;
    }
  }
  const long long total = (long long)N * Ho * Wo * cv;
  long long g = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  DYK_DISPATCH_DTYPE(dtype, (dwconv_kernel<kBf16><<<(unsigned)g, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                (const uint8_t*)x, xs, w, scale, bias, (uint8_t*)y, ys, N, H, W, cv, k, stride, pad, Ho,
                                Wo, act)));
  DYK_LAUNCH_OK("dwconv_kernel");
  return DYK_OK;
}
