// Stem convolution, 3x3 / Cin = 3 / pad 1, 16 or 32 output channels, stride 1 or 2 (models.py:35-36: the MobileNet stems),
// reading the caller's NCHW fp32 / uint8 frames like the generic stem kernel in conv_direct.cu.
#include "common.h"
#include <cstdlib>
#include "act.cuh"
#include "vec.cuh"

namespace dyk {

__device__ __forceinline__ float2 stem_ffma2(float2 x, float2 y, float2 z) {   // two IEEE fp32 FMAs per instruction (FFMA2)
  uint64_t ux = *reinterpret_cast<uint64_t*>(&x), uy = *reinterpret_cast<uint64_t*>(&y), uz = *reinterpret_cast<uint64_t*>(&z), ud;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ux), "l"(uy), "l"(uz));
  return *reinterpret_cast<float2*>(&ud);
}

// ------------------------------------------------------------------ stem, 3x3 / Cin = 3 / pad 1 specialisation
// The generic stem kernel (conv_direct.cu) runs the MobileNet stems (3 -> 16 / 32, stride 2; stem_tc covers only 3 -> 32 stride 1) at a
// tenth of their HBM rate (ncu, MobileNetV3-dual bs 64: 468 us per modality for 231 MB = 11.8 % of the step): run-time
// loops over k and Cin, 64-bit index math and an IEEE division per input byte, one scalar FMA per weight read.  Here
// everything is unrolled for k = 3, Cin = 3; a thread owns two W-adjacent output pixels (their windows share columns, the
// weight reads are shared), the byte -> v / 255 conversion is a 256-entry table of the same IEEE quotients, the weights
// are read as float4 broadcasts and multiplied with FFMA2.  Accumulation order (r, s, ci) and arithmetic per element are
// those of the generic kernel; taps outside the frame contribute fmaf(0, w, acc) = acc: bit-identical results.
template <int COUT_T, int S, bool kBf16, typename TIn>
__global__ void __launch_bounds__(128)
stem3x3_kernel(const TIn* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
               const float* __restrict__ bias, uint8_t* __restrict__ y, long long ys, int N, int H, int W, int Ho, int Wo, int act, int words) {
  __shared__ __align__(16) float wsm[27 * COUT_T];   // [(r*3+s)*3+ci][co]
  __shared__ __align__(16) float ssm[2 * COUT_T];
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 27 * COUT_T; i += blockDim.x) {
    const int t = i / COUT_T, co = i - t * COUT_T;
    wsm[i] = __ldg(&w[co * 27 + t]);
  }
  for (int i = threadIdx.x; i < COUT_T; i += blockDim.x) {
    ssm[i] = scale ? __ldg(&scale[i]) : 1.f;
    ssm[COUT_T + i] = bias ? __ldg(&bias[i]) : 0.f;
  }
  if constexpr (sizeof(TIn) == 1)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = __fdiv_rn((float)i, 255.f);
  __syncthreads();
  constexpr int kCols = S + 3;                      // input columns under two adjacent outputs
  const int pairs_w = (Wo + 1) >> 1;
  const unsigned total = (unsigned)N * Ho * pairs_w;
  const int plane = H * W;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int pw = (int)(i % (unsigned)pairs_w);
    const unsigned t2 = i / (unsigned)pairs_w;
    const int ho = (int)(t2 % (unsigned)Ho), n = (int)(t2 / (unsigned)Ho);
    const int wo = pw * 2, h0 = ho * S - 1, w0 = wo * S - 1;
    const TIn* xn = x + (long long)n * 3 * plane;
    float2 acc[2][COUT_T / 2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < COUT_T / 2; ++c) acc[j][c] = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = h0 + r;
      const bool hok = h >= 0 && h < H;
      float v[kCols][3];
      bool loaded = false;
      if constexpr (sizeof(TIn) == 1 && S == 2) {
        // uint8 frames, stride 2, W % 4 == 0 and a 4-byte aligned frame pointer: the five columns under the two outputs are
        // the last byte of one aligned word and the four bytes of the next — two loads per (row, plane) instead of five
        // (ncu on the byte-load version: long-scoreboard 2.0 of 7.4 stall cycles per issued instruction)
        if (words) {
          loaded = true;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            uint32_t word = 0;
            float left = 0.f;
            if (hok) {
              const uint8_t* rowp = reinterpret_cast<const uint8_t*>(xn) + ci * plane + h * W + 4 * pw;
              word = __ldg(reinterpret_cast<const uint32_t*>(rowp));
              if (pw > 0) left = lut[__ldg(rowp - 1)];
            }
            v[0][ci] = left;
#pragma unroll
            for (int b = 0; b < 4; ++b) v[1 + b][ci] = hok ? lut[(word >> (8 * b)) & 0xffu] : 0.f;
          }
        }
      }
      if (!loaded)
#pragma unroll
      for (int col = 0; col < kCols; ++col) {
        const int ww = w0 + col;
        const bool ok = hok && ww >= 0 && ww < W;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          float f = 0.f;
          if (ok) {
            const TIn raw = __ldg(xn + ci * plane + h * W + ww);
            if constexpr (sizeof(TIn) == 1) f = lut[raw];
            else f = raw;
          }
          v[col][ci] = f;
        }
      }
#pragma unroll
      for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float4* wp = reinterpret_cast<const float4*>(wsm + ((r * 3 + q) * 3 + ci) * COUT_T);
          const float a0 = v[q][ci], a1 = v[q + S][ci];
#pragma unroll
          for (int c4 = 0; c4 < COUT_T / 4; ++c4) {
            const float4 w4 = wp[c4];
            acc[0][2 * c4] = stem_ffma2(make_float2(a0, a0), make_float2(w4.x, w4.y), acc[0][2 * c4]);
            acc[0][2 * c4 + 1] = stem_ffma2(make_float2(a0, a0), make_float2(w4.z, w4.w), acc[0][2 * c4 + 1]);
            acc[1][2 * c4] = stem_ffma2(make_float2(a1, a1), make_float2(w4.x, w4.y), acc[1][2 * c4]);
            acc[1][2 * c4 + 1] = stem_ffma2(make_float2(a1, a1), make_float2(w4.z, w4.w), acc[1][2 * c4 + 1]);
          }
        }
    }
    const long long pix = ((long long)n * Ho + ho) * Wo + wo;
    // the activation is a run-time value, switched once per pixel pair (not per element), so that one instantiation serves all
    // seven activations (this kernel is FMA-bound: 432+ FMAs per pixel against 16 - 32 activations)
    DYK_DISPATCH_ACT(act, {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (wo + j >= Wo) break;
        uint8_t* yp = y + (pix + j) * ys * 2;
#pragma unroll
        for (int c8 = 0; c8 < COUT_T / 8; ++c8) {
          float o[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[2 * q] = act_apply<kAct>(fmaf(acc[j][c8 * 4 + q].x, ssm[c8 * 8 + 2 * q], ssm[COUT_T + c8 * 8 + 2 * q]));
            o[2 * q + 1] = act_apply<kAct>(fmaf(acc[j][c8 * 4 + q].y, ssm[c8 * 8 + 2 * q + 1], ssm[COUT_T + c8 * 8 + 2 * q + 1]));
          }
          *(reinterpret_cast<uint4*>(yp) + c8) = pack8<kBf16>(o);
        }
      }
    });
  }
}

// returns DYK_OK after a launch, 1 when the shape is not handled here (the caller keeps its generic kernel), < 0 on error
int stem3x3_try(const void* x, const float* w, const float* scale, const float* bias, void* y, int64_t ys, int N, int H, int W,
                int Cin, int Cout, int k, int stride, int pad, int act, int dtype, int x_kind, cudaStream_t stream) {
  const char* env = getenv("DYK_STEM_FAST");          // read per call: the parity test switches kernels inside one process
  if (env != nullptr && env[0] == '0') return 1;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (!(k == 3 && Cin == 3 && pad == 1 && (stride == 1 || stride == 2) && (Cout == 16 || Cout == 32) &&
        (dtype == DYK_F16 || dtype == DYK_BF16) && (long long)N * Ho * Wo < (1ll << 31)))
    return 1;
  const int words = (x_kind == 1 && stride == 2 && W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0) ? 1 : 0;
  const long long units = (long long)N * Ho * ((Wo + 1) / 2);
  long long g = (units + 127) / 128;
  const long long gcap = (long long)num_sms() * 16;
  if (g > gcap) g = gcap;
#define DYK_STEM3(CT, SS, TIN)                                                                                      \
  DYK_DISPATCH_DTYPE(dtype, (stem3x3_kernel<CT, SS, kBf16, TIN><<<(unsigned)g, 128, 0, stream>>>(                    \
                                static_cast<const TIN*>(x), w, scale, bias, (uint8_t*)y, ys, N, H, W, Ho, Wo, act, words)))
#define DYK_STEM3_S(CT, TIN) do { if (stride == 1) DYK_STEM3(CT, 1, TIN); else DYK_STEM3(CT, 2, TIN); } while (0)
  if (Cout == 16) { if (x_kind == 0) DYK_STEM3_S(16, float); else DYK_STEM3_S(16, uint8_t); }
  else { if (x_kind == 0) DYK_STEM3_S(32, float); else DYK_STEM3_S(32, uint8_t); }
#undef DYK_STEM3_S
#undef DYK_STEM3
  DYK_LAUNCH_OK("stem3x3_kernel");
  return DYK_OK;
}

}  // namespace dyk
