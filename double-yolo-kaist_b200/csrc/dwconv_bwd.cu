// Backward of the depthwise convolutions: data gradient and weight gradient of
// nn.Conv2d(groups = C) as the reference builds it for the MobileNet backbones (models.py:41, `groups` key) and inside
// DepthwiseSeparableConv2d (build_utils/layers.py:224).  PyTorch autograd does both for the reference.
//
// Both are HBM / L1-bound CUDA-core kernels on NHWC 16-bit tensors with 8-channel (16-byte) vectors, like the forward
// kernel in conv_direct.cu; the filter is fp32 [k][k][C] (the forward kernel's layout).
//  * dgrad: gather form — one thread per input pixel and channel vector sums the taps whose output pixel exists
//           (stride 2: only taps of matching parity), so there are no atomics and dx can be overwritten or accumulated.
//  * wgrad: dW[c][r][s] = sum over output pixels dz[p][c] * x[p*stride + (r, s) - pad][c]: two-stage reduction with a
//           fixed slab order (bit-reproducible).  grid = (channel groups of 64, slabs, k filter rows); a thread keeps
//           k x 8 accumulators (one filter row of one channel vector).
#include "common.h"
#include "vec.cuh"

namespace dyk {

template <bool kBf16>
__global__ void __launch_bounds__(256)
dwconv_dgrad_kernel(const uint8_t* __restrict__ dz, long long dzs, const float* __restrict__ w, uint8_t* __restrict__ dx,
                    long long dxs, int N, int H, int W, int cv, int k, int stride, int pad, int Ho, int Wo, int accumulate,
                    unsigned total) {
  const int C = cv * 8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % (unsigned)cv;
    unsigned t = i / (unsigned)cv;
    const int wi = t % (unsigned)W; t /= (unsigned)W;
    const int hi = t % (unsigned)H;
    const int n = t / (unsigned)H;
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int r = 0; r < k; ++r) {
      const int hh = hi + pad - r;
      if (hh < 0 || hh % stride) continue;
      const int ho = hh / stride;
      if (ho >= Ho) continue;
      for (int s = 0; s < k; ++s) {
        const int ww = wi + pad - s;
        if (ww < 0 || ww % stride) continue;
        const int wo = ww / stride;
        if (wo >= Wo) continue;
        float g[8];
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(dz + (((long long)n * Ho + ho) * Wo + wo) * dzs * 2) + c), g);
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w + (long long)(r * k + s) * C + c * 8));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(w + (long long)(r * k + s) * C + c * 8) + 1);
        acc[0] = fmaf(g[0], wa.x, acc[0]); acc[1] = fmaf(g[1], wa.y, acc[1]);
        acc[2] = fmaf(g[2], wa.z, acc[2]); acc[3] = fmaf(g[3], wa.w, acc[3]);
        acc[4] = fmaf(g[4], wb.x, acc[4]); acc[5] = fmaf(g[5], wb.y, acc[5]);
        acc[6] = fmaf(g[6], wb.z, acc[6]); acc[7] = fmaf(g[7], wb.w, acc[7]);
      }
    }
    uint4* out = reinterpret_cast<uint4*>(dx + (((long long)n * H + hi) * W + wi) * dxs * 2) + c;
    if (accumulate) {
      float old[8];
      unpack8<kBf16>(*out, old);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += old[q];
    }
    *out = pack8<kBf16>(acc);
  }
}

// part[slab][r*K + s][c]
template <bool kBf16, int K>
__global__ void __launch_bounds__(256)
dwconv_wgrad_kernel(const uint8_t* __restrict__ x, long long xs, const uint8_t* __restrict__ dz, long long dzs, int N, int H,
                    int W, int C, int stride, int pad, int Ho, int Wo, int slabs, float* __restrict__ part) {
  const int cl = threadIdx.x & 7, plane = threadIdx.x >> 3;
  const int cvec = blockIdx.x * 8 + cl;
  const int r = blockIdx.z;
  const unsigned npix = (unsigned)N * Ho * Wo;
  const unsigned per = (npix + slabs - 1) / slabs;
  const unsigned p0 = blockIdx.y * per;
  const unsigned p1 = p0 + per < npix ? p0 + per : npix;
  float acc[K][8];
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[s][q] = 0.f;
  if (cvec * 8 < C) {
    for (unsigned p = p0 + plane; p < p1; p += 32) {
      const int wo = p % (unsigned)Wo;
      const unsigned t = p / (unsigned)Wo;
      const int ho = t % (unsigned)Ho;
      const int n = t / (unsigned)Ho;
      const int hh = ho * stride + r - pad;
      if (hh < 0 || hh >= H) continue;
      float g[8];
      unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(dz + (long long)p * dzs * 2) + cvec), g);
      const uint8_t* xrow = x + (((long long)n * H + hh) * W) * xs * 2;
#pragma unroll
      for (int s = 0; s < K; ++s) {
        const int ww = wo * stride + s - pad;
        if (ww < 0 || ww >= W) continue;
        float f[8];
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(xrow + (long long)ww * xs * 2) + cvec), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[s][q] = fmaf(g[q], f[q], acc[s][q]);
      }
    }
  }
  __shared__ float red[32][8][K * 8 + 1];
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int q = 0; q < 8; ++q) red[plane][cl][s * 8 + q] = acc[s][q];
  __syncthreads();
  for (int j = threadIdx.x; j < 8 * K * 8; j += 256) {
    const int q = j & 7, cvl = (j >> 3) & 7, s = j >> 6;     // channel fastest: coalesced partial writes
    float sum = 0.f;
    for (int l = 0; l < 32; ++l) sum += red[l][cvl][s * 8 + q];
    const int c = (blockIdx.x * 8 + cvl) * 8 + q;
    if (c < C) part[((long long)blockIdx.y * K * K + r * K + s) * C + c] = sum;
  }
}

// grad_w (OIHW with I = 1: [C][k][k]) (+)= sum over slabs, fixed order
__global__ void dwconv_wgrad_finalize_kernel(const float* __restrict__ part, int slabs, int C, int KK, float* __restrict__ gw,
                                             int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * KK) return;
  const int c = i % C, t = i / C;
  float s = 0.f;
  for (int l = 0; l < slabs; ++l) s += part[((long long)l * KK + t) * C + c];
  float* o = gw + (long long)c * KK + t;
  *o = accumulate ? *o + s : s;
}

}  // namespace dyk

using namespace dyk;
#define DYK_EXPORT extern "C" __attribute__((visibility("default")))

DYK_EXPORT int dyk_dwconv2d_dgrad(const void* dz, int64_t dzs, const float* w, void* dx, int64_t dxs, int32_t N, int32_t H,
                                  int32_t W, int32_t C, int32_t k, int32_t stride, int32_t pad, int32_t accumulate,
                                  int32_t dtype, void* stream_) {
  DYK_REQUIRE(dz && w && dx, "dyk_dwconv2d_dgrad: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && dzs % 8 == 0 && dxs % 8 == 0 && k >= 1 && stride >= 1 && pad >= 0,
              "dyk_dwconv2d_dgrad: bad shape");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_dwconv2d_dgrad: empty output");
  const int cv = C / 8;
  const long long total = (long long)N * H * W * cv;
  DYK_REQUIRE(total < (1ll << 31), "dyk_dwconv2d_dgrad: tensor too large for 32-bit indexing");
  long long g = (total + 255) / 256;
  if (g > (long long)num_sms() * 16) g = (long long)num_sms() * 16;
  DYK_DISPATCH_DTYPE(dtype, (dwconv_dgrad_kernel<kBf16><<<(unsigned)g, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                (const uint8_t*)dz, dzs, w, (uint8_t*)dx, dxs, N, H, W, cv, k, stride, pad, Ho, Wo, accumulate,
                                (unsigned)total)));
  DYK_LAUNCH_OK("dwconv_dgrad_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_dwconv2d_wgrad(const void* x, int64_t xs, const void* dz, int64_t dzs, float* grad_w, int32_t N, int32_t H,
                                  int32_t W, int32_t C, int32_t k, int32_t stride, int32_t pad, int32_t accumulate,
                                  int32_t dtype, float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x && dz && grad_w && workspace, "dyk_dwconv2d_wgrad: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && dzs % 8 == 0 && stride >= 1 && pad >= 0, "dyk_dwconv2d_wgrad: bad shape");
  DYK_REQUIRE(k == 3 || k == 5, "dyk_dwconv2d_wgrad: only 3x3 and 5x5 depthwise filters are on the path");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_dwconv2d_wgrad: empty output");
  const long long npix = (long long)N * Ho * Wo;
  DYK_REQUIRE(npix < (1ll << 31), "dyk_dwconv2d_wgrad: tensor too large for 32-bit indexing");
  long long slabs = npix / 256;
  if (slabs < 1) slabs = 1;
  if (slabs > DYK_DW_WGRAD_SLABS) slabs = DYK_DW_WGRAD_SLABS;
  const dim3 grid((C + 63) / 64, (unsigned)slabs, k);
  if (k == 3) {
    DYK_DISPATCH_DTYPE(dtype, (dwconv_wgrad_kernel<kBf16, 3><<<grid, 256, 0, stream>>>(
                                  (const uint8_t*)x, xs, (const uint8_t*)dz, dzs, N, H, W, C, stride, pad, Ho, Wo, (int)slabs,
                                  workspace)));
  } else {
    DYK_DISPATCH_DTYPE(dtype, (dwconv_wgrad_kernel<kBf16, 5><<<grid, 256, 0, stream>>>(
                                  (const uint8_t*)x, xs, (const uint8_t*)dz, dzs, N, H, W, C, stride, pad, Ho, Wo, (int)slabs,
                                  workspace)));
  }
  DYK_LAUNCH_OK("dwconv_wgrad_kernel");
  const int tot = C * k * k;
  dwconv_wgrad_finalize_kernel<<<(tot + 255) / 256, 256, 0, stream>>>(workspace, (int)slabs, C, k * k, grad_w, accumulate);
  DYK_LAUNCH_OK("dwconv_wgrad_finalize_kernel");
  return DYK_OK;
}
