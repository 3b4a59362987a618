// HBM-bound NHWC kernels of the hot path: weighted fusion add, channel-slice copy, max-pool, nearest
// upsample, squeeze-excitation, layout conversion and weight packing.  All of them move 16-byte
// (8-channel) vectors with the channel index fastest so that a warp touches contiguous memory.
#include "common.h"
#include "vec.cuh"

namespace dyk {

static inline int grid_for(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------ WeightedFeatureFusion
// reference: build_utils/layers.py:63-85   x*w[0] + a*w[1]  (or plain x + a)
// gate != nullptr: operand a is a SqueezeExcitation input whose `scale * x` (layers.py:190) has not been materialised: it is
// formed here, rounded to the storage type like dyk_scale_channels would have stored it, and then enters the sum.
template <bool kBf16>
__global__ void fused_add_kernel(const uint8_t* __restrict__ a, long long as, const uint8_t* __restrict__ b,
                                 long long bs, uint8_t* __restrict__ y, long long ys, long long npix, int cv,
                                 const float* __restrict__ wts, const float* __restrict__ gate = nullptr,
                                 long long gate_stride = 0, int HW = 1) {
  float w0 = 1.f, w1 = 1.f;
  if (wts) { w0 = __ldg(wts); w1 = __ldg(wts + 1); }
  const long long total = npix * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int c = (int)(i - pix * cv);
    const uint4 va = __ldg(reinterpret_cast<const uint4*>(a + (pix * as) * 2) + c);
    const uint4 vb = __ldg(reinterpret_cast<const uint4*>(b + (pix * bs) * 2) + c);
    float fa[8], fb[8], fo[8];
    unpack8<kBf16>(va, fa);
    unpack8<kBf16>(vb, fb);
    if (gate) {
      const float* g = gate + (pix / HW) * gate_stride + c * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) fa[k] *= __ldg(g + k);
      unpack8<kBf16>(pack8<kBf16>(fa), fa);      // the rounding of the (never stored) gated tensor
    }
    if (wts) {
      // reference order: x = x * w0 (rounded to the tensor dtype), a = a * w1 (rounded), then x + a
#pragma unroll
      for (int k = 0; k < 8; ++k) fo[k] = fuse2(fa[k], fb[k], w0, w1);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) fo[k] = fa[k] + fb[k];
    }
    *(reinterpret_cast<uint4*>(y + (pix * ys) * 2) + c) = pack8<kBf16>(fo);
  }
}

__global__ void fusion_weights_kernel(const float* __restrict__ w_raw, float* __restrict__ w_out, int n) {
  const int i = threadIdx.x;
  if (i < n) w_out[i] = (1.f / (1.f + expf(-w_raw[i]))) * (2.f / n);
}

// ------------------------------------------------------------------ FeatureConcat fallback copy
__global__ void copy_slice_kernel(const uint8_t* __restrict__ src, long long ss, uint8_t* __restrict__ dst,
                                  long long ds, long long npix, int cv) {
  const long long total = npix * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int c = (int)(i - pix * cv);
    *(reinterpret_cast<uint4*>(dst + (pix * ds) * 2) + c) = __ldg(reinterpret_cast<const uint4*>(src + (pix * ss) * 2) + c);
  }
}

// ------------------------------------------------------------------ MaxPool2d(k, stride, (k-1)//2)
// reference: models.py:91-94; padding behaves as -inf (window clipped to the image).
template <bool kBf16>
__global__ void maxpool_kernel(const uint8_t* __restrict__ x, long long xs, uint8_t* __restrict__ y, long long ys,
                               int N, int H, int W, int cv, int k, int stride, int pad, int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long t = i / cv;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float m[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) m[q] = -INFINITY;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    for (int r = 0; r < k; ++r) {
      const int h = h0 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int w = w0 + s;
        if (w < 0 || w >= W) continue;
        const long long pix = ((long long)n * H + h) * W + w;
        float f[8];
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(x + pix * xs * 2) + c), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) m[q] = fmaxf(m[q], f[q]);
      }
    }
    const long long opix = ((long long)n * Ho + ho) * Wo + wo;
    *(reinterpret_cast<uint4*>(y + opix * ys * 2) + c) = pack8<kBf16>(m);
  }
}

// 5x5 / stride 1 (every SPP pool: k = 9 and 13 are cascades of this one): the kernel above issues 25 predicated 16-byte loads,
// 200 conversions and 200 fp32 max per 8-channel output with 64-bit index math (21 us for 21 MB in + 21 MB out at bs 16,
// a tenth of the HBM rate).  Here a thread owns one 8-channel vector and a strip of four outputs along W: the 5 x 8 input
// window is loaded once and slides, the loops are unrolled, and the maximum is taken on the packed 16-bit pairs (HMNMX2 —
// exact: the result is one of the inputs).
template <bool kBf16>
__device__ __forceinline__ uint32_t max2_packed(uint32_t a, uint32_t b) {
  if constexpr (kBf16) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}
template <bool kBf16>
__global__ void __launch_bounds__(256)
maxpool5_kernel(const uint8_t* __restrict__ x, long long xs, uint8_t* __restrict__ y, long long ys, int N, int H, int W, int cv,
                int strips_w, unsigned total) {
  constexpr uint32_t kNegInf = kBf16 ? 0xFF80FF80u : 0xFC00FC00u;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned c = i % (unsigned)cv;
    unsigned t = i / (unsigned)cv;
    const unsigned sw = t % (unsigned)strips_w; t /= (unsigned)strips_w;
    const int ho = (int)(t % (unsigned)H), n = (int)(t / (unsigned)H);
    const int wo0 = (int)sw * 4;
    uint4 m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = make_uint4(kNegInf, kNegInf, kNegInf, kNegInf);
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int h = ho - 2 + r;
      const bool hok = h >= 0 && h < H;
      const uint8_t* xrow = x + (((long long)n * H + (hok ? h : 0)) * W) * xs * 2;
#pragma unroll
      for (int col = 0; col < 8; ++col) {
        const int w = wo0 - 2 + col;
        uint4 v = make_uint4(kNegInf, kNegInf, kNegInf, kNegInf);
        if (hok && w >= 0 && w < W) v = __ldg(reinterpret_cast<const uint4*>(xrow + (long long)w * xs * 2) + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (col - j >= 0 && col - j < 5) {
            m[j].x = max2_packed<kBf16>(m[j].x, v.x); m[j].y = max2_packed<kBf16>(m[j].y, v.y);
            m[j].z = max2_packed<kBf16>(m[j].z, v.z); m[j].w = max2_packed<kBf16>(m[j].w, v.w);
          }
        }
      }
    }
    uint8_t* yp = y + ((((long long)n * H + ho) * W + wo0) * ys) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (wo0 + j < W) *(reinterpret_cast<uint4*>(yp + (long long)j * ys * 2) + c) = m[j];
  }
}

// ------------------------------------------------------------------ nn.Upsample(scale_factor=s), nearest
__global__ void upsample_kernel(const uint8_t* __restrict__ x, long long xs, uint8_t* __restrict__ y, long long ys,
                                int N, int H, int W, int cv, int s) {
  const int Ho = H * s, Wo = W * s;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long t = i / cv;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const long long pix = ((long long)n * H + ho / s) * W + wo / s;
    const long long opix = ((long long)n * Ho + ho) * Wo + wo;
    *(reinterpret_cast<uint4*>(y + opix * ys * 2) + c) = __ldg(reinterpret_cast<const uint4*>(x + pix * xs * 2) + c);
  }
}

// ------------------------------------------------------------------ SqueezeExcitation
// reference: build_utils/layers.py:184-190
// (1) pooled[n][slab][c] = sum over a slab of pixels; the slabs are summed in a fixed order by (2), so the result is
//     deterministic (no atomics)
template <bool kBf16>
__global__ void se_pool_kernel(const uint8_t* __restrict__ x, long long xs, int HW, int C, int slabs,
                               float* __restrict__ pooled) {
  // block: 32 pixel lanes x 8 channel-vectors (64 channels); grid: (C/64 ceil, slabs, N)
  const int cvec = blockIdx.x * 8 + (threadIdx.x & 7);
  const int plane = threadIdx.x >> 3;  // 0..31
  const int n = blockIdx.z;
  const int per = (HW + slabs - 1) / slabs;
  const int p0 = blockIdx.y * per;
  const int p1 = min(HW, p0 + per);
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  if (cvec * 8 < C) {
#pragma unroll 4          // four independent 16-byte loads in flight per thread (the sums stay in pixel order)
    for (int pidx = p0 + plane; pidx < p1; pidx += 32) {
      const long long pix = (long long)n * HW + pidx;
      float f[8];
      unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(x + pix * xs * 2) + cvec), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += f[q];
    }
  }
  __shared__ float red[32][8][8];
#pragma unroll
  for (int q = 0; q < 8; ++q) red[plane][threadIdx.x & 7][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int cv = threadIdx.x >> 3, q = threadIdx.x & 7;
    float s = 0.f;
    for (int l = 0; l < 32; ++l) s += red[l][cv][q];
    const int c = (blockIdx.x * 8 + cv) * 8 + q;
    if (c < C) pooled[((long long)n * slabs + blockIdx.y) * C + c] = s;
  }
}
// (2a) hid[n][j] = relu(W1[j] . mean[n] + b1[j]); one warp per hidden unit, grid (N, ceil(Csq / 8))
//      (one block per image was latency bound: 123 us for C = 1024 — 16 CTAs on 148 SMs)
__global__ void __launch_bounds__(256)
se_fc1_kernel(const float* __restrict__ pooled, int slabs, float inv_hw, int C, int Csq, const float* __restrict__ w1,
              const float* __restrict__ b1, float* __restrict__ hid) {
  extern __shared__ float mean[];   // [C]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int sl = 0; sl < slabs; ++sl) s += pooled[((long long)n * slabs + sl) * C + c];
    mean[c] = s * inv_hw;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.y * 8 + warp;
  if (j >= Csq) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += __ldg(&w1[(long long)j * C + c]) * mean[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) hid[(long long)n * Csq + j] = fmaxf(s + __ldg(&b1[j]), 0.f);
}
// (2b) gate[n][c] = hardsigmoid(W2[c] . hid[n] + b2[c]); one warp per channel, grid (N, ceil(C / 8))
__global__ void __launch_bounds__(256)
se_fc2_kernel(const float* __restrict__ hid, int C, int Csq, const float* __restrict__ w2, const float* __restrict__ b2,
              float* __restrict__ gate) {
  extern __shared__ float h[];      // [Csq]
  const int n = blockIdx.x;
  for (int j = threadIdx.x; j < Csq; j += blockDim.x) h[j] = hid[(long long)n * Csq + j];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * 8 + warp;
  if (c >= C) return;
  float s = 0.f;
  for (int j = lane; j < Csq; j += 32) s += __ldg(&w2[(long long)c * Csq + j]) * h[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const float v = s + __ldg(&b2[c]);
    gate[(long long)n * C + c] = fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
  }
}
// (3) y = x * gate[n][c]
template <bool kBf16>
__global__ void scale_channels_kernel(const uint8_t* __restrict__ x, long long xs, const float* __restrict__ gate,
                                      uint8_t* __restrict__ y, long long ys, int N, int HW, int C) {
  const int cv = C / 8;
  const long long total = (long long)N * HW * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long pix = i / cv;
    const int n = (int)(pix / HW);
    float f[8];
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(x + pix * xs * 2) + c), f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + (long long)n * C + c * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + (long long)n * C + c * 8) + 1);
    f[0] *= g0.x; f[1] *= g0.y; f[2] *= g0.z; f[3] *= g0.w;
    f[4] *= g1.x; f[5] *= g1.y; f[6] *= g1.z; f[7] *= g1.w;
    *(reinterpret_cast<uint4*>(y + pix * ys * 2) + c) = pack8<kBf16>(f);
  }
}

// ------------------------------------------------------------------ packing / layout
template <bool kBf16>
__global__ void pack_ohwi_kernel(const float* __restrict__ w, void* __restrict__ out, int O, int I, int kh, int kw) {
  const long long total = (long long)O * I * kh * kw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes the output [o][r][s][ci]
    const int ci = (int)(i % I);
    long long t = i / I;
    const int s = (int)(t % kw); t /= kw;
    const int r = (int)(t % kh);
    const int o = (int)(t / kh);
    store1<kBf16>(out, i, __ldg(&w[(((long long)o * I + ci) * kh + r) * kw + s]));
  }
}

// Eval-mode BatchNorm folded into per-channel (scale, bias) for every convolution of a model in ONE launch (one block per
// layer).  descs: device int64 [n][10] = (gamma or 0, beta or 0, running_mean, running_var or 0 when the layer has no BN,
// conv bias or 0, scale_out or 0, bias_out or 0, C, Cpad, float bits of eps).  Entries C..Cpad of the outputs are zeroed.
__global__ void __launch_bounds__(256) fold_bn_multi_kernel(const long long* __restrict__ descs) {
  const long long* d = descs + (long long)blockIdx.x * 10;
  const float* gamma = reinterpret_cast<const float*>(d[0]);
  const float* beta = reinterpret_cast<const float*>(d[1]);
  const float* mean = reinterpret_cast<const float*>(d[2]);
  const float* var = reinterpret_cast<const float*>(d[3]);
  const float* cbias = reinterpret_cast<const float*>(d[4]);
  float* scale_out = reinterpret_cast<float*>(d[5]);
  float* bias_out = reinterpret_cast<float*>(d[6]);
  const int C = (int)d[7], Cpad = (int)d[8];
  const float eps = __int_as_float((int)d[9]);
  for (int c = threadIdx.x; c < Cpad; c += blockDim.x) {
    float sc = 0.f, bi = 0.f;
    if (c < C) {
      if (var != nullptr) {
        sc = (gamma ? gamma[c] : 1.f) * __frsqrt_rn(var[c] + eps);
        bi = (beta ? beta[c] : 0.f) - mean[c] * sc;
      } else {
        sc = 1.f;
        bi = cbias ? cbias[c] : 0.f;
      }
    }
    if (scale_out) scale_out[c] = sc;
    if (bias_out) bias_out[c] = bi;
  }
}

// SqueezeExcitation gate folded into the weights of the consuming 1x1 convolution: out[n][co][ci] = w[co][ci] * gate[n][ci]
template <bool kBf16>
__global__ void scale_weights_per_image_kernel(const uint8_t* __restrict__ w, const float* __restrict__ gate, long long gs,
                                               uint8_t* __restrict__ out, int N, int Cout, int cv) {
  const long long per = (long long)Cout * cv;
  const long long total = per * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / per, r = i - n * per;
    const int c = (int)(r % cv);
    float f[8];
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(w) + r), f);
    const float* g = gate + n * gs + c * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= __ldg(g + k);
    reinterpret_cast<uint4*>(out)[i] = pack8<kBf16>(f);
  }
}

// [N][C][HW] fp32 -> [N][HW][ys] dtype through a 32x32 smem transpose
template <bool kBf16>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, void* __restrict__ y, long long ys, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? __ldg(&x[((long long)n * C + c) * HW + p]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (c < C && p < HW) store1<kBf16>(y, ((long long)n * HW + p) * ys + c, tile[threadIdx.x][j]);
  }
}
template <bool kBf16>
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ x, long long xs, float* __restrict__ y, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? load1<kBf16>(x, ((long long)n * HW + p) * xs + c) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (c < C && p < HW) y[((long long)n * C + c) * HW + p] = tile[threadIdx.x][j];
  }
}

}  // namespace dyk

using namespace dyk;

#define DYK_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" __attribute__((visibility("default"))) int dyk_fused_add(const void* a, int64_t as, const void* b, int64_t bs, void* y, int64_t ys,
                             int64_t npix, int32_t C, const float* wts, int32_t dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(a && b && y, "dyk_fused_add: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && as % 8 == 0 && bs % 8 == 0 && ys % 8 == 0,
              "dyk_fused_add: C and strides must be multiples of 8 (C=%d)", C);
  DYK_REQUIRE(DYK_ALIGNED16(a) && DYK_ALIGNED16(b) && DYK_ALIGNED16(y), "dyk_fused_add: 16-byte alignment");
  if (npix == 0) return DYK_OK;
  const int cv = C / 8;
  const int grid = grid_for(npix * cv, 256);
  DYK_DISPATCH_DTYPE(dtype, (fused_add_kernel<kBf16><<<grid, 256, 0, stream>>>(
                                (const uint8_t*)a, as, (const uint8_t*)b, bs, (uint8_t*)y, ys, npix, cv, wts)));
  DYK_LAUNCH_OK("fused_add_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_fused_add_gated(const void* a, int64_t as, const float* gate, int64_t gate_stride,
                             int32_t HW, const void* b, int64_t bs, void* y, int64_t ys, int64_t npix, int32_t C,
                             const float* wts, int32_t dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(a && b && y && gate && HW > 0, "dyk_fused_add_gated: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && as % 8 == 0 && bs % 8 == 0 && ys % 8 == 0 && gate_stride >= C,
              "dyk_fused_add_gated: C and strides must be multiples of 8 (C=%d)", C);
  DYK_REQUIRE(DYK_ALIGNED16(a) && DYK_ALIGNED16(b) && DYK_ALIGNED16(y), "dyk_fused_add_gated: 16-byte alignment");
  if (npix == 0) return DYK_OK;
  const int cv = C / 8;
  const int grid = grid_for(npix * cv, 256);
  DYK_DISPATCH_DTYPE(dtype, (fused_add_kernel<kBf16><<<grid, 256, 0, stream>>>(
                                (const uint8_t*)a, as, (const uint8_t*)b, bs, (uint8_t*)y, ys, npix, cv, wts, gate, gate_stride, HW)));
  DYK_LAUNCH_OK("fused_add_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_scale_weights_per_image(const void* w_packed, const float* gate, int64_t gate_stride,
                             void* out, int32_t N, int32_t Cout, int32_t Cin, int32_t dtype, void* stream_) {
  DYK_REQUIRE(w_packed && gate && out && N > 0 && Cout > 0, "dyk_scale_weights_per_image: bad arguments");
  DYK_REQUIRE(Cin > 0 && Cin % 8 == 0 && gate_stride >= Cin && DYK_ALIGNED16(w_packed) && DYK_ALIGNED16(out),
              "dyk_scale_weights_per_image: Cin must be a multiple of 8, 16-byte aligned tensors");
  const int cv = Cin / 8;
  const int grid = grid_for((long long)N * Cout * cv, 256);
  DYK_DISPATCH_DTYPE(dtype, (scale_weights_per_image_kernel<kBf16><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                (const uint8_t*)w_packed, gate, gate_stride, (uint8_t*)out, N, Cout, cv)));
  DYK_LAUNCH_OK("scale_weights_per_image_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_fusion_weights(const float* w_raw, float* w_out, int32_t n, void* stream_) {
  DYK_REQUIRE(w_raw && w_out && n > 0 && n <= 32, "dyk_fusion_weights: bad arguments");
  fusion_weights_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(w_raw, w_out, n);
  DYK_LAUNCH_OK("fusion_weights_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_copy_slice(const void* src, int64_t ss, void* dst, int64_t ds, int64_t npix, int32_t C,
                              int32_t dtype, void* stream_) {
  (void)dtype;
  DYK_REQUIRE(src && dst, "dyk_copy_slice: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && ss % 8 == 0 && ds % 8 == 0, "dyk_copy_slice: C/strides must be multiples of 8");
  DYK_REQUIRE(DYK_ALIGNED16(src) && DYK_ALIGNED16(dst), "dyk_copy_slice: 16-byte alignment");
  if (npix == 0) return DYK_OK;
  const int cv = C / 8;
  copy_slice_kernel<<<grid_for(npix * cv, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      (const uint8_t*)src, ss, (uint8_t*)dst, ds, npix, cv);
  DYK_LAUNCH_OK("copy_slice_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_maxpool2d(const void* x, int64_t xs, void* y, int64_t ys, int32_t N, int32_t H, int32_t W,
                             int32_t C, int32_t k, int32_t stride, int32_t dtype, void* stream_) {
  DYK_REQUIRE(x && y, "dyk_maxpool2d: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && ys % 8 == 0, "dyk_maxpool2d: C/strides must be multiples of 8");
  DYK_REQUIRE(k >= 1 && stride >= 1, "dyk_maxpool2d: k=%d stride=%d", k, stride);
  const int pad = (k - 1) / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_maxpool2d: empty output");
  const int cv = C / 8;
  if (k == 5 && stride == 1 && (dtype == DYK_F16 || dtype == DYK_BF16)) {          // SPP pools (models.py:91-94 with size=5, stride=1)
    const int strips = (W + 3) / 4;
    const long long tot = (long long)N * H * strips * cv;
    if (tot < (1ll << 31)) {
      DYK_DISPATCH_DTYPE(dtype, (maxpool5_kernel<kBf16><<<grid_for(tot, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                    (const uint8_t*)x, xs, (uint8_t*)y, ys, N, H, W, cv, strips, (unsigned)tot)));
      DYK_LAUNCH_OK("maxpool5_kernel");
      return DYK_OK;
    }
  }
  const int grid = grid_for((long long)N * Ho * Wo * cv, 256);
  DYK_DISPATCH_DTYPE(dtype, (maxpool_kernel<kBf16><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                (const uint8_t*)x, xs, (uint8_t*)y, ys, N, H, W, cv, k, stride, pad, Ho, Wo)));
  DYK_LAUNCH_OK("maxpool_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_upsample_nearest(const void* x, int64_t xs, void* y, int64_t ys, int32_t N, int32_t H,
                                    int32_t W, int32_t C, int32_t s, int32_t dtype, void* stream_) {
  (void)dtype;
  DYK_REQUIRE(x && y, "dyk_upsample_nearest: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && ys % 8 == 0 && s >= 1, "dyk_upsample_nearest: bad shape");
  const int cv = C / 8;
  upsample_kernel<<<grid_for((long long)N * H * s * W * s * cv, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      (const uint8_t*)x, xs, (uint8_t*)y, ys, N, H, W, cv, s);
  DYK_LAUNCH_OK("upsample_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_se_gate(const void* x, int64_t xs, int32_t N, int32_t HW, int32_t C, const float* w1,
                           const float* b1, const float* w2, const float* b2, int32_t Csq, float* pooled,
                           float* gate, int32_t dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x && w1 && b1 && w2 && b2 && pooled && gate, "dyk_se_gate: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && Csq > 0 && N > 0 && HW > 0, "dyk_se_gate: bad shape");
  DYK_REQUIRE((size_t)(C + Csq) * 4 <= 48 * 1024, "dyk_se_gate: C + Csq too large");
  int slabs = HW / 256;
  if (slabs < 1) slabs = 1;
  if (slabs > 31) slabs = 31;   // one slab's worth of the scratch holds the hidden activations
  const dim3 grid((C + 63) / 64, slabs, N);
  DYK_DISPATCH_DTYPE(dtype, (se_pool_kernel<kBf16><<<grid, 256, 0, stream>>>((const uint8_t*)x, xs, HW, C, slabs, pooled)));
  DYK_LAUNCH_OK("se_pool_kernel");
  return dyk_se_mlp(pooled, slabs, N, HW, C, w1, b1, w2, b2, Csq, gate, stream_);
}

extern "C" __attribute__((visibility("default"))) int dyk_se_mlp(float* pooled, int32_t slabs, int32_t N, int32_t HW, int32_t C, const float* w1,
                          const float* b1, const float* w2, const float* b2, int32_t Csq, float* gate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(pooled && w1 && b1 && w2 && b2 && gate, "dyk_se_mlp: null pointer");
  DYK_REQUIRE((size_t)(C + Csq) * 4 <= 48 * 1024, "dyk_se_mlp: C + Csq too large");
  // hidden activations live behind the per-slab partial sums in the caller's scratch (N * 32 * C floats >= N * (slabs * C + Csq))
  DYK_REQUIRE(Csq > 0 && Csq <= C && slabs >= 1 && slabs <= 31, "dyk_se_mlp: Csq=%d must not exceed C=%d (slabs=%d)", Csq, C, slabs);
  float* hid = pooled + (size_t)N * slabs * C;
  se_fc1_kernel<<<dim3(N, (Csq + 7) / 8), 256, C * sizeof(float), stream>>>(pooled, slabs, 1.f / (float)HW, C, Csq, w1, b1, hid);
  DYK_LAUNCH_OK("se_fc1_kernel");
  se_fc2_kernel<<<dim3(N, (C + 7) / 8), 256, Csq * sizeof(float), stream>>>(hid, C, Csq, w2, b2, gate);
  DYK_LAUNCH_OK("se_fc2_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_scale_channels(const void* x, int64_t xs, const float* gate, void* y, int64_t ys, int32_t N,
                                  int32_t HW, int32_t C, int32_t dtype, void* stream_) {
  DYK_REQUIRE(x && gate && y, "dyk_scale_channels: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && ys % 8 == 0, "dyk_scale_channels: bad shape");
  const int grid = grid_for((long long)N * HW * (C / 8), 256);
  DYK_DISPATCH_DTYPE(dtype, (scale_channels_kernel<kBf16><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                (const uint8_t*)x, xs, gate, (uint8_t*)y, ys, N, HW, C)));
  DYK_LAUNCH_OK("scale_channels_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_pack_weights_ohwi(const float* w, void* out, int32_t O, int32_t I, int32_t kh, int32_t kw,
                                     int32_t dtype, void* stream_) {
  DYK_REQUIRE(w && out && O > 0 && I > 0 && kh > 0 && kw > 0, "dyk_pack_weights_ohwi: bad arguments");
  const int grid = grid_for((long long)O * I * kh * kw, 256);
  DYK_DISPATCH_DTYPE(dtype, (pack_ohwi_kernel<kBf16><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(w, out, O, I, kh, kw)));
  DYK_LAUNCH_OK("pack_ohwi_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_nchw_f32_to_nhwc(const float* x, void* y, int64_t ys, int32_t N, int32_t C, int32_t H, int32_t W,
                                    int32_t dtype, void* stream_) {
  DYK_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0 && ys >= C, "dyk_nchw_f32_to_nhwc: bad arguments");
  const int HW = H * W;
  const dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  DYK_DISPATCH_DTYPE(dtype, (nchw_to_nhwc_kernel<kBf16><<<grid, block, 0, static_cast<cudaStream_t>(stream_)>>>(x, y, ys, C, HW)));
  DYK_LAUNCH_OK("nchw_to_nhwc_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_nhwc_to_nchw_f32(const void* x, int64_t xs, float* y, int32_t N, int32_t C, int32_t H, int32_t W,
                                    int32_t dtype, void* stream_) {
  DYK_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0 && xs >= C, "dyk_nhwc_to_nchw_f32: bad arguments");
  const int HW = H * W;
  const dim3 grid((HW + 31) / 32, (C + 31) / 32, N), block(32, 8);
  DYK_DISPATCH_DTYPE(dtype, (nhwc_to_nchw_kernel<kBf16><<<grid, block, 0, static_cast<cudaStream_t>(stream_)>>>(x, xs, y, C, HW)));
  DYK_LAUNCH_OK("nhwc_to_nchw_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_fold_bn_multi(const int64_t* descs, int32_t n, void* stream_) {
  DYK_REQUIRE(descs && n > 0, "dyk_fold_bn_multi: bad arguments");
  fold_bn_multi_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const long long*>(descs));
  DYK_LAUNCH_OK("fold_bn_multi_kernel");
  return DYK_OK;
}
