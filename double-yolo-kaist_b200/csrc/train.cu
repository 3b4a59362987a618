// Training-side HBM-bound kernels: train-mode BatchNorm (batch statistics, running-stat update), BN + activation
// backward, and the backward of the element-wise / pooling / SE / head-permute ops of the Darknet graph.
//
// Replaces what PyTorch autograd does for the reference's modules: nn.BatchNorm2d in
// training mode + activation (models.py:47-64), WeightedFeatureFusion (layers.py:63-85), FeatureConcat (:32-44),
// nn.MaxPool2d (models.py:91-94), nn.Upsample (:100-101), SqueezeExcitation (layers.py:175-190) and the permute of
// YOLOLayer.forward (models.py:229).  Convolution gradients are in conv_wgrad.cu (tcgen05) and, for dgrad, the
// forward kernels run on flipped weights.
//
// All per-channel reductions are two-stage with a fixed slab order (no floating-point atomics), so training runs are
// bit-reproducible.  Tensors are NHWC 16-bit with (base, pixel stride) addressing like the forward kernels.
#include "common.h"
#include <cstdlib>
#include "vec.cuh"
#include "act.cuh"
#include "frame_sample.cuh"

namespace dyk {

constexpr int kMaxSlabs = DYK_TRAIN_MAX_SLABS;
constexpr int kFusedFinalizeMaxSlabs = 256;

static inline int grid_for_t(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
// enough slabs that (C / 64) x slabs blocks fill the GPU a few times over, capped by the workspace contract
static inline int slabs_for(long long npix) {
  static const int slab_pix = getenv("DYK_SLAB_PIX") ? atoi(getenv("DYK_SLAB_PIX")) : 256;
  long long s = npix / slab_pix;
  if (s < 1) s = 1;
  if (s > kMaxSlabs) s = kMaxSlabs;
  return (int)s;
}

// d act(x) / dx as PyTorch's backward formulas define it (incl. the values chosen at the kinks)
__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case DYK_ACT_LEAKY: return x > 0.f ? 1.f : 0.1f;
    case DYK_ACT_MISH: {
      // mish'(x) = t + x*sigmoid(x)*(1 - t^2), t = tanh(softplus(x)) = n/(n+2), n = e^x (e^x + 2).  With n + 1 = (e^x + 1)^2
      // this is (n (n+2) + 4 x e^x (e^x + 1)) / (n+2)^2: one exponential and one reciprocal per element.
      const float e = ex2_approx(fminf(x, 20.f) * 1.4426950408889634f);    // bare MUFU ops, see mish_f (act.cuh)
      const float n = e * (e + 2.f);
      const float r = rcp_approx(n + 2.f);
      return x > 20.f ? 1.f : fmaf(n, n + 2.f, 4.f * x * e * (e + 1.f)) * r * r;
    }
    case DYK_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case DYK_ACT_RELU6: return (x > 0.f && x < 6.f) ? 1.f : 0.f;
    case DYK_ACT_HARDSWISH: return x < -3.f ? 0.f : (x <= 3.f ? x * (1.f / 3.f) + 0.5f : 1.f);
    case DYK_ACT_HARDSIGMOID: return (x > -3.f && x < 3.f) ? (1.f / 6.f) : 0.f;
    default: return 1.f;
  }
}

template <int kAct>
__device__ __forceinline__ float act_grad_t(float x) { return act_grad(x, kAct); }     // the switch folds at compile time

// ------------------------------------------------------------------ channel-fixed thread layout
// Every per-channel kernel below gives a thread ONE 8-channel vector for its whole life, so per-channel parameters are
// loaded once (registers / shared memory) instead of once per element, and walks pixels.  A block of 256 threads is
// (256 >> lg) pixel lanes x (1 << lg) channel-vector lanes; lg is chosen per channel count so that few lanes idle
// (C = 32 -> 4 lanes, C >= 64 -> 8 lanes, MobileNet's 40 / 72 / 120 ... -> whatever wastes < 15 %).
struct ChanGeom {
  int lg;   // log2(channel-vector lanes per block)
  int gx;   // blocks along the channel dimension
};
static inline ChanGeom chan_geom(int C) {
  const int cv = C / 8;
  int best = 0;
  for (int lg = 3; lg >= 0; --lg) {
    const int L = 1 << lg, padded = (cv + L - 1) / L * L;
    if ((padded - cv) * 100 <= 15 * padded) { best = lg; break; }
  }
  return ChanGeom{best, (cv + (1 << best) - 1) >> best};
}

// ------------------------------------------------------------------ per-channel slab reductions
// grid: (geom.gx, slabs).
//   kMode 0: part[slab][0][c] = sum z            part[slab][1][c] = sum z*z                 (BN forward statistics)
//   kMode 1: part[slab][0][c] = sum g            part[slab][1][c] = sum g*(z - mean)        (BN backward), g = dy*act'(zhat);
//                                                 the finalize kernel multiplies the second sum by invstd
//   kMode 2: part[slab][0][c] = sum a            (bias gradient)
//   kMode 3: part[slab][0][c] = sum a*b          (per-channel dot, reduced over channels later: fusion-weight grads,
//                                                 and with pixel ranges restricted to one image: SE gate grads)
//   kMode 4: kMode 1 that also stores g (16-bit) to gout — for activations with an expensive derivative (Mish: the
//            reduce and the apply pass were both bound by instruction issue, not by HBM), so the apply pass is a plain
//            dz = cA*g + cD + cE*z over (g, z) and the exponential is evaluated once per element
// Two pixels per thread are in flight per iteration; lanes of a warp that hold the same channels are combined with
// shuffles, the 8 warps through shared memory — all in a fixed order.
// Finalisation fused into the reduction ("last block done"): every block publishes its slab partials, takes a ticket on
// a per-channel-group counter, and the block that draws the last ticket sums the partials of its channels — in a fixed
// order that does not depend on which block it is, in double — and writes the per-channel results.  Saves one launch
// (and its ~3 us of launch + drain latency on the critical path) per BatchNorm per direction: 364 per dyolov4 step.
struct ReduceFin {
  unsigned* counter;        // [grid.x] tickets, zero on entry, reset to zero by the last block; nullptr = no fused finalize
  float count;
  // forward statistics (kMode 0)
  const float* gamma; const float* beta; float eps, momentum;
  float* running_mean; float* running_var; float* scale_out; float* shift_out; float* mean_out; float* invstd_out;
  // backward (kMode 1 / 4)
  const float* invstd; float* dgamma; float* dbeta; float* coef;
};

template <bool kBf16, int kMode, int kAct>
__global__ void __launch_bounds__(256, (kMode == 1 || kMode == 4) ? 3 : 4)
chan_reduce_kernel(const uint8_t* __restrict__ a, long long as, const uint8_t* __restrict__ b, long long bs,
                   long long pix0, long long npix, int C, int slabs, int lg, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ mean, int act, float* __restrict__ part,
                   uint8_t* __restrict__ gout, long long gs, const ReduceFin fin) {
  constexpr bool kBn = kMode == 1 || kMode == 4;
  const int L = 1 << lg, nplanes = 256 >> lg;
  const int cl = threadIdx.x & (L - 1), plane = threadIdx.x >> lg;
  const int cvec = blockIdx.x * L + cl;
  const bool live = cvec * 8 < C;
  const long long per = (npix + slabs - 1) / slabs;
  const long long p0 = blockIdx.y * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  __shared__ __align__(16) float prm[3][64];
  if constexpr (kBn) {
    if (threadIdx.x < 3 * L * 8) {
      const int k = threadIdx.x / (L * 8), j = threadIdx.x - k * (L * 8);
      const int c = blockIdx.x * L * 8 + j;
      const float* src = k == 0 ? scale : (k == 1 ? shift : mean);
      prm[k][j] = c < C ? __ldg(src + c) : 0.f;
    }
    __syncthreads();
  }
  float s0[8], s1[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) s0[q] = s1[q] = 0.f;
  if (live) {
    const uint8_t* pa = a + (pix0 + p0 + plane) * as * 2 + cvec * 16;
    const uint8_t* pb = kBn || kMode == 3 ? b + (pix0 + p0 + plane) * bs * 2 + cvec * 16 : nullptr;
    uint8_t* pg = kMode == 4 ? gout + (pix0 + p0 + plane) * gs * 2 + cvec * 16 : nullptr;
    const long long sa = (long long)nplanes * as * 2, sb = (long long)nplanes * bs * 2, sg = (long long)nplanes * gs * 2;
    auto accum = [&](const uint4& va, const uint4& vb, uint8_t* gdst) {
      float fa[8];
      unpack8<kBf16>(va, fa);
      if constexpr (kMode == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { s0[q] += fa[q]; s1[q] = fmaf(fa[q], fa[q], s1[q]); }
      } else if constexpr (kMode == 2) {
#pragma unroll
        for (int q = 0; q < 8; ++q) s0[q] += fa[q];
      } else {
        float fb[8];
        unpack8<kBf16>(vb, fb);
        if constexpr (kMode == 3) {
#pragma unroll
          for (int q = 0; q < 8; ++q) s0[q] = fmaf(fa[q], fb[q], s0[q]);
        } else {   // a = dy, b = z
          float gv[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 sc = *reinterpret_cast<const float4*>(&prm[0][cl * 8 + h * 4]);
            const float4 sh = *reinterpret_cast<const float4*>(&prm[1][cl * 8 + h * 4]);
            const float4 mu = *reinterpret_cast<const float4*>(&prm[2][cl * 8 + h * 4]);
            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w}, muv[4] = {mu.x, mu.y, mu.z, mu.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float z = fb[h * 4 + q];
              const float g = fa[h * 4 + q] * act_grad_t<kAct>(fmaf(z, scv[q], shv[q]));
              if constexpr (kMode == 4) gv[h * 4 + q] = g;    // stored rounded to 16 bits; the sums keep fp32
              s0[h * 4 + q] += g;
              s1[h * 4 + q] = fmaf(g, z - muv[q], s1[h * 4 + q]);
            }
          }
          if constexpr (kMode == 4) *reinterpret_cast<uint4*>(gdst) = pack8<kBf16>(gv);
        }
      }
    };
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    long long pidx = p0 + plane;
    for (; pidx + nplanes < p1; pidx += 2 * nplanes) {          // two independent pixels in flight
      const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(pa));
      const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(pa + sa));
      uint4 b0 = zero, b1 = zero;
      if constexpr (kBn || kMode == 3) {
        b0 = __ldg(reinterpret_cast<const uint4*>(pb));
        b1 = __ldg(reinterpret_cast<const uint4*>(pb + sb));
        pb += 2 * sb;
      }
      pa += 2 * sa;
      accum(a0, b0, pg);
      accum(a1, b1, pg + sg);
      if constexpr (kMode == 4) pg += 2 * sg;
    }
    if (pidx < p1) {
      const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(pa));
      uint4 b0 = zero;
      if constexpr (kBn || kMode == 3) b0 = __ldg(reinterpret_cast<const uint4*>(pb));
      accum(a0, b0, pg);
    }
  }
  // lanes of a warp with equal cl -> lane cl (fixed xor tree), then the 8 warps in order
  for (int off = L; off < 32; off <<= 1) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      s0[q] += __shfl_xor_sync(0xffffffffu, s0[q], off);
      if constexpr (kMode <= 1 || kMode == 4) s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], off);
    }
  }
  __shared__ float red[2][8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < L) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      red[0][warp][lane * 8 + q] = s0[q];
      red[1][warp][lane * 8 + q] = s1[q];
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < 2 * L * 8) {
    const int which = threadIdx.x / (L * 8), j = threadIdx.x - which * (L * 8);
    float s = 0.f;
    // with L == 8 a warp holds 4 pixel lanes; with smaller L the warps still partition the pixel lanes
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[which][w][j];
    const int c = blockIdx.x * L * 8 + j;
    if (c < C) part[((long long)blockIdx.y * 2 + which) * C + c] = s;
  }
  if constexpr (kMode == 0 || kMode == 1 || kMode == 4) {
    if (fin.counter == nullptr) return;
    __shared__ int is_last;
    __threadfence();                      // this block's partials are visible device-wide before its ticket
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned t = atomicAdd(&fin.counter[blockIdx.x], 1u);
      is_last = (t == (unsigned)slabs - 1u);
      if (is_last) fin.counter[blockIdx.x] = 0u;       // ready for the next launch that uses this counter
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // nch channels of this block, 256 / nch slab lanes per channel; lane sums (double) -> shared -> lanes added in order
    const int nch = L * 8, lanes = 256 / nch;
    const int ch = threadIdx.x % nch, ln = threadIdx.x / nch;
    const int c = blockIdx.x * nch + ch;
    __shared__ double fred[2][256];
    double a0 = 0.0, a1 = 0.0;
    if (c < C) {
      for (int sl = ln; sl < slabs; sl += lanes) {
        a0 += (double)__ldcg(part + ((long long)sl * 2 + 0) * C + c);
        a1 += (double)__ldcg(part + ((long long)sl * 2 + 1) * C + c);
      }
    }
    fred[0][threadIdx.x] = a0;
    fred[1][threadIdx.x] = a1;
    __syncthreads();
    if (ln != 0 || c >= C) return;
    double s0 = 0.0, s1 = 0.0;
    for (int l2 = 0; l2 < lanes; ++l2) { s0 += fred[0][l2 * nch + ch]; s1 += fred[1][l2 * nch + ch]; }
    const double count = (double)fin.count;
    if constexpr (kMode == 0) {           // same arithmetic as bn_fwd_finalize_kernel
      const double m = s0 / count;
      double var = s1 / count - m * m;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)fin.eps));
      const float g = fin.gamma ? fin.gamma[c] : 1.f, bt = fin.beta ? fin.beta[c] : 0.f;
      fin.scale_out[c] = g * invstd;
      fin.shift_out[c] = bt - (float)m * g * invstd;
      fin.mean_out[c] = (float)m;
      fin.invstd_out[c] = invstd;
      if (fin.running_mean) fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)m;
      if (fin.running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unbiased;
      }
    } else {                              // same arithmetic as bn_bwd_finalize_kernel
      const double is = (double)fin.invstd[c];
      const double sgx = s1 * is;
      if (fin.dgamma) fin.dgamma[c] += (float)sgx;
      if (fin.dbeta) fin.dbeta[c] += (float)s0;
      const double cA = (double)(fin.gamma ? fin.gamma[c] : 1.f) * is;
      const double cC = -cA * sgx / count;
      fin.coef[c] = (float)cA;
      fin.coef[C + c] = (float)(-cA * s0 / count - cC * is * (double)mean[c]);
      fin.coef[2 * C + c] = (float)(cC * is);
    }
  }
}

template <int kMode>
static int launch_chan_reduce(int dtype, const void* a, long long as, const void* b, long long bs, long long pix0,
                              long long npix, int C, int slabs, const float* scale, const float* shift, const float* mean,
                              int act, float* part, cudaStream_t stream, void* gout = nullptr, long long gs = 0,
                              const ReduceFin& fin = ReduceFin{}) {
  const ChanGeom g = chan_geom(C);
  const dim3 grid(g.gx, slabs);
  if constexpr (kMode == 1 || kMode == 4) {
    DYK_DISPATCH_ACT(act, DYK_DISPATCH_DTYPE(dtype, (chan_reduce_kernel<kBf16, kMode, kAct><<<grid, 256, 0, stream>>>(
                                (const uint8_t*)a, as, (const uint8_t*)b, bs, pix0, npix, C, slabs, g.lg, scale, shift, mean,
                                act, part, (uint8_t*)gout, gs, fin))));
  } else {
    DYK_DISPATCH_DTYPE(dtype, (chan_reduce_kernel<kBf16, kMode, DYK_ACT_LINEAR><<<grid, 256, 0, stream>>>(
                                (const uint8_t*)a, as, (const uint8_t*)b, bs, pix0, npix, C, slabs, g.lg, scale, shift, mean,
                                act, part, (uint8_t*)gout, gs, fin)));
  }
  DYK_LAUNCH_OK("chan_reduce_kernel");
  return DYK_OK;
}

// Sum of the slab partials of 8 channels by one block: 32 slab lanes x 8 channels, each lane adds its slabs in order
// (double), then the 32 lanes are added in order.  Returns the two sums to threads 0..7 (channel = threadIdx.x).
__device__ __forceinline__ void slab_totals(const float* __restrict__ part, int slabs, int C, int c0, double& t0, double& t1) {
  __shared__ double red[2][32][8];
  const int ch = threadIdx.x & 7, sl0 = threadIdx.x >> 3;
  const int c = c0 + ch;
  double s0 = 0.0, s1 = 0.0;
  if (c < C) {
    for (int sl = sl0; sl < slabs; sl += 32) {
      s0 += (double)part[((long long)sl * 2 + 0) * C + c];
      s1 += (double)part[((long long)sl * 2 + 1) * C + c];
    }
  }
  red[0][sl0][ch] = s0;
  red[1][sl0][ch] = s1;
  __syncthreads();
  t0 = t1 = 0.0;
  if (threadIdx.x < 8) {
    for (int l = 0; l < 32; ++l) { t0 += red[0][l][threadIdx.x]; t1 += red[1][l][threadIdx.x]; }
  }
}

// BN forward finalize: batch mean / biased variance -> (scale, shift) for y = z*scale + shift, saved (mean, invstd),
// running statistics updated in place exactly like nn.BatchNorm2d(momentum) (unbiased variance in running_var).
// grid: C/8 blocks of 256 threads.
__global__ void __launch_bounds__(256)
bn_fwd_finalize_kernel(const float* __restrict__ part, int slabs, float count, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                       float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                       float* __restrict__ mean_out, float* __restrict__ invstd_out, int C) {
  double s0, s1;
  slab_totals(part, slabs, C, blockIdx.x * 8, s0, s1);
  const int c = blockIdx.x * 8 + threadIdx.x;
  if (threadIdx.x >= 8 || c >= C) return;
  const double m = s0 / count;
  double var = s1 / count - m * m;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - (float)m * g * invstd;
  mean_out[c] = (float)m;
  invstd_out[c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
  if (running_var) {
    const double unbiased = count > 1.f ? var * (double)count / ((double)count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// y = act(z*scale[c] + shift[c]); grid (geom.gx, Y): channel-fixed threads.  kF pixels per thread are in flight per
// iteration; kContig: a block walks its own contiguous range of pixels (sequential streams per block) instead of
// interleaving with all other blocks.
template <bool kBf16, int kAct, int kF = 2, bool kContig = false>
__global__ void __launch_bounds__(256, 4)
bn_act_apply_kernel(const uint8_t* __restrict__ z, long long zs, const float* __restrict__ scale,
                    const float* __restrict__ shift, int act, uint8_t* __restrict__ y, long long ys, long long npix, int C,
                    int lg) {
  const int L = 1 << lg, nplanes = 256 >> lg;
  const int cvec = blockIdx.x * L + (threadIdx.x & (L - 1));
  if (cvec * 8 >= C) return;
  float sc[8], sh[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + cvec * 8)), a1 = __ldg(reinterpret_cast<const float4*>(scale + cvec * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + cvec * 8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + cvec * 8) + 1);
    sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
    sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
  }
  long long step, pix, end;
  if constexpr (kContig) {
    const long long per = ((npix + gridDim.y - 1) / gridDim.y + nplanes - 1) / nplanes * nplanes;
    step = nplanes;
    pix = (long long)blockIdx.y * per + (threadIdx.x >> lg);
    end = (long long)(blockIdx.y + 1) * per < npix ? (long long)(blockIdx.y + 1) * per : npix;
  } else {
    step = (long long)gridDim.y * nplanes;
    pix = (long long)blockIdx.y * nplanes + (threadIdx.x >> lg);
    end = npix;
  }
  auto one = [&](const uint4& v, long long p) {
    float f[8];
    unpack8<kBf16>(v, f);
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = act_apply<kAct>(fmaf(f[q], sc[q], sh[q]));
    *(reinterpret_cast<uint4*>(y + p * ys * 2) + cvec) = pack8<kBf16>(f);
  };
  for (; pix + (kF - 1) * step < end; pix += kF * step) {
    uint4 v[kF];
#pragma unroll
    for (int u = 0; u < kF; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(z + (pix + u * step) * zs * 2) + cvec);
#pragma unroll
    for (int u = 0; u < kF; ++u) one(v[u], pix + u * step);
  }
  for (; pix < end; pix += step) one(__ldg(reinterpret_cast<const uint4*>(z + pix * zs * 2) + cvec), pix);
}

// BN backward finalize: dgamma / dbeta (accumulated into the fp32 parameter gradients) and the per-channel coefficients of
//   dz = cA*g + cD + cE*z   with  g = dy*act'(zhat),  cA = gamma*invstd,
//   cD = -cA*(sum g)/M + cA*invstd^2*mean*(sum g*(z-mean))/M,   cE = -cA*invstd^2*(sum g*(z-mean))/M
// (the usual  cA*(g - mean(g) - xhat*mean(g*xhat))  with xhat = (z-mean)*invstd expanded in z).
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(const float* __restrict__ part, int slabs, float count, const float* __restrict__ gamma,
                       const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, float* __restrict__ coef, int C) {
  double s0, s1;
  slab_totals(part, slabs, C, blockIdx.x * 8, s0, s1);
  const int c = blockIdx.x * 8 + threadIdx.x;
  if (threadIdx.x >= 8 || c >= C) return;
  const double is = (double)invstd[c];
  const double sgx = s1 * is;                       // sum g*xhat
  if (dgamma) dgamma[c] += (float)sgx;
  if (dbeta) dbeta[c] += (float)s0;
  const double cA = (double)(gamma ? gamma[c] : 1.f) * is;
  const double cC = -cA * sgx / count;              // multiplies xhat
  coef[c] = (float)cA;
  coef[C + c] = (float)(-cA * s0 / count - cC * is * (double)mean[c]);
  coef[2 * C + c] = (float)(cC * is);
}

// kFromG: dy already holds g = dy*act'(zhat) (written by chan_reduce mode 4, possibly the dz buffer itself: in place)
template <bool kBf16, bool kFromG, int kAct, int kF = 2, bool kContig = false>
__global__ void __launch_bounds__(256, 3)
bn_act_bwd_apply_kernel(const uint8_t* dy, long long dys, const uint8_t* __restrict__ z, long long zs,
                        const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ coef,
                        int C, int act, uint8_t* dz, long long dzs, long long npix, int lg) {
  const int L = 1 << lg, nplanes = 256 >> lg;
  const int cvec = blockIdx.x * L + (threadIdx.x & (L - 1));
  if (cvec * 8 >= C) return;
  float ps[5][8];   // scale, shift, cA, cD, cE of these 8 channels
  {
    const float* srcs[5] = {scale, shift, coef, coef + C, coef + 2 * C};
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(srcs[v] + cvec * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(srcs[v] + cvec * 8) + 1);
      ps[v][0] = a.x; ps[v][1] = a.y; ps[v][2] = a.z; ps[v][3] = a.w;
      ps[v][4] = b.x; ps[v][5] = b.y; ps[v][6] = b.z; ps[v][7] = b.w;
    }
  }
  long long step, pix, end;
  if constexpr (kContig) {
    const long long per = ((npix + gridDim.y - 1) / gridDim.y + nplanes - 1) / nplanes * nplanes;
    step = nplanes;
    pix = (long long)blockIdx.y * per + (threadIdx.x >> lg);
    end = (long long)(blockIdx.y + 1) * per < npix ? (long long)(blockIdx.y + 1) * per : npix;
  } else {
    step = (long long)gridDim.y * nplanes;
    pix = (long long)blockIdx.y * nplanes + (threadIdx.x >> lg);
    end = npix;
  }
  auto one = [&](const uint4& vg, const uint4& vz, long long p) {
    float g[8], fz[8], o[8];
    unpack8<kBf16>(vg, g);
    unpack8<kBf16>(vz, fz);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float gg = kFromG ? g[q] : g[q] * act_grad_t<kAct>(fmaf(fz[q], ps[0][q], ps[1][q]));
      o[q] = fmaf(ps[2][q], gg, fmaf(ps[4][q], fz[q], ps[3][q]));
    }
    *(reinterpret_cast<uint4*>(dz + p * dzs * 2) + cvec) = pack8<kBf16>(o);
  };
  for (; pix + (kF - 1) * step < end; pix += kF * step) {
    uint4 g[kF], zz[kF];
#pragma unroll
    for (int u = 0; u < kF; ++u) {
      g[u] = *(reinterpret_cast<const uint4*>(dy + (pix + u * step) * dys * 2) + cvec);      // plain loads: may alias dz
      zz[u] = __ldg(reinterpret_cast<const uint4*>(z + (pix + u * step) * zs * 2) + cvec);
    }
#pragma unroll
    for (int u = 0; u < kF; ++u) one(g[u], zz[u], pix + u * step);
  }
  for (; pix < end; pix += step)
    one(*(reinterpret_cast<const uint4*>(dy + pix * dys * 2) + cvec), __ldg(reinterpret_cast<const uint4*>(z + pix * zs * 2) + cvec), pix);
}

// grid.y of the channel-fixed apply kernels: enough blocks to fill the GPU ~8 times over, at least 4 pixels per thread
static inline int apply_grid_y(long long npix, const ChanGeom& g) {
  const int nplanes = 256 >> g.lg;
  long long y = (npix + (long long)nplanes * 4 - 1) / ((long long)nplanes * 4);
  static const int per_sm = getenv("DYK_BN_BLOCKS") ? atoi(getenv("DYK_BN_BLOCKS")) : 8;
  const long long cap = ((long long)num_sms() * per_sm + g.gx - 1) / g.gx;
  if (y > cap) y = cap;
  if (y < 1) y = 1;
  return (int)y;
}
// out[c] (+)= sum over slabs of part[slab][0][c]   (bias gradients); grid: C/8 blocks of 256 threads
__global__ void __launch_bounds__(256)
slab_sum_kernel(const float* __restrict__ part, int slabs, int C, float* __restrict__ out, int accumulate) {
  double s0, s1;
  slab_totals(part, slabs, C, blockIdx.x * 8, s0, s1);
  const int c = blockIdx.x * 8 + threadIdx.x;
  if (threadIdx.x >= 8 || c >= C) return;
  out[c] = accumulate ? out[c] + (float)s0 : (float)s0;
}

// dst = alpha*src (+ dst)
template <bool kBf16>
__global__ void axpby_kernel(const uint8_t* __restrict__ src, long long ss, const float* __restrict__ alpha_ptr,
                             uint8_t* __restrict__ dst, long long ds, long long npix, int cv, int accumulate) {
  const float alpha = alpha_ptr ? __ldg(alpha_ptr) : 1.f;
  const long long total = npix * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / cv;
    const int c = (int)(i - pix * cv);
    float f[8];
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(src + pix * ss * 2) + c), f);
    uint4* dp = reinterpret_cast<uint4*>(dst + pix * ds * 2) + c;
    if (accumulate) {
      float d[8];
      unpack8<kBf16>(*dp, d);
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = fmaf(alpha, f[q], d[q]);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] *= alpha;
    }
    *dp = pack8<kBf16>(f);
  }
}

// WeightedFeatureFusion backward for the raw parameter w: dots[i] = sum(dy * operand_i) was reduced per channel into
// part_i[slab][0][c];  grad_w[i] += dots[i] * (2/n) * sigmoid(w_i) * (1 - sigmoid(w_i))      (layers.py:66)
__global__ void fusion_weights_bwd_kernel(const float* __restrict__ w_raw, const float* __restrict__ part0,
                                          const float* __restrict__ part1, int slabs, int C, int n,
                                          float* __restrict__ grad_w) {
  __shared__ double red[2][256];
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < slabs * C; i += blockDim.x) {
    const int sl = i / C, c = i - sl * C;
    s0 += (double)part0[((long long)sl * 2) * C + c];
    s1 += (double)part1[((long long)sl * 2) * C + c];
  }
  red[0][threadIdx.x] = s0;
  red[1][threadIdx.x] = s1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x < 2) {
    const float sg = 1.f / (1.f + expf(-w_raw[threadIdx.x]));
    grad_w[threadIdx.x] += (float)red[threadIdx.x][0] * (2.f / n) * sg * (1.f - sg);
  }
}

// ------------------------------------------------------------------ MaxPool2d backward
// (1) idx[n][ho][wo][c] = position (hi*W + wi) of the first maximum of the window (scan order, strict '>').
//     One thread = one output pixel x 8 channels (16-byte loads; the scalar version issued one 2-byte load per element
//     and tap: 169 per element for the 13x13 SPP pool).
template <bool kBf16>
__global__ void maxpool_argmax_kernel(const uint8_t* __restrict__ x, long long xs, int N, int H, int W, int C, int k,
                                      int stride, int pad, int Ho, int Wo, int* __restrict__ idx) {
  const int cv = C / 8;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long t = i / cv;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8];
    int bi[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bi[q] = -1; }
    for (int r = 0; r < k; ++r) {
      const int hi = ho * stride - pad + r;
      if (hi < 0 || hi >= H) continue;
      for (int s_ = 0; s_ < k; ++s_) {
        const int wi = wo * stride - pad + s_;
        if (wi < 0 || wi >= W) continue;
        float f[8];
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + hi) * W + wi) * xs * 2) + c), f);
        const int pos = hi * W + wi;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (f[q] > best[q] || bi[q] < 0) { best[q] = f[q]; bi[q] = pos; }
      }
    }
    int4* o = reinterpret_cast<int4*>(idx + ((((long long)n * Ho + ho) * Wo + wo) * C + c * 8));
    o[0] = make_int4(bi[0], bi[1], bi[2], bi[3]);
    o[1] = make_int4(bi[4], bi[5], bi[6], bi[7]);
  }
}
// (2) dx[n][hi][wi][c] (+)= sum of dy over the windows whose argmax is this pixel (fixed order -> deterministic);
//     one thread = one input pixel x 8 channels
template <bool kBf16>
__global__ void maxpool_bwd_kernel(const uint8_t* __restrict__ dy, long long dys, const int* __restrict__ idx, int N, int H,
                                   int W, int C, int k, int stride, int pad, int Ho, int Wo, uint8_t* __restrict__ dx,
                                   long long dxs, int accumulate) {
  const int cv = C / 8;
  const long long total = (long long)N * H * W * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long t = i / cv;
    const int wi = (int)(t % W); t /= W;
    const int hi = (int)(t % H);
    const int n = (int)(t / H);
    const int me = hi * W + wi;
    float s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = 0.f;
    // outputs whose window covers (hi, wi): ho*stride - pad <= hi <= ho*stride - pad + k - 1
    int ho_lo = (hi + pad - k + 1 + stride - 1) / stride; if (hi + pad - k + 1 < 0) ho_lo = 0;
    int wo_lo = (wi + pad - k + 1 + stride - 1) / stride; if (wi + pad - k + 1 < 0) wo_lo = 0;
    const int ho_hi = min(Ho - 1, (hi + pad) / stride), wo_hi = min(Wo - 1, (wi + pad) / stride);
    for (int ho = ho_lo; ho <= ho_hi; ++ho)
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const long long o = (((long long)n * Ho + ho) * Wo + wo);
        const int4* ip = reinterpret_cast<const int4*>(idx + o * C + c * 8);
        const int4 i0 = __ldg(ip), i1 = __ldg(ip + 1);
        const int id[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
        bool any = false;
#pragma unroll
        for (int q = 0; q < 8; ++q) any |= id[q] == me;
        if (!any) continue;
        float f[8];
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(dy + o * dys * 2) + c), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) s[q] += id[q] == me ? f[q] : 0.f;
      }
    uint4* dp = reinterpret_cast<uint4*>(dx + (((long long)n * H + hi) * W + wi) * dxs * 2) + c;
    if (accumulate) {
      float d[8];
      unpack8<kBf16>(*dp, d);
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] += d[q];
    }
    *dp = pack8<kBf16>(s);
  }
}

// ------------------------------------------------------------------ nearest Upsample backward: sum of the s x s block
template <bool kBf16>
__global__ void upsample_bwd_kernel(const uint8_t* __restrict__ dy, long long dys, uint8_t* __restrict__ dx, long long dxs,
                                    int N, int H, int W, int cv, int s, int accumulate) {
  const long long total = (long long)N * H * W * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long t = i / cv;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int a = 0; a < s; ++a)
      for (int b = 0; b < s; ++b) {
        const long long op = ((long long)n * H * s + h * s + a) * W * s + w * s + b;
        float f[8];
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(dy + op * dys * 2) + c), f);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += f[q];
      }
    uint4* dp = reinterpret_cast<uint4*>(dx + (((long long)n * H + h) * W + w) * dxs * 2) + c;
    if (accumulate) {
      float d[8];
      unpack8<kBf16>(*dp, d);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += d[q];
    }
    *dp = pack8<kBf16>(acc);
  }
}

// ------------------------------------------------------------------ SqueezeExcitation backward (layers.py:184-190)
// (1) one block per image: recompute the tiny MLP and back-propagate the gate gradient through it
//       dv2 = dgate * [ -3 < v2 < 3 ] / 6, dhid = W2^T dv2, dv1 = dhid * [v1 > 0], dmean = W1^T dv1
//     keeps mean / relu(hid) / dv2 / dv1 per image for (2);  dmean_out[n][c] = dmean / HW (added to every pixel of dx)
__global__ void __launch_bounds__(1024)
se_mlp_bwd_kernel(const float* __restrict__ pooled, int slabs, float inv_hw, const float* __restrict__ dgate_part,
                  int dslabs, int C, int Csq, const float* __restrict__ w1, const float* __restrict__ b1,
                  const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ keep,
                  float* __restrict__ dmean_out) {
  extern __shared__ float sm[];
  float* mean = sm;            // [C]
  float* dv2 = mean + C;       // [C]
  float* hid = dv2 + C;        // [Csq]  pre-activation v1
  float* dv1 = hid + Csq;      // [Csq]
  const int n = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  for (int c = tid; c < C; c += nt) {
    float s = 0.f;
    for (int sl = 0; sl < slabs; ++sl) s += pooled[((long long)n * slabs + sl) * C + c];
    mean[c] = s * inv_hw;
  }
  __syncthreads();
  for (int j = warp; j < Csq; j += nwarps) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += w1[(long long)j * C + c] * mean[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) hid[j] = s + b1[j];
  }
  __syncthreads();
  for (int c = warp; c < C; c += nwarps) {
    float s = 0.f;
    for (int j = lane; j < Csq; j += 32) s += w2[(long long)c * Csq + j] * fmaxf(hid[j], 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float v2 = s + b2[c];
      float dg = 0.f;
      for (int sl = 0; sl < dslabs; ++sl) dg += dgate_part[(((long long)n * dslabs + sl) * 2) * C + c];
      dv2[c] = (v2 > -3.f && v2 < 3.f) ? dg * (1.f / 6.f) : 0.f;
    }
  }
  __syncthreads();
  for (int j = warp; j < Csq; j += nwarps) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += w2[(long long)c * Csq + j] * dv2[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dv1[j] = hid[j] > 0.f ? s : 0.f;
  }
  __syncthreads();
  float* kn = keep + (long long)n * (2 * C + 2 * Csq);     // [mean C | dv2 C | relu(hid) Csq | dv1 Csq]
  for (int c = tid; c < C; c += nt) {
    kn[c] = mean[c];
    kn[C + c] = dv2[c];
    float s = 0.f;
    for (int j = 0; j < Csq; ++j) s += w1[(long long)j * C + c] * dv1[j];
    dmean_out[(long long)n * C + c] = s * inv_hw;
  }
  for (int j = tid; j < Csq; j += nt) {
    kn[2 * C + j] = fmaxf(hid[j], 0.f);
    kn[2 * C + Csq + j] = dv1[j];
  }
}
// (2) weight / bias gradients, one thread per element, images summed in a fixed order:
//       gw2[c][j] += sum_n dv2[n][c] * relu(hid)[n][j]     gb2[c] += sum_n dv2[n][c]
//       gw1[j][c] += sum_n dv1[n][j] * mean[n][c]          gb1[j] += sum_n dv1[n][j]
__global__ void se_wgrad_kernel(const float* __restrict__ keep, int N, int C, int Csq, float* __restrict__ gw1,
                                float* __restrict__ gb1, float* __restrict__ gw2, float* __restrict__ gb2) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nw = (long long)C * Csq;
  const int stride = 2 * C + 2 * Csq;
  if (i < nw) {
    const int c = (int)(i / Csq), j = (int)(i - (long long)c * Csq);       // gw2 [C][Csq]
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += keep[(long long)n * stride + C + c] * keep[(long long)n * stride + 2 * C + j];
    gw2[i] += s;
  } else if (i < 2 * nw) {
    const long long e = i - nw;
    const int j = (int)(e / C), c = (int)(e - (long long)j * C);           // gw1 [Csq][C]
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += keep[(long long)n * stride + 2 * C + Csq + j] * keep[(long long)n * stride + c];
    gw1[e] += s;
  } else if (i < 2 * nw + C) {
    const int c = (int)(i - 2 * nw);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += keep[(long long)n * stride + C + c];
    gb2[c] += s;
  } else if (i < 2 * nw + C + Csq) {
    const int j = (int)(i - 2 * nw - C);
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += keep[(long long)n * stride + 2 * C + Csq + j];
    gb1[j] += s;
  }
}
// dx (+)= dy * gate[n][c] + dmean[n][c]
template <bool kBf16>
__global__ void se_bwd_apply_kernel(const uint8_t* __restrict__ dy, long long dys, const float* __restrict__ gate,
                                    const float* __restrict__ dmean, uint8_t* __restrict__ dx, long long dxs, int N, int HW,
                                    int C, int accumulate) {
  const int cv = C / 8;
  const long long total = (long long)N * HW * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    const long long pix = i / cv;
    const int n = (int)(pix / HW);
    float f[8];
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(dy + pix * dys * 2) + c), f);
    uint4* dp = reinterpret_cast<uint4*>(dx + pix * dxs * 2) + c;
    float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (accumulate) unpack8<kBf16>(*dp, d);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const long long gi = (long long)n * C + c * 8 + q;
      f[q] = f[q] * __ldg(gate + gi) + __ldg(dmean + gi) + d[q];
    }
    *dp = pack8<kBf16>(f);
  }
}

// ------------------------------------------------------------------ YOLOLayer (training) backward
// dz[n][y][x][a*no + o] = dp[n][a][y][x][o]  (inverse of the view+permute of models.py:229), written as 16-bit NHWC
// with the channel dimension zero-padded to Cpad so the tensor-core dgrad / wgrad kernels can consume it.
template <bool kBf16>
__global__ void yolo_train_bwd_kernel(const float* __restrict__ dp, int N, int na, int ny, int nx, int no, void* __restrict__ dz,
                                      int Cpad) {
  const long long total = (long long)N * ny * nx * Cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % Cpad);
    long long t = i / Cpad;
    const int x = (int)(t % nx); t /= nx;
    const int y = (int)(t % ny);
    const int n = (int)(t / ny);
    float v = 0.f;
    if (ch < na * no) {
      const int a = ch / no, o = ch - a * no;
      v = dp[((((long long)n * na + a) * ny + y) * nx + x) * no + o];
    }
    store1<kBf16>(dz, i, v);
  }
}

// ------------------------------------------------------------------ stem frames -> NHWC, 8 channels (zero padded)
template <bool kBf16, typename TIn>
__global__ void frames_to_nhwc8_kernel(const TIn* __restrict__ x, uint8_t* __restrict__ y, int N, int Cin, int HW) {
  const long long total = (long long)N * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < Cin; ++c) {
      const TIn v = x[(n * Cin + c) * HW + p];
      if constexpr (sizeof(TIn) == 1) f[c] = (float)v / 255.0f;
      else f[c] = (float)v;
    }
    *reinterpret_cast<uint4*>(y + i * 16) = pack8<kBf16>(f);
  }
}

// ------------------------------------------------------------------ stem frames -> im2col rows (3x3, pad 1, Cin = 3)
// y[n][h][w][ci*9 + r*3 + s] = frame[n][ci][h + r - 1][w + s - 1] (0 outside, channels 27..31 zero): with the 3x3
// neighbourhood as 32 "channels" the stem weight gradient is the weight gradient of a 1x1 convolution, and the channel
// order makes its [Cout][27] result the OIHW gradient itself.
// Hs > 0: x holds Hs x Ws frames and the rows are taken from their bilinear resize to H x W (frame_sample.cuh).
template <bool kBf16, typename TIn>
__global__ void frames_to_im2col32_kernel(const TIn* __restrict__ x, uint8_t* __restrict__ y, int N, int H, int W, int Hs, int Ws,
                                          float rs_h, float rs_w) {
  const long long total = (long long)N * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long t = i / W;
    const int h = (int)(t % H);
    const long long n = t / H;
    float f[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) f[k] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int ww = w + q - 1;
          if (ww < 0 || ww >= W) continue;
          if (Hs > 0) {
            f[ci * 9 + r * 3 + q] = frame_bilinear(x + (n * 3 + ci) * (long long)Hs * Ws, Ws, resize_axis(hh, rs_h, Hs),
                                                   resize_axis(ww, rs_w, Ws));
            continue;
          }
          const TIn v = __ldg(&x[((n * 3 + ci) * H + hh) * W + ww]);
          if constexpr (sizeof(TIn) == 1) f[ci * 9 + r * 3 + q] = (float)v / 255.0f;
          else f[ci * 9 + r * 3 + q] = (float)v;
        }
      }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float g[8] = {f[8 * c], f[8 * c + 1], f[8 * c + 2], f[8 * c + 3], f[8 * c + 4], f[8 * c + 5], f[8 * c + 6], f[8 * c + 7]};
      *reinterpret_cast<uint4*>(y + i * 64 + c * 16) = pack8<kBf16>(g);
    }
  }
}

// ------------------------------------------------------------------ dgrad weight packing
// OIHW fp32 -> [I][kh][kw][Opad] 16-bit with the taps rotated by 180 degrees: W'[i][r][s][o] = W[o][i][kh-1-r][kw-1-s]
template <bool kBf16>
__global__ void pack_dgrad_kernel(const float* __restrict__ w, void* __restrict__ out, int O, int I, int kh, int kw, int Opad) {
  const long long total = (long long)I * kh * kw * Opad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % Opad);
    long long t = i / Opad;
    const int s = (int)(t % kw); t /= kw;
    const int r = (int)(t % kh);
    const int ci = (int)(t / kh);
    float v = 0.f;
    if (o < O) v = w[(((long long)o * I + ci) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)];
    store1<kBf16>(out, i, v);
  }
}

// All convolution weights of a model in ONE launch: OIHW fp32 -> the forward layout [O][kh][kw][I] and the data-gradient
// layout [I][kh][kw][Opad] (taps rotated by 180 degrees) from a single coalesced read.  A block owns a 32 x 32
// (out-channel x in-channel) tile of one layer with all its taps in shared memory and writes >= 64-byte segments to both
// destinations.  descs: device int64 [n][8] = (w, fwd, dgrad or 0, O, I, taps (<= 9), Opad, first tile of this layer).
constexpr int kPackTaps = 9;
template <bool kBf16>
__global__ void __launch_bounds__(256)
pack_multi_kernel(const long long* __restrict__ descs, int n) {
  __shared__ float tile[32][32 * kPackTaps + 1];
  int lo = 0, hi = n - 1;                                    // last layer whose first tile is <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid * 8 + 7] <= (long long)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const long long* d = descs + lo * 8;
  const float* w = reinterpret_cast<const float*>(d[0]);
  void* fwd = reinterpret_cast<void*>(d[1]);
  void* dg = reinterpret_cast<void*>(d[2]);
  const int O = (int)d[3], I = (int)d[4], taps = (int)d[5], Opad = (int)d[6];
  const int t = blockIdx.x - (int)d[7];
  const int tiles_i = (I + 31) / 32;
  const int o0 = (t / tiles_i) * 32, i0 = (t % tiles_i) * 32;
  const int no = min(32, O - o0), ni = min(32, I - i0);
  const int row = ni * taps;
  for (int e = threadIdx.x; e < no * row; e += 256) {
    const int o = e / row, j = e - o * row;
    tile[o][j] = __ldg(w + ((long long)(o0 + o) * I + i0) * taps + j);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < no * row; e += 256) {       // forward layout, in-channel fastest
    const int ci = e % ni, r = e / ni;
    const int tap = r % taps, o = r / taps;
    store1<kBf16>(fwd, ((long long)(o0 + o) * taps + tap) * I + i0 + ci, tile[o][ci * taps + tap]);
  }
  if (dg != nullptr) {
    for (int e = threadIdx.x; e < no * row; e += 256) {     // dgrad layout, out-channel fastest, taps reversed
      const int o = e % no, r = e / no;
      const int tap = r % taps, ci = r / taps;
      store1<kBf16>(dg, ((long long)(i0 + ci) * taps + (taps - 1 - tap)) * Opad + o0 + o, tile[o][ci * taps + tap]);
    }
  }
}

}  // namespace dyk

using namespace dyk;

#define DYK_EXPORT extern "C" __attribute__((visibility("default")))
#define DYK_AL16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

DYK_EXPORT int dyk_bn_train_stats(const void* z, int64_t zs, int64_t npix, int32_t C, int32_t dtype, const float* gamma,
                                  const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                  float* scale, float* shift, float* mean, float* invstd, float* workspace, uint32_t* counters,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(z && scale && shift && mean && invstd && workspace, "dyk_bn_train_stats: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && zs % 8 == 0 && npix > 0 && DYK_AL16(z), "dyk_bn_train_stats: bad shape");
  const int slabs = slabs_for(npix);
  // finalisation fused into the reduction kernel (one launch) while the last block's serial sum over the slabs is short;
  // with ~1000 slabs (the 128x160 and larger layers) the separate, wider finalize kernel is faster (measured +20 us / layer)
  if (slabs > kFusedFinalizeMaxSlabs) counters = nullptr;
  if (counters != nullptr) {
    ReduceFin fin{};
    fin.counter = counters; fin.count = (float)npix; fin.gamma = gamma; fin.beta = beta; fin.eps = eps; fin.momentum = momentum;
    fin.running_mean = running_mean; fin.running_var = running_var; fin.scale_out = scale; fin.shift_out = shift;
    fin.mean_out = mean; fin.invstd_out = invstd;
    return launch_chan_reduce<0>(dtype, z, zs, nullptr, 0, 0, npix, C, slabs, nullptr, nullptr, nullptr, 0, workspace, stream,
                                 nullptr, 0, fin);
  }
  if (int rc = launch_chan_reduce<0>(dtype, z, zs, nullptr, 0, 0, npix, C, slabs, nullptr, nullptr, nullptr, 0, workspace, stream))
    return rc;
  bn_fwd_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(workspace, slabs, (float)npix, gamma, beta, eps, momentum,
                                                         running_mean, running_var, scale, shift, mean, invstd, C);
  DYK_LAUNCH_OK("bn_fwd_finalize_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_bn_act_apply(const void* z, int64_t zs, const float* scale, const float* shift, int32_t act, void* y,
                                int64_t ys, int64_t npix, int32_t C, int32_t dtype, void* stream_) {
  DYK_REQUIRE(z && scale && shift && y, "dyk_bn_act_apply: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && zs % 8 == 0 && ys % 8 == 0 && DYK_AL16(z) && DYK_AL16(y), "dyk_bn_act_apply: bad shape");
  if (npix == 0) return DYK_OK;
  const ChanGeom g = chan_geom(C);
  const dim3 grid(g.gx, apply_grid_y(npix, g));
  DYK_DISPATCH_ACT(act, DYK_DISPATCH_DTYPE(dtype, (bn_act_apply_kernel<kBf16, kAct><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                (const uint8_t*)z, zs, scale, shift, act, (uint8_t*)y, ys, npix, C, g.lg))));
  DYK_LAUNCH_OK("bn_act_apply_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_bn_act_bwd(const void* dy, int64_t dys, const void* z, int64_t zs, const float* scale, const float* shift,
                              const float* mean, const float* invstd, const float* gamma, int32_t act, int64_t npix,
                              int32_t C, int32_t dtype, void* dz, int64_t dzs, float* dgamma, float* dbeta,
                              float* workspace, uint32_t* counters, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(dy && z && scale && shift && mean && invstd && dz && workspace, "dyk_bn_act_bwd: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && zs % 8 == 0 && dys % 8 == 0 && dzs % 8 == 0 && npix > 0, "dyk_bn_act_bwd: bad shape");
  DYK_REQUIRE(DYK_AL16(dy) && DYK_AL16(z) && DYK_AL16(dz), "dyk_bn_act_bwd: 16-byte alignment");
  const int slabs = slabs_for(npix);
  float* part = workspace;                                  // [slabs][2][C]
  float* coef = workspace + (size_t)kMaxSlabs * 2 * C;      // [3][C]
  // expensive derivative: evaluate it once, store g in the dz buffer, apply in place (see chan_reduce mode 4)
  static const bool g_off = getenv("DYK_BN_GSTORE") != nullptr && getenv("DYK_BN_GSTORE")[0] == '0';
  const bool store_g = act == DYK_ACT_MISH && !g_off;
  if (slabs > kFusedFinalizeMaxSlabs) counters = nullptr;
  ReduceFin fin{};
  fin.counter = counters; fin.count = (float)npix; fin.gamma = gamma; fin.invstd = invstd; fin.dgamma = dgamma; fin.dbeta = dbeta;
  fin.coef = coef;
  if (store_g) {
    if (int rc = launch_chan_reduce<4>(dtype, dy, dys, z, zs, 0, npix, C, slabs, scale, shift, mean, act, part, stream, dz, dzs, fin))
      return rc;
  } else {
    if (int rc = launch_chan_reduce<1>(dtype, dy, dys, z, zs, 0, npix, C, slabs, scale, shift, mean, act, part, stream, nullptr, 0, fin))
      return rc;
  }
  if (counters == nullptr) {
    bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(part, slabs, (float)npix, gamma, mean, invstd, dgamma, dbeta, coef, C);
    DYK_LAUNCH_OK("bn_bwd_finalize_kernel");
  }
  const ChanGeom g = chan_geom(C);
  const dim3 grid(g.gx, apply_grid_y(npix, g));
  if (store_g) {
    DYK_DISPATCH_DTYPE(dtype, (bn_act_bwd_apply_kernel<kBf16, true, DYK_ACT_LINEAR><<<grid, 256, 0, stream>>>(
                                  (const uint8_t*)dz, dzs, (const uint8_t*)z, zs, scale, shift, coef, C, act, (uint8_t*)dz, dzs,
                                  npix, g.lg)));
  } else {
    DYK_DISPATCH_ACT(act, DYK_DISPATCH_DTYPE(dtype, (bn_act_bwd_apply_kernel<kBf16, false, kAct><<<grid, 256, 0, stream>>>(
                                  (const uint8_t*)dy, dys, (const uint8_t*)z, zs, scale, shift, coef, C, act, (uint8_t*)dz, dzs,
                                  npix, g.lg))));
  }
  DYK_LAUNCH_OK("bn_act_bwd_apply_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_chan_sum(const void* x, int64_t xs, int64_t npix, int32_t C, int32_t dtype, float* out,
                            int32_t accumulate, float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x && out && workspace, "dyk_chan_sum: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && npix > 0 && DYK_AL16(x), "dyk_chan_sum: bad shape");
  const int slabs = slabs_for(npix);
  if (int rc = launch_chan_reduce<2>(dtype, x, xs, nullptr, 0, 0, npix, C, slabs, nullptr, nullptr, nullptr, 0, workspace, stream))
    return rc;
  slab_sum_kernel<<<(C + 7) / 8, 256, 0, stream>>>(workspace, slabs, C, out, accumulate);
  DYK_LAUNCH_OK("slab_sum_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_axpby(const void* src, int64_t ss, const float* alpha, void* dst, int64_t ds, int64_t npix, int32_t C,
                         int32_t accumulate, int32_t dtype, void* stream_) {
  DYK_REQUIRE(src && dst, "dyk_axpby: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && ss % 8 == 0 && ds % 8 == 0 && DYK_AL16(src) && DYK_AL16(dst), "dyk_axpby: bad shape");
  if (npix == 0) return DYK_OK;
  const int cv = C / 8;
  DYK_DISPATCH_DTYPE(dtype, (axpby_kernel<kBf16><<<grid_for_t(npix * cv, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                (const uint8_t*)src, ss, alpha, (uint8_t*)dst, ds, npix, cv, accumulate)));
  DYK_LAUNCH_OK("axpby_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_fusion_weights_bwd(const void* dy, int64_t dys, const void* a, int64_t as, const void* b, int64_t bs,
                                      int64_t npix, int32_t C, int32_t dtype, const float* w_raw, int32_t n, float* grad_w,
                                      float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(dy && a && b && w_raw && grad_w && workspace, "dyk_fusion_weights_bwd: null pointer");
  DYK_REQUIRE(n == 2, "dyk_fusion_weights_bwd: only two-operand fusion is supported (n=%d)", n);
  DYK_REQUIRE(C > 0 && C % 8 == 0 && dys % 8 == 0 && as % 8 == 0 && bs % 8 == 0 && npix > 0, "dyk_fusion_weights_bwd: bad shape");
  const int slabs = slabs_for(npix);
  float* part0 = workspace;
  float* part1 = workspace + (size_t)kMaxSlabs * 2 * C;
  if (int rc = launch_chan_reduce<3>(dtype, dy, dys, a, as, 0, npix, C, slabs, nullptr, nullptr, nullptr, 0, part0, stream)) return rc;
  if (int rc = launch_chan_reduce<3>(dtype, dy, dys, b, bs, 0, npix, C, slabs, nullptr, nullptr, nullptr, 0, part1, stream)) return rc;
  fusion_weights_bwd_kernel<<<1, 256, 0, stream>>>(w_raw, part0, part1, slabs, C, n, grad_w);
  DYK_LAUNCH_OK("fusion_weights_bwd_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_maxpool2d_bwd(const void* x, int64_t xs, const void* dy, int64_t dys, void* dx, int64_t dxs, int32_t N,
                                 int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride, int32_t accumulate, int32_t dtype,
                                 int32_t* idx_workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x && dy && dx && idx_workspace, "dyk_maxpool2d_bwd: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && dys % 8 == 0 && dxs % 8 == 0 && k >= 1 && stride >= 1,
              "dyk_maxpool2d_bwd: C and the pixel strides must be multiples of 8");
  DYK_REQUIRE(DYK_AL16(x) && DYK_AL16(dy) && DYK_AL16(dx) && DYK_AL16(idx_workspace), "dyk_maxpool2d_bwd: 16-byte alignment");
  const int pad = (k - 1) / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_maxpool2d_bwd: empty output");
  DYK_DISPATCH_DTYPE(dtype, (maxpool_argmax_kernel<kBf16><<<grid_for_t((long long)N * Ho * Wo * (C / 8), 128), 128, 0, stream>>>(
                                (const uint8_t*)x, xs, N, H, W, C, k, stride, pad, Ho, Wo, idx_workspace)));
  DYK_LAUNCH_OK("maxpool_argmax_kernel");
  DYK_DISPATCH_DTYPE(dtype, (maxpool_bwd_kernel<kBf16><<<grid_for_t((long long)N * H * W * (C / 8), 128), 128, 0, stream>>>(
                                (const uint8_t*)dy, dys, idx_workspace, N, H, W, C, k, stride, pad, Ho, Wo, (uint8_t*)dx, dxs,
                                accumulate)));
  DYK_LAUNCH_OK("maxpool_bwd_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_upsample_nearest_bwd(const void* dy, int64_t dys, void* dx, int64_t dxs, int32_t N, int32_t H, int32_t W,
                                        int32_t C, int32_t s, int32_t accumulate, int32_t dtype, void* stream_) {
  DYK_REQUIRE(dy && dx, "dyk_upsample_nearest_bwd: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && dys % 8 == 0 && dxs % 8 == 0 && s >= 1, "dyk_upsample_nearest_bwd: bad shape");
  const int cv = C / 8;
  DYK_DISPATCH_DTYPE(dtype, (upsample_bwd_kernel<kBf16><<<grid_for_t((long long)N * H * W * cv, 256), 256, 0,
                                                        static_cast<cudaStream_t>(stream_)>>>(
                                (const uint8_t*)dy, dys, (uint8_t*)dx, dxs, N, H, W, cv, s, accumulate)));
  DYK_LAUNCH_OK("upsample_bwd_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_se_bwd(const void* x, int64_t xs, const void* dy, int64_t dys, void* dx, int64_t dxs, int32_t N, int32_t HW,
                          int32_t C, const float* w1, const float* b1, const float* w2, const float* b2, int32_t Csq,
                          const float* pooled, const float* gate, float* gw1, float* gb1, float* gw2, float* gb2,
                          int32_t accumulate, int32_t dtype, float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x && dy && dx && w1 && b1 && w2 && b2 && pooled && gate && gw1 && gb1 && gw2 && gb2 && workspace,
              "dyk_se_bwd: null pointer");
  DYK_REQUIRE(C > 0 && C % 8 == 0 && xs % 8 == 0 && dys % 8 == 0 && dxs % 8 == 0 && Csq > 0 && N > 0 && HW > 0, "dyk_se_bwd: bad shape");
  DYK_REQUIRE((size_t)(2 * C + 2 * Csq) * 4 <= 48 * 1024, "dyk_se_bwd: C + Csq too large");
  // forward used slabs = clamp(HW/256, 1, 32) per image for `pooled` (dyk_se_gate)
  int fslabs = HW / 256; if (fslabs < 1) fslabs = 1; if (fslabs > 31) fslabs = 31;
  int dslabs = slabs_for(HW); if (dslabs > 32) dslabs = 32;
  float* dgate_part = workspace;                                   // [N][dslabs][2][C]
  float* dmean = workspace + (size_t)N * 32 * 2 * C;               // [N][C]
  float* keep = dmean + (size_t)N * C;                             // [N][2C + 2Csq]  (<= 4 N C floats)
  for (int n = 0; n < N; ++n) {   // dgate[n][c] = sum_hw dy * x
    if (int rc = launch_chan_reduce<3>(dtype, dy, dys, x, xs, (long long)n * HW, HW, C, dslabs, nullptr, nullptr, nullptr, 0,
                                       dgate_part + (size_t)n * dslabs * 2 * C, stream))
      return rc;
  }
  DYK_REQUIRE(Csq <= C, "dyk_se_bwd: Csq > C");
  se_mlp_bwd_kernel<<<N, 1024, (2 * C + 2 * Csq) * sizeof(float), stream>>>(pooled, fslabs, 1.f / (float)HW, dgate_part, dslabs, C,
                                                                          Csq, w1, b1, w2, b2, keep, dmean);
  DYK_LAUNCH_OK("se_mlp_bwd_kernel");
  const long long nel = 2ll * C * Csq + C + Csq;
  se_wgrad_kernel<<<(int)((nel + 255) / 256), 256, 0, stream>>>(keep, N, C, Csq, gw1, gb1, gw2, gb2);
  DYK_LAUNCH_OK("se_wgrad_kernel");
  DYK_DISPATCH_DTYPE(dtype, (se_bwd_apply_kernel<kBf16><<<grid_for_t((long long)N * HW * (C / 8), 256), 256, 0, stream>>>(
                                (const uint8_t*)dy, dys, gate, dmean, (uint8_t*)dx, dxs, N, HW, C, accumulate)));
  DYK_LAUNCH_OK("se_bwd_apply_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_yolo_train_bwd(const float* dp, int32_t N, int32_t na, int32_t ny, int32_t nx, int32_t no, void* dz,
                                  int32_t Cpad, int32_t dtype, void* stream_) {
  DYK_REQUIRE(dp && dz && Cpad >= na * no && Cpad % 8 == 0, "dyk_yolo_train_bwd: bad arguments");
  DYK_DISPATCH_DTYPE(dtype, (yolo_train_bwd_kernel<kBf16><<<grid_for_t((long long)N * ny * nx * Cpad, 256), 256, 0,
                                                          static_cast<cudaStream_t>(stream_)>>>(dp, N, na, ny, nx, no, dz, Cpad)));
  DYK_LAUNCH_OK("yolo_train_bwd_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_pack_weights_dgrad(const float* w_oihw, void* w_packed, int32_t O, int32_t I, int32_t kh, int32_t kw,
                                      int32_t Opad, int32_t dtype, void* stream_) {
  DYK_REQUIRE(w_oihw && w_packed && O > 0 && I > 0 && kh > 0 && kw > 0 && Opad >= O, "dyk_pack_weights_dgrad: bad arguments");
  DYK_DISPATCH_DTYPE(dtype, (pack_dgrad_kernel<kBf16><<<grid_for_t((long long)I * kh * kw * Opad, 256), 256, 0,
                                                      static_cast<cudaStream_t>(stream_)>>>(w_oihw, w_packed, O, I, kh, kw, Opad)));
  DYK_LAUNCH_OK("pack_dgrad_kernel");
  return DYK_OK;
}

static int frames_to_im2col32(const void* x_nchw, void* y, int32_t N, int32_t Hs, int32_t Ws, int32_t H, int32_t W, int32_t dtype,
                              int32_t x_kind, cudaStream_t stream) {
  DYK_REQUIRE(x_nchw && y && N > 0 && H > 0 && W > 0 && DYK_AL16(y), "dyk_frames_to_im2col32: bad arguments");
  const int grid = grid_for_t((long long)N * H * W, 128);
  const float rs_h = Hs > 0 ? (float)Hs / (float)H : 0.f, rs_w = Hs > 0 ? (float)Ws / (float)W : 0.f;
  if (x_kind == 1) {
    DYK_DISPATCH_DTYPE(dtype, (frames_to_im2col32_kernel<kBf16, uint8_t><<<grid, 128, 0, stream>>>((const uint8_t*)x_nchw, (uint8_t*)y, N, H, W, Hs, Ws, rs_h, rs_w)));
  } else {
    DYK_DISPATCH_DTYPE(dtype, (frames_to_im2col32_kernel<kBf16, float><<<grid, 128, 0, stream>>>((const float*)x_nchw, (uint8_t*)y, N, H, W, Hs, Ws, rs_h, rs_w)));
  }
  DYK_LAUNCH_OK("frames_to_im2col32_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_frames_to_im2col32(const void* x_nchw, void* y, int32_t N, int32_t H, int32_t W, int32_t dtype, int32_t x_kind,
                                      void* stream_) {
  return frames_to_im2col32(x_nchw, y, N, 0, 0, H, W, dtype, x_kind, static_cast<cudaStream_t>(stream_));
}

DYK_EXPORT int dyk_frames_to_im2col32_resize(const void* x_nchw, void* y, int32_t N, int32_t Hs, int32_t Ws, int32_t H, int32_t W,
                                             int32_t dtype, int32_t x_kind, void* stream_) {
  DYK_REQUIRE(Hs > 0 && Ws > 0, "dyk_frames_to_im2col32_resize: bad source size");
  return frames_to_im2col32(x_nchw, y, N, Hs, Ws, H, W, dtype, x_kind, static_cast<cudaStream_t>(stream_));
}

DYK_EXPORT int dyk_pack_weights_multi(const int64_t* descs, int32_t n, int32_t total_tiles, int32_t dtype, void* stream_) {
  DYK_REQUIRE(descs && n > 0 && total_tiles > 0, "dyk_pack_weights_multi: bad arguments");
  DYK_DISPATCH_DTYPE(dtype, (pack_multi_kernel<kBf16><<<total_tiles, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
                                reinterpret_cast<const long long*>(descs), n)));
  DYK_LAUNCH_OK("pack_multi_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_frames_to_nhwc8(const void* x_nchw, void* y, int32_t N, int32_t Cin, int32_t H, int32_t W, int32_t dtype,
                                   int32_t x_kind, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x_nchw && y && N > 0 && Cin >= 1 && Cin <= 8 && H > 0 && W > 0 && DYK_AL16(y), "dyk_frames_to_nhwc8: bad arguments");
  const int grid = grid_for_t((long long)N * H * W, 256);
  if (x_kind == 1) {
    DYK_DISPATCH_DTYPE(dtype, (frames_to_nhwc8_kernel<kBf16, uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)x_nchw, (uint8_t*)y, N, Cin, H * W)));
  } else {
    DYK_DISPATCH_DTYPE(dtype, (frames_to_nhwc8_kernel<kBf16, float><<<grid, 256, 0, stream>>>((const float*)x_nchw, (uint8_t*)y, N, Cin, H * W)));
  }
  DYK_LAUNCH_OK("frames_to_nhwc8_kernel");
  return DYK_OK;
}
