// Fused multi-tensor optimizer step: ONE launch updates every parameter tensor of a param group.
//
// Replaces torch.optim.SGD(momentum, nesterov=True, weight_decay) / torch.optim.Adam(betas, weight_decay) as the
// reference builds them (train.py:85-91) and steps them through GradScaler (train_utils/kaist_train_eval_utils.py:
// 103-108): 568 parameter tensors in dyolov4_fshare, i.e. 568 x (3..6) elementwise launches per step in eager PyTorch, a
// handful of multi-tensor launches with foreach=True.  Here a device descriptor table (one row per tensor, like
// pack_multi_kernel) lets a block find its tensor by binary search; the GradScaler unscale (g / scale) and the
// "skip the step when a gradient overflowed" decision are read from device memory, so the step never synchronises.
//
// Arithmetic follows torch/optim/sgd.py and adam.py (single-tensor formulas, fp32):
//   SGD : g += wd*p;  buf = first ? g : momentum*buf + (1-dampening)*g;  g = nesterov ? g + momentum*buf : buf;  p -= lr*g
//   Adam: g += wd*p;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// Memory-bound: SGD moves 5 floats per element (read p, g, buf; write p, buf), Adam 7.
#include "common.h"

namespace dyk {

constexpr int kOptThreads = 256;
constexpr int kOptPerBlock = 4096;     // elements per block: 4 float4 per thread

struct OptRow {            // one row of the int64 [n][6] table
  float* p;
  const float* g;
  float* s1;               // SGD momentum buffer / Adam exp_avg
  float* s2;               // Adam exp_avg_sq (unused for SGD)
  long long numel;
  long long first_block;
};

__device__ __forceinline__ const long long* opt_find(const long long* __restrict__ descs, int n, long long block) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid * 6 + 5] <= block) lo = mid; else hi = mid - 1;
  }
  return descs + lo * 6;
}

struct SgdArgs { float lr, momentum, dampening, wd; int nesterov, first; };
struct AdamArgs { float step_size, b1, b2, eps, wd, inv_sqrt_bc2; };

__device__ __forceinline__ void sgd_one(float& p, float g, float& buf, const SgdArgs& a, float inv_scale) {
  g *= inv_scale;
  if (a.wd != 0.f) g = fmaf(a.wd, p, g);
  if (a.momentum != 0.f) {
    buf = a.first ? g : fmaf(a.momentum, buf, (1.f - a.dampening) * g);
    g = a.nesterov ? fmaf(a.momentum, buf, g) : buf;
  }
  p = fmaf(-a.lr, g, p);
}
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float inv_scale) {
  g *= inv_scale;
  if (a.wd != 0.f) g = fmaf(a.wd, p, g);
  m = fmaf(a.b1, m, (1.f - a.b1) * g);
  v = fmaf(a.b2, v, (1.f - a.b2) * g * g);
  const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
  p -= a.step_size * (m / denom);
}

template <bool kAdam>
__global__ void __launch_bounds__(kOptThreads)
optim_multi_kernel(const long long* __restrict__ descs, int n, SgdArgs sa, AdamArgs aa, const float* __restrict__ grad_scale,
                   const float* __restrict__ found_inf) {
  if (found_inf != nullptr && __ldg(found_inf) != 0.f) return;          // GradScaler: an overflowed step is skipped
  const float inv_scale = grad_scale != nullptr ? 1.f / __ldg(grad_scale) : 1.f;
  const long long* d = opt_find(descs, n, (long long)blockIdx.x);
  float* p = reinterpret_cast<float*>(d[0]);
  const float* g = reinterpret_cast<const float*>(d[1]);
  float* s1 = reinterpret_cast<float*>(d[2]);
  float* s2 = reinterpret_cast<float*>(d[3]);
  const long long numel = d[4];
  const long long e0 = ((long long)blockIdx.x - d[5]) * kOptPerBlock;
  const long long e1 = e0 + kOptPerBlock < numel ? e0 + kOptPerBlock : numel;
  const bool vec = ((d[0] | d[1] | d[2] | (kAdam ? d[3] : 0)) & 15) == 0;
  if (vec) {
    const long long v1 = e0 + ((e1 - e0) & ~3ll);
    for (long long e = e0 + threadIdx.x * 4; e < v1; e += kOptThreads * 4) {
      float4 pv = *reinterpret_cast<float4*>(p + e);
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g + e));
      float4 m = *reinterpret_cast<float4*>(s1 + e);
      if constexpr (kAdam) {
        float4 v = *reinterpret_cast<float4*>(s2 + e);
        adam_one(pv.x, gv.x, m.x, v.x, aa, inv_scale); adam_one(pv.y, gv.y, m.y, v.y, aa, inv_scale);
        adam_one(pv.z, gv.z, m.z, v.z, aa, inv_scale); adam_one(pv.w, gv.w, m.w, v.w, aa, inv_scale);
        *reinterpret_cast<float4*>(s2 + e) = v;
      } else {
        sgd_one(pv.x, gv.x, m.x, sa, inv_scale); sgd_one(pv.y, gv.y, m.y, sa, inv_scale);
        sgd_one(pv.z, gv.z, m.z, sa, inv_scale); sgd_one(pv.w, gv.w, m.w, sa, inv_scale);
      }
      *reinterpret_cast<float4*>(s1 + e) = m;
      *reinterpret_cast<float4*>(p + e) = pv;
    }
    for (long long e = v1 + threadIdx.x; e < e1; e += kOptThreads) {
      if constexpr (kAdam) adam_one(p[e], g[e], s1[e], s2[e], aa, inv_scale);
      else sgd_one(p[e], g[e], s1[e], sa, inv_scale);
    }
  } else {
    for (long long e = e0 + threadIdx.x; e < e1; e += kOptThreads) {
      if constexpr (kAdam) adam_one(p[e], g[e], s1[e], s2[e], aa, inv_scale);
      else sgd_one(p[e], g[e], s1[e], sa, inv_scale);
    }
  }
}

}  // namespace dyk

using namespace dyk;
#define DYK_EXPORT extern "C" __attribute__((visibility("default")))

DYK_EXPORT int32_t dyk_optim_block_elems(void) { return kOptPerBlock; }

DYK_EXPORT int dyk_optim_sgd_multi(const int64_t* descs, int32_t n, int64_t total_blocks, float lr, float momentum, float dampening,
                                   float weight_decay, int32_t nesterov, int32_t first_step, const float* grad_scale,
                                   const float* found_inf, void* stream_) {
  DYK_REQUIRE(descs && n > 0 && total_blocks > 0 && total_blocks < (1ll << 31), "dyk_optim_sgd_multi: bad arguments");
  DYK_REQUIRE(!(nesterov && (momentum <= 0.f || dampening != 0.f)), "dyk_optim_sgd_multi: nesterov needs momentum > 0 and zero dampening");
  SgdArgs sa{lr, momentum, dampening, weight_decay, nesterov, first_step};
  AdamArgs aa{};
  optim_multi_kernel<false><<<(unsigned)total_blocks, kOptThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const long long*>(descs), n, sa, aa, grad_scale, found_inf);
  DYK_LAUNCH_OK("optim_multi_kernel<sgd>");
  return DYK_OK;
}

DYK_EXPORT int dyk_optim_adam_multi(const int64_t* descs, int32_t n, int64_t total_blocks, float lr, float beta1, float beta2,
                                    float eps, float weight_decay, int64_t step, const float* grad_scale, const float* found_inf,
                                    void* stream_) {
  DYK_REQUIRE(descs && n > 0 && total_blocks > 0 && total_blocks < (1ll << 31) && step >= 1, "dyk_optim_adam_multi: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  SgdArgs sa{};
  AdamArgs aa{(float)((double)lr / bc1), beta1, beta2, eps, weight_decay, (float)(1.0 / sqrt(bc2))};
  optim_multi_kernel<true><<<(unsigned)total_blocks, kOptThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const long long*>(descs), n, sa, aa, grad_scale, found_inf);
  DYK_LAUNCH_OK("optim_multi_kernel<adam>");
  return DYK_OK;
}
