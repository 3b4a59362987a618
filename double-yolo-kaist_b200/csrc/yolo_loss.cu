// Training loss of the three YOLO heads with its gradient, fused: target building, box (GIoU / CIoU) loss, IoU-aware
// objectness BCE and class BCE, forward and backward in one pass per head, everything on the device (no host sync: the
// number of matched targets only exists in device memory).
//
// Replaces, for the reference (the step that follows the forward in every training iteration):
//   build_utils/utils.py:305-384  build_targets  (wh-IoU anchor matching, positives in anchor-major order)
//   build_utils/utils.py:209-302  compute_loss   (dozens of gather / scatter / elementwise launches + autograd)
//   build_utils/utils.py:95-138   bbox_iou       (xywh, GIoU and CIoU)     build_utils/utils.py:166-172 wh_iou
//
// Determinism: no floating-point atomics.  Several labels can match the same (image, anchor, cell): the reference's
// `tobj[b, a, gj, gi] = ...` keeps the LAST of them and autograd's index backward SUMS their gradients; here the first
// positive of a cell ("owner") walks its duplicates in index order and does both.
#include "common.h"
#include <cstdint>

namespace dyk {

constexpr int kLossThreads = 1024;
constexpr int kMaxNo = 96;   // 5 + up to 91 classes per anchor

// ------------------------------------------------------------------ build_targets
// One block.  Candidate (a, t) in anchor-major order is a positive when wh_iou(anchor a, target t on this grid) > iou_t
// (same fp32 operation order as the reference so the threshold decisions are identical).
__global__ void __launch_bounds__(kLossThreads)
loss_match_kernel(const float* __restrict__ targets, int nt, const float* __restrict__ anchors, int na, int ny, int nx,
                  float iou_t, int* __restrict__ count, int* __restrict__ idx, float* __restrict__ tbox, int* __restrict__ tcls) {
  __shared__ int warp_tot[kLossThreads / 32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int total = na * nt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < total; c0 += kLossThreads) {
    const int c = c0 + threadIdx.x;
    bool keep = false;
    int a = 0, t = 0;
    float gx = 0.f, gy = 0.f, gw = 0.f, gh = 0.f;
    if (c < total) {
      a = c / nt;
      t = c - a * nt;
      const float* tt = targets + (long long)t * 6;
      gx = __fmul_rn(tt[2], (float)nx); gy = __fmul_rn(tt[3], (float)ny);
      gw = __fmul_rn(tt[4], (float)nx); gh = __fmul_rn(tt[5], (float)ny);
      const float aw = anchors[a * 2], ah = anchors[a * 2 + 1];
      const float inter = __fmul_rn(fminf(aw, gw), fminf(ah, gh));
      const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(__fmul_rn(aw, ah), __fmul_rn(gw, gh)), inter));
      keep = iou > iou_t;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int before = base_s;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (keep) {
      const int k = before + __popc(m & ((1u << lane) - 1u));
      const float* tt = targets + (long long)t * 6;
      const int gi = (int)gx, gj = (int)gy;           // .long(): truncation towards zero
      idx[k * 4 + 0] = (int)tt[0];
      idx[k * 4 + 1] = a;
      idx[k * 4 + 2] = gj;
      idx[k * 4 + 3] = gi;
      tbox[k * 4 + 0] = __fsub_rn(gx, (float)gi);
      tbox[k * 4 + 1] = __fsub_rn(gy, (float)gj);
      tbox[k * 4 + 2] = gw;
      tbox[k * 4 + 3] = gh;
      tcls[k] = (int)tt[1];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = base_s;
      for (int w = 0; w < kLossThreads / 32; ++w) s += warp_tot[w];
      base_s = s;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base_s;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// BCEWithLogits element and its derivative, torch's stable form:
//   l = (1 - t) x + w (log1p(exp(-|x|)) + max(-x, 0)),  w = 1 + (pos_weight - 1) t ;  dl/dx = (1 - t) - w sigmoid(-x)
__device__ __forceinline__ float bce_logits(float x, float t, float pw, float* dldx) {
  const float w = 1.f + (pw - 1.f) * t;
  const float l = (1.f - t) * x + w * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f));
  *dldx = (1.f - t) - w * sigmoidf_(-x);
  return l;
}

// ------------------------------------------------------------------ positives: box + class loss, tobj, gradients
// One block.  ws layout (floats): gvec[max_pos][no] | tobjv[max_pos] | lbox[max_pos] | lcls[max_pos] | sums[2].
__global__ void __launch_bounds__(kLossThreads)
loss_pos_kernel(const float* __restrict__ p, float* __restrict__ dp, int B, int na, int ny, int nx, int no,
                const int* __restrict__ count, const int* __restrict__ idx, const float* __restrict__ tbox,
                const int* __restrict__ tcls, const float* __restrict__ anchors, int max_pos, int v4, int ciou, float gr,
                float cls_pw, float w_box, float w_cls, float* __restrict__ tobj, float* __restrict__ ws, int* __restrict__ err) {
  const int nb = *count;
  float* gvec = ws;
  float* tobjv = gvec + (long long)max_pos * no;
  float* lboxv = tobjv + max_pos;
  float* lclsv = lboxv + max_pos;
  float* sums = lclsv + max_pos;
  const int nc = no - 5;
  const float g_out = nb > 0 ? -w_box / (float)nb : 0.f;                 // d(w_box * mean(1 - iou)) / d iou_k
  const float g_cls = (nb > 0 && nc > 1) ? w_cls / ((float)nb * (float)nc) : 0.f;
  for (int k = threadIdx.x; k < nb; k += kLossThreads) {
    const int b = idx[k * 4], a = idx[k * 4 + 1], gj = idx[k * 4 + 2], gi = idx[k * 4 + 3];
    float* g = gvec + (long long)k * no;
    for (int j = 0; j < no; ++j) g[j] = 0.f;
    if (b < 0 || b >= B || gj < 0 || gj >= ny || gi < 0 || gi >= nx) {   // the reference would raise an IndexError
      atomicExch(err, 1);
      tobjv[k] = 0.f; lboxv[k] = 0.f; lclsv[k] = 0.f;
      continue;
    }
    const float* ps = p + ((((long long)b * na + a) * ny + gj) * nx + gi) * no;
    const float aw = anchors[a * 2], ah = anchors[a * 2 + 1];
    // predicted box and d(box)/d(logits)
    float px, py, pw, ph, dpx, dpy, dpw, dph;
    if (v4) {
      const float sx = sigmoidf_(ps[0]), sy = sigmoidf_(ps[1]), sw = sigmoidf_(ps[2]), sh = sigmoidf_(ps[3]);
      px = sx * 2.f - 0.5f; py = sy * 2.f - 0.5f;
      pw = (sw * 2.f) * (sw * 2.f) * aw; ph = (sh * 2.f) * (sh * 2.f) * ah;
      dpx = 2.f * sx * (1.f - sx); dpy = 2.f * sy * (1.f - sy);
      dpw = 8.f * sw * sw * (1.f - sw) * aw; dph = 8.f * sh * sh * (1.f - sh) * ah;
    } else {
      const float sx = sigmoidf_(ps[0]), sy = sigmoidf_(ps[1]);
      const float ew = expf(ps[2]), eh = expf(ps[3]);
      px = sx; py = sy;
      pw = fminf(ew, 1e3f) * aw; ph = fminf(eh, 1e3f) * ah;
      dpx = sx * (1.f - sx); dpy = sy * (1.f - sy);
      dpw = ew <= 1e3f ? ew * aw : 0.f; dph = eh <= 1e3f ? eh * ah : 0.f;
    }
    const float tx = tbox[k * 4], ty = tbox[k * 4 + 1], tw = tbox[k * 4 + 2], th = tbox[k * 4 + 3];
    const float b1x1 = px - pw / 2.f, b1x2 = px + pw / 2.f, b1y1 = py - ph / 2.f, b1y2 = py + ph / 2.f;
    const float b2x1 = tx - tw / 2.f, b2x2 = tx + tw / 2.f, b2y1 = ty - th / 2.f, b2y2 = ty + th / 2.f;
    const float ix = fminf(b1x2, b2x2) - fmaxf(b1x1, b2x1), iy = fminf(b1y2, b2y2) - fmaxf(b1y1, b2y1);
    const float cix = fmaxf(ix, 0.f), ciy = fmaxf(iy, 0.f);
    const float inter = cix * ciy;
    const float w1 = b1x2 - b1x1, h1 = b1y2 - b1y1, w2 = b2x2 - b2x1, h2 = b2y2 - b2y1;
    const float uni = (w1 * h1 + 1e-16f) + w2 * h2 - inter;
    const float iou = inter / uni;
    const float cw = fmaxf(b1x2, b2x2) - fminf(b1x1, b2x1), ch = fmaxf(b1y2, b2y2) - fminf(b1y1, b2y1);
    float out, g_iou = g_out, g_uni = 0.f, g_cw, g_ch, g_w1 = 0.f, g_h1 = 0.f, g_cx = 0.f, g_cy = 0.f;
    if (!ciou) {
      const float c_area = cw * ch + 1e-16f;
      out = iou - (c_area - uni) / c_area;
      g_uni = g_out / c_area;
      const float g_ca = -g_out * uni / (c_area * c_area);
      g_cw = g_ca * ch; g_ch = g_ca * cw;
    } else {
      const float c2 = cw * cw + ch * ch + 1e-16f;
      const float dx = (b2x1 + b2x2) - (b1x1 + b1x2), dy = (b2y1 + b2y2) - (b1y1 + b1y2);
      const float rho2 = dx * dx / 4.f + dy * dy / 4.f;
      const float kv = 4.f / (3.14159265358979323846f * 3.14159265358979323846f);
      const float r1 = w1 / h1;
      const float da = atanf(w2 / h2) - atanf(r1);
      const float v = kv * da * da;
      const float alpha = v / (1.f - iou + v);                           // no_grad in the reference
      out = iou - (rho2 / c2 + v * alpha);
      const float g_c2 = g_out * rho2 / (c2 * c2);
      g_cw = g_c2 * 2.f * cw; g_ch = g_c2 * 2.f * ch;
      const float g_rho2 = -g_out / c2;
      g_cx = -g_rho2 * dx / 2.f;      // d rho2 / d(b1x1 + b1x2) = -dx / 2  (applied to both corners below)
      g_cy = -g_rho2 * dy / 2.f;
      const float g_a1 = -(-g_out * alpha) * kv * 2.f * da;             // g_v * dv/dA1, dv/dA1 = -2 kv da, g_v = -g_out alpha
      const float g_r1 = g_a1 / (1.f + r1 * r1);
      g_w1 = g_r1 / h1;
      g_h1 = -g_r1 * w1 / (h1 * h1);
    }
    // iou = inter / uni ; uni = w1 h1 + eps + w2 h2 - inter
    g_uni += g_iou * (-inter / (uni * uni));
    const float g_inter = g_iou / uni - g_uni;
    g_w1 += g_uni * h1;
    g_h1 += g_uni * w1;
    const float g_ix = ix >= 0.f ? g_inter * ciy : 0.f, g_iy = iy >= 0.f ? g_inter * cix : 0.f;
    float gx1 = g_cx, gx2 = g_cx, gy1 = g_cy, gy2 = g_cy;               // grads of b1x1, b1x2, b1y1, b1y2
    if (b1x2 < b2x2) gx2 += g_ix;
    if (b1x1 > b2x1) gx1 -= g_ix;
    if (b1y2 < b2y2) gy2 += g_iy;
    if (b1y1 > b2y1) gy1 -= g_iy;
    if (b1x2 > b2x2) gx2 += g_cw;
    if (b1x1 < b2x1) gx1 -= g_cw;
    if (b1y2 > b2y2) gy2 += g_ch;
    if (b1y1 < b2y1) gy1 -= g_ch;
    gx2 += g_w1; gx1 -= g_w1;
    gy2 += g_h1; gy1 -= g_h1;
    g[0] = (gx1 + gx2) * dpx;
    g[1] = (gy1 + gy2) * dpy;
    g[2] = (gx2 - gx1) * 0.5f * dpw;
    g[3] = (gy2 - gy1) * 0.5f * dph;
    lboxv[k] = 1.f - out;
    tobjv[k] = (1.f - gr) + gr * fmaxf(out, 0.f);
    float lc = 0.f;
    if (nc > 1) {
      const int cls = tcls[k];
      if (cls < 0 || cls >= nc) atomicExch(err, 2);                      // the reference asserts c.max() < model.nc
      for (int j = 0; j < nc; ++j) {
        float d;
        lc += bce_logits(ps[5 + j], j == cls ? 1.f : 0.f, cls_pw, &d);
        g[5 + j] = d * g_cls;
      }
    }
    lclsv[k] = lc;
  }
  __threadfence_block();
  __syncthreads();
  // owners: first positive of each (b, a, gj, gi); sum the gradients of its duplicates in order, keep the last tobj
  for (int k = threadIdx.x; k < nb; k += kLossThreads) {
    const int4 me = *reinterpret_cast<const int4*>(idx + k * 4);
    if (me.x < 0 || me.x >= B || me.z < 0 || me.z >= ny || me.w < 0 || me.w >= nx) continue;
    bool owner = true;
    for (int j = 0; j < k; ++j) {
      const int4 o = *reinterpret_cast<const int4*>(idx + j * 4);
      if (o.x == me.x && o.y == me.y && o.z == me.z && o.w == me.w) { owner = false; break; }
    }
    if (!owner) continue;
    const long long cell = (((long long)me.x * na + me.y) * ny + me.z) * nx + me.w;
    float acc[kMaxNo];
    for (int c = 0; c < no; ++c) acc[c] = gvec[(long long)k * no + c];
    float tv = tobjv[k];
    for (int j = k + 1; j < nb; ++j) {
      const int4 o = *reinterpret_cast<const int4*>(idx + j * 4);
      if (o.x == me.x && o.y == me.y && o.z == me.z && o.w == me.w) {
        for (int c = 0; c < no; ++c) acc[c] += gvec[(long long)j * no + c];
        tv = tobjv[j];
      }
    }
    tobj[cell] = tv;
    for (int c = 0; c < no; ++c)
      if (c != 4) dp[cell * no + c] = acc[c];
  }
  // loss sums in index order (one thread: nb is small and the order is fixed)
  if (threadIdx.x == 0) {
    float sb = 0.f, sc = 0.f;
    for (int k = 0; k < nb; ++k) { sb += lboxv[k]; sc += lclsv[k]; }
    sums[0] = nb > 0 ? sb / (float)nb : 0.f;
    sums[1] = (nb > 0 && nc > 1) ? sc / ((float)nb * (float)nc) : 0.f;
  }
}

// ------------------------------------------------------------------ objectness BCE over every cell
__global__ void __launch_bounds__(256)
loss_obj_kernel(const float* __restrict__ p, float* __restrict__ dp, long long cells, int no, const float* __restrict__ tobj,
                float obj_pw, float g_scale, float* __restrict__ part) {
  float s = 0.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < cells; i += (long long)gridDim.x * 256) {
    float d;
    s += bce_logits(p[i * no + 4], tobj[i], obj_pw, &d);
    dp[i * no + 4] = d * g_scale;
  }
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// acc[0..2] += (box mean, obj mean, cls mean) of this head; the last head also writes out = acc * weights
__global__ void loss_finish_kernel(const float* __restrict__ part, int blocks, long long cells, const float* __restrict__ sums,
                                   float* __restrict__ acc, int finalize, float w_box, float w_obj, float w_cls,
                                   float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < blocks; ++i) s += (double)part[i];
  acc[0] += sums[0];
  acc[1] += (float)(s / (double)cells);
  acc[2] += sums[1];
  if (finalize) {
    out[0] = acc[0] * w_box;
    out[1] = acc[1] * w_obj;
    out[2] = acc[2] * w_cls;
  }
}

// backward of the autograd node: dp_out[cell][c] = dp[cell][c] * upstream[group(c)], groups: box 0-3, obj 4, cls 5+
__global__ void loss_scale_grad_kernel(const float* __restrict__ dp, float* __restrict__ out, long long n, int no,
                                       const float* __restrict__ up) {
  const float u0 = up[0], u1 = up[1], u2 = up[2];
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % no);
    out[i] = dp[i] * (c < 4 ? u0 : (c == 4 ? u1 : u2));
  }
}

}  // namespace dyk

using namespace dyk;
#define DYK_EXPORT extern "C" __attribute__((visibility("default")))

DYK_EXPORT int dyk_yolo_build_targets(const float* targets, int32_t nt, const float* anchor_vec, int32_t na, int32_t ny,
                                      int32_t nx, float iou_t, int32_t* count, int32_t* idx, float* tbox, int32_t* tcls,
                                      void* stream_) {
  DYK_REQUIRE(anchor_vec && count && (nt == 0 || (targets && idx && tbox && tcls)), "dyk_yolo_build_targets: null pointer");
  DYK_REQUIRE(nt >= 0 && na > 0 && ny > 0 && nx > 0 && (long long)na * nt < (1ll << 24), "dyk_yolo_build_targets: bad shape");
  loss_match_kernel<<<1, kLossThreads, 0, static_cast<cudaStream_t>(stream_)>>>(targets, nt, anchor_vec, na, ny, nx, iou_t, count,
                                                                             idx, tbox, tcls);
  DYK_LAUNCH_OK("loss_match_kernel");
  return DYK_OK;
}

DYK_EXPORT int64_t dyk_yolo_loss_workspace_floats(int32_t max_pos, int32_t no, int64_t cells) {
  long long blocks = (cells + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  return (long long)max_pos * (no + 3) + 2 + cells /* tobj */ + blocks + 8;
}

DYK_EXPORT int dyk_yolo_loss_head(const float* p, float* dp, int32_t B, int32_t na, int32_t ny, int32_t nx, int32_t no,
                                  const int32_t* count, const int32_t* idx, const float* tbox, const int32_t* tcls,
                                  const float* anchor_vec, int32_t max_pos, int32_t v4, int32_t ciou, float gr, float obj_pw,
                                  float cls_pw, float w_box, float w_obj, float w_cls, float* acc, int32_t finalize,
                                  float* out, int32_t* err, float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(p && dp && count && anchor_vec && acc && out && err && workspace, "dyk_yolo_loss_head: null pointer");
  DYK_REQUIRE(max_pos == 0 || (idx && tbox && tcls), "dyk_yolo_loss_head: null target arrays");
  DYK_REQUIRE(B > 0 && na > 0 && ny > 0 && nx > 0 && no >= 5 && no <= kMaxNo, "dyk_yolo_loss_head: bad shape (no = %d)", no);
  const long long cells = (long long)B * na * ny * nx;
  int blocks = (int)((cells + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  float* posws = workspace;                                         // max_pos * (no + 3) + 2
  float* sums = posws + (long long)max_pos * (no + 3);
  float* tobj = sums + 2;
  float* part = tobj + cells;
  DYK_CUDA_OK(cudaMemsetAsync(dp, 0, (size_t)cells * no * sizeof(float), stream));
  DYK_CUDA_OK(cudaMemsetAsync(tobj, 0, (size_t)cells * sizeof(float), stream));
  loss_pos_kernel<<<1, kLossThreads, 0, stream>>>(p, dp, B, na, ny, nx, no, count, idx, tbox, tcls, anchor_vec, max_pos, v4, ciou,
                                                 gr, cls_pw, w_box, w_cls, tobj, posws, err);
  DYK_LAUNCH_OK("loss_pos_kernel");
  loss_obj_kernel<<<blocks, 256, 0, stream>>>(p, dp, cells, no, tobj, obj_pw, w_obj / (float)cells, part);
  DYK_LAUNCH_OK("loss_obj_kernel");
  loss_finish_kernel<<<1, 32, 0, stream>>>(part, blocks, cells, sums, acc, finalize, w_box, w_obj, w_cls, out);
  DYK_LAUNCH_OK("loss_finish_kernel");
  return DYK_OK;
}

DYK_EXPORT int dyk_yolo_loss_scale_grad(const float* dp, float* out, int64_t n, int32_t no, const float* upstream,
                                        void* stream_) {
  DYK_REQUIRE(dp && out && upstream && no >= 5, "dyk_yolo_loss_scale_grad: bad arguments");
  if (n == 0) return DYK_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  loss_scale_grad_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(dp, out, n, no, upstream);
  DYK_LAUNCH_OK("loss_scale_grad_kernel");
  return DYK_OK;
}
