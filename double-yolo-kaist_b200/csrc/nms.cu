// Batched non-maximum suppression for all images of a batch in three launches.
//
// Replaces non_max_suppression (reference build_utils/utils.py:387-464) including xywh2xyxy
// (:50-57) and the torchvision.ops.nms call at :448 (stable descending score sort + greedy IoU
// suppression, strict '>' threshold) and the [:max_num] prefix at :449.
//
//   1. nms_filter_kernel   one CTA per image: confidence / width-height / class filters, score =
//                          obj*cls, order-preserving compaction into 64-bit sort keys
//                          key = (~ordered(score) << 32) | candidate_id   (ascending key order ==
//                          descending score, ties by ascending candidate id == stable sort).
//   2. nms_sort_kernel     one CTA per image: bitonic sort of the keys (shared-memory tiles of 4096
//                          keys; strides >= 4096 run through L2-resident global memory).
//   3. nms_suppress_kernel one CTA per image: greedy suppression in sorted order, 256 candidates per
//                          round, stopping as soon as max_num boxes are kept — only the first max_num
//                          survivors are ever observable through the reference API.
//
// All IoU arithmetic uses explicitly rounded fp32 operations (no FMA contraction) in the same order as
// torchvision's CPU kernel, so kept indices are bit-exact against it.
#include "common.h"
#include <cmath>

namespace dyk {

constexpr int kSortTile = 4096;
constexpr float kMaxWh = 4096.f;
constexpr float kMinWh = 2.f;

__device__ __forceinline__ uint32_t float_ordered(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_ordered(uint32_t o) {
  const uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(b);
}

// ------------------------------------------------------------------ 1. filter + compaction
__global__ void __launch_bounds__(1024)
nms_filter_kernel(const float* __restrict__ pred, int rows, int nc, float conf_thres, int multi_label,
                  unsigned long long classes_mask, unsigned long long* __restrict__ keys, int n_pad,
                  int* __restrict__ counts) {
  const int b = blockIdx.x;
  const int no = nc + 5;
  const float* P = pred + (long long)b * rows * no;
  unsigned long long* K = keys + (long long)b * n_pad;
  const int per_row = multi_label ? nc : 1;
  const long long ncand = (long long)rows * per_row;
  __shared__ int warp_tot[32];
  __shared__ int running;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < ncand; base += blockDim.x) {
    const long long cid = base + threadIdx.x;
    bool keep = false;
    float score = 0.f;
    unsigned int out_id = 0;
    if (cid < ncand) {
      const int row = (int)(cid / per_row);
      const int k = (int)(cid - (long long)row * per_row);
      const float* r = P + (long long)row * no;
      const float obj = r[4];
      const float w = r[2], h = r[3];
      // utils.py:408-409
      if (obj > conf_thres && w > kMinWh && w < kMaxWh && h > kMinWh && h < kMaxWh) {
        int label;
        if (multi_label) {
          label = k;
          score = __fmul_rn(r[5 + k], obj);       // utils.py:416
        } else {
          label = 0;
          score = __fmul_rn(r[5], obj);
          for (int c = 1; c < nc; ++c) {           // utils.py:426 max over classes (first maximum)
            const float s = __fmul_rn(r[5 + c], obj);
            if (s > score) { score = s; label = c; }
          }
        }
        keep = score > conf_thres;                 // utils.py:423 / :427
        if (keep && classes_mask != 0ull) keep = (label < 64) && ((classes_mask >> label) & 1ull);  // :430-431
        out_id = (unsigned int)row * (unsigned int)nc + (unsigned int)label;
      }
    }
    const unsigned int ballot = __ballot_sync(0xffffffffu, keep);
    const int prefix = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      const int v = warp_tot[i];
      if (i < warp) woff += v;
      tot += v;
    }
    const int start = running;
    if (keep) {
      const unsigned long long key =
          ((unsigned long long)(~float_ordered(score)) << 32) | (unsigned long long)out_id;
      K[start + woff + prefix] = key;
    }
    __syncthreads();
    if (threadIdx.x == 0) running = start + tot;
    __syncthreads();
  }
  const int count = running;
  if (threadIdx.x == 0) counts[b] = count;
  // pad to the next power of two (>= 2) for the bitonic network
  int n = 2;
  while (n < count) n <<= 1;
  for (int i = count + threadIdx.x; i < n; i += blockDim.x) K[i] = ~0ull;
}

// ------------------------------------------------------------------ 2. bitonic sort (ascending keys)
__device__ __forceinline__ void cmpswap(unsigned long long& a, unsigned long long& b, bool up) {
  if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

__global__ void __launch_bounds__(1024)
nms_sort_kernel(unsigned long long* keys, int n_pad, const int* __restrict__ counts) {
  __shared__ unsigned long long sm[kSortTile];
  const int b = blockIdx.x;
  unsigned long long* K = keys + (long long)b * n_pad;
  const int count = counts[b];
  if (count <= 1) return;
  int n = 2;
  while (n < count) n <<= 1;
  const int tile = n < kSortTile ? n : kSortTile;
  // phase A: fully sort each tile in shared memory (alternating directions as the network requires)
  for (int t0 = 0; t0 < n; t0 += tile) {
    for (int i = threadIdx.x; i < tile; i += blockDim.x) sm[i] = K[t0 + i];
    __syncthreads();
    for (int k = 2; k <= tile; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < (tile >> 1); i += blockDim.x) {
          const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
          const int hi = lo | j;
          const bool up = (((t0 + lo) & k) == 0);
          unsigned long long a = sm[lo], c = sm[hi];
          cmpswap(a, c, up);
          sm[lo] = a; sm[hi] = c;
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < tile; i += blockDim.x) K[t0 + i] = sm[i];
    __syncthreads();
  }
  // phase B: merge stages wider than a tile
  for (int k = tile << 1; k <= n; k <<= 1) {
    int j = k >> 1;
    for (; j >= tile; j >>= 1) {  // partners live in different tiles: exchange through global memory
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        const int hi = lo | j;
        const bool up = ((lo & k) == 0);
        unsigned long long a = K[lo], c = K[hi];
        cmpswap(a, c, up);
        K[lo] = a; K[hi] = c;
      }
      __syncthreads();
    }
    for (int t0 = 0; t0 < n; t0 += tile) {  // remaining strides fit in a tile
      for (int i = threadIdx.x; i < tile; i += blockDim.x) sm[i] = K[t0 + i];
      __syncthreads();
      for (int jj = tile >> 1; jj > 0; jj >>= 1) {
        for (int i = threadIdx.x; i < (tile >> 1); i += blockDim.x) {
          const int lo = ((i & ~(jj - 1)) << 1) | (i & (jj - 1));
          const int hi = lo | jj;
          const bool up = (((t0 + lo) & k) == 0);
          unsigned long long a = sm[lo], c = sm[hi];
          cmpswap(a, c, up);
          sm[lo] = a; sm[hi] = c;
        }
        __syncthreads();
      }
      for (int i = threadIdx.x; i < tile; i += blockDim.x) K[t0 + i] = sm[i];
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ 3. greedy suppression
struct Box { float x1, y1, x2, y2; };

__device__ __forceinline__ float box_area(const Box& a) {
  return __fmul_rn(__fsub_rn(a.x2, a.x1), __fsub_rn(a.y2, a.y1));
}
// torchvision CPU nms: inter / (area_i + area_j - inter) > thr, each operation rounded to fp32.
__device__ __forceinline__ bool iou_gt(const Box& a, float area_a, const Box& b, float area_b, float thr) {
  const float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1);
  const float xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return ovr > thr;
}

constexpr int kChunk = 256;
constexpr int kMaxKeep = 1024;

__global__ void __launch_bounds__(kChunk)
nms_suppress_kernel(const float* __restrict__ pred, int rows, int nc, const unsigned long long* __restrict__ keys,
                    int n_pad, const int* __restrict__ counts, float iou_thres, int agnostic, int max_num,
                    float* __restrict__ out, int* __restrict__ out_count) {
  extern __shared__ unsigned char smraw[];
  Box* kept = reinterpret_cast<Box*>(smraw);                          // [max_num] class-offset boxes
  float* kept_area = reinterpret_cast<float*>(kept + max_num);        // [max_num]
  __shared__ Box cbox[kChunk];
  __shared__ float carea[kChunk];
  __shared__ unsigned long long cmask[kChunk][kChunk / 64];
  __shared__ unsigned long long alive_words[kChunk / 64];
  __shared__ int slot_of[kChunk];   // output slot claimed by a chunk entry in this round, or -1
  __shared__ int nk_sh;

  const int b = blockIdx.x;
  const int no = nc + 5;
  const float* P = pred + (long long)b * rows * no;
  const unsigned long long* K = keys + (long long)b * n_pad;
  const int count = counts[b];
  float* O = out + (long long)b * max_num * 6;
  const int tid = threadIdx.x;
  if (tid == 0) nk_sh = 0;
  __syncthreads();
  int nk = 0;
  for (int base = 0; base < count && nk < max_num; base += kChunk) {
    const int i = base + tid;
    const bool valid = i < count;
    Box ub = {0.f, 0.f, 0.f, 0.f};   // un-offset box (output)
    Box ob = {0.f, 0.f, 0.f, 0.f};   // class-offset box (IoU), utils.py:446-447
    float area = 0.f, score = 0.f, labelf = 0.f;
    if (valid) {
      const unsigned long long key = K[i];
      const unsigned int id = (unsigned int)(key & 0xffffffffull);
      score = float_from_ordered(~(unsigned int)(key >> 32));
      const int row = (int)(id / (unsigned int)nc), label = (int)(id % (unsigned int)nc);
      labelf = (float)label;
      const float* r = P + (long long)row * no;
      const float cx = r[0], cy = r[1], hw = __fdiv_rn(r[2], 2.f), hh = __fdiv_rn(r[3], 2.f);
      ub.x1 = __fsub_rn(cx, hw); ub.y1 = __fsub_rn(cy, hh);     // utils.py:50-57
      ub.x2 = __fadd_rn(cx, hw); ub.y2 = __fadd_rn(cy, hh);
      const float off = agnostic ? 0.f : __fmul_rn(labelf, kMaxWh);
      ob.x1 = __fadd_rn(ub.x1, off); ob.y1 = __fadd_rn(ub.y1, off);
      ob.x2 = __fadd_rn(ub.x2, off); ob.y2 = __fadd_rn(ub.y2, off);
      area = box_area(ob);
    }
    bool alive = valid;
    for (int j = 0; j < nk && alive; ++j)
      if (iou_gt(kept[j], kept_area[j], ob, area, iou_thres)) alive = false;
    cbox[tid] = ob;
    carea[tid] = area;
    slot_of[tid] = -1;
    const unsigned int bal = __ballot_sync(0xffffffffu, alive);
    if ((tid & 31) == 0) reinterpret_cast<unsigned int*>(alive_words)[tid >> 5] = bal;
    __syncthreads();
    // suppression row of this candidate against later candidates of the same chunk
#pragma unroll
    for (int wd = 0; wd < kChunk / 64; ++wd) {
      unsigned long long m = 0ull;
      if (alive) {
        const int t0 = wd * 64;
        for (int t = (tid + 1 > t0 ? tid + 1 : t0); t < t0 + 64; ++t) {
          if (base + t >= count) break;
          if (iou_gt(ob, area, cbox[t], carea[t], iou_thres)) m |= (1ull << (t - t0));
        }
      }
      cmask[tid][wd] = m;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long removed[kChunk / 64];
#pragma unroll
      for (int wd = 0; wd < kChunk / 64; ++wd) removed[wd] = 0ull;
      int n_now = nk;
      for (int wd = 0; wd < kChunk / 64 && n_now < max_num; ++wd) {
        unsigned long long avail = alive_words[wd] & ~removed[wd];
        while (avail && n_now < max_num) {
          const int bit = __ffsll((long long)avail) - 1;
          const int t = wd * 64 + bit;
          kept[n_now] = cbox[t];
          kept_area[n_now] = carea[t];
          slot_of[t] = n_now;  // the owning thread emits the output row
#pragma unroll
          for (int w2 = 0; w2 < kChunk / 64; ++w2) removed[w2] |= cmask[t][w2];
          ++n_now;
          const unsigned long long higher = (bit == 63) ? 0ull : (~0ull << (bit + 1));
          avail = alive_words[wd] & ~removed[wd] & higher;
        }
      }
      nk_sh = n_now;
    }
    __syncthreads();
    const int slot = slot_of[tid];
    if (valid && slot >= 0) {
      float* o = O + slot * 6;
      o[0] = ub.x1; o[1] = ub.y1; o[2] = ub.x2; o[3] = ub.y2; o[4] = score; o[5] = labelf;
    }
    nk = nk_sh;
    __syncthreads();
  }
  if (tid == 0) out_count[b] = nk;
}

static int next_pow2(long long v) {
  long long n = 2;
  while (n < v) n <<= 1;
  return (int)n;
}

}  // namespace dyk

using namespace dyk;

extern "C" __attribute__((visibility("default"))) int64_t dyk_nms_workspace_bytes(int32_t B, int32_t rows, int32_t nc, int32_t multi_label) {
  if (B <= 0 || rows <= 0 || nc <= 0) return 0;
  const long long ncand = (long long)rows * ((multi_label && nc > 1) ? nc : 1);
  const long long n_pad = next_pow2(ncand);
  return (int64_t)B * n_pad * 8 + (int64_t)B * 4 + 256;
}

extern "C" __attribute__((visibility("default"))) int dyk_nms_batched(const float* pred, int32_t B, int32_t rows, int32_t nc, float conf_thres,
                               double iou_thres_d, int32_t multi_label, uint64_t classes_mask, int32_t agnostic,
                               int32_t max_num, float* out, int32_t* out_count, void* workspace,
                               int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(pred && out && out_count && workspace, "dyk_nms_batched: null pointer");
  DYK_REQUIRE(B > 0 && rows > 0 && nc > 0, "dyk_nms_batched: B=%d rows=%d nc=%d", B, rows, nc);
  DYK_REQUIRE(max_num > 0 && max_num <= kMaxKeep, "dyk_nms_batched: max_num=%d (1..%d)", max_num, kMaxKeep);
  multi_label = (multi_label && nc > 1) ? 1 : 0;  // utils.py:404
  // torchvision's CPU kernel evaluates `float iou > double threshold`.  For fp32 x that is equivalent to
  // x > t with t the largest fp32 value <= threshold, so the device compares in fp32 against t.
  float iou_thres = static_cast<float>(iou_thres_d);
  if (static_cast<double>(iou_thres) > iou_thres_d) iou_thres = nextafterf(iou_thres, -INFINITY);
  DYK_REQUIRE((long long)rows * nc < (1ll << 31), "dyk_nms_batched: too many candidates");
  const int64_t need = dyk_nms_workspace_bytes(B, rows, nc, multi_label);
  DYK_REQUIRE(workspace_bytes >= need, "dyk_nms_batched: workspace %lld < %lld bytes", (long long)workspace_bytes,
              (long long)need);
  DYK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "dyk_nms_batched: workspace must be 8-byte aligned");
  const long long ncand = (long long)rows * (multi_label ? nc : 1);
  const int n_pad = next_pow2(ncand);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(workspace);
  int* counts = reinterpret_cast<int*>(keys + (long long)B * n_pad);

  nms_filter_kernel<<<B, 1024, 0, stream>>>(pred, rows, nc, conf_thres, multi_label, classes_mask, keys, n_pad, counts);
  DYK_LAUNCH_OK("nms_filter_kernel");
  nms_sort_kernel<<<B, 1024, 0, stream>>>(keys, n_pad, counts);
  DYK_LAUNCH_OK("nms_sort_kernel");
  const size_t dyn = (size_t)max_num * (sizeof(Box) + sizeof(float));
  nms_suppress_kernel<<<B, kChunk, dyn, stream>>>(pred, rows, nc, keys, n_pad, counts, iou_thres, agnostic, max_num, out,
                                                  out_count);
  DYK_LAUNCH_OK("nms_suppress_kernel");
  return DYK_OK;
}
