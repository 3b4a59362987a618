// YOLO head post-processing: the (bs, na*no, ny, nx) -> (bs, na, ny, nx, no) permute and, in eval mode,
// the box decode, for one head per launch, writing straight into the concatenated prediction tensor.
//
// Replaces YOLOLayer.forward (reference models.py:218-258) and the torch.cat of the three heads at
// models.py:315.  Arithmetic order follows the reference so fp32 results agree to rounding:
//   v3 (models.py:243-246): xy = (sigmoid(t) + grid) * stride ; wh = (exp(t) * anchor_vec) * stride
//   v4 (models.py:249-252): s = sigmoid(t); xy = (s*2 - 0.5 + grid) * stride ; wh = ((s*2)^2 * anchor_vec) * stride
#include "common.h"
#include "vec.cuh"

namespace dyk {

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// in_kind: 0 = fp16, 1 = bf16, 2 = fp32 head tensor
template <int kIn>
__device__ __forceinline__ float load_head(const void* p, long long idx) {
  if constexpr (kIn == 2) return reinterpret_cast<const float*>(p)[idx];
  else return load1<kIn == 1>(p, idx);
}

template <int kIn>
__global__ void yolo_decode_kernel(const void* __restrict__ p, long long ps, float* __restrict__ p_out,
                                   float* __restrict__ io_out, int N, int ny, int nx, int na, int no,
                                   const float* __restrict__ anchor_vec, float stride, int v4,
                                   long long rows_total, long long row_off) {
  const long long total = (long long)N * na * ny * nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int gx = (int)(i % nx);
    long long t = i / nx;
    const int gy = (int)(t % ny); t /= ny;
    const int a = (int)(t % na);
    const int n = (int)(t / na);
    const long long pix = ((long long)n * ny + gy) * nx + gx;
    const long long src = pix * ps + (long long)a * no;
    float* po = p_out + i * no;                 // [n][a][gy][gx][:]
    float* io = io_out ? io_out + (((long long)n * rows_total + row_off + ((long long)a * ny + gy) * nx + gx) * no)
                       : nullptr;
    const float aw = __ldg(&anchor_vec[a * 2]), ah = __ldg(&anchor_vec[a * 2 + 1]);
    for (int o = 0; o < no; ++o) {
      const float v = load_head<kIn>(p, src + o);
      po[o] = v;
      if (io) {
        float r;
        if (v4) {
          const float s = sigmoid_f(v);
          if (o < 2) r = __fmul_rn(__fadd_rn(__fsub_rn(__fmul_rn(s, 2.f), 0.5f), (o == 0 ? (float)gx : (float)gy)), stride);
          else if (o < 4) { const float d = __fmul_rn(s, 2.f); r = __fmul_rn(__fmul_rn(__fmul_rn(d, d), (o == 2 ? aw : ah)), stride); }
          else r = s;
        } else {
          if (o < 2) r = __fmul_rn(__fadd_rn(sigmoid_f(v), (o == 0 ? (float)gx : (float)gy)), stride);
          else if (o < 4) r = __fmul_rn(__fmul_rn(expf(v), (o == 2 ? aw : ah)), stride);
          else r = sigmoid_f(v);
        }
        io[o] = r;
      }
    }
  }
}

}  // namespace dyk

using namespace dyk;

extern "C" __attribute__((visibility("default"))) int dyk_yolo_decode(const void* p, int64_t ps, float* p_out, float* io_out, int32_t N, int32_t ny,
                               int32_t nx, int32_t na, int32_t no, const float* anchor_vec, float stride, int32_t v4,
                               int64_t rows_total, int64_t row_off, int32_t in_kind, void* stream_) {
  DYK_REQUIRE(p && p_out && anchor_vec, "dyk_yolo_decode: null pointer");
  DYK_REQUIRE(N > 0 && ny > 0 && nx > 0 && na > 0 && no >= 5 && ps >= (int64_t)na * no, "dyk_yolo_decode: bad shape");
  DYK_REQUIRE(in_kind >= 0 && in_kind <= 2, "dyk_yolo_decode: in_kind=%d", in_kind);
  if (io_out) DYK_REQUIRE(row_off >= 0 && row_off + (int64_t)na * ny * nx <= rows_total, "dyk_yolo_decode: row range");
  const long long total = (long long)N * na * ny * nx;
  long long g = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  switch (in_kind) {
    case 0: yolo_decode_kernel<0><<<(unsigned)g, 256, 0, stream>>>(p, ps, p_out, io_out, N, ny, nx, na, no, anchor_vec, stride, v4, rows_total, row_off); break;
    case 1: yolo_decode_kernel<1><<<(unsigned)g, 256, 0, stream>>>(p, ps, p_out, io_out, N, ny, nx, na, no, anchor_vec, stride, v4, rows_total, row_off); break;
    default: yolo_decode_kernel<2><<<(unsigned)g, 256, 0, stream>>>(p, ps, p_out, io_out, N, ny, nx, na, no, anchor_vec, stride, v4, rows_total, row_off); break;
  }
  DYK_LAUNCH_OK("yolo_decode_kernel");
  return DYK_OK;
}

// ---- post-NMS box rescaling: scale_coords + clip_coords (build_utils/utils.py:60-92), in place on n rows of xyxy boxes.
// Same fp32 operation order as the reference's in-place tensor ops: (x - pad) / gain, then clamp to [0, size].
namespace dyk {
__global__ void scale_coords_kernel(float* __restrict__ boxes, long long row_stride, int n, float pad_x, float pad_y, float gain,
                                    float img_w, float img_h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 4) return;
  const int r = i >> 2, c = i & 3;
  float* p = boxes + (long long)r * row_stride + c;
  const bool is_x = (c & 1) == 0;
  float v = __fdiv_rn(__fsub_rn(*p, is_x ? pad_x : pad_y), gain);
  v = fminf(fmaxf(v, 0.f), is_x ? img_w : img_h);
  *p = v;
}
}  // namespace dyk

extern "C" __attribute__((visibility("default"))) int dyk_scale_coords(float* boxes, int64_t row_stride, int32_t n, float pad_x,
                                                                        float pad_y, float gain, float img_w, float img_h,
                                                                        void* stream_) {
  DYK_REQUIRE(n == 0 || boxes, "dyk_scale_coords: null pointer");
  DYK_REQUIRE(n >= 0 && row_stride >= 4 && gain != 0.f, "dyk_scale_coords: bad arguments");
  if (n == 0) return DYK_OK;
  dyk::scale_coords_kernel<<<(n * 4 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(boxes, row_stride, n, pad_x, pad_y,
                                                                                            gain, img_w, img_h);
  DYK_LAUNCH_OK("scale_coords_kernel");
  return DYK_OK;
}
