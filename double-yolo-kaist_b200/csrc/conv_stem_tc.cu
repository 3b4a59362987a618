// Stem convolution (Cin <= 4, 3x3, stride 1, pad 1 -> 32 channels) on the tensor cores, reading the caller's NCHW frames.
//
// Replaces the first nn.Conv2d + BatchNorm2d + activation of each modality (reference models.py:35-36, in_channels = 3,
// also at second_index) together with the callers' `imgs.float() / 255.0` (train_utils/kaist_train_eval_utils.py:54-55,
// evaluate.py:67-68) and the NCHW -> NHWC change.
//
// Why: the CUDA-core stem (conv_direct.cu) needs one shared-memory weight read per FMA and takes ~0.55 ms per modality
// at 16 x 512 x 640 — 15 % of the whole inference step for 0.2 % of its FLOPs, 9x the HBM time of this layer
// (profiles/r01_launches_bench_summary.txt).  Here the CUDA cores only gather: thread t of a 128-thread CTA builds the
// K = 27 (padded to 32) im2col row of output pixel wo0 + t in registers (coalesced reads of the three image planes),
// rounds it to the 16-bit compute type and writes it as one 64-byte row of the canonical K-major SWIZZLE_64B operand
// layout; two tcgen05.mma (128 x 32 x 16) against the 32 x 32 weight tile produce the tile in TMEM; the epilogue
// (scale / bias / activation) stores 64 contiguous bytes per pixel.  Many small CTAs per SM (8 KB of shared memory and
// 32 TMEM columns each) hide the gather latency; nothing is pipelined inside a CTA.
//
// Numerics: frames and stem weights are rounded to fp16 / bf16 here (the CUDA-core kernel kept them in fp32) — the same
// rounding every later layer applies to its operands; oracle/darknet_ref.py models it for the parity tests.
#include "common.h"
#include "ptx.cuh"
#include "act.cuh"
#include "conv_common.cuh"
#include "frame_sample.cuh"

namespace dyk {

template <bool kBf16>
__device__ __forceinline__ uint32_t pack2h(float a, float b) { return pack2<kBf16>(a, b); }

template <bool kBf16, typename TIn>
__global__ void __launch_bounds__(128, 8)
stem_tc_kernel(const TIn* __restrict__ x, const float* __restrict__ w /* [32][3][3][Cin] fp32 */,
               const float* __restrict__ scale, const float* __restrict__ bias, uint8_t* __restrict__ y, long long ys,
               int N, int H, int W, int Cin, int act, int tiles_w, long long num_tiles, int Hs, int Ws, float rs_h, float rs_w) {
  // Hs > 0: x holds Hs x Ws frames and the convolution runs over their bilinear resize to H x W (frame_sample.cuh)
  __shared__ __align__(1024) uint8_t sa[128 * 64];     // A: 128 pixels x 32 K (16-bit), SWIZZLE_64B
  __shared__ __align__(1024) uint8_t sb[32 * 64];      // B: 32 out channels x 32 K
  __shared__ uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float ssc[32], sbi[32];
  const int t = threadIdx.x, warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // (warp-uniform for the compiler)
  constexpr int kCin = 3;                              // the only stem shape routed here (K = 27 padded to 32)
  constexpr int KK = 9 * kCin;
  (void)Cin;

  // uint8 frames: v * (1/255) instead of the IEEE division v / 255 — the two differ in fp32 for 126 byte values but are
  // identical for all 256 after the rounding to fp16 / bf16 that follows (checked exhaustively), so the 256-entry table
  // of the first version (27 bank-conflicting shared-memory lookups per pixel) is not needed.
  constexpr float kInv255 = 1.0f / 255.0f;
  // weights -> sb (row = out channel, 64 B per row, 16-byte chunk c stored at c ^ ((row >> 1) & 3))
  if (t < 32) {
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k0 = 2 * j, k1 = 2 * j + 1;            // K index = (r*3+s)*Cin + ci == flat index into w[co][...]
      const float a = k0 < KK ? __ldg(&w[t * KK + k0]) : 0.f;
      const float b = k1 < KK ? __ldg(&w[t * KK + k1]) : 0.f;
      pk[j] = pack2h<kBf16>(a, b);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(sb + t * 64 + ((c ^ ((t >> 1) & 3)) * 16)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
  }
  if (t >= 32 && t < 64) {
    ssc[t - 32] = scale ? __ldg(&scale[t - 32]) : 1.f;
    sbi[t - 32] = bias ? __ldg(&bias[t - 32]) : 0.f;
  }
  if (t == 0) {
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<32>(&tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tw = (int)(tile % tiles_w);
    const int ho = (int)((tile / tiles_w) % H);
    const int n = (int)(tile / ((long long)tiles_w * H));
    const int wo = tw * 128 + t;
    // ---- gather the im2col row of this thread's pixel
    float v[32];
#pragma unroll
    for (int k = 27; k < 32; ++k) v[k] = 0.f;
    auto cvt = [&](TIn raw) -> float {
      if constexpr (sizeof(TIn) == 1) return (float)raw * kInv255;
      else return (float)raw;
    };
    const long long plane = (long long)H * W;
    const TIn* xb = x + ((long long)n * kCin * H + (ho - 1)) * W + (wo - 1);     // top-left tap of channel 0
    if (Hs > 0) {                                                                // multi-scale: sample the resized frame
#pragma unroll
      for (int k = 0; k < 27; ++k) v[k] = 0.f;
      if (wo < W) {
        const TIn* xs = x + (long long)n * kCin * Hs * Ws;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int h = ho - 1 + r;
          if (h < 0 || h >= H) continue;
          const ResizeAxis ah = resize_axis(h, rs_h, Hs);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int ww = wo - 1 + q;
            if (ww < 0 || ww >= W) continue;
            const ResizeAxis aw = resize_axis(ww, rs_w, Ws);
#pragma unroll
            for (int ci = 0; ci < kCin; ++ci)
              v[(r * 3 + q) * kCin + ci] = frame_bilinear(xs + (long long)ci * Hs * Ws, Ws, ah, aw);
          }
        }
      }
    } else if (ho >= 1 && ho + 1 < H && wo >= 1 && wo + 1 < W) {                 // interior pixel: no bounds checks
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int ci = 0; ci < kCin; ++ci) {
          const TIn* row = xb + ci * plane + (long long)r * W;
#pragma unroll
          for (int q = 0; q < 3; ++q) v[(r * 3 + q) * kCin + ci] = cvt(__ldg(row + q));
        }
    } else {
#pragma unroll
      for (int k = 0; k < 27; ++k) v[k] = 0.f;
      if (wo < W) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int h = ho - 1 + r;
          if (h < 0 || h >= H) continue;
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int ww = wo - 1 + q;
            if (ww < 0 || ww >= W) continue;
#pragma unroll
            for (int ci = 0; ci < kCin; ++ci) v[(r * 3 + q) * kCin + ci] = cvt(__ldg(xb + ci * plane + (long long)r * W + q));
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 q = make_uint4(pack2h<kBf16>(v[8 * c + 0], v[8 * c + 1]), pack2h<kBf16>(v[8 * c + 2], v[8 * c + 3]),
                                 pack2h<kBf16>(v[8 * c + 4], v[8 * c + 5]), pack2h<kBf16>(v[8 * c + 6], v[8 * c + 7]));
      *reinterpret_cast<uint4*>(sa + t * 64 + ((c ^ ((t >> 1) & 3)) * 16)) = q;
    }
    fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (t == 0) {
      tc_fence_after_sync();
      constexpr uint32_t idesc = umma_idesc_f16(128, 32, kBf16 ? 1 : 0);
      const uint64_t ad = umma_desc_kmajor<64>(smem_u32(sa));
      const uint64_t bd = umma_desc_kmajor<64>(smem_u32(sb));
      umma_f16_ss(tmem, ad, bd, idesc, 0u);
      umma_f16_ss(tmem, ad + 2, bd + 2, idesc, 1u);
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after_sync();
    uint32_t acc[32];
    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), acc);
    tmem_ld_wait();
    tc_fence_before_sync();
    if (wo < W) {
      const long long pix = ((long long)n * H + ho) * W + wo;
      uint4* yp = reinterpret_cast<uint4*>(y + pix * ys * 2);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float o[8];
        const float4 s0 = *reinterpret_cast<const float4*>(ssc + 8 * c), s1 = *reinterpret_cast<const float4*>(ssc + 8 * c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sbi + 8 * c), b1 = *reinterpret_cast<const float4*>(sbi + 8 * c + 4);
        const float scv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float biv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = apply_act(fmaf(__uint_as_float(acc[8 * c + q]), scv[q], biv[q]), act);
        yp[c] = make_uint4(pack2h<kBf16>(o[0], o[1]), pack2h<kBf16>(o[2], o[3]), pack2h<kBf16>(o[4], o[5]), pack2h<kBf16>(o[6], o[7]));
      }
    }
    __syncthreads();                   // sa and the accumulator are free for the next tile
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}

// Returns DYK_OK after launching, 1 when the shape is not covered (the caller falls back to the CUDA-core kernel).
int stem_tc_try(const void* x, const float* w, const float* scale, const float* bias, void* y, int64_t ys, int N, int H,
                int W, int Cin, int Cout, int k, int stride, int pad, int act, int dtype, int x_kind, cudaStream_t stream,
                int Hs, int Ws) {
  if (!(k == 3 && stride == 1 && pad == 1 && Cout == 32 && Cin == 3)) return 1;
  const float rs_h = Hs > 0 ? (float)Hs / (float)H : 0.f, rs_w = Hs > 0 ? (float)Ws / (float)W : 0.f;
  const int tiles_w = ceil_div(W, 128);
  const long long num_tiles = (long long)N * H * tiles_w;
  long long grid = num_tiles;
  const long long cap = (long long)num_sms() * 16;
  if (grid > cap) grid = cap;
#define DYK_STEM_TC(BF, TIN)                                                                                             \
  stem_tc_kernel<BF, TIN><<<(unsigned)grid, 128, 0, stream>>>(static_cast<const TIN*>(x), w, scale, bias, (uint8_t*)y, ys, N, \
                                                              H, W, Cin, act, tiles_w, num_tiles, Hs, Ws, rs_h, rs_w)
  if (dtype == DYK_BF16) {
    if (x_kind == 0) DYK_STEM_TC(true, float); else DYK_STEM_TC(true, uint8_t);
  } else {
    if (x_kind == 0) DYK_STEM_TC(false, float); else DYK_STEM_TC(false, uint8_t);
  }
#undef DYK_STEM_TC
  DYK_LAUNCH_OK("stem_tc_kernel");
  return DYK_OK;
}

}  // namespace dyk
