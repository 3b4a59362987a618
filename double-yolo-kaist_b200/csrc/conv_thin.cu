// 1x1 convolutions with few input channels (the 16->16, 16->64, 24->72, 40->120 expand and 72->24 project layers at the
// top of the MobileNet backbones, models.py:28-64 with size=1) as a warp-level tensor-core kernel.
//
// The tcgen05 implicit GEMM (conv_tc.cu) spends one 128-pixel tile round trip (TMA load -> one MMA -> TMEM read -> TMA
// store, two accumulator stages) on 4 - 6 KB of input and fetches 32- or 48-byte pixel rows per TMA request:
// tools/layer_times.py on MobileNetV3-dual bs 64 shows these layers at 1.2 - 2.3 TB/s (16->16 @256x320: 288 us for a
// 52 us HBM floor; 24->72 @128x160: 198 us for 39 us).  A first CUDA-core version (one thread per pixel, fp32 weights
// broadcast from shared memory, FFMA2) fixed 16->16 (118 us) but is FMA-pipe bound beyond ~256 FMAs per pixel (24->72: 170 us,
// 64->24: 166 us against 62 us on the tcgen05 kernel) and was dropped for the kernel below.
#include "common.h"
#include <cstdlib>
#include "act.cuh"
#include "ptx.cuh"
#include "vec.cuh"
#include "conv_common.cuh"

namespace dyk {

struct ThinArgs {
  const uint8_t* x;
  long long xs;
  const uint8_t* w;      // [Cout][Cin] dtype
  const float* scale;
  const float* bias;
  uint8_t* y;
  long long ys;
  const uint8_t* res;
  long long rs;
  int H, W, Ho, Wo, Cin, Cout, stride;
  unsigned npix;         // N * Ho * Wo
};

// Warp-level tensor-core MMAs (mma.sync m16n8k16, fp32 accumulate): a warp owns 32 consecutive output pixels; every lane
// fetches one pixel's channels with 16-byte loads into a warp-private shared-memory tile (cp.async, double-buffered: the next group's pixels arrive while the current group is
// multiplied, finished and stored), ldmatrix turns the tile into A fragments, the weights sit in shared memory in
// B-fragment order (one conflict-free 8-byte read per lane per MMA pair), the epilogue (scale, bias, activation,
// residual in fp32, one rounding) writes the 16-bit results back into the warp's tile in fragment order and the lanes
// store their pixel's channels with 16-byte writes.  These layers are HBM-bound by two orders of magnitude in tensor
// throughput, so the legacy MMA path costs nothing; what matters is that no 128-pixel tile round trip
// (TMA -> tcgen05 -> TMEM -> TMA) is paid per 4 - 6 KB of input.
struct MmaArgs {
  ThinArgs t;
  int KS, NT;            // 16-channel K steps (Cin rounded up), 8-channel output tiles
  int a_pitch, o_pitch;  // bytes per row of the warp tile when it holds the input / the output
  int tile_bytes;        // per warp
  unsigned ngroups;      // ceil(npix / 32)
};

template <bool kBf16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (kBf16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool kBf16, int KS_MAX, int kAct, bool kRes>
__global__ void __launch_bounds__(256)
conv1x1_mma_kernel(const MmaArgs m) {
  const ThinArgs& a = m.t;
  extern __shared__ __align__(16) uint8_t mma_smem[];
  uint2* bsm = reinterpret_cast<uint2*>(mma_smem);                                   // [KS][NT][32 lanes] B fragments
  float* ssm = reinterpret_cast<float*>(mma_smem + (size_t)m.KS * m.NT * 256);       // scale [Cout], bias [Cout]
  uint8_t* tiles = reinterpret_cast<uint8_t*>(ssm + 2 * a.Cout);                     // [8 warps][tile_bytes]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;

  for (int i = threadIdx.x; i < m.KS * m.NT * 32; i += blockDim.x) {
    const int ln = i & 31, nt = (i >> 5) % m.NT, ks = (i >> 5) / m.NT;
    const int n = nt * 8 + (ln >> 2), k = ks * 16 + (ln & 3) * 2;
    uint2 v = make_uint2(0u, 0u);
    const uint32_t* wrow = reinterpret_cast<const uint32_t*>(a.w + ((long long)n * a.Cin) * 2);   // Cin % 8 == 0: 4-byte aligned pairs
    if (k < a.Cin) v.x = __ldg(wrow + (k >> 1));
    if (k + 8 < a.Cin) v.y = __ldg(wrow + ((k + 8) >> 1));
    bsm[i] = v;
  }
  for (int i = threadIdx.x; i < a.Cout; i += blockDim.x) {
    ssm[i] = a.scale ? __ldg(&a.scale[i]) : 1.f;
    ssm[a.Cout + i] = a.bias ? __ldg(&a.bias[i]) : 0.f;
  }
  uint8_t* tile0 = tiles + warp * 2 * m.tile_bytes;      // two tiles per warp: the next group's pixels arrive (cp.async)
  __syncthreads();                                        // while the current group is multiplied, finished and stored

  const int c8n = a.Cin >> 3, c8pad = m.KS * 2;
  const int KS = m.KS, NT = m.NT, Cout = a.Cout, o_pitch = m.o_pitch;
  // this lane's pixel of group `grp` -> row `lane` of `dst` (asynchronous; channels beyond Cin and pixels beyond the tensor
  // zero-filled)
  auto fetch = [&](unsigned grp, uint8_t* dst) {
    const unsigned pix = grp * 32 + lane;
    const bool ok = pix < a.npix;
    long long ipix = ok ? pix : 0;
    if (a.stride != 1 && ok) {
      const unsigned wo = pix % (unsigned)a.Wo, tt = pix / (unsigned)a.Wo;
      const unsigned ho = tt % (unsigned)a.Ho, n = tt / (unsigned)a.Ho;
      ipix = ((long long)n * a.H + ho * a.stride) * a.W + wo * a.stride;
    }
    const uint8_t* xp = a.x + ipix * a.xs * 2;
    const uint32_t row = smem_u32(dst + lane * m.a_pitch);
    for (int c = 0; c < c8n; ++c)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(row + c * 16), "l"(xp + c * 16), "r"(ok ? 16 : 0) : "memory");
    for (int c = c8n; c < c8pad; ++c) *reinterpret_cast<uint4*>(dst + lane * m.a_pitch + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const unsigned gstep = gridDim.x * 8;
  unsigned grp = blockIdx.x * 8 + warp;
  if (grp < m.ngroups) fetch(grp, tile0);
  for (int cur = 0; grp < m.ngroups; grp += gstep, cur ^= 1) {
    uint8_t* tile = tile0 + cur * m.tile_bytes;
    const unsigned pix = grp * 32 + lane;
    const bool ok = pix < a.npix;
    if (grp + gstep < m.ngroups) {
      fetch(grp + gstep, tile0 + (cur ^ 1) * m.tile_bytes);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    // ---- A fragments of both 16-row halves for every K step
    uint32_t af[2][KS_MAX][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < KS_MAX; ++ks) {
        if (ks < KS) {
          const uint32_t addr = smem_u32(tile + (mt * 16 + (lane & 15)) * m.a_pitch + (ks * 16 + (lane >> 4) * 8) * 2);
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(af[mt][ks][0]), "=r"(af[mt][ks][1]), "=r"(af[mt][ks][2]), "=r"(af[mt][ks][3]) : "r"(addr));
        }
      }
    __syncwarp();                                  // the tile is free: it now collects the output
    // per-lane row bases of the four (half, 8-row block) fragments, computed once per group: the output rows in the warp tile
    // and (kRes: compiled out otherwise — predicated-off residual code was half of the epilogue's issue slots) the residual rows
    uint8_t* orow[4];
    const uint8_t* rrow[4];
    bool rok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      orow[k] = tile + (k * 8 + g) * o_pitch + t4 * 4;
      if constexpr (kRes) {
        const unsigned rp = grp * 32 + k * 8 + g;
        rok[k] = rp < a.npix;
        rrow[k] = a.res + ((long long)(rok[k] ? rp : 0) * a.rs) * 2 + t4 * 4;
      }
    }
    // ---- 4 output tiles (32 channels) at a time
    for (int n0 = 0; n0 < NT; n0 += 4) {
      float acc[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][j][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS_MAX; ++ks) {
        if (ks < KS) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n0 + j < NT) {
              const uint2 b = bsm[(ks * NT + n0 + j) * 32 + lane];
              mma16816<kBf16>(acc[0][j], af[0][ks], b.x, b.y);
              mma16816<kBf16>(acc[1][j], af[1][ks], b.x, b.y);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0 + j < NT) {
          const int cb = (n0 + j) * 16;            // byte offset of the tile's first channel in a 16-bit row
          const float2 sc = *reinterpret_cast<const float2*>(ssm + (n0 + j) * 8 + t4 * 2);
          const float2 bi = *reinterpret_cast<const float2*>(ssm + Cout + (n0 + j) * 8 + t4 * 2);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {          // rows g + 16*mt + 8*h
              float o0 = act_apply<kAct>(fmaf(acc[mt][j][2 * h], sc.x, bi.x));
              float o1 = act_apply<kAct>(fmaf(acc[mt][j][2 * h + 1], sc.y, bi.y));
              if constexpr (kRes) {
                if (rok[mt * 2 + h]) {
                  const float2 r = unpack2<kBf16>(__ldg(reinterpret_cast<const uint32_t*>(rrow[mt * 2 + h] + cb)));
                  o0 += r.x;
                  o1 += r.y;
                }
              }
              *reinterpret_cast<uint32_t*>(orow[mt * 2 + h] + cb) = pack2<kBf16>(o0, o1);
            }
        }
      }
    }
    __syncwarp();
    if (ok) {
      const uint4* row = reinterpret_cast<const uint4*>(tile + lane * o_pitch);
      uint4* yp = reinterpret_cast<uint4*>(a.y + (long long)pix * a.ys * 2);
#pragma unroll 4
      for (int c = 0; c < NT; ++c) yp[c] = row[c];
    }
    __syncwarp();                                  // before the group after next is fetched into this tile
  }
}

// returns DYK_OK after a launch, 1 when the layer is not handled here, < 0 on error
int conv1x1_thin_try(const dyk_conv_params* p, cudaStream_t stream) {
  const char* env = getenv("DYK_THIN");        // read per call: tests switch kernels inside one process
  if (env != nullptr && env[0] == '0') return 1;
  if (p->kh != 1 || p->kw != 1 || p->pad != 0 || p->x2 || p->w_image_stride || p->out_f32 || p->upsample2x || p->y_plane ||
      p->out_h || p->out_w || p->Cout_store != p->Cout || p->Cout % 8 != 0)
    return 1;
  // measured against conv_tc.cu on MobileNetV3-dual bs 64: ahead for Cin < 64 and for wide-in / narrow-out projections;
  // Cin == 64 (128-byte TMA rows) stays on the tcgen05 kernel (64->24 @128x160: 62 us there, 84 us here)
  const bool mma_ok = p->Cin <= 128 && p->Cout <= 128 && p->Cin % 64 != 0 && (p->Cin < 64 || p->Cout <= 48);
  if (!mma_ok) return 1;
  const int Ho = (p->H - 1) / p->stride + 1, Wo = (p->W - 1) / p->stride + 1;
  const long long npix = (long long)p->N * Ho * Wo;
  if (npix >= (1ll << 31) - 64) return 1;
  ThinArgs a;
  a.x = static_cast<const uint8_t*>(p->x); a.xs = p->x_pix_stride;
  a.w = static_cast<const uint8_t*>(p->w); a.scale = p->scale; a.bias = p->bias;
  a.y = static_cast<uint8_t*>(p->y); a.ys = p->y_pix_stride;
  a.res = static_cast<const uint8_t*>(p->res); a.rs = p->res_pix_stride;
  a.H = p->H; a.W = p->W; a.Ho = Ho; a.Wo = Wo; a.Cin = p->Cin; a.Cout = p->Cout; a.stride = p->stride;
  a.npix = (unsigned)npix;
  {
    MmaArgs m;
    m.t = a;
    m.KS = ceil_div(p->Cin, 16);
    m.NT = p->Cout / 8;
    // row pitches: an odd number of 16-byte units, so that ldmatrix rows, the lanes' 16-byte row accesses and the 4-byte
    // fragment writes all spread over the banks
    m.a_pitch = (m.KS * 2 + 1) * 16;
    m.o_pitch = ((m.NT + 1) | 1) * 16;
    m.tile_bytes = 32 * (m.a_pitch > m.o_pitch ? m.a_pitch : m.o_pitch);
    m.ngroups = (unsigned)((npix + 31) / 32);
    const size_t smem = (size_t)m.KS * m.NT * 256 + 2 * (size_t)p->Cout * sizeof(float) + 16 * (size_t)m.tile_bytes;
    long long grid = ceil_div64(m.ngroups, 8);
    const long long cap = (long long)num_sms() * 4;    // persistent-ish: the weight fragments are staged once per CTA
    if (grid > cap) grid = cap;
#define DYK_MMA(KSM)                                                                                                \
  DYK_DISPATCH_ACT(p->act, DYK_DISPATCH_DTYPE(p->dtype, {                                                            \
    auto kern = p->res ? conv1x1_mma_kernel<kBf16, KSM, kAct, true> : conv1x1_mma_kernel<kBf16, KSM, kAct, false>;   \
    static bool configured[2] = {false, false};                                                                    \
    if (!configured[p->res ? 1 : 0]) {                                                                             \
      DYK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));            \
      configured[p->res ? 1 : 0] = true;                                                                           \
    }                                                                                                              \
    kern<<<(unsigned)grid, 256, smem, stream>>>(m);                                                                \
  }))
    if (smem > 112 * 1024) return 1;
    if (m.KS <= 2) DYK_MMA(2);
    else if (m.KS <= 4) DYK_MMA(4);
    else DYK_MMA(8);
#undef DYK_MMA
    DYK_LAUNCH_OK("conv1x1_mma_kernel");
    return DYK_OK;
  }
}

}  // namespace dyk
