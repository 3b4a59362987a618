// Depthwise 3x3 / 5x5 convolution (groups == C; models.py:41, build_utils/layers.py:224) with TMA-staged input tiles.
//
// The strip kernel in conv_direct.cu reads its input window straight from global memory: ncu on the MobileNetV3-dual
// layers (bs 64) shows DRAM traffic equal to the algorithmic bytes but only 1.4 - 1.9 TB/s (3x3) and 0.6 TB/s (5x5):
// 80 - 113 registers leave 16 - 24 warps per SM, each waiting on k dependent rounds of predicated 16-byte loads
// (long-scoreboard stalls; 274 instructions per 8-channel output of which 72 are FMAs).  Here a persistent CTA walks
// over (rows x columns x channel-block) tiles of one channel block: the input window of the NEXT tile, halo included, is
// fetched by one cp.async.bulk.tensor (4-D map over the NHWC view; the out-of-bounds fill supplies the zero padding, so
// the inner loops carry no bounds predicate and no 64-bit address math) while the threads compute the current tile from
// shared memory.  Memory-level parallelism comes from the copy engine instead of from occupancy.  The per-output
// arithmetic (r outer, s inner, fp32 FMAs; zero taps add exactly 0) is unchanged, so results are bit-identical to the
// strip kernel.
#include "common.h"
#include <array>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include "act.cuh"
#include "ptx.cuh"
#include "vec.cuh"

namespace dyk {

int encode_map_generic(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes, const char* what);

// 3x3 keeps the whole filter in registers (72) next to 32 accumulators: 192 threads x 2 CTAs leave 168 registers per
// thread (256 threads = 128 registers spilled ~30 of them); 5x5 reads one filter row at a time from shared memory
constexpr int dwt_threads(int K, int S) { return (K == 3 || S == 1) ? 192 : 256; }
// outputs along W per thread unit (the window slides: k + (strip-1)*stride columns per filter row).  5x5 stride 1 takes 8:
// its filter rows come from shared memory, and with 4 outputs those reads outnumber the activation reads
constexpr int dwt_strip(int K, int S) { return (K == 5 && S == 1) ? 8 : 4; }
constexpr int kDwtStages = 2;
constexpr int kDwtSbBytes = 2048;   // scale + bias of one channel block (<= 256 channels) in shared memory

struct DwtArgs {
  const float* w;        // [K*K][C]
  const float* scale;    // [C] or null
  const float* bias;     // [C] or null
  uint8_t* y;
  long long ys;          // output pixel stride (elements)
  int C, Ho, Wo, pad;
  int CB, cvb;           // channels per block, 8-channel vectors per block
  int TH, TW, SW;        // output tile rows / columns, strips per tile row
  int in_rows, in_cols;  // input box
  int tiles_w, tiles_h;
  unsigned ntiles;       // N * tiles_h * tiles_w
  unsigned stage_bytes;  // in_rows * in_cols * CB * 2, rounded up to 128
  int PT;                // pixel slots: threads with tid >= PT * cvb idle
};

// two fp32 FMAs per instruction (FFMA2 on sm_100): the inner loops are issue-bound, not FMA-pipe-bound; each half is an
// IEEE fma.rn, so results are those of two fmaf calls
__device__ __forceinline__ float2 ffma2(float2 x, float2 y, float2 z) {
  uint64_t ux = *reinterpret_cast<uint64_t*>(&x), uy = *reinterpret_cast<uint64_t*>(&y), uz = *reinterpret_cast<uint64_t*>(&z), ud;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ux), "l"(uy), "l"(uz));
  return *reinterpret_cast<float2*>(&ud);
}
template <bool kBf16>
__device__ __forceinline__ void unpack8x2(const uint4& v, float2 (&f)[4]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kBf16) f[i] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    else f[i] = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
  }
}

template <bool kBf16, int K, int S, int kAct>
__global__ void __launch_bounds__(dwt_threads(K, S), 2)
dwconv_tile_kernel(const __grid_constant__ CUtensorMap xmap, const DwtArgs a) {
  constexpr int kThreads = dwt_threads(K, S);
  constexpr int kStrip = dwt_strip(K, S);
  extern __shared__ __align__(128) uint8_t dwt_smem[];
  // pointer arithmetic on the __shared__ symbol (not an integer round trip) keeps the address space known to the compiler:
  // LDS / STS with 32-bit addresses instead of generic LD / ST with 64-bit address math
  uint8_t* base = dwt_smem + ((128u - (smem_u32(dwt_smem) & 127u)) & 127u);
  uint64_t* full = reinterpret_cast<uint64_t*>(base);                       // [kDwtStages]
  // folded BN scale / bias and (5x5) the filter, each as [..][cvb] float4 for channels 0-3 ("lo") and 4-7 ("hi") of a
  // vector: the 8 threads of a quarter-warp read 128 contiguous bytes (the natural [..][CB] float layout makes every
  // 16-byte read of a thread 32 bytes from its neighbour's: 2-way bank conflicts on ten loads per filter row)
  float4* sbs = reinterpret_cast<float4*>(base + 128);                      // scale lo, scale hi, bias lo, bias hi
  float4* wsm = reinterpret_cast<float4*>(base + 128 + kDwtSbBytes);        // [K*K][2][cvb] (K == 5 only)
  uint8_t* stage0 = base + 128 + kDwtSbBytes + (K == 5 ? ((K * K * a.CB * 4 + 127) & ~127) : 0);
  constexpr int kWin = (kStrip - 1) * S + K;

  const int tid = threadIdx.x;
  const int cb = blockIdx.y;
  const int c8 = tid % a.cvb, p = tid / a.cvb;
  const int ch = cb * a.CB + c8 * 8;            // first of this thread's 8 channels
  const bool active = p < a.PT && ch < a.C;

  if (tid == 0) {
    tma_prefetch_desc(&xmap);
    for (int s = 0; s < kDwtStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  auto issue = [&](unsigned tile, int stage) {
    const unsigned tw = tile % (unsigned)a.tiles_w;
    const unsigned t2 = tile / (unsigned)a.tiles_w;
    const unsigned th = t2 % (unsigned)a.tiles_h, n = t2 / (unsigned)a.tiles_h;
    mbar_arrive_expect_tx(&full[stage], (unsigned)(a.in_rows * a.in_cols * a.CB * 2));
    tma_load_4d(stage0 + (size_t)stage * a.stage_bytes, &xmap, &full[stage], cb * a.CB, (int)tw * a.TW * S - a.pad,
                (int)th * a.TH * S - a.pad, (int)n);
  };

  // per-CTA constants: filter (3x3: registers; 5x5: shared memory), folded BN scale / bias (shared memory)
  float2 wreg[K == 3 ? 9 : 1][4];
  {
    float* sb = reinterpret_cast<float*>(sbs);
    for (int i = tid; i < a.CB; i += kThreads) {
      const int c = cb * a.CB + i, v = i >> 3, q = i & 7;
      const int slot = ((q >> 2) * a.cvb + v) * 4 + (q & 3);
      sb[slot] = (a.scale && c < a.C) ? __ldg(&a.scale[c]) : 1.f;
      sb[a.CB + slot] = (a.bias && c < a.C) ? __ldg(&a.bias[c]) : 0.f;
    }
  }
  if (active) {
    if constexpr (K == 3) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 wa = __ldg(reinterpret_cast<const float4*>(a.w + (long long)t * a.C + ch));
        const float4 wb = __ldg(reinterpret_cast<const float4*>(a.w + (long long)t * a.C + ch) + 1);
        wreg[t][0] = make_float2(wa.x, wa.y); wreg[t][1] = make_float2(wa.z, wa.w);
        wreg[t][2] = make_float2(wb.x, wb.y); wreg[t][3] = make_float2(wb.z, wb.w);
      }
    }
  }
  if constexpr (K == 5) {
    float* wf = reinterpret_cast<float*>(wsm);
    for (int i = tid; i < K * K * a.CB; i += kThreads) {
      const int t = i / a.CB, c = i - t * a.CB, v = c >> 3, q = c & 7;
      wf[((t * 2 + (q >> 2)) * a.cvb + v) * 4 + (q & 3)] =
          (cb * a.CB + c < a.C) ? __ldg(&a.w[(long long)t * a.C + cb * a.CB + c]) : 0.f;
    }
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x < a.ntiles) issue(blockIdx.x, 0);

  const unsigned pix_bytes = (unsigned)a.CB * 2u;
  const unsigned row_bytes = (unsigned)a.in_cols * pix_bytes;
  const int units = a.TH * a.SW;
  unsigned it = 0;
  for (unsigned tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    __syncthreads();                                   // everyone is done with the other stage (tile it-1)
    if (tid == 0 && tile + gridDim.x < a.ntiles) issue(tile + gridDim.x, stage ^ 1);
    mbar_wait(&full[stage], (it >> 1) & 1);
    if (!active) continue;
    const unsigned tw = tile % (unsigned)a.tiles_w;
    const unsigned t2 = tile / (unsigned)a.tiles_w;
    const unsigned th = t2 % (unsigned)a.tiles_h, n = t2 / (unsigned)a.tiles_h;
    const int ho0 = (int)th * a.TH, wo0 = (int)tw * a.TW;
    const uint8_t* sm = stage0 + (size_t)stage * a.stage_bytes + c8 * 16;
    for (int u = p; u < units; u += a.PT) {
      const int sx = u / a.TH, row = u - sx * a.TH;    // rows fastest: neighbouring threads read different rows
      const int ho = ho0 + row, wo = wo0 + sx * kStrip;
      if (ho >= a.Ho || wo >= a.Wo) continue;
      float2 acc[kStrip][4];
#pragma unroll
      for (int j = 0; j < kStrip; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = make_float2(0.f, 0.f);
      const uint8_t* win = sm + (unsigned)(row * S) * row_bytes + (unsigned)(sx * kStrip * S) * pix_bytes;
#pragma unroll
      for (int r = 0; r < K; ++r) {
        float2 wrow[K == 5 ? 5 : 1][4];
        if constexpr (K == 5) {
#pragma unroll
          for (int s = 0; s < K; ++s) {
            const float4 wa = wsm[((r * K + s) * 2) * a.cvb + c8];
            const float4 wb = wsm[((r * K + s) * 2 + 1) * a.cvb + c8];
            wrow[s][0] = make_float2(wa.x, wa.y); wrow[s][1] = make_float2(wa.z, wa.w);
            wrow[s][2] = make_float2(wb.x, wb.y); wrow[s][3] = make_float2(wb.z, wb.w);
          }
        }
#pragma unroll
        for (int col = 0; col < kWin; ++col) {
          float2 f[4];
          unpack8x2<kBf16>(*reinterpret_cast<const uint4*>(win + r * row_bytes + col * pix_bytes), f);
#pragma unroll
          for (int j = 0; j < kStrip; ++j) {
            const int s = col - j * S;
            if (s >= 0 && s < K) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if constexpr (K == 3) acc[j][q] = ffma2(f[q], wreg[r * 3 + s][q], acc[j][q]);
                else acc[j][q] = ffma2(f[q], wrow[s][q], acc[j][q]);
              }
            }
          }
        }
      }
      float2 sc[4], bi[4];
      {
        const float4 s0 = sbs[c8], s1 = sbs[a.cvb + c8], b0 = sbs[2 * a.cvb + c8], b1 = sbs[3 * a.cvb + c8];
        sc[0] = make_float2(s0.x, s0.y); sc[1] = make_float2(s0.z, s0.w); sc[2] = make_float2(s1.x, s1.y); sc[3] = make_float2(s1.z, s1.w);
        bi[0] = make_float2(b0.x, b0.y); bi[1] = make_float2(b0.z, b0.w); bi[2] = make_float2(b1.x, b1.y); bi[3] = make_float2(b1.z, b1.w);
      }
      uint8_t* yp = a.y + ((((long long)n * a.Ho + ho) * a.Wo + wo) * a.ys + ch) * 2;
#pragma unroll
      for (int j = 0; j < kStrip; ++j) {
        if (wo + j < a.Wo) {
          float o[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 t = ffma2(acc[j][q], sc[q], bi[q]);
            o[2 * q] = act_apply<kAct>(t.x);
            o[2 * q + 1] = act_apply<kAct>(t.y);
          }
          *reinterpret_cast<uint4*>(yp + (long long)j * a.ys * 2) = pack8<kBf16>(o);
        }
      }
    }
  }
}

// Tile geometry.  Every candidate (channel block CB | C, output tile TH x TW) whose input box fits the stage budget is
// scored with a small model of what the sweep in tools/dw_bench.py --sweep measures:
//   compute efficiency = active threads / threads  x  units / (rounds x pixel slots)  x  useful outputs / covered outputs,
//                        with a fixed per-tile cost (barrier, index math, pipeline bubble) of one round
//                        (fitted to the sweep: 1/3 round picked tiles that were too small on the C = 16 and C = 128 layers);
//   memory cost        = (input re-read through the halo + output) / algorithmic bytes.
// Picks are cached per (C, Ho, Wo, k, stride, N).
struct DwtGeom { int CB, TW, TH, in_cols, in_rows; };

static inline int dwt_in_cols(int TW, int stride, int k, int cvb) {
  int in_cols = (TW - 1) * stride + k;
  if (cvb < 8) {                         // short pixel rows: a row pitch of cvb (mod 8) 16-byte units keeps a quarter-warp
    const int m = 8 / cvb;               // that spans several rows on distinct banks
    while (in_cols % m != 1 % m) ++in_cols;
  }
  return in_cols;
}
static inline size_t dwt_wbytes(int k, int CB) { return k == 5 ? (size_t)((k * k * CB * 4 + 127) & ~127) : 0; }
static inline size_t dwt_budget(int k, int CB) { return (size_t)52 * 1024 - dwt_wbytes(k, CB) / 2; }

static bool dwt_search(int C, int Ho, int Wo, int k, int S, int N, DwtGeom* out) {
  const int T = dwt_threads(k, S), strip = dwt_strip(k, S);
  bool have_mid = false;
  for (int cb = 32; cb <= 128; cb += 8) have_mid |= (C % cb == 0);
  double best = -1.0;
  for (int CB = 8; CB <= 256; CB += 8) {
    const bool ragged = (C % CB != 0);
    if (ragged && !(CB == 64 && C > 256 && !have_mid)) continue;
    if (!ragged && CB != C && (CB < 32 || (CB > 128 && have_mid))) continue;   // pixel rows of 64..256 bytes when C allows
    if (CB > C) break;
    const int cvb = CB / 8, PT = T / cvb;
    if (PT < 1) continue;
    const int ncb = ceil_div(C, CB);
    const size_t budget = dwt_budget(k, CB);
    const double chan = (double)C / ((double)ncb * CB);
    for (int TW = strip; TW <= 64 && TW < Wo + strip; TW += strip) {
      const int in_cols = dwt_in_cols(TW, S, k, cvb);
      if (in_cols > 256) break;
      for (int TH = 1; TH <= 16 && TH <= Ho; ++TH) {
        const int in_rows = (TH - 1) * S + k;
        if ((size_t)in_rows * in_cols * CB * 2 > budget) break;
        const int tiles_h = ceil_div(Ho, TH), tiles_w = ceil_div(Wo, TW);
        const int units = TH * (TW / strip), rounds = ceil_div(units, PT);
        const double cover = ((double)Ho * Wo) / ((double)tiles_h * TH * tiles_w * TW);
        const double comp = ((double)PT * cvb / T) * ((double)units / ((double)rounds * PT)) * cover * chan * rounds / (rounds + 1.0);
        const double amp = ((double)in_rows * in_cols) / ((double)TH * TW * S * S) / cover / chan;
        const double mem = (amp * S * S + 1.0) / (S * S + 1.0);
        const double nt = (double)N * tiles_h * tiles_w * ncb;
        const double fill = nt >= 2.0 * num_sms() ? 1.0 : nt / (2.0 * num_sms());
        const double score = comp * fill / (0.5 + 0.5 * mem);
        if (score > best) { best = score; *out = DwtGeom{CB, TW, TH, in_cols, in_rows}; }
      }
    }
  }
  return best > 0.0;
}

// returns DYK_OK after a launch, 1 when the shape is not handled here (the caller keeps its own kernel), < 0 on error
int dwconv_tile_try(const void* x, int64_t xs, const float* w, const float* scale, const float* bias, void* y, int64_t ys,
                    int N, int H, int W, int C, int k, int stride, int pad, int act, int dtype, cudaStream_t stream) {
  const char* env = getenv("DYK_DW_TILE");     // read per call: the parity test switches kernels inside one process
  const bool off = env != nullptr && env[0] == '0';
  if (off || !(k == 3 || k == 5) || !(stride == 1 || stride == 2) || !(dtype == DYK_F16 || dtype == DYK_BF16)) return 1;
  if (pad < 0 || pad > k) return 1;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DwtArgs a = {};
  a.w = w; a.scale = scale; a.bias = bias; a.y = static_cast<uint8_t*>(y); a.ys = ys;
  a.C = C; a.Ho = Ho; a.Wo = Wo; a.pad = pad;
  DwtGeom g;
  const int strip = dwt_strip(k, stride);
  if (const char* cfg = getenv("DYK_DW_TILE_CFG")) {       // "CB,TW,TH": tools/dw_bench.py --sweep
    if (sscanf(cfg, "%d,%d,%d", &g.CB, &g.TW, &g.TH) != 3 || g.CB < 8 || g.CB % 8 || g.CB > 256 || g.TW < strip ||
        g.TW % strip || g.TH < 1 || g.CB > ((C + 7) & ~7))
      return fail(DYK_EINVAL, "DYK_DW_TILE_CFG=%s: expects CB,TW,TH", cfg);
    g.in_cols = dwt_in_cols(g.TW, stride, k, g.CB / 8);
    g.in_rows = (g.TH - 1) * stride + k;
    if (g.in_cols > 256 || (size_t)g.in_rows * g.in_cols * g.CB * 2 > dwt_budget(k, g.CB))
      return fail(DYK_EINVAL, "DYK_DW_TILE_CFG=%s: box does not fit", cfg);
  } else {
    static std::mutex mu;
    static std::map<std::array<int, 6>, DwtGeom> cache;
    const std::array<int, 6> key = {C, Ho, Wo, k, stride, N};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
      if (!dwt_search(C, Ho, Wo, k, stride, N, &g)) g.CB = 0;
      if (getenv("DYK_DW_TILE_DEBUG"))
        fprintf(stderr, "dyk: depthwise C=%d %dx%d k%d s%d N=%d -> CB=%d TW=%d TH=%d (box %dx%d)\n", C, Ho, Wo, k, stride, N,
                g.CB, g.TW, g.TH, g.in_rows, g.in_cols);
      it = cache.emplace(key, g).first;
    }
    g = it->second;
    if (g.CB == 0) return 1;
  }
  a.CB = g.CB; a.TW = g.TW; a.TH = g.TH; a.in_cols = g.in_cols; a.in_rows = g.in_rows;
  a.cvb = a.CB / 8;
  a.PT = dwt_threads(k, stride) / a.cvb;
  a.SW = a.TW / strip;
  const int ncb = ceil_div(C, a.CB);
  const size_t wbytes = dwt_wbytes(k, a.CB);
  a.tiles_h = ceil_div(Ho, a.TH);
  a.tiles_w = ceil_div(Wo, a.TW);
  const size_t box_bytes = (size_t)a.in_rows * a.in_cols * a.CB * 2;
  a.stage_bytes = (unsigned)((box_bytes + 127) & ~(size_t)127);
  const long long nt = (long long)N * a.tiles_h * a.tiles_w;
  if (nt >= (1ll << 31) || ncb > 65535) return 1;
  a.ntiles = (unsigned)nt;

  CUtensorMap xmap;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)xs * 2, (cuuint64_t)W * xs * 2, (cuuint64_t)H * W * xs * 2};
  const cuuint32_t box[4] = {(cuuint32_t)a.CB, (cuuint32_t)a.in_cols, (cuuint32_t)a.in_rows, 1};
  int rc = encode_map_generic(&xmap, x, 4, dims, strides, box, 0, "depthwise input");
  if (rc != DYK_OK) return rc;

  const size_t smem = 128 + 128 + kDwtSbBytes + wbytes + (size_t)kDwtStages * a.stage_bytes;
  // all CTAs must be co-resident (two per SM): one CTA too many would run its whole share of the tiles as a second wave
  long long gx = 2ll * num_sms() / ncb;
  if (gx < 1) gx = 1;
  if (gx > nt) gx = nt;
  const dim3 grid((unsigned)gx, (unsigned)ncb);
#define DYK_DWT(KK, SS)                                                                                              \
  DYK_DISPATCH_ACT(act, DYK_DISPATCH_DTYPE(dtype, {                                                                  \
    auto kern = dwconv_tile_kernel<kBf16, KK, SS, kAct>;                                                            \
    static bool configured = false;                                                                                \
    if (!configured) {                                                                                             \
      DYK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));            \
      configured = true;                                                                                           \
    }                                                                                                              \
    kern<<<grid, dwt_threads(KK, SS), smem, stream>>>(xmap, a);                                                       \
  }))
  if (k == 3 && stride == 1) DYK_DWT(3, 1);
  else if (k == 3) DYK_DWT(3, 2);
  else if (stride == 1) DYK_DWT(5, 1);
  else DYK_DWT(5, 2);
#undef DYK_DWT
  DYK_LAUNCH_OK("dwconv_tile_kernel");
  return DYK_OK;
}

}  // namespace dyk
