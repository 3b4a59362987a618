// Dense convolution (+ per-channel scale/bias = folded BatchNorm, activation, optional residual add,
// optional fused 2x nearest upsample of the output) as a persistent, warp-specialised tcgen05
// implicit GEMM for sm_100a.
//
// Replaces: nn.Conv2d -> nn.BatchNorm2d -> activation of a [convolutional] block (reference
// models.py:28-64), the unweighted [shortcut] add that follows it (build_utils/layers.py:63-85) and
// nn.Upsample(scale_factor=2) after it (models.py:100-101).
//
// GEMM view:  D[M = 128 output pixels][N = BLOCK_N out channels] += A[M][K] * B[N][K]^T,
//             K runs over (filter tap r,s) x (input-channel chunk of BLOCK_K).
//  * A tile  : one TMA box {BLOCK_K ch, tw, th, tn} of the NHWC input, shifted by the tap offset; the
//              zero padding of the convolution is TMA out-of-bounds fill.  tw*th*tn == 128, so the box
//              lands in shared memory as 128 rows of BLOCK_K channels in the canonical K-major
//              SWIZZLE_{128,64}B layout that tcgen05.mma consumes directly.  Stride-2 convolutions read
//              four "parity plane" tensor maps (even/odd rows x even/odd columns of the input), which
//              turns every tap of a strided conv back into a dense box.
//  * B tile  : TMA box {BLOCK_K, 1 tap, BLOCK_N} of the packed weights [Cout][kh*kw][Cin].
//  * D       : fp32 in TMEM, double buffered (2 x BLOCK_N columns) so the epilogue of tile i overlaps
//              the main loop of tile i+1.
//  * epilogue: tcgen05.ld -> scale/bias/activation (+ residual) in fp32 -> fp16/bf16 -> swizzled smem
//              staging -> TMA store (clipped at tensor edges; 4 parity stores when upsampling).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = epilogue (TMEM lane quarter = warp_idx & 3; two warps per quarter split the columns).
#include "common.h"
#include "ptx.cuh"
#include "act.cuh"

namespace dyk {

struct ConvTmaps {
  CUtensorMap a[4];  // input, by (row parity*2 + col parity); stride-1 convs use a[0] only
  CUtensorMap b;     // packed weights (Cin, taps, Cout)
  CUtensorMap y[4];  // output; upsample2x uses the 4 parity planes of the 2x map, else y[0]
};

struct ConvKArgs {
  int tw, th, tn;                  // A/Y box spatial extents, tw*th*tn == 128
  int tiles_w, tiles_h, tiles_b;   // tile grid over (Wo, Ho, N)
  int n_blocks;                    // ceil(Cout_store / BLOCK_N)
  int num_tiles;
  int k_chunks;                    // ceil(Cin / BLOCK_K)
  int kh, kw, stride, pad;
  int Ho, Wo, N;
  int Cout_store;
  int act;
  int upsample2x;
  int out_f32;                     // 1: y is fp32, written with plain stores (head convs feeding decode)
  void* y_f32;
  long long y_pix_stride;
  const float* scale;              // may be null
  const float* bias;               // may be null
  const void* res;                 // may be null
  long long res_pix_stride;
};

constexpr int kNumEpiWarps = 8;
constexpr int kNumEpiThreads = kNumEpiWarps * 32;
constexpr int kNumThreads = 64 + kNumEpiThreads;   // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue
constexpr int kBlockM = 128;

template <int BLOCK_N, int BLOCK_K>
struct ConvSmem {
  static constexpr int kStoreC = BLOCK_N < 64 ? BLOCK_N : 64;          // channels per TMA store box
  static constexpr int kABytes = kBlockM * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kBlockM * kStoreC * 2;          // one store group
  static constexpr int kNumStaging = 2;
  static constexpr int kVecBytes = 2 * BLOCK_N * 4;                    // per-tile scale + bias vectors
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kBudget = 227 * 1024 - kBarrierBytes - kVecBytes - kNumStaging * kStagingBytes - 1024 /*align*/;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTotal =
      kStages * kStageBytes + kNumStaging * kStagingBytes + kVecBytes + kBarrierBytes + 1024;
  static_assert(kStages >= 3, "pipeline too shallow");
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kNumEpiThreads) : "memory"); }

template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (kBf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool kBf16>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (kBf16) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&v));
  }
}

struct TileCoord {
  int nblk, w0, h0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvKArgs& p, int tile) {
  TileCoord t;
  t.nblk = tile % p.n_blocks;
  const int mt = tile / p.n_blocks;
  const int tww = mt % p.tiles_w;
  const int thh = (mt / p.tiles_w) % p.tiles_h;
  const int tb = mt / (p.tiles_w * p.tiles_h);
  t.w0 = tww * p.tw;
  t.h0 = thh * p.th;
  t.n0 = tb * p.tn;
  return t;
}

// Epilogue of one 128 x BLOCK_N accumulator tile, executed by the 8 epilogue warps.
// Thread (q = warp & 3, lane, half = epilogue warp >> 2) owns tile row q*32+lane and, in every store group of
// kStoreC channels, the `half`-th half of the group's columns.  kAct is a compile-time activation so that
// only one code path is resident in the instruction cache.
template <int BLOCK_N, int BLOCK_K, bool kBf16, int kAct>
__device__ __forceinline__ void epilogue_tile(const ConvTmaps& tm, const ConvKArgs& p, const TileCoord& tc,
                                              uint32_t t_row, uint8_t* staging, float* svec, int& sbuf,
                                              uint64_t* tempty, int q, int lane, int half, bool store_thread) {
  using S = ConvSmem<BLOCK_N, BLOCK_K>;
  constexpr int kStoreC = S::kStoreC;
  constexpr int kStoreRowBytes = kStoreC * 2;       // 128 or 64
  constexpr int kCols = kStoreC / 2;                // columns per thread per group: 32 or 16
  constexpr int kGroups = BLOCK_N / kStoreC;
  const int row = q * 32 + lane;
  const int wi = row % p.tw;
  const int hi = (row / p.tw) % p.th;
  const int ni = row / (p.tw * p.th);
  const int wo = tc.w0 + wi, ho = tc.h0 + hi, nn = tc.n0 + ni;
  const bool pix_ok = (wo < p.Wo) && (ho < p.Ho) && (nn < p.N);
  const long long pix = (static_cast<long long>(nn) * p.Ho + ho) * p.Wo + wo;
  const int n_base = tc.nblk * BLOCK_N;
  const bool has_res = p.res != nullptr;

  // per-tile scale / bias vectors -> shared memory (one element per epilogue thread)
  {
    const int t = (half * 4 + q) * 32 + lane;
    if (t < BLOCK_N) {
      svec[t] = p.scale ? __ldg(p.scale + n_base + t) : 1.f;
      svec[BLOCK_N + t] = p.bias ? __ldg(p.bias + n_base + t) : 0.f;
    }
  }
  // residual values of this thread's columns in group 0 are fetched while the MMAs still run
  uint4 rres[kCols / 8];
  auto load_res = [&](int g) {
    const int c0 = n_base + g * kStoreC + half * kCols;
    const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.res) + (pix * p.res_pix_stride + c0) * 2;
#pragma unroll
    for (int j = 0; j < kCols / 8; ++j) {
      rres[j] = make_uint4(0u, 0u, 0u, 0u);
      if (has_res && pix_ok && c0 + j * 8 < p.Cout_store) rres[j] = __ldg(reinterpret_cast<const uint4*>(rp) + j);
    }
  };
  load_res(0);

#pragma unroll 1
  for (int g = 0; g < kGroups; ++g) {
    const int cg0 = n_base + g * kStoreC;   // first output channel of this store group
    if (cg0 >= p.Cout_store) break;          // uniform: whole group is beyond the tensor
    const bool last_group = (g == kGroups - 1) || (cg0 + kStoreC >= p.Cout_store);
    if (!p.out_f32 && store_thread) tma_store_wait_read<1>();  // staging[sbuf] no longer read by an older store
    epi_bar_sync();                                            // (also publishes svec on g == 0)
    uint8_t* sbase = staging + sbuf * S::kStagingBytes;

    uint32_t v[kCols];
    if constexpr (kCols == 32) tmem_ld_32x32b_x32(t_row + g * kStoreC + half * kCols, v);
    else tmem_ld_32x32b_x16(t_row + g * kStoreC + half * kCols, v);
    tmem_ld_wait();
    if (last_group) {
      // all TMEM reads of this accumulator stage are done -> hand it back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
    const int cl = g * kStoreC + half * kCols;   // first column (within the tile) of this thread
    float o[kCols];
#pragma unroll
    for (int j = 0; j < kCols; j += 4) {
      const float4 sc = *reinterpret_cast<const float4*>(svec + cl + j);
      const float4 bi = *reinterpret_cast<const float4*>(svec + BLOCK_N + cl + j);
      o[j + 0] = act_apply<kAct>(fmaf(__uint_as_float(v[j + 0]), sc.x, bi.x));
      o[j + 1] = act_apply<kAct>(fmaf(__uint_as_float(v[j + 1]), sc.y, bi.y));
      o[j + 2] = act_apply<kAct>(fmaf(__uint_as_float(v[j + 2]), sc.z, bi.z));
      o[j + 3] = act_apply<kAct>(fmaf(__uint_as_float(v[j + 3]), sc.w, bi.w));
    }
    if (has_res) {
#pragma unroll
      for (int j = 0; j < kCols / 8; ++j) {
        const uint32_t rr[4] = {rres[j].x, rres[j].y, rres[j].z, rres[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2<kBf16>(rr[e]);
          o[j * 8 + e * 2 + 0] += f.x;
          o[j * 8 + e * 2 + 1] += f.y;
        }
      }
      if (!last_group) load_res(g + 1);
    }
    if (p.out_f32) {
      if (pix_ok) {
        float* yo = reinterpret_cast<float*>(p.y_f32) + pix * p.y_pix_stride + n_base + cl;
#pragma unroll
        for (int j = 0; j < kCols; ++j)
          if (n_base + cl + j < p.Cout_store) yo[j] = o[j];
      }
      continue;
    }
    // swizzled staging write: row `row`, this thread's 16-byte chunks of the store row
#pragma unroll
    for (int ch = 0; ch < kCols / 8; ++ch) {
      const int chunk = half * (kCols / 8) + ch;
      int phys;
      if constexpr (kStoreRowBytes == 128) phys = chunk ^ (row & 7);
      else phys = chunk ^ ((row >> 1) & 3);
      const uint4 val = make_uint4(pack2<kBf16>(o[ch * 8 + 0], o[ch * 8 + 1]), pack2<kBf16>(o[ch * 8 + 2], o[ch * 8 + 3]),
                                   pack2<kBf16>(o[ch * 8 + 4], o[ch * 8 + 5]), pack2<kBf16>(o[ch * 8 + 6], o[ch * 8 + 7]));
      *reinterpret_cast<uint4*>(sbase + row * kStoreRowBytes + phys * 16) = val;
    }
    fence_proxy_async_smem();
    epi_bar_sync();
    if (store_thread) {
      if (p.upsample2x) {
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) tma_store_4d(&tm.y[pp], sbase, cg0, tc.w0, tc.h0, tc.n0);
      } else {
        tma_store_4d(&tm.y[0], sbase, cg0, tc.w0, tc.h0, tc.n0);
      }
      tma_store_commit();
    }
    sbuf ^= 1;
  }
}

template <int BLOCK_N, int BLOCK_K, bool kBf16>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_tc_kernel(const __grid_constant__ ConvTmaps tm, const ConvKArgs p) {
  using S = ConvSmem<BLOCK_N, BLOCK_K>;
  constexpr int kSwz = BLOCK_K * 2;            // bytes per smem operand row == swizzle span (128 / 64)
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
  static_assert(BLOCK_K == 64 || BLOCK_K == 32, "BLOCK_K");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kStages * S::kStageBytes;
  float* svec = reinterpret_cast<float*>(staging + S::kNumStaging * S::kStagingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(svec) + S::kVecBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;        // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.kh * p.kw;
  const int num_kb = taps * p.k_chunks;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a[0]);
    tma_prefetch_desc(&tm.b);
    if (!p.out_f32) tma_prefetch_desc(&tm.y[0]);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kNumEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(p, tile);
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.kw, s = tap - r * p.kw;
          int dh = r - p.pad, dw = s - p.pad;
          int map_idx = 0;
          if (p.stride == 2) {
            const int ph = dh & 1, pw = dw & 1;
            map_idx = ph * 2 + pw;
            dh = (dh - ph) >> 1;
            dw = (dw - pw) >> 1;
          }
          const CUtensorMap* amap = &tm.a[map_idx];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
            tma_load_4d(sa, amap, &full_bar[stage], kc * BLOCK_K, tc.w0 + dw, tc.h0 + dh, tc.n0);
            tma_load_3d(sb, &tm.b, &full_bar[stage], kc * BLOCK_K, tap, tc.nblk * BLOCK_N);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N, kBf16 ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int tl = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
        const int as = tl & 1;
        const uint32_t aphase = (tl >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(stage_base + stage * S::kStageBytes);
          const uint64_t adesc = umma_desc_kmajor<kSwz>(sa);
          const uint64_t bdesc = umma_desc_kmajor<kSwz>(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the >>4 address field
            umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp_idx - 2;
    const int q = warp_idx & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;              // which half of each store group's columns
    const bool store_thread = (threadIdx.x == 64);
    int tl = 0;
    int sbuf = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      const int as = tl & 1;
      const uint32_t aphase = (tl >> 1) & 1;
      const TileCoord tc = tile_coord(p, tile);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
      // The accumulator wait sits inside the activation-specialised body's caller so that residual / vector
      // prefetches of the *next* tile are not hoisted above it by accident.
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after_sync();
#define DYK_EPI(ACT)                                                                                         \
  epilogue_tile<BLOCK_N, BLOCK_K, kBf16, ACT>(tm, p, tc, t_row, staging, svec, sbuf, &tempty_bar[as], q, lane, \
                                              half, store_thread)
      switch (p.act) {
        case DYK_ACT_LEAKY: DYK_EPI(DYK_ACT_LEAKY); break;
        case DYK_ACT_MISH: DYK_EPI(DYK_ACT_MISH); break;
        case DYK_ACT_RELU: DYK_EPI(DYK_ACT_RELU); break;
        case DYK_ACT_RELU6: DYK_EPI(DYK_ACT_RELU6); break;
        case DYK_ACT_HARDSWISH: DYK_EPI(DYK_ACT_HARDSWISH); break;
        case DYK_ACT_HARDSIGMOID: DYK_EPI(DYK_ACT_HARDSIGMOID); break;
        default: DYK_EPI(DYK_ACT_LINEAR); break;
      }
#undef DYK_EPI
    }
    if (store_thread) tma_store_wait_all<0>();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------- host

static int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes /* rank-1 */, const cuuint32_t* box, int swizzle_bytes,
                      const char* what) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(DYK_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, rank, const_cast<void*>(base), dims, strides_bytes, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(DYK_ECUDA,
                "cuTensorMapEncodeTiled(%s) failed with CUresult %d (rank %d dims %llu,%llu,%llu,%llu box "
                "%u,%u,%u,%u stride1 %llu base %p)",
                what, (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
                (unsigned long long)strides_bytes[0], base);
  }
  return DYK_OK;
}

// Spatial box (tw, th, tn) with tw*th*tn == 128 that wastes the fewest output pixels.
static void pick_tile(int Wo, int Ho, int N, int* tw, int* th, int* tn) {
  double best = -1;
  for (int w = 1; w <= 128; w *= 2) {
    for (int h = 1; w * h <= 128; h *= 2) {
      const int n = 128 / (w * h);
      const long long covered = (long long)ceil_div(Wo, w) * w * (long long)ceil_div(Ho, h) * h *
                                (long long)ceil_div(N, n) * n;
      double eff = (double)Wo * Ho * N / (double)covered;
      // prefer wide boxes (longer contiguous runs per TMA row) on ties
      eff += 1e-6 * w - 1e-7 * n;
      if (eff > best) { best = eff; *tw = w; *th = h; *tn = n; }
    }
  }
}

template <int BLOCK_N, int BLOCK_K, bool kBf16>
static int launch_conv(const ConvTmaps& tm, const ConvKArgs& ka, cudaStream_t stream) {
  using S = ConvSmem<BLOCK_N, BLOCK_K>;
  auto kern = conv_tc_kernel<BLOCK_N, BLOCK_K, kBf16>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    DYK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  int grid = ka.num_tiles < num_sms() ? ka.num_tiles : num_sms();
  kern<<<grid, kNumThreads, S::kTotal, stream>>>(tm, ka);
  DYK_LAUNCH_OK("conv_tc_kernel");
  return DYK_OK;
}

template <int BLOCK_K, bool kBf16>
static int dispatch_n(int block_n, const ConvTmaps& tm, const ConvKArgs& ka, cudaStream_t stream) {
  switch (block_n) {
    case 32: return launch_conv<32, BLOCK_K, kBf16>(tm, ka, stream);
    case 64: return launch_conv<64, BLOCK_K, kBf16>(tm, ka, stream);
    case 128: return launch_conv<128, BLOCK_K, kBf16>(tm, ka, stream);
    case 256: return launch_conv<256, BLOCK_K, kBf16>(tm, ka, stream);
  }
  return fail(DYK_EINVAL, "bad BLOCK_N %d", block_n);
}

// Picks the N tile by a small cycle model: a k-block costs max(MMA issue, L2->SM operand fetch at ~42 B/clk/SM),
// a tile adds a fixed epilogue/hand-off cost, and the grid runs in waves over the SMs.
static int pick_block_n(int cout_store, long long m_tiles, int num_kb, int block_k) {
  const int sms = num_sms();
  int best_n = 32;
  double best_cost = 1e30;
  const int cands[4] = {256, 128, 64, 32};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (bn > 32 && bn / 2 >= cout_store) continue;  // more than 2x wider than the layer
    const long long tiles = m_tiles * ceil_div(cout_store, bn);
    const long long waves = (tiles + sms - 1) / sms;
    const double l2 = (128.0 * block_k * 2 + (double)bn * block_k * 2) / 42.0;
    const double mma = (block_k / 16) * (bn < 64 ? 32.0 : bn / 2.0);
    const double tile_cost = num_kb * (l2 > mma ? l2 : mma) + 1200.0 + 6.0 * bn;
    const double cost = (double)waves * tile_cost;
    if (cost < best_cost - 1e-9) { best_cost = cost; best_n = bn; }
  }
  return best_n;
}

}  // namespace dyk

using namespace dyk;

extern "C" __attribute__((visibility("default"))) int dyk_conv2d_fwd(const dyk_conv_params* p, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(p != nullptr, "dyk_conv2d_fwd: null params");
  DYK_REQUIRE(p->x && p->w && p->y, "dyk_conv2d_fwd: null tensor pointer");
  DYK_REQUIRE(p->dtype == DYK_F16 || p->dtype == DYK_BF16, "dyk_conv2d_fwd: dtype %d", p->dtype);
  DYK_REQUIRE(p->stride == 1 || p->stride == 2, "dyk_conv2d_fwd: stride %d unsupported", p->stride);
  DYK_REQUIRE(p->kh >= 1 && p->kw >= 1 && p->kh * p->kw <= 49, "dyk_conv2d_fwd: kernel %dx%d", p->kh, p->kw);
  DYK_REQUIRE(p->Cin % 8 == 0 && p->Cin > 0, "dyk_conv2d_fwd: Cin=%d must be a positive multiple of 8", p->Cin);
  DYK_REQUIRE(p->x_pix_stride % 8 == 0 && (p->out_f32 || p->y_pix_stride % 8 == 0),
              "dyk_conv2d_fwd: pixel strides must be multiples of 8 elements (16 B)");
  DYK_REQUIRE(p->Cout > 0 && p->Cout_store >= p->Cout && (p->out_f32 || p->Cout_store % 8 == 0),
              "dyk_conv2d_fwd: Cout=%d Cout_store=%d (Cout_store must be a multiple of 8, >= Cout)", p->Cout,
              p->Cout_store);
  DYK_REQUIRE(!(p->out_f32 && p->upsample2x), "dyk_conv2d_fwd: out_f32 + upsample2x not supported");
  DYK_REQUIRE(p->y_pix_stride >= p->Cout_store && p->x_pix_stride >= p->Cin, "dyk_conv2d_fwd: stride < channels");
  DYK_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
              "dyk_conv2d_fwd: x, w, y must be 16-byte aligned");
  if (p->res) {
    DYK_REQUIRE(p->res_pix_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p->res) & 15) == 0,
                "dyk_conv2d_fwd: residual must be 16-byte aligned with stride %% 8 == 0");
    DYK_REQUIRE(!p->upsample2x, "dyk_conv2d_fwd: residual + upsample2x not supported");
  }
  const int Ho = (p->H + 2 * p->pad - p->kh) / p->stride + 1;
  const int Wo = (p->W + 2 * p->pad - p->kw) / p->stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_conv2d_fwd: empty output");

  const int BK = (p->Cin <= 32) ? 32 : 64;
  const int swz = BK * 2;
  ConvTmaps tm;
  ConvKArgs ka;
  memset(&tm, 0, sizeof(tm));
  memset(&ka, 0, sizeof(ka));

  // 1x1 stride-1 convolutions are plain GEMMs over the flattened pixel index.
  const bool flat = (p->kh == 1 && p->kw == 1 && p->stride == 1 && p->pad == 0 && !p->upsample2x);
  int tw, th, tn;
  int gW = Wo, gH = Ho, gN = p->N;  // logical output grid the tiles run over
  if (flat) {
    gW = (int)((long long)p->N * p->H * p->W);
    DYK_REQUIRE((long long)p->N * p->H * p->W < (1ll << 31), "dyk_conv2d_fwd: too many pixels");
    gH = 1; gN = 1; tw = 128; th = 1; tn = 1;
  } else {
    pick_tile(Wo, Ho, p->N, &tw, &th, &tn);
  }
  const cuuint32_t abox[4] = {(cuuint32_t)BK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  const long long xs = p->x_pix_stride * 2;  // bytes per pixel step
  int rc;
  if (flat) {
    const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)gW, 1, 1};
    const cuuint64_t str[3] = {(cuuint64_t)xs, (cuuint64_t)xs * gW, (cuuint64_t)xs * gW};
    if ((rc = encode_map(&tm.a[0], p->x, 4, dims, str, abox, swz, "A/flat"))) return rc;
  } else if (p->stride == 1) {
    const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)p->W, (cuuint64_t)p->H, (cuuint64_t)p->N};
    const cuuint64_t str[3] = {(cuuint64_t)xs, (cuuint64_t)xs * p->W, (cuuint64_t)xs * p->W * p->H};
    if ((rc = encode_map(&tm.a[0], p->x, 4, dims, str, abox, swz, "A"))) return rc;
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        const int Hp = (p->H - ph + 1) / 2, Wp = (p->W - pw + 1) / 2;
        if (Hp <= 0 || Wp <= 0) continue;
        const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)p->N};
        const cuuint64_t str[3] = {(cuuint64_t)xs * 2, (cuuint64_t)xs * p->W * 2, (cuuint64_t)xs * p->W * p->H};
        const uint8_t* base = reinterpret_cast<const uint8_t*>(p->x) + ((long long)ph * p->W + pw) * xs;
        if ((rc = encode_map(&tm.a[ph * 2 + pw], base, 4, dims, str, abox, swz, "A/parity"))) return rc;
      }
  }
  const int taps = p->kh * p->kw;
  const long long m_tiles = (long long)ceil_div(gW, tw) * ceil_div(gH, th) * ceil_div(gN, tn);
  const int BN = pick_block_n(p->Cout_store, m_tiles, taps * ceil_div(p->Cin, BK), BK);
  {
    const cuuint64_t dims[3] = {(cuuint64_t)p->Cin, (cuuint64_t)taps, (cuuint64_t)p->Cout};
    const cuuint64_t str[2] = {(cuuint64_t)p->Cin * 2, (cuuint64_t)p->Cin * 2 * taps};
    const cuuint32_t box[3] = {(cuuint32_t)BK, 1, (cuuint32_t)BN};
    if ((rc = encode_map(&tm.b, p->w, 3, dims, str, box, swz, "B"))) return rc;
  }
  if (!p->out_f32) {
    const int storeC = BN < 64 ? BN : 64;
    const cuuint32_t ybox[4] = {(cuuint32_t)storeC, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    const long long ys = p->y_pix_stride * 2;
    if (p->upsample2x) {
      const int W2 = 2 * Wo, H2 = 2 * Ho;
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const cuuint64_t dims[4] = {(cuuint64_t)p->Cout_store, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)p->N};
          const cuuint64_t str[3] = {(cuuint64_t)ys * 2, (cuuint64_t)ys * W2 * 2, (cuuint64_t)ys * W2 * H2};
          uint8_t* base = reinterpret_cast<uint8_t*>(p->y) + ((long long)ph * W2 + pw) * ys;
          if ((rc = encode_map(&tm.y[ph * 2 + pw], base, 4, dims, str, ybox, storeC * 2, "Y/up"))) return rc;
        }
    } else {
      const cuuint64_t dims[4] = {(cuuint64_t)p->Cout_store, (cuuint64_t)gW, (cuuint64_t)gH, (cuuint64_t)gN};
      const cuuint64_t str[3] = {(cuuint64_t)ys, (cuuint64_t)ys * gW, (cuuint64_t)ys * gW * gH};
      if ((rc = encode_map(&tm.y[0], p->y, 4, dims, str, ybox, storeC * 2, "Y"))) return rc;
    }
  }

  ka.tw = tw; ka.th = th; ka.tn = tn;
  ka.tiles_w = ceil_div(gW, tw); ka.tiles_h = ceil_div(gH, th); ka.tiles_b = ceil_div(gN, tn);
  ka.n_blocks = ceil_div(p->Cout_store, BN);
  const long long nt = m_tiles * ka.n_blocks;
  DYK_REQUIRE(nt < (1ll << 31), "dyk_conv2d_fwd: too many tiles");
  ka.num_tiles = (int)nt;
  ka.k_chunks = ceil_div(p->Cin, BK);
  ka.kh = p->kh; ka.kw = p->kw; ka.stride = flat ? 1 : p->stride; ka.pad = p->pad;
  ka.Ho = gH; ka.Wo = gW; ka.N = gN;
  ka.Cout_store = p->Cout_store;
  ka.act = p->act;
  ka.upsample2x = p->upsample2x;
  ka.out_f32 = p->out_f32;
  ka.y_f32 = p->y;
  ka.y_pix_stride = p->y_pix_stride;
  ka.scale = p->scale; ka.bias = p->bias;
  ka.res = p->res; ka.res_pix_stride = p->res_pix_stride;

  const bool bf = p->dtype == DYK_BF16;
  if (BK == 64) return bf ? dispatch_n<64, true>(BN, tm, ka, stream) : dispatch_n<64, false>(BN, tm, ka, stream);
  return bf ? dispatch_n<32, true>(BN, tm, ka, stream) : dispatch_n<32, false>(BN, tm, ka, stream);
}
