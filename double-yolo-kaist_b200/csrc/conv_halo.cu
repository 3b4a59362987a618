// 3x3 stride-1 pad-1 convolution (+ folded BN, activation, optional residual) as a persistent tcgen05 implicit GEMM
// that loads every input pixel ONCE per (tile, 64-channel chunk): the "halo" kernel.
//
// Replaces the same reference calls as conv_tc.cu (nn.Conv2d -> nn.BatchNorm2d -> activation, models.py:28-64, and
// the unweighted [shortcut] that follows, build_utils/layers.py:63-85) for the 3x3 / stride-1 layers, which carry
// ~90 % of the FLOPs of the Darknet53 / CSPDarknet53 models.
//
// Why: conv_tc.cu re-fetches the 128-pixel A tile from L2 for each of the nine filter taps, so a 128x256 tile needs
// 48 KB of operands per 512 MMA cycles (96 B/clk/SM) while L2 -> SM delivers ~50 B/clk/SM: the tensor pipe idles ~45 %
// of the time (role-cycle counters, profiles/).  Here:
//   * a sub-tile is 8 (w) x 16 (h) output pixels of one image; ONE TMA box {64 ch, 10, 18} brings its input window
//     including the 1-pixel halo (180 pixel rows of 128 B, SWIZZLE_128B; conv padding = TMA out-of-bounds zero fill);
//   * filter tap (r, s) is not another load but another *view* of that box: the A descriptor starts at pixel row
//     r*10 + s and uses a stride of 10 rows between 8-row core-matrix groups, so MMA row m = h*8 + w reads pixel
//     (h + r, w + s).  The 128B swizzle is a function of the absolute shared-memory address, so a view shifted by
//     whole 128-byte rows stays consistent with what TMA wrote (verified on B200: tools/exp/halo_desc.cu);
//   * a CTA tile is kSub (1 or 2) sub-tiles x BLOCK_N (64 / 128) channels: with kSub = 2 the weight tile of a tap is
//     reused by two 128-row MMAs, i.e. per 512 MMA cycles the SM ingests 16 KB of weights + 47 KB / 9 of activations
//     (~42 B/clk/SM) instead of 96.
// Pipelines: A ring (2 slots, one per 64-channel chunk) and B ring (one slot per tap), fp32 accumulators double
// buffered in TMEM (2 x kSub x BLOCK_N columns), epilogue as in conv_tc.cu (8 independent warps, own TMA stores).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue.
#include "common.h"
#include "ptx.cuh"
#include "act.cuh"
#include "conv_common.cuh"
#include <cstdlib>
#include <cstring>

namespace dyk {

struct HaloTmaps {
  CUtensorMap a;  // input  (Cin, W, H, N), box {64, 10, 18, 1}
  CUtensorMap b;  // packed weights (Cin, 9, Cout), box {64, 1, BLOCK_N}
  CUtensorMap y;  // output (Cout_store, W, H, N), box {32, 8, 4, 1}
};

struct HaloKArgs {
  int H, W, N;
  int num_subs;                 // subs_w * subs_h * N
  int n_blocks, num_tiles;      // tiles = ceil(num_subs / kSub) * n_blocks, n-block fastest
  int k_chunks;                 // ceil(Cin / 64)
  int k_last_steps;             // K = 16 MMA steps of the last chunk (Cin = 32: 2 — the zero-filled half of the box is skipped)
  int Cout_store;
  int act;
  FastDiv fd_nblocks, fd_subs_w, fd_subs_h;
  const float* scale;           // may be null
  const float* bias;            // may be null
  const void* res;              // may be null
  long long res_pix_stride;
  unsigned long long* prof;     // role-cycle counters (-DDYK_CONV_PROFILE builds only), may be null
  int grid_cap;                 // SM budget of this launch (sm_budget)
  int res_pf;                   // residual L2 prefetch mode (conv_common.cuh)
};

#ifdef DYK_CONV_PROFILE
constexpr bool kHProf = true;
#else
constexpr bool kHProf = false;
#endif
#define HPROF_T0() ((kHProf && p.prof) ? clock64() : 0ll)
#define HPROF_ADD(var, t0) do { if (kHProf && p.prof) var += clock64() - (t0); } while (0)

constexpr int kHaloEpiWarps = 8;
constexpr int kHaloThreads = 64 + kHaloEpiWarps * 32;
constexpr int kSubW = 8, kSubH = 16;                      // output pixels of a sub-tile (128 GEMM rows)
constexpr int kHaloW = kSubW + 2, kHaloH = kSubH + 2;     // input window
constexpr int kHaloRows = kHaloW * kHaloH;                // 180
constexpr int kSubBytes = 23552;                          // 180 * 128 B rounded up to 1024
constexpr int kAStages = 2;

template <int BLOCK_N, int kSub>
struct HaloSmem {
  static constexpr int kASlot = kSub * kSubBytes;
  static constexpr int kBSlot = BLOCK_N * 128;
  static constexpr int kStagingBytes = kHaloEpiWarps * 2 * 2048;
  static constexpr int kVecFloats = 2 * BLOCK_N;                   // per warp: scale[BLOCK_N] + bias[BLOCK_N]
  static constexpr int kVecBytes = kHaloEpiWarps * kVecFloats * 4;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kBudget = 227 * 1024 - kAStages * kASlot - kStagingBytes - kVecBytes - kBarrierBytes - 1024;
  static constexpr int kBStagesRaw = kBudget / kBSlot;
  static constexpr int kBStages = kBStagesRaw > 9 ? 9 : kBStagesRaw;
  static constexpr int kTotal = kAStages * kASlot + kBStages * kBSlot + kStagingBytes + kVecBytes + kBarrierBytes + 1024;
  static_assert(kBStages >= 4, "B ring too shallow");
};

struct SubCoord {
  int w0, h0, n;   // first output pixel of the sub-tile; n >= N marks a padding sub-tile
};
__device__ __forceinline__ SubCoord sub_coord(const HaloKArgs& p, unsigned sub) {
  SubCoord c;
  const unsigned rowt = fd_div(sub, p.fd_subs_w);
  const unsigned sw = sub - rowt * p.fd_subs_w.div;
  const unsigned n = fd_div(rowt, p.fd_subs_h);
  const unsigned sh = rowt - n * p.fd_subs_h.div;
  c.w0 = sw * kSubW;
  c.h0 = sh * kSubH;
  c.n = sub < (unsigned)p.num_subs ? (int)n : p.N;   // out-of-range sub-tiles address image N (fully out of bounds)
  return c;
}

template <int BLOCK_N, int kSub, bool kBf16, int kAct>
__device__ __forceinline__ void halo_epilogue(const HaloTmaps& tm, const HaloKArgs& p, const SubCoord& sc, int n_base,
                                              uint32_t t_row, uint8_t* wstage, float* wvec, int& sbuf,
                                              uint64_t* tempty, int q, int lane, int c_begin, int c_step, int& staged_base) {
  constexpr int kCols = 32;
  constexpr int kChunks = BLOCK_N / kCols;
  const int row = q * 32 + lane;
  const int wo = sc.w0 + (row & (kSubW - 1)), ho = sc.h0 + (row >> 3);
  const bool sub_ok = sc.n < p.N;
  const bool pix_ok = sub_ok && (wo < p.W) && (ho < p.H);
  const long long pix = (static_cast<long long>(sc.n) * p.H + ho) * p.W + wo;
  const bool has_res = p.res != nullptr;

  uint4 rres[kCols / 8];
  auto load_res = [&](int c) {
    const int c0 = n_base + c * kCols;
    const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.res) + (pix * p.res_pix_stride + c0) * 2;
#pragma unroll
    for (int j = 0; j < kCols / 8; ++j) {
      rres[j] = make_uint4(0u, 0u, 0u, 0u);
      if (pix_ok && c0 + j * 8 < p.Cout_store) rres[j] = __ldg(reinterpret_cast<const uint4*>(rp) + j);
    }
  };
  // scale / bias of the whole n-block -> warp-private shared memory (only when the channel block changes)
  if (n_base != staged_base) {
    staged_base = n_base;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < BLOCK_N / 32; ++j) {
      const int col = n_base + j * 32 + lane;
      wvec[j * 32 + lane] = p.scale ? __ldg(p.scale + col) : 1.f;
      wvec[BLOCK_N + j * 32 + lane] = p.bias ? __ldg(p.bias + col) : 0.f;
    }
    __syncwarp();
  }

#pragma unroll 1
  for (int c = c_begin; c < kChunks; c += c_step) {
    const int cl = c * kCols;
    const int cg0 = n_base + cl;
    const bool beyond = cg0 >= p.Cout_store || !sub_ok;   // warp-uniform
    const bool last = (c + c_step >= kChunks) || (cg0 + c_step * kCols >= p.Cout_store);
    if (beyond) {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      break;
    }
    if (has_res) load_res(c);
    uint32_t v[kCols];
    tmem_ld_32x32b_x32(t_row + cl, v);
    tmem_ld_wait();
    if (last) {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
    uint8_t* sbase = wstage + sbuf * 2048;
    if (lane == 0) tma_store_wait_read<1>();
    __syncwarp();
#pragma unroll
    for (int ch = 0; ch < kCols / 8; ++ch) {
      float o[8];
      const float4 sc0 = *reinterpret_cast<const float4*>(wvec + cl + ch * 8);
      const float4 sc1 = *reinterpret_cast<const float4*>(wvec + cl + ch * 8 + 4);
      const float4 bi0 = *reinterpret_cast<const float4*>(wvec + BLOCK_N + cl + ch * 8);
      const float4 bi1 = *reinterpret_cast<const float4*>(wvec + BLOCK_N + cl + ch * 8 + 4);
      o[0] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 0]), sc0.x, bi0.x));
      o[1] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 1]), sc0.y, bi0.y));
      o[2] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 2]), sc0.z, bi0.z));
      o[3] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 3]), sc0.w, bi0.w));
      o[4] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 4]), sc1.x, bi1.x));
      o[5] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 5]), sc1.y, bi1.y));
      o[6] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 6]), sc1.z, bi1.z));
      o[7] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 7]), sc1.w, bi1.w));
      if (has_res) {
        const uint32_t rr[4] = {rres[ch].x, rres[ch].y, rres[ch].z, rres[ch].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2<kBf16>(rr[e]);
          o[e * 2 + 0] += f.x;
          o[e * 2 + 1] += f.y;
        }
      }
      const int phys = ch ^ ((lane >> 1) & 3);   // SWIZZLE_64B of the 64-byte staging rows
      const uint4 val = make_uint4(pack2<kBf16>(o[0], o[1]), pack2<kBf16>(o[2], o[3]), pack2<kBf16>(o[4], o[5]),
                                   pack2<kBf16>(o[6], o[7]));
      *reinterpret_cast<uint4*>(sbase + lane * 64 + phys * 16) = val;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(&tm.y, sbase, cg0, sc.w0, sc.h0 + 4 * q, sc.n);
      tma_store_commit();
    }
    sbuf ^= 1;
    if (last) break;
  }
}

template <int BLOCK_N, int kSub, bool kBf16>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ HaloTmaps tm, const HaloKArgs p) {
  using S = HaloSmem<BLOCK_N, kSub>;
  constexpr int kBStages = S::kBStages;
  constexpr uint32_t kAccCols = kSub * BLOCK_N;          // TMEM columns of one accumulator stage
  constexpr uint32_t kTmemCols = 2 * kAccCols;           // 128 / 256 / 512
  static_assert(BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");
  static_assert(kTmemCols <= 512, "accumulators do not fit TMEM");
  static_assert(kSub == 1 || kSub == 2, "kSub");

  griddep_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_base = smem;
  uint8_t* b_base = a_base + kAStages * S::kASlot;
  uint8_t* staging = b_base + kBStages * S::kBSlot;
  float* vecs = reinterpret_cast<float*>(staging + S::kStagingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(vecs) + S::kVecBytes);
  uint64_t* a_full = bars;                          // [kAStages]
  uint64_t* a_empty = a_full + kAStages;            // [kAStages]
  uint64_t* b_full = a_empty + kAStages;            // [kBStages]
  uint64_t* b_empty = b_full + kBStages;            // [kBStages]
  uint64_t* tfull = b_empty + kBStages;             // [2]
  uint64_t* tempty = tfull + 2;                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  // warp index through a shuffle (as CUTLASS' canonical_warp_idx_sync): the compiler then knows it is warp-uniform and keeps
  // everything derived from it (role, TMEM lane quarter, staging addresses, TMA-store operands) in uniform registers
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.b);
    tma_prefetch_desc(&tm.y);
    for (int i = 0; i < kAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kHaloEpiWarps); }
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // (same value in every lane: tell the compiler)
  griddep_wait();   // PDL: everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      long long t_wait = 0;
      const long long t_begin = HPROF_T0();
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const unsigned mt = fd_div((unsigned)tile, p.fd_nblocks);
        const int nblk = tile - (int)(mt * p.fd_nblocks.div);
        SubCoord sc[kSub];
#pragma unroll
        for (int j = 0; j < kSub; ++j) sc[j] = sub_coord(p, mt * kSub + j);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          { const long long t0 = HPROF_T0(); mbar_wait(&a_empty[as], aph ^ 1); HPROF_ADD(t_wait, t0); }
          uint8_t* sa = a_base + as * S::kASlot;
          if (elect_one()) {
            mbar_arrive_expect_tx(&a_full[as], kSub * kHaloRows * 128);
#pragma unroll
            for (int j = 0; j < kSub; ++j)
              tma_load_4d(sa + j * kSubBytes, &tm.a, &a_full[as], kc * 64, sc[j].w0 - 1, sc[j].h0 - 1, sc[j].n);
          }
          __syncwarp();
          if (++as == kAStages) { as = 0; aph ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            { const long long t0 = HPROF_T0(); mbar_wait(&b_empty[bs], bph ^ 1); HPROF_ADD(t_wait, t0); }
            if (elect_one()) {
              mbar_arrive_expect_tx(&b_full[bs], S::kBSlot);
              tma_load_3d(b_base + bs * S::kBSlot, &tm.b, &b_full[bs], kc * 64, tap, nblk * BLOCK_N);
            }
            __syncwarp();
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
          }
        }
      }
      if (kHProf && p.prof && lane == 0) {
        atomicAdd(p.prof + 0, (unsigned long long)t_wait);
        atomicAdd(p.prof + 1, (unsigned long long)(clock64() - t_begin));
        atomicAdd(p.prof + 7, 1ull);
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer: whole warp in the loop, one elected
    // lane per tcgen05 instruction (uniform-register operands; see conv_halo2.cu)
    {
      constexpr uint32_t idesc = umma_idesc_f16(128, BLOCK_N, kBf16 ? 1 : 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int tl = 0;
      long long t_wdata = 0, t_wacc = 0;
      const long long t_begin = HPROF_T0();
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
        const int acc = tl & 1;
        { const long long t0 = HPROF_T0(); mbar_wait(&tempty[acc], ((tl >> 1) & 1) ^ 1); HPROF_ADD(t_wacc, t0); }
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          { const long long t0 = HPROF_T0(); mbar_wait(&a_full[as], aph); HPROF_ADD(t_wdata, t0); }
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(a_base + as * S::kASlot);
          for (int tap = 0; tap < 9; ++tap) {
            { const long long t0 = HPROF_T0(); mbar_wait(&b_full[bs], bph); HPROF_ADD(t_wdata, t0); }
            tc_fence_after_sync();
            const uint64_t bdesc = umma_desc_kmajor<128>(smem_u32(b_base + bs * S::kBSlot));
            const int r = tap / 3, s = tap - 3 * r;
            const int ksteps = kc == p.k_chunks - 1 ? p.k_last_steps : 4;
            const bool el = elect_one();
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
              // shifted view of the halo tile: first row r*10 + s, 10 rows between 8-row groups
              uint64_t adesc = 0;
              adesc |= static_cast<uint64_t>(((sa + j * kSubBytes + (r * kHaloW + s) * 128) & 0x3FFFF) >> 4);
              adesc |= static_cast<uint64_t>(1) << 16;
              adesc |= static_cast<uint64_t>((kHaloW * 128) >> 4) << 32;
              adesc |= static_cast<uint64_t>(1) << 46;
              adesc |= static_cast<uint64_t>(2) << 61;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (el && k < ksteps)
                  umma_f16_ss(d_tmem + j * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
            }
            if (el) umma_commit(&b_empty[bs]);
            __syncwarp();
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
          }
          if (elect_one()) {
            umma_commit(&a_empty[as]);
            if (kc == p.k_chunks - 1) umma_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++as == kAStages) { as = 0; aph ^= 1; }
        }
      }
      if (kHProf && p.prof && lane == 0) {
        atomicAdd(p.prof + 2, (unsigned long long)t_wdata);
        atomicAdd(p.prof + 3, (unsigned long long)t_wacc);
        atomicAdd(p.prof + 4, (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp_idx - 2;
    const int q = warp_idx & 3;                 // TMEM lane quarter this warp may access
    const int grp = ew >> 2;                    // kSub == 2: sub-tile; kSub == 1: column-chunk parity
    const int j = kSub == 2 ? grp : 0;
    const int c_begin = kSub == 2 ? 0 : grp;
    const int c_step = kSub == 2 ? 1 : 2;
    uint8_t* wstage = staging + ew * 4096;
    float* wvec = vecs + ew * S::kVecFloats;
    int tl = 0, sbuf = 0;
    int staged_base = -1;
    long long t_wtfull = 0;
    const long long t_begin = HPROF_T0();
    // residual rows of the tiles this CTA reaches kResPrefetchTiles iterations from now -> L2 (conv_common.cuh)
    auto prefetch_res = [&](int t) {
      if (p.res_pf && p.res != nullptr && t < p.num_tiles && (kSub == 2 || grp == 0)) {
        const unsigned mt = fd_div((unsigned)t, p.fd_nblocks);
        const int nb = (t - (int)(mt * p.fd_nblocks.div)) * BLOCK_N;
        const SubCoord f = sub_coord(p, mt * kSub + j);
        const int row = q * 32 + lane;
        const int wo = f.w0 + (row & (kSubW - 1)), ho = f.h0 + (row >> 3);
        if (f.n < p.N && wo < p.W && ho < p.H) {
          const long long pix = (static_cast<long long>(f.n) * p.H + ho) * p.W + wo;
          const int cols = p.Cout_store - nb < BLOCK_N ? p.Cout_store - nb : BLOCK_N;
          l2_prefetch_row(reinterpret_cast<const uint8_t*>(p.res) + (pix * p.res_pix_stride + nb) * 2, (unsigned)cols * 2u, p.res_pf);
        }
      }
    };
    for (int d = 1; d < kResPrefetchTiles; ++d) prefetch_res((int)blockIdx.x + d * (int)gridDim.x);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      prefetch_res(tile + kResPrefetchTiles * (int)gridDim.x);
      const int acc = tl & 1;
      const unsigned mt = fd_div((unsigned)tile, p.fd_nblocks);
      const int nblk = tile - (int)(mt * p.fd_nblocks.div);
      const SubCoord sc = sub_coord(p, mt * kSub + j);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + j * BLOCK_N;
      { const long long t0 = HPROF_T0(); mbar_wait(&tfull[acc], (tl >> 1) & 1); HPROF_ADD(t_wtfull, t0); }
      tc_fence_after_sync();
#define DYK_HEPI(ACT)                                                                                          \
  halo_epilogue<BLOCK_N, kSub, kBf16, ACT>(tm, p, sc, nblk * BLOCK_N, t_row, wstage, wvec, sbuf, &tempty[acc], q, \
                                           lane, c_begin, c_step, staged_base)
      switch (p.act) {
        case DYK_ACT_LEAKY: DYK_HEPI(DYK_ACT_LEAKY); break;
        case DYK_ACT_MISH: DYK_HEPI(DYK_ACT_MISH); break;
        case DYK_ACT_RELU: DYK_HEPI(DYK_ACT_RELU); break;
        case DYK_ACT_RELU6: DYK_HEPI(DYK_ACT_RELU6); break;
        case DYK_ACT_HARDSWISH: DYK_HEPI(DYK_ACT_HARDSWISH); break;
        case DYK_ACT_HARDSIGMOID: DYK_HEPI(DYK_ACT_HARDSIGMOID); break;
        default: DYK_HEPI(DYK_ACT_LINEAR); break;
      }
#undef DYK_HEPI
    }
    if (lane == 0) tma_store_wait_all<0>();
    if (kHProf && p.prof && ew == 0 && lane == 0) {
      atomicAdd(p.prof + 5, (unsigned long long)t_wtfull);
      atomicAdd(p.prof + 6, (unsigned long long)(clock64() - t_begin));
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------- host
extern unsigned long long* g_conv_prof;   // conv_tc.cu (dyk_conv_set_profile)
int encode_map_generic(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes, const char* what);

template <int BLOCK_N, int kSub, bool kBf16>
static int launch_halo(const HaloTmaps& tm, const HaloKArgs& ka, cudaStream_t stream) {
  using S = HaloSmem<BLOCK_N, kSub>;
  auto kern = conv3x3_halo_kernel<BLOCK_N, kSub, kBf16>;
  static bool configured = false;
  if (!configured) {
    DYK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  const int grid = ka.num_tiles < ka.grid_cap ? ka.num_tiles : ka.grid_cap;
  DYK_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(kHaloThreads), S::kTotal, stream, tm, ka));
  DYK_LAUNCH_OK("conv3x3_halo_kernel");
  return DYK_OK;
}

// Returns DYK_OK after launching, or 1 when the layer is not eligible (the caller then uses conv_tc_kernel).
int conv3x3_halo_try(const dyk_conv_params* p, cudaStream_t stream) {
  if (!(p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && !p->upsample2x && !p->out_f32 && !p->y_plane &&
        p->out_h == 0 && p->out_w == 0))
    return 1;
  if (p->Cout_store < 64 || p->Cin < 16) return 1;
  // Measured on B200 (profiles/): the operand path is bound by shared-memory bandwidth (~100 B/clk/SM effective for
  // MMA operand reads + TMA writes), so removing the 9x re-load of A pays off when the A share is large or the layer
  // is deep enough to amortise the per-tile cost: Cin >= 256 (+9..12 %) and Cin <= 32 (+40 %).  In between the
  // generic kernel is as fast or faster.  DYK_HALO=all forces the halo kernel wherever it is eligible.
  static const bool halo_all = getenv("DYK_HALO") != nullptr && !strcmp(getenv("DYK_HALO"), "all");
  if (!halo_all && !((p->Cin >= 256 && p->Cout_store >= 256) || p->Cin <= 32)) return 1;
  const int H = p->H, W = p->W, N = p->N;
  const int subs_w = ceil_div(W, kSubW), subs_h = ceil_div(H, kSubH);
  const long long num_subs = (long long)subs_w * subs_h * N;
  if (num_subs >= (1ll << 30)) return 1;
  // spatial waste of the fixed 8x16 sub-tile; tiny maps are better served by the generic kernel's tile picker
  const double eff = (double)W * H / ((double)subs_w * kSubW * subs_h * kSubH);
  if (eff < 0.6) return 1;
  // Shared-memory bandwidth (128 B/clk/SM) is what bounds an SS-mode tcgen05.mma at these tile sizes: every MMA of
  // K = 16 re-reads (128 + N) rows of 32 B, i.e. 96 B/clk at N = 256 and 128 B/clk at N = 128, on top of the TMA
  // writes.  So: the widest N the layer allows; two sub-tiles per CTA only for N <= 128 (TMEM holds 512 columns).
  static const int force_bn = getenv("DYK_HALO_BN") ? atoi(getenv("DYK_HALO_BN")) : 0;
  int BN = p->Cout_store > 128 ? 256 : (p->Cout_store > 64 ? 128 : 64);
  if (force_bn == 64 || force_bn == 128 || force_bn == 256) BN = force_bn;
  const int n_blocks = ceil_div(p->Cout_store, BN);
  const int k_chunks = ceil_div(p->Cin, 64);
  const int kSub = BN == 256 ? 1 : ((ceil_div64(num_subs, 2) * n_blocks >= sm_budget(p->sm_limit)) ? 2 : 1);

  HaloTmaps tm;
  HaloKArgs ka;
  memset(&tm, 0, sizeof(tm));
  memset(&ka, 0, sizeof(ka));
  int rc;
  {
    const long long xs = p->x_pix_stride * 2;
    const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t str[3] = {(cuuint64_t)xs, (cuuint64_t)xs * W, (cuuint64_t)xs * W * H};
    const cuuint32_t box[4] = {64, kHaloW, kHaloH, 1};
    if ((rc = encode_map_generic(&tm.a, p->x, 4, dims, str, box, 128, "halo A"))) return rc;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)p->Cin, 9, (cuuint64_t)p->Cout};
    const cuuint64_t str[2] = {(cuuint64_t)p->Cin * 2, (cuuint64_t)p->Cin * 2 * 9};
    const cuuint32_t box[3] = {64, 1, (cuuint32_t)BN};
    if ((rc = encode_map_generic(&tm.b, p->w, 3, dims, str, box, 128, "halo B"))) return rc;
  }
  {
    const long long ys = p->y_pix_stride * 2;
    const cuuint64_t dims[4] = {(cuuint64_t)p->Cout_store, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t str[3] = {(cuuint64_t)ys, (cuuint64_t)ys * W, (cuuint64_t)ys * W * H};
    const cuuint32_t box[4] = {32, kSubW, 4, 1};
    if ((rc = encode_map_generic(&tm.y, p->y, 4, dims, str, box, 64, "halo Y"))) return rc;
  }
  ka.H = H; ka.W = W; ka.N = N;
  ka.num_subs = (int)num_subs;
  ka.n_blocks = n_blocks;
  ka.num_tiles = (int)(ceil_div64(num_subs, kSub) * n_blocks);
  ka.k_chunks = k_chunks;
  ka.k_last_steps = ceil_div(p->Cin - (k_chunks - 1) * 64, 16);
  ka.Cout_store = p->Cout_store;
  ka.act = p->act;
  ka.fd_nblocks = make_fastdiv((unsigned)n_blocks);
  ka.fd_subs_w = make_fastdiv((unsigned)subs_w);
  ka.fd_subs_h = make_fastdiv((unsigned)subs_h);
  ka.scale = p->scale; ka.bias = p->bias;
  ka.res = p->res; ka.res_pix_stride = p->res_pix_stride;
  ka.prof = g_conv_prof;
  ka.grid_cap = sm_budget(p->sm_limit);
  ka.res_pf = res_prefetch_mode();

  const bool bf = p->dtype == DYK_BF16;
  if (BN == 256) return bf ? launch_halo<256, 1, true>(tm, ka, stream) : launch_halo<256, 1, false>(tm, ka, stream);
  if (BN == 128) {
    if (kSub == 2) return bf ? launch_halo<128, 2, true>(tm, ka, stream) : launch_halo<128, 2, false>(tm, ka, stream);
    return bf ? launch_halo<128, 1, true>(tm, ka, stream) : launch_halo<128, 1, false>(tm, ka, stream);
  }
  if (kSub == 2) return bf ? launch_halo<64, 2, true>(tm, ka, stream) : launch_halo<64, 2, false>(tm, ka, stream);
  return bf ? launch_halo<64, 1, true>(tm, ka, stream) : launch_halo<64, 1, false>(tm, ka, stream);
}

}  // namespace dyk
