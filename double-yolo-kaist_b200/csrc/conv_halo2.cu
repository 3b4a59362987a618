// CTA-pair (tcgen05 cta_group::2) version of the 3x3 stride-1 halo convolution (conv_halo.cu): two SMs of one TPC
// compute a 256-pixel x 256-channel tile together.
//
// Why pairs: an SS-mode tcgen05.mma re-reads its operands from shared memory for every K = 16 step, and the measured
// ceiling of that path is ~100 B/clk/SM for MMA operand reads + TMA writes together (profiles/README.md).  A single CTA
// at 128 x 256 needs 96 B/clk for operand reads alone plus 68 B/clk of TMA writes -> ~60 % tensor pipe.  In a pair each
// CTA holds its own 128 pixel rows (A, as a halo window: loaded once per 64-channel chunk for all nine taps) and HALF of
// the 256-row weight tile (B); the pair's MMA (M = 256) reads A locally and each B half once, so per CTA and tap
// (512 MMA cycles) shared memory moves 32 KB of operand reads + 18.6 KB of TMA writes = 99 B/clk.
//
// Protocol (as in CUTLASS' 2-SM kernels): both CTAs run a TMA producer for their own halves, crediting the bytes to the
// LEADER CTA's "full" barriers; only the leader issues tcgen05.mma.cta_group::2; tcgen05.commit multicasts the "empty"
// and "accumulator full" arrivals to both CTAs; the epilogue warps of both CTAs arrive on the leader's "accumulator
// empty" barrier.  Accumulators: 128 lanes x 256 fp32 columns per CTA, double buffered (all 512 TMEM columns).
//
// Warp roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner (+ MMA issuer in the leader),
// warps 2..9 = epilogue.
//
// Dual-source variant (kDual; WeightedFeatureFusion, build_utils/layers.py:63-85, fused into the convolution that
// consumes it): the A operand is w0 * x + w1 * x2, formed in shared memory.  Each CTA's producer loads the halo windows
// of BOTH tensors (same box, same swizzle, so element i of one slot corresponds to element i of the other) and eight extra
// "combiner" warps overwrite the x slot with the rounded weighted sum — once per 64-channel chunk, for all nine taps —
// then publish it to the MMA issuer (fence.proxy.async + arrive on the leader's "full" barrier).  The fused sum is never
// written to memory: both modality tensors are read once.  The two extra halo slots are paid for with three weight stages.
//
// Resident-weights variant (kResB; the early high-resolution layers: Cin <= 64, 64..128 output channels, e.g. 32->64 @256x320
// and 64->128 @128x160): these layers are HBM-bound (Cin + 2 Cout bytes per pixel against 18 Cin Cout FLOPs) and were
// running at 40 % of the HBM rate for two reasons measured with the role counters / ncu: (1) the generic kernel re-fetches
// the input tile for each of the nine taps and the weight tile for every pixel tile, 288 KB of L2 -> SM traffic per 128
// pixels, which is bound by the ~50 B/clk/SM L2 delivery rate, not by HBM; (2) with two input slots only one tile's load is
// in flight per SM, too few bytes to cover the HBM latency.  Here the whole filter (9 taps x Cout x 128 B, half per CTA of
// the pair: <= 72 KB) is loaded ONCE per CTA and stays in shared memory, the input halo window is the only operand
// traffic (23 KB per 128 pixels), and the freed shared memory holds a ring of four input windows (three loads in flight).
#include "common.h"
#include "ptx.cuh"
#include "act.cuh"
#include "conv_common.cuh"
#include "vec.cuh"
#include <cstdlib>
#include <cstring>

namespace dyk {

int encode_map_generic(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes, const char* what);
extern unsigned long long* g_conv_prof;   // conv_tc.cu (dyk_conv_set_profile)

struct Halo2Tmaps {
  CUtensorMap a;  // input  (Cin, W, H, N), box {64, 10, 18, 1}
  CUtensorMap b;  // packed weights (Cin, 9, Cout), box {64, 1, 128}
  CUtensorMap bs; // the same tensor with box {64, 1, 128 / tail_split} for the split items of the last round
  CUtensorMap y;  // output (Cout_store, W, H, N), box {32, 8, 4, 1}
  CUtensorMap a2; // second source of the dual-source variant (same geometry as a)
};

#ifdef DYK_CONV_PROFILE
constexpr bool kH2Prof = true;
#else
constexpr bool kH2Prof = false;
#endif
// counters (dyk_conv_set_profile): 0 MMA wait data, 1 MMA wait accumulator, 2 MMA issue-loop total, 3 epilogue (warp 2,
// leader) wait accumulator, 4 epilogue total, 5 epilogue cycles inside tcgen05.ld + wait, 6 epilogue cycles waiting for a
// staging buffer (TMA store read), 7 clusters
struct Halo2KArgs {
  unsigned long long* prof;
  int res_pf;                   // residual L2 prefetch mode (conv_common.cuh)
  int cluster_cap;              // SM pairs this launch may occupy (sm_budget / 2)
  int H, W, N;
  int num_subs;
  int n_blocks, num_tiles;      // pair tiles = ceil(num_subs / 2) * n_blocks, n-block fastest
  // Work items: the first full_items tiles are computed whole (N = 256); the tiles of the last, partial round are
  // split into (1 << tail_lg) items of 256 >> tail_lg output channels each, so that the round occupies all clusters for
  // a fraction of a tile time instead of a few clusters for a whole one (the MMA's N is a run-time descriptor field).
  int full_items, tail_lg, num_items;
  int k_chunks;
  int Cout_store;
  int act;
  FastDiv fd_nblocks, fd_subs_w, fd_subs_h;
  const float* scale;
  const float* bias;
  const void* res;
  long long res_pix_stride;
  const float* x_wts_raw;       // dual source: the fusion module's raw parameter w[2]; A = sigmoid(w)[0] * x + sigmoid(w)[1] * x2
  int res_cols;                 // resident-weights variant: output channels (MMA N), multiple of 32, <= 128
  int k_last_steps;             // K = 16 MMA steps of the last 64-channel chunk (Cin = 32: 2)
};

constexpr int kH2EpiWarps = 8;
constexpr int kH2Threads = 64 + kH2EpiWarps * 32;
constexpr int kH2SubW = 8, kH2SubH = 16, kH2HaloW = 10, kH2HaloH = 18;
constexpr int kH2HaloBytes = kH2HaloW * kH2HaloH * 128;   // 23040
constexpr int kH2ASlot = 23552;
constexpr int kH2BSlot = 128 * 128;                       // half of the 256-row weight tile
constexpr int kH2AStages = 2, kH2BStages = 8;
constexpr int kH2BlockN = 256;
constexpr int kH2StagingBytes = kH2EpiWarps * 2 * 2048;
constexpr int kH2VecFloats = 2 * kH2BlockN;
constexpr int kH2VecBytes = kH2EpiWarps * kH2VecFloats * 4;
constexpr int kH2Total = kH2AStages * kH2ASlot + kH2BStages * kH2BSlot + kH2StagingBytes + kH2VecBytes + 1024 + 1024;
static_assert(kH2Total <= 227 * 1024, "halo2 shared memory budget");
constexpr int kH2CombWarps = 8;                                // dual-source variant: combiner warps per CTA
constexpr int kH2BStagesDual = 5;
constexpr int kH2TotalDual = 2 * kH2AStages * kH2ASlot + kH2BStagesDual * kH2BSlot + kH2StagingBytes + kH2VecBytes + 1024 + 1024;
static_assert(kH2TotalDual <= 227 * 1024, "halo2 (dual source) shared memory budget");
constexpr int kH2AStagesRes = 4, kH2BSlotRes = 64 * 128;       // resident weights: 9 taps x (<= 64 rows per CTA) x 128 B
constexpr int kH2TotalRes = kH2AStagesRes * kH2ASlot + 9 * kH2BSlotRes + kH2StagingBytes + kH2VecBytes + 1024 + 1024;
static_assert(kH2TotalRes <= 227 * 1024, "halo2 (resident weights) shared memory budget");
constexpr int kH2MaxAStages = 4, kH2MaxBStages = 9;

struct Sub2 {
  int w0, h0, n;
};
struct Item2 {
  int tile, ncol0, ncols;
};
__device__ __forceinline__ Item2 item2_decode(const Halo2KArgs& p, int item) {
  if (p.res_cols) return Item2{item, 0, p.res_cols};
  if (item < p.full_items) return Item2{item, 0, kH2BlockN};
  const int j = item - p.full_items;
  const int ncols = kH2BlockN >> p.tail_lg;
  return Item2{p.full_items + (j >> p.tail_lg), (j & ((1 << p.tail_lg) - 1)) * ncols, ncols};
}
__device__ __forceinline__ Sub2 sub2_coord(const Halo2KArgs& p, unsigned sub) {
  Sub2 c;
  const unsigned rowt = fd_div(sub, p.fd_subs_w);
  const unsigned sw = sub - rowt * p.fd_subs_w.div;
  const unsigned n = fd_div(rowt, p.fd_subs_h);
  const unsigned sh = rowt - n * p.fd_subs_h.div;
  c.w0 = sw * kH2SubW;
  c.h0 = sh * kH2SubH;
  c.n = sub < (unsigned)p.num_subs ? (int)n : p.N;
  return c;
}

template <bool kBf16, int kAct, bool kNoRes>
__device__ __forceinline__ void halo2_epilogue(const Halo2Tmaps& tm, const Halo2KArgs& p, const Sub2& sc, int n_base,
                                               int nchunks, uint32_t t_row, uint8_t* wstage, float* wvec, int& sbuf,
                                               uint32_t tempty_leader, int q, int lane, int half, bool prof_warp,
                                               long long& prof_ld, long long& prof_st, const uint4 (&rres_all)[4][4],
                                               int& staged_base) {
  constexpr int kCols = 32;
  constexpr int kChunks = kH2BlockN / kCols;   // up to 8 (nchunks of them in this item); this warp handles chunks half, half+2, ...
  const bool sub_ok = sc.n < p.N;
  const bool has_res = !kNoRes && p.res != nullptr;   // (the dual-source variant never carries a residual)

  // scale / bias of this item's output channels -> warp-private shared memory; only when they differ from what is staged
  // (never again on layers with a single channel block: the 16 global loads + syncs per tile were a fifth of the
  // epilogue's instructions on the short tiles of the early layers)
  if (n_base * 16 + nchunks != staged_base) {
    staged_base = n_base * 16 + nchunks;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kH2BlockN / 32; ++j) {
      if (j < nchunks) {
        const int col = n_base + j * 32 + lane;
        wvec[j * 32 + lane] = p.scale ? __ldg(p.scale + col) : 1.f;
        wvec[kH2BlockN + j * 32 + lane] = p.bias ? __ldg(p.bias + col) : 0.f;
      }
    }
    __syncwarp();
  }

#pragma unroll
  for (int ci = 0; ci < kChunks / 2; ++ci) {
    const int c = half + 2 * ci;
    const int cl = c * kCols;
    const int cg0 = n_base + cl;
    const bool beyond = cg0 >= p.Cout_store || !sub_ok;
    const bool last = (c + 2 >= nchunks) || (cg0 + 2 * kCols >= p.Cout_store);
    if (beyond) {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_leader);
      break;
    }
    uint32_t v[kCols];
    const long long t_ld0 = (kH2Prof && p.prof) ? clock64() : 0;
    tmem_ld_32x32b_x32(t_row + cl, v);
    tmem_ld_wait();
    if (kH2Prof && p.prof && prof_warp) prof_ld += clock64() - t_ld0;
    if (last) {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_leader);
    }
    uint8_t* sbase = wstage + sbuf * 2048;
    const long long t_st0 = (kH2Prof && p.prof) ? clock64() : 0;
    if (lane == 0) tma_store_wait_read<1>();
    __syncwarp();
    if (kH2Prof && p.prof && prof_warp) prof_st += clock64() - t_st0;
#pragma unroll
    for (int ch = 0; ch < kCols / 8; ++ch) {
      float o[8];
      const float4 sc0 = *reinterpret_cast<const float4*>(wvec + cl + ch * 8);
      const float4 sc1 = *reinterpret_cast<const float4*>(wvec + cl + ch * 8 + 4);
      const float4 bi0 = *reinterpret_cast<const float4*>(wvec + kH2BlockN + cl + ch * 8);
      const float4 bi1 = *reinterpret_cast<const float4*>(wvec + kH2BlockN + cl + ch * 8 + 4);
      o[0] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 0]), sc0.x, bi0.x));
      o[1] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 1]), sc0.y, bi0.y));
      o[2] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 2]), sc0.z, bi0.z));
      o[3] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 3]), sc0.w, bi0.w));
      o[4] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 4]), sc1.x, bi1.x));
      o[5] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 5]), sc1.y, bi1.y));
      o[6] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 6]), sc1.z, bi1.z));
      o[7] = act_apply<kAct>(fmaf(__uint_as_float(v[ch * 8 + 7]), sc1.w, bi1.w));
      if (has_res) {
        const uint4 rv = rres_all[ci][ch];
        const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2<kBf16>(rr[e]);
          o[e * 2 + 0] += f.x;
          o[e * 2 + 1] += f.y;
        }
      }
      const int phys = ch ^ ((lane >> 1) & 3);
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

template <bool kBf16, bool kDual, bool kResB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kH2Threads + (kDual ? 32 * kH2CombWarps : 0), 1)
conv3x3_halo2_kernel(const __grid_constant__ Halo2Tmaps tm, const Halo2KArgs p) {
  static_assert(!(kDual && kResB), "variants are exclusive");
  constexpr int kBStages = kResB ? 9 : (kDual ? kH2BStagesDual : kH2BStages);
  constexpr int kAStg = kResB ? kH2AStagesRes : kH2AStages;
  constexpr int kBSlotBytes = kResB ? kH2BSlotRes : kH2BSlot;
  griddep_launch_dependents();
  // profile build: wall-clock (globaltimer, ns) of kernel entry / prologue end / role-loop end / exit, min and max over CTAs
  auto stamp = [&](int lo, int hi) {
    if (kH2Prof && p.prof && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (lo >= 0) atomicMin(p.prof + lo, t);
      if (hi >= 0) atomicMax(p.prof + hi, t);
    }
  };
  stamp(8, 9);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_base = smem;
  uint8_t* a2_base = a_base + kAStg * kH2ASlot;                             // dual source only
  uint8_t* b_base = a_base + (kDual ? 2 : 1) * kAStg * kH2ASlot;
  uint8_t* staging = b_base + kBStages * kBSlotBytes;
  float* vecs = reinterpret_cast<float*>(staging + kH2StagingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(vecs) + kH2VecBytes);
  uint64_t* a_full = bars;                          // [2]  (used in the leader)
  uint64_t* a_empty = a_full + kH2MaxAStages;       // [2]  (per CTA, multicast arrivals)
  uint64_t* b_full = a_empty + kH2MaxAStages;       // [8]  (leader)
  uint64_t* b_empty = b_full + kH2MaxBStages;       // [8]  (per CTA)
  uint64_t* tfull = b_empty + kH2MaxBStages;        // [2]  (per CTA, multicast arrivals)
  uint64_t* tempty = tfull + 2;                     // [2]  (leader; 16 arrivals)
  uint64_t* a_raw = tempty + 2;                     // [2]  (per CTA; dual source: both halo windows landed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_raw + kH2MaxAStages);

  // warp index through a shuffle (as CUTLASS' canonical_warp_idx_sync): the compiler then knows it is warp-uniform and keeps
  // everything derived from it (role, TMEM lane quarter, staging addresses, TMA-store operands) in uniform registers
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.b);
    tma_prefetch_desc(&tm.bs);
    tma_prefetch_desc(&tm.y);
    if (kDual) tma_prefetch_desc(&tm.a2);
    // dual source: "full" = the combiner warps of both CTAs have published the weighted sum (no TMA bytes credited)
    for (int i = 0; i < kAStg; ++i) {
      mbar_init(&a_full[i], kDual ? 2 * kH2CombWarps : 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&a_raw[i], 1);
    }
    for (int i = 0; i < kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * kH2EpiWarps); }
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();         // (the cluster barrier below already orders the CTA; this one is what compute-sanitizer models)
  cluster_sync_all();      // barriers of both CTAs initialised before any remote arrive / TMA credit
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // (same value in every lane: tell the compiler)
  griddep_wait();   // PDL: everything above overlapped the previous kernel's tail; its results are visible from here on
  stamp(-1, 10);

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, own halves; the whole
    // warp walks the loop, one elected lane issues)
    {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      if (kResB && lane == 0) {
        // the whole filter, once: tap t -> slot t (this CTA's half of the output channels)
        const uint32_t b_bytes = (uint32_t)p.res_cols * 128u;
        for (int tap = 0; tap < 9; ++tap) {
          if (leader) mbar_arrive_expect_tx(&b_full[tap], b_bytes);
          tma_load_3d_2sm(b_base + tap * kBSlotBytes, &tm.bs, leader_smem_addr(&b_full[tap]), 0, tap, (int)rank * (p.res_cols >> 1));
        }
      }
      for (int item = cluster_id; item < p.num_items; item += num_clusters) {
        const Item2 it = item2_decode(p, item);
        const int tile = it.tile;
        const unsigned mt = fd_div((unsigned)tile, p.fd_nblocks);
        const int nblk = tile - (int)(mt * p.fd_nblocks.div);
        const Sub2 sc = sub2_coord(p, mt * 2 + rank);
        const bool whole = it.ncols == kH2BlockN;
        const CUtensorMap* bmap = whole ? &tm.b : &tm.bs;
        const uint32_t b_bytes = (uint32_t)it.ncols * 128u;              // both halves: ncols rows of 128 B
        const int b_row = nblk * kH2BlockN + it.ncol0 + (int)rank * (it.ncols >> 1);
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(&a_empty[as], aph ^ 1);
          if (elect_one()) {
            if constexpr (kDual) {
              mbar_arrive_expect_tx(&a_raw[as], 2 * kH2HaloBytes);
              tma_load_4d(a_base + as * kH2ASlot, &tm.a, &a_raw[as], kc * 64, sc.w0 - 1, sc.h0 - 1, sc.n);
              tma_load_4d(a2_base + as * kH2ASlot, &tm.a2, &a_raw[as], kc * 64, sc.w0 - 1, sc.h0 - 1, sc.n);
            } else {
              if (leader) mbar_arrive_expect_tx(&a_full[as], 2 * kH2HaloBytes);
              tma_load_4d_2sm(a_base + as * kH2ASlot, &tm.a, leader_smem_addr(&a_full[as]), kc * 64, sc.w0 - 1, sc.h0 - 1, sc.n);
            }
          }
          __syncwarp();
          if (++as == kAStg) { as = 0; aph ^= 1; }
          if constexpr (!kResB) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_empty[bs], bph ^ 1);
              if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(&b_full[bs], b_bytes);
                tma_load_3d_2sm(b_base + bs * kBSlotBytes, bmap, leader_smem_addr(&b_full[bs]), kc * 64, tap, b_row);
              }
              __syncwarp();
              if (++bs == kBStages) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // The WHOLE warp walks the loop (waits included) and one elected lane issues each tcgen05 instruction.  With a single
    // thread inside `if (lane == 0)` every operand reached the MMA through a per-lane loop (ELECT + 5 R2UR.BROADCAST +
    // BRA.U.ANY, ~24 SASS instructions per MMA, ncu source view): the issue loop itself ran at ~150 cycles per MMA and
    // paced the tensor pipe (128 cycles per 256x256x16 pair MMA).  Warp-uniform control flow keeps descriptors and
    // barrier addresses in uniform registers.
    if (leader) {
      const uint32_t idesc_whole = umma_idesc_f16(256, kH2BlockN, kBf16 ? 1 : 0);
      const uint32_t idesc_split = umma_idesc_f16(256, kResB ? p.res_cols : (kH2BlockN >> p.tail_lg), kBf16 ? 1 : 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int tl = 0;
      long long w_data = 0, w_acc = 0;
      const long long t_begin = (kH2Prof && p.prof) ? clock64() : 0;
#define H2WAIT(bar, ph, var) do { const long long t0 = (kH2Prof && p.prof) ? clock64() : 0; mbar_wait(bar, ph); \
                                  if (kH2Prof && p.prof) var += clock64() - t0; } while (0)
      for (int item = cluster_id; item < p.num_items; item += num_clusters, ++tl) {
        const uint32_t idesc = (!kResB && item < p.full_items) ? idesc_whole : idesc_split;
        const int acc = tl & 1;
        H2WAIT(&tempty[acc], ((tl >> 1) & 1) ^ 1, w_acc);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kH2BlockN;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          H2WAIT(&a_full[as], aph, w_data);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(a_base + as * kH2ASlot);
          const int ksteps = kc == p.k_chunks - 1 ? p.k_last_steps : 4;
          // descriptor of tap (0, 0); a tap only adds a compile-time row offset once the loop is unrolled (the issue loop
          // is a chain of dependent uniform-datapath instructions: every one removed is ~7 cycles per tap)
          uint64_t adesc0 = 0;
          adesc0 |= static_cast<uint64_t>((sa & 0x3FFFF) >> 4);
          adesc0 |= static_cast<uint64_t>(1) << 16;
          adesc0 |= static_cast<uint64_t>((kH2HaloW * 128) >> 4) << 32;
          adesc0 |= static_cast<uint64_t>(1) << 46;
          adesc0 |= static_cast<uint64_t>(2) << 61;
          // (unrolled only in the resident-weights variant, whose taps are a dozen instructions each once the slot index and
          // the row offset are constants; for the streamed-weights variants the 9x larger loop body measured slightly slower)
          constexpr int kTapUnroll = kResB ? 9 : 1;
#pragma unroll kTapUnroll
          for (int tap = 0; tap < 9; ++tap) {
            if constexpr (kResB) {
              if (tl == 0) { H2WAIT(&b_full[tap], 0, w_data); tc_fence_after_sync(); }     // the filter arrives once
              bs = tap;
            } else {
              H2WAIT(&b_full[bs], bph, w_data);
              tc_fence_after_sync();
            }
            const uint64_t bdesc = umma_desc_kmajor<128>(smem_u32(b_base + bs * kBSlotBytes));
            const int r = tap / 3, s = tap - 3 * r;
            // (the operand ring lies below 256 KB of shared memory: the 14-bit address field cannot carry into bit 16)
            const uint64_t adesc = adesc0 + static_cast<uint64_t>(((r * kH2HaloW + s) * 128) >> 4);
            if (elect_one()) {
              if (ksteps == 4) {           // the common case without per-MMA predicates (shorter dependent chain)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_ss_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < ksteps) umma_f16_ss_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
              }
              if constexpr (!kResB) umma_commit_2sm(&b_empty[bs]);
            }
            __syncwarp();
            if constexpr (!kResB) {
              if (++bs == kBStages) { bs = 0; bph ^= 1; }
            }
          }
          if (elect_one()) {
            umma_commit_2sm(&a_empty[as]);
            if (kc == p.k_chunks - 1) umma_commit_2sm(&tfull[acc]);
          }
          __syncwarp();
          if (++as == kAStg) { as = 0; aph ^= 1; }
        }
      }
#undef H2WAIT
      if (kH2Prof && p.prof && lane == 0) {
        atomicAdd(p.prof + 0, (unsigned long long)w_data);
        atomicAdd(p.prof + 1, (unsigned long long)w_acc);
        atomicAdd(p.prof + 2, (unsigned long long)(clock64() - t_begin));
      }
    }
  } else if (kDual && warp_idx >= 2 + kH2EpiWarps) {
    // ------------------------------------------------------------------ combiner (dual source): A = w0 * x + w1 * x2
    const int ct = threadIdx.x - kH2Threads;
    // fusion_weights_kernel's arithmetic (elementwise.cu): sigmoid(w) * (2 / n), n = 2
    const float w0 = (1.f / (1.f + expf(-__ldg(p.x_wts_raw)))) * (2.f / 2);
    const float w1 = (1.f / (1.f + expf(-__ldg(p.x_wts_raw + 1)))) * (2.f / 2);
    int as = 0;
    uint32_t aph = 0;
    for (int item = cluster_id; item < p.num_items; item += num_clusters) {
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        mbar_wait(&a_raw[as], aph);
        uint4* pa = reinterpret_cast<uint4*>(a_base + as * kH2ASlot);
        const uint4* pb = reinterpret_cast<const uint4*>(a2_base + as * kH2ASlot);
#pragma unroll 4
        for (int i = ct; i < kH2HaloBytes / 16; i += 32 * kH2CombWarps) {
          float fa[8], fb[8], fo[8];
          unpack8<kBf16>(pa[i], fa);
          unpack8<kBf16>(pb[i], fb);
#pragma unroll
          for (int k = 0; k < 8; ++k) fo[k] = fuse2(fa[k], fb[k], w0, w1);
          pa[i] = pack8<kBf16>(fo);
        }
        fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_smem_addr(&a_full[as]));
        if (++as == kAStg) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps per CTA, own 128 rows)
    const int ew = warp_idx - 2;
    const int q = warp_idx & 3;
    const int half = ew >> 2;
    uint8_t* wstage = staging + ew * 4096;
    float* wvec = vecs + ew * kH2VecFloats;
    int tl = 0, sbuf = 0;
    int staged_base = -1;      // (first output channel, chunk count) whose scale / bias are staged in wvec
    const bool prof_warp = leader && ew == 0;
    long long prof_ld = 0, prof_st = 0, prof_wait = 0;
    const long long t_begin = (kH2Prof && p.prof) ? clock64() : 0;
    // residual rows of the items this CTA reaches kResPrefetchTiles iterations from now -> L2 (conv_common.cuh)
    auto prefetch_res = [&](int item_f) {
      if (!kDual && p.res_pf && p.res != nullptr && half == 0 && item_f < p.num_items) {
        const Item2 f = item2_decode(p, item_f);
        const unsigned mt = fd_div((unsigned)f.tile, p.fd_nblocks);
        const int nb = (f.tile - (int)(mt * p.fd_nblocks.div)) * kH2BlockN + f.ncol0;
        const Sub2 fs = sub2_coord(p, mt * 2 + rank);
        const int row = q * 32 + lane;
        const int wo = fs.w0 + (row & 7), ho = fs.h0 + (row >> 3);
        if (fs.n < p.N && wo < p.W && ho < p.H && nb < p.Cout_store) {
          const long long pix = (static_cast<long long>(fs.n) * p.H + ho) * p.W + wo;
          const int cols = p.Cout_store - nb < f.ncols ? p.Cout_store - nb : f.ncols;
          l2_prefetch_row(reinterpret_cast<const uint8_t*>(p.res) + (pix * p.res_pix_stride + nb) * 2, (unsigned)cols * 2u, p.res_pf);
        }
      }
    };
    for (int d = 1; d < kResPrefetchTiles; ++d) prefetch_res(cluster_id + d * num_clusters);
    for (int item = cluster_id; item < p.num_items; item += num_clusters, ++tl) {
      prefetch_res(item + kResPrefetchTiles * num_clusters);
      const Item2 it = item2_decode(p, item);
      const int tile = it.tile;
      const int nchunks = it.ncols >> 5;
      const int acc = tl & 1;
      const unsigned mt = fd_div((unsigned)tile, p.fd_nblocks);
      const int nblk = tile - (int)(mt * p.fd_nblocks.div);
      const Sub2 sc = sub2_coord(p, mt * 2 + rank);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kH2BlockN;
      const uint32_t tempty_leader = leader_smem_addr(&tempty[acc]);
      // residual operand of this thread's row: all four column chunks are requested before the accumulator wait, so
      // their global-memory latency hides behind the main loop instead of being paid inside the epilogue
      uint4 rres_all[4][4];
      if (!kDual && p.res != nullptr) {
        const int row = q * 32 + lane;
        const int wo = sc.w0 + (row & 7), ho = sc.h0 + (row >> 3);
        const bool pix_ok = sc.n < p.N && wo < p.W && ho < p.H;
        const long long pix = (static_cast<long long>(sc.n) * p.H + ho) * p.W + wo;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c0 = nblk * kH2BlockN + it.ncol0 + (half + 2 * ci) * 32;
          const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.res) + (pix * p.res_pix_stride + c0) * 2;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            rres_all[ci][j] = make_uint4(0u, 0u, 0u, 0u);
            if (pix_ok && half + 2 * ci < nchunks && c0 + j * 8 < p.Cout_store)
              rres_all[ci][j] = __ldg(reinterpret_cast<const uint4*>(rp) + j);
          }
        }
      }
      {
        const long long t0 = (kH2Prof && p.prof) ? clock64() : 0;
        mbar_wait(&tfull[acc], (tl >> 1) & 1);
        if (kH2Prof && p.prof) prof_wait += clock64() - t0;
      }
      tc_fence_after_sync();
#define DYK_H2EPI(ACT) \
  halo2_epilogue<kBf16, ACT, kDual>(tm, p, sc, nblk * kH2BlockN + it.ncol0, nchunks, t_row, wstage, wvec, sbuf, tempty_leader, q, lane, half, \
                             prof_warp, prof_ld, prof_st, rres_all, staged_base)
      switch (p.act) {
        case DYK_ACT_LEAKY: DYK_H2EPI(DYK_ACT_LEAKY); break;
        case DYK_ACT_MISH: DYK_H2EPI(DYK_ACT_MISH); break;
        case DYK_ACT_RELU: DYK_H2EPI(DYK_ACT_RELU); break;
        case DYK_ACT_RELU6: DYK_H2EPI(DYK_ACT_RELU6); break;
        case DYK_ACT_HARDSWISH: DYK_H2EPI(DYK_ACT_HARDSWISH); break;
        case DYK_ACT_HARDSIGMOID: DYK_H2EPI(DYK_ACT_HARDSIGMOID); break;
        default: DYK_H2EPI(DYK_ACT_LINEAR); break;
      }
#undef DYK_H2EPI
    }
    if (lane == 0) tma_store_wait_all<0>();
    if (kH2Prof && p.prof && prof_warp && lane == 0) {
      atomicAdd(p.prof + 3, (unsigned long long)prof_wait);
      atomicAdd(p.prof + 4, (unsigned long long)(clock64() - t_begin));
      atomicAdd(p.prof + 5, (unsigned long long)prof_ld);
      atomicAdd(p.prof + 6, (unsigned long long)prof_st);
      atomicAdd(p.prof + 7, 1ull);
    }
  }

  tc_fence_before_sync();
  if (kH2Prof) { __syncthreads(); stamp(12, 11); }
  cluster_sync_all();      // the peer's barriers / smem stay valid until the leader's last multicast commit landed
  if (warp_idx == 1) tmem_dealloc_2sm<512>(tmem_base);
  stamp(-1, 13);
}

template <bool kBf16, bool kDual, bool kResB>
static int launch_halo2(const Halo2Tmaps& tm, const Halo2KArgs& ka, cudaStream_t stream) {
  auto kern = conv3x3_halo2_kernel<kBf16, kDual, kResB>;
  constexpr int kSmem = kResB ? kH2TotalRes : (kDual ? kH2TotalDual : kH2Total);
  static bool configured = false;
  if (!configured) {
    DYK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  int clusters = ka.cluster_cap;
  if (ka.num_items < clusters) clusters = ka.num_items;
  DYK_CUDA_OK(launch_pdl(kern, dim3(2 * clusters), dim3(kH2Threads + (kDual ? 32 * kH2CombWarps : 0)), (size_t)kSmem, stream, tm, ka));
  DYK_LAUNCH_OK("conv3x3_halo2_kernel");
  return DYK_OK;
}

// Resident-weights variant: 3x3 / stride 1 / pad 1 with one 64-channel K chunk and 64..128 output channels.
static bool halo2_resident_eligible(const dyk_conv_params* p) {
  static const bool off = getenv("DYK_HALO2_RES") != nullptr && getenv("DYK_HALO2_RES")[0] == '0';
  if (off) return false;
  if (!(p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && !p->upsample2x && !p->out_f32 && !p->y_plane &&
        p->out_h == 0 && p->out_w == 0 && !p->x2))
    return false;
  if (p->Cin > 64 || p->Cin < 16 || p->Cout_store < 64 || p->Cout_store > 128 || p->Cout_store % 32 != 0) return false;
  const int subs_w = ceil_div(p->W, kH2SubW), subs_h = ceil_div(p->H, kH2SubH);
  const long long num_subs = (long long)subs_w * subs_h * p->N;
  if (num_subs >= (1ll << 30) || num_subs < num_sms()) return false;      // at least one tile per SM, or the one-off filter load dominates
  const double eff = (double)p->W * p->H / ((double)subs_w * kH2SubW * subs_h * kH2SubH);
  return eff >= 0.6;
}

bool conv3x3_halo2_eligible(const dyk_conv_params* p) {
  static const bool off = getenv("DYK_HALO2") != nullptr && getenv("DYK_HALO2")[0] == '0';
  if (off) return false;
  if (!(p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && !p->upsample2x && !p->out_f32 && !p->y_plane &&
        p->out_h == 0 && p->out_w == 0))
    return false;
  if (p->Cout_store < 256 || p->Cin < 64) return false;
  const int subs_w = ceil_div(p->W, kH2SubW), subs_h = ceil_div(p->H, kH2SubH);
  const long long num_subs = (long long)subs_w * subs_h * p->N;
  if (num_subs >= (1ll << 30)) return false;
  const double eff = (double)p->W * p->H / ((double)subs_w * kH2SubW * subs_h * kH2SubH);
  return eff >= 0.6;
}

int conv3x3_halo2_try(const dyk_conv_params* p, cudaStream_t stream) {
  const bool resident = halo2_resident_eligible(p);
  if (!resident && !conv3x3_halo2_eligible(p)) return 1;
  const int H = p->H, W = p->W, N = p->N;
  const int subs_w = ceil_div(W, kH2SubW), subs_h = ceil_div(H, kH2SubH);
  const long long num_subs = (long long)subs_w * subs_h * N;
  const int n_blocks = ceil_div(p->Cout_store, kH2BlockN);
  const bool dual = p->x2 != nullptr;

  Halo2Tmaps tm;
  Halo2KArgs ka;
  memset(&tm, 0, sizeof(tm));
  memset(&ka, 0, sizeof(ka));
  int rc;
  {
    const long long xs = p->x_pix_stride * 2;
    const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t str[3] = {(cuuint64_t)xs, (cuuint64_t)xs * W, (cuuint64_t)xs * W * H};
    const cuuint32_t box[4] = {64, kH2HaloW, kH2HaloH, 1};
    if ((rc = encode_map_generic(&tm.a, p->x, 4, dims, str, box, 128, "halo2 A"))) return rc;
    if (dual) {
      const long long xs2 = p->x2_pix_stride * 2;
      const cuuint64_t str2[3] = {(cuuint64_t)xs2, (cuuint64_t)xs2 * W, (cuuint64_t)xs2 * W * H};
      if ((rc = encode_map_generic(&tm.a2, p->x2, 4, dims, str2, box, 128, "halo2 A (second source)"))) return rc;
    }
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)p->Cin, 9, (cuuint64_t)p->Cout};
    const cuuint64_t str[2] = {(cuuint64_t)p->Cin * 2, (cuuint64_t)p->Cin * 2 * 9};
    const cuuint32_t box[3] = {64, 1, 128};
    if ((rc = encode_map_generic(&tm.b, p->w, 3, dims, str, box, 128, "halo2 B"))) return rc;
  }
  {
    const long long ys = p->y_pix_stride * 2;
    const cuuint64_t dims[4] = {(cuuint64_t)p->Cout_store, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t str[3] = {(cuuint64_t)ys, (cuuint64_t)ys * W, (cuuint64_t)ys * W * H};
    const cuuint32_t box[4] = {32, kH2SubW, 4, 1};
    if ((rc = encode_map_generic(&tm.y, p->y, 4, dims, str, box, 64, "halo2 Y"))) return rc;
  }
  ka.H = H; ka.W = W; ka.N = N;
  ka.num_subs = (int)num_subs;
  ka.n_blocks = n_blocks;
  ka.num_tiles = (int)(ceil_div64(num_subs, 2) * n_blocks);
  {
    // split the tiles of the last partial round.  Measured (tools/conv_bench.py, DYK_H2_TAIL=1|2|4): an item of N = 128
    // still takes ~0.9 of a whole tile's time and one of N = 64 more than a whole tile on deep layers — an item walks
    // the same number of pipeline steps (k-chunks x 9 taps) and each step has a latency floor of ~400 cycles (TMA ->
    // full barrier -> MMA -> commit -> empty barrier), so narrowing N barely shortens it.  Worth 2 us on the 55 us
    // layers (256->512 @32x40: 57.3 -> 55.3 us); the real fix for the partial round is a split along K.
    static const int force = getenv("DYK_H2_TAIL") ? atoi(getenv("DYK_H2_TAIL")) : 0;   // 1, 2, 4 force; 0 = cost model
    const int clusters = sm_budget(p->sm_limit) / 2;
    ka.cluster_cap = clusters;
    const int tail = resident ? 0 : ka.num_tiles % clusters;
    int lg = 0;
    if (tail > 0) {
      const double cost[3] = {1.0, 0.9, 1.2};
      double best = 1e9;
      for (int l = 0; l < 3; ++l) {
        const double est = ceil_div(tail << l, clusters) * cost[l];
        if (est < best - 1e-9) { best = est; lg = l; }
      }
      if (force == 1) lg = 0; else if (force == 2) lg = 1; else if (force == 4) lg = 2;
    }
    ka.tail_lg = lg;
    ka.full_items = ka.num_tiles - tail;
    ka.num_items = ka.full_items + (tail << lg);
    const cuuint64_t dims[3] = {(cuuint64_t)p->Cin, 9, (cuuint64_t)p->Cout};
    const cuuint64_t str[2] = {(cuuint64_t)p->Cin * 2, (cuuint64_t)p->Cin * 2 * 9};
    const cuuint32_t box[3] = {64, 1, (cuuint32_t)(resident ? p->Cout_store / 2 : (128 >> lg))};
    if ((rc = encode_map_generic(&tm.bs, p->w, 3, dims, str, box, 128, "halo2 B (split)"))) return rc;
  }
  ka.k_chunks = ceil_div(p->Cin, 64);
  ka.Cout_store = p->Cout_store;
  ka.act = p->act;
  ka.fd_nblocks = make_fastdiv((unsigned)n_blocks);
  ka.fd_subs_w = make_fastdiv((unsigned)subs_w);
  ka.fd_subs_h = make_fastdiv((unsigned)subs_h);
  ka.scale = p->scale; ka.bias = p->bias;
  ka.res = p->res; ka.res_pix_stride = p->res_pix_stride;
  ka.prof = g_conv_prof;
  ka.res_pf = res_prefetch_mode();
  ka.x_wts_raw = p->x_wts_raw;
  ka.k_last_steps = ceil_div(p->Cin - (ka.k_chunks - 1) * 64, 16);
  if (resident) {
    ka.res_cols = p->Cout_store;
    return p->dtype == DYK_BF16 ? launch_halo2<true, false, true>(tm, ka, stream) : launch_halo2<false, false, true>(tm, ka, stream);
  }
  if (dual)
    return p->dtype == DYK_BF16 ? launch_halo2<true, true, false>(tm, ka, stream) : launch_halo2<false, true, false>(tm, ka, stream);
  return p->dtype == DYK_BF16 ? launch_halo2<true, false, false>(tm, ka, stream) : launch_halo2<false, false, false>(tm, ka, stream);
}

}  // namespace dyk
