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
#include "common.h"
#include "ptx.cuh"
#include "act.cuh"
#include "conv_common.cuh"
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

struct Sub2 {
  int w0, h0, n;
};
struct Item2 {
  int tile, ncol0, ncols;
};
__device__ __forceinline__ Item2 item2_decode(const Halo2KArgs& p, int item) {
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

template <bool kBf16, int kAct>
__device__ __forceinline__ void halo2_epilogue(const Halo2Tmaps& tm, const Halo2KArgs& p, const Sub2& sc, int n_base,
                                               int nchunks, uint32_t t_row, uint8_t* wstage, float* wvec, int& sbuf,
                                               uint32_t tempty_leader, int q, int lane, int half, bool prof_warp,
                                               long long& prof_ld, long long& prof_st, const uint4 (&rres_all)[4][4]) {
  constexpr int kCols = 32;
  constexpr int kChunks = kH2BlockN / kCols;   // up to 8 (nchunks of them in this item); this warp handles chunks half, half+2, ...
  const int row = q * 32 + lane;
  const bool sub_ok = sc.n < p.N;
  const bool has_res = p.res != nullptr;

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
      if (lane == 0) mbar_arrive_cluster(tempty_leader);
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
      if (lane == 0) mbar_arrive_cluster(tempty_leader);
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

template <bool kBf16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kH2Threads, 1)
conv3x3_halo2_kernel(const __grid_constant__ Halo2Tmaps tm, const Halo2KArgs p) {
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
  uint8_t* b_base = a_base + kH2AStages * kH2ASlot;
  uint8_t* staging = b_base + kH2BStages * kH2BSlot;
  float* vecs = reinterpret_cast<float*>(staging + kH2StagingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(vecs) + kH2VecBytes);
  uint64_t* a_full = bars;                          // [2]  (used in the leader)
  uint64_t* a_empty = a_full + kH2AStages;          // [2]  (per CTA, multicast arrivals)
  uint64_t* b_full = a_empty + kH2AStages;          // [8]  (leader)
  uint64_t* b_empty = b_full + kH2BStages;          // [8]  (per CTA)
  uint64_t* tfull = b_empty + kH2BStages;           // [2]  (per CTA, multicast arrivals)
  uint64_t* tempty = tfull + 2;                     // [2]  (leader; 16 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp_idx = threadIdx.x >> 5;
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
    for (int i = 0; i < kH2AStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kH2BStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * kH2EpiWarps); }
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before_sync();
  cluster_sync_all();      // barriers of both CTAs initialised before any remote arrive / TMA credit
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();   // PDL: everything above overlapped the previous kernel's tail; its results are visible from here on
  stamp(-1, 10);

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, own halves)
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
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
          if (leader) mbar_arrive_expect_tx(&a_full[as], 2 * kH2HaloBytes);
          tma_load_4d_2sm(a_base + as * kH2ASlot, &tm.a, leader_smem_addr(&a_full[as]), kc * 64, sc.w0 - 1, sc.h0 - 1, sc.n);
          if (++as == kH2AStages) { as = 0; aph ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            if (leader) mbar_arrive_expect_tx(&b_full[bs], b_bytes);
            tma_load_3d_2sm(b_base + bs * kH2BSlot, bmap, leader_smem_addr(&b_full[bs]), kc * 64, tap, b_row);
            if (++bs == kH2BStages) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader && lane == 0) {
      const uint32_t idesc_whole = umma_idesc_f16(256, kH2BlockN, kBf16 ? 1 : 0);
      const uint32_t idesc_split = umma_idesc_f16(256, kH2BlockN >> p.tail_lg, kBf16 ? 1 : 0);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int tl = 0;
      long long w_data = 0, w_acc = 0;
      const long long t_begin = (kH2Prof && p.prof) ? clock64() : 0;
#define H2WAIT(bar, ph, var) do { const long long t0 = (kH2Prof && p.prof) ? clock64() : 0; mbar_wait(bar, ph); \
                                  if (kH2Prof && p.prof) var += clock64() - t0; } while (0)
      for (int item = cluster_id; item < p.num_items; item += num_clusters, ++tl) {
        const uint32_t idesc = item < p.full_items ? idesc_whole : idesc_split;
        const int acc = tl & 1;
        H2WAIT(&tempty[acc], ((tl >> 1) & 1) ^ 1, w_acc);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kH2BlockN;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          H2WAIT(&a_full[as], aph, w_data);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(a_base + as * kH2ASlot);
          for (int tap = 0; tap < 9; ++tap) {
            H2WAIT(&b_full[bs], bph, w_data);
            tc_fence_after_sync();
            const uint64_t bdesc = umma_desc_kmajor<128>(smem_u32(b_base + bs * kH2BSlot));
            const int r = tap / 3, s = tap - 3 * r;
            uint64_t adesc = 0;
            adesc |= static_cast<uint64_t>(((sa + (r * kH2HaloW + s) * 128) & 0x3FFFF) >> 4);
            adesc |= static_cast<uint64_t>(1) << 16;
            adesc |= static_cast<uint64_t>((kH2HaloW * 128) >> 4) << 32;
            adesc |= static_cast<uint64_t>(1) << 46;
            adesc |= static_cast<uint64_t>(2) << 61;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&b_empty[bs]);
            if (++bs == kH2BStages) { bs = 0; bph ^= 1; }
          }
          umma_commit_2sm(&a_empty[as]);
          if (++as == kH2AStages) { as = 0; aph ^= 1; }
        }
        umma_commit_2sm(&tfull[acc]);
      }
#undef H2WAIT
      if (kH2Prof && p.prof) {
        atomicAdd(p.prof + 0, (unsigned long long)w_data);
        atomicAdd(p.prof + 1, (unsigned long long)w_acc);
        atomicAdd(p.prof + 2, (unsigned long long)(clock64() - t_begin));
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
    const bool prof_warp = leader && ew == 0;
    long long prof_ld = 0, prof_st = 0, prof_wait = 0;
    const long long t_begin = (kH2Prof && p.prof) ? clock64() : 0;
    for (int item = cluster_id; item < p.num_items; item += num_clusters, ++tl) {
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
      if (p.res != nullptr) {
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
  halo2_epilogue<kBf16, ACT>(tm, p, sc, nblk * kH2BlockN + it.ncol0, nchunks, t_row, wstage, wvec, sbuf, tempty_leader, q, lane, half, \
                             prof_warp, prof_ld, prof_st, rres_all)
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

template <bool kBf16>
static int launch_halo2(const Halo2Tmaps& tm, const Halo2KArgs& ka, cudaStream_t stream) {
  auto kern = conv3x3_halo2_kernel<kBf16>;
  static bool configured = false;
  if (!configured) {
    DYK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kH2Total));
    configured = true;
  }
  int clusters = num_sms() / 2;
  if (ka.num_items < clusters) clusters = ka.num_items;
  DYK_CUDA_OK(launch_pdl(kern, dim3(2 * clusters), dim3(kH2Threads), (size_t)kH2Total, stream, tm, ka));
  DYK_LAUNCH_OK("conv3x3_halo2_kernel");
  return DYK_OK;
}

// Returns DYK_OK after launching, 1 when the layer is not eligible, < 0 on error.
int conv3x3_halo2_try(const dyk_conv_params* p, cudaStream_t stream) {
  static const bool off = getenv("DYK_HALO2") != nullptr && getenv("DYK_HALO2")[0] == '0';
  if (off) return 1;
  if (!(p->kh == 3 && p->kw == 3 && p->stride == 1 && p->pad == 1 && !p->upsample2x && !p->out_f32 && !p->y_plane &&
        p->out_h == 0 && p->out_w == 0))
    return 1;
  if (p->Cout_store < 256 || p->Cin < 64) return 1;
  const int H = p->H, W = p->W, N = p->N;
  const int subs_w = ceil_div(W, kH2SubW), subs_h = ceil_div(H, kH2SubH);
  const long long num_subs = (long long)subs_w * subs_h * N;
  if (num_subs >= (1ll << 30)) return 1;
  const double eff = (double)W * H / ((double)subs_w * kH2SubW * subs_h * kH2SubH);
  if (eff < 0.6) return 1;
  const int n_blocks = ceil_div(p->Cout_store, kH2BlockN);

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
    const int clusters = num_sms() / 2;
    const int tail = ka.num_tiles % clusters;
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
    const cuuint32_t box[3] = {64, 1, (cuuint32_t)(128 >> lg)};
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
  return p->dtype == DYK_BF16 ? launch_halo2<true>(tm, ka, stream) : launch_halo2<false>(tm, ka, stream);
}

}  // namespace dyk
