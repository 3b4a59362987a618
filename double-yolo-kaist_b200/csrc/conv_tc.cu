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
// Warp roles (320 or 576 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2.. = 8 or 16 epilogue warps (TMEM lane quarter = warp_idx & 3; kSplit warps per quarter split the columns).
#include "common.h"
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "ptx.cuh"
#include "act.cuh"
#include "conv_common.cuh"

namespace dyk {

struct ConvTmaps {
  CUtensorMap a[4];  // input, by (row parity*2 + col parity); stride-1 convs use a[0] only
  CUtensorMap b;     // packed weights (Cin, taps, Cout)
  CUtensorMap y[4];  // output; upsample2x uses the 4 parity planes of the 2x map, else y[0]
};

struct ConvKArgs {
  int tw, th, tn;                  // A/Y box spatial extents, tw*th*tn == 128 (all powers of two)
  int tw_log2, th_log2;
  FastDiv fd_nblocks, fd_tiles_w, fd_tiles_h;
  int tiles_w, tiles_h, tiles_b;   // tile grid over (Wo, Ho, N)
  int n_blocks;                    // ceil(Cout_store / BLOCK_N)
  int num_tiles;
  int k_chunks;                    // ceil(Cin / BLOCK_K)
  int kh, kw, stride, pad;
  int pair_w;                      // stride-2 Cin = 32 layers: two W-adjacent pixels form one 64-channel K block (see host)
  int Ho, Wo, N;
  int Cout_store;
  int act;
  int upsample2x;
  int out_f32;                     // 1: y is fp32, written with plain stores (head convs feeding decode)
  int acc_main_k;                  // acc32 variant: K-elements of a tap whose MMAs go to the main accumulator
  void* y_f32;
  long long y_pix_stride;
  const float* scale;              // may be null
  const float* bias;               // may be null
  const void* res;                 // may be null
  long long res_pix_stride;
  unsigned long long* prof;        // optional role-cycle counters (dyk_conv_set_profile), may be null
  int res_pf;                      // residual L2 prefetch mode (conv_common.cuh)
  int grid_cap;                    // SM budget of this launch (sm_budget)
  int flat;                        // 1x1 stride-1 convolution run as a GEMM over the flattened pixel index (tw = 128)
  int b_img_rows;                  // per-image weights (SE gate folded into the consumer): weight rows per image, 0 = shared
};

// Role-cycle counters (diagnostics): where the three pipelines of the kernel wait.
enum { PROF_PRODUCER_WAIT_EMPTY = 0, PROF_PRODUCER_TOTAL, PROF_MMA_WAIT_FULL, PROF_MMA_WAIT_TEMPTY, PROF_MMA_TOTAL,
       PROF_EPI_WAIT_TFULL, PROF_EPI_TOTAL, PROF_CTAS, PROF_COUNT };

// Role-cycle counters are compiled in only with -DDYK_CONV_PROFILE (libdyk_b200_prof.so, tools/conv_bench.py):
// they cost registers in the single-thread producer / MMA roles.
#ifdef DYK_CONV_PROFILE
constexpr bool kProf = true;
#else
constexpr bool kProf = false;
#endif
// Epilogue warps: 4 TMEM lane quarters x kSplit column parts.  kSplit = 2 (8 warps, 168 registers each) for layers with a
// long main loop; kSplit = 4 (16 warps, 112 registers) for short main loops (1x1 layers, small Cin), whose time is the
// epilogue's instruction issue: more warps per scheduler hide the MUFU / TMEM-load latencies (Mish layers: ~1.5x).
template <int kSplit> constexpr int epi_warps() { return 4 * kSplit; }
template <int kSplit> constexpr int num_threads() { return 64 + 32 * epi_warps<kSplit>(); }
constexpr int kBlockM = 128;

template <int BLOCK_N, int BLOCK_K, int kSplit>
struct ConvSmem {
  static constexpr int kNumEpiWarps = 4 * kSplit;
  // Every epilogue warp owns rows [32q, 32q+32) of the tile and every second chunk of kColsW output channels;
  // it stages and TMA-stores its own 32 x kColsW sub-boxes, so the epilogue needs no CTA-wide barrier.
  static constexpr int kColsW = BLOCK_N >= 32 * kSplit ? 32 : BLOCK_N / kSplit;     // channels per warp chunk (32 or 16)
  static_assert(kColsW == 32 || kColsW == 16, "chunk width");
  static constexpr int kChunkBytes = 32 * kColsW * 2;                   // one staged sub-box: 2 KB or 1 KB
  static constexpr int kABytes = kBlockM * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kNumEpiWarps * 2 * 2048;        // 2 buffers per warp, 2 KB apart
  static constexpr int kVecFloats = 2 * BLOCK_N / kSplit;               // per warp: scale + bias of its BLOCK_N/kSplit columns
  static constexpr int kVecBytes = kNumEpiWarps * kVecFloats * 4;
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kBudget = 227 * 1024 - kBarrierBytes - kStagingBytes - kVecBytes - 1024 /*align*/;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kVecBytes + kBarrierBytes + 1024;
  static_assert(kStages >= 3, "pipeline too shallow");
};

struct TileCoord {
  int nblk, w0, h0, n0;
};
__device__ __forceinline__ TileCoord tile_coord(const ConvKArgs& p, int tile) {
  TileCoord t;
  if (p.flat) {      // 1x1 convolutions over the flattened pixel index: no spatial decomposition (three dependent
    t.h0 = 0;        // multiply-high chains per tile and role otherwise, ~6 % of the epilogue's stall samples on short tiles)
    t.n0 = 0;
    if (p.n_blocks == 1) {
      t.nblk = 0;
      t.w0 = tile << 7;
    } else {
      const unsigned m = fd_div((unsigned)tile, p.fd_nblocks);
      t.nblk = tile - (int)(m * p.fd_nblocks.div);
      t.w0 = (int)m << 7;
    }
    return t;
  }
  const unsigned mt = fd_div((unsigned)tile, p.fd_nblocks);
  t.nblk = tile - (int)(mt * p.fd_nblocks.div);
  const unsigned rowt = fd_div(mt, p.fd_tiles_w);            // tile row index over (tiles_h * tiles_b)
  const unsigned tww = mt - rowt * p.fd_tiles_w.div;
  const unsigned tb = fd_div(rowt, p.fd_tiles_h);
  const unsigned thh = rowt - tb * p.fd_tiles_h.div;
  t.w0 = tww * p.tw;
  t.h0 = thh * p.th;
  t.n0 = tb * p.tn;
  return t;
}

// Epilogue of one 128 x BLOCK_N accumulator tile, executed independently by each of the 8 epilogue warps.
// Warp (q = warp & 3, half = epilogue warp >> 2): thread `lane` owns tile row q*32+lane (TMEM lane) and the
// column chunks half, half+2, ... of kColsW channels.  Per chunk: tcgen05.ld -> scale/bias/activation
// (+ residual) in fp32 -> 16-bit -> swizzled warp-private staging -> one TMA store of the 32-row sub-box issued
// by lane 0.  kAct is a compile-time activation so that only one code path is resident in the instruction cache.
// Per-thread geometry of the epilogue that does not change from tile to tile (hoisted out of the tile loop: on the
// short-main-loop 1x1 layers a warp handles only 16 - 32 elements per thread and tile, and ncu showed ~290 of its 448
// instructions per tile were this bookkeeping, not arithmetic).
struct EpiRow {
  int wi, hi, ni;          // offset of this thread's row (pixel) inside the tile
  int sw, sh, sn;          // offset of this warp's 32-row sub-box inside the tile
};
__device__ __forceinline__ EpiRow epi_row(const ConvKArgs& p, int q, int lane) {
  EpiRow r;
  const int row = q * 32 + lane, r0 = q * 32;
  r.wi = row & (p.tw - 1);
  r.hi = (row >> p.tw_log2) & (p.th - 1);
  r.ni = row >> (p.tw_log2 + p.th_log2);
  r.sw = r0 & (p.tw - 1);
  r.sh = (r0 >> p.tw_log2) & (p.th - 1);
  r.sn = r0 >> (p.tw_log2 + p.th_log2);
  return r;
}

template <int BLOCK_N, int BLOCK_K, int kSplit, bool kBf16, int kAct>
__device__ __forceinline__ void epilogue_tile(const ConvTmaps& tm, const ConvKArgs& p, const TileCoord& tc, const EpiRow& er,
                                              uint32_t t_row, uint8_t* wstage, float* wvec, int& sbuf, int& staged_nblk,
                                              uint64_t* tempty, int lane, int half) {
  using S = ConvSmem<BLOCK_N, BLOCK_K, kSplit>;
  constexpr int kCols = S::kColsW;                  // 32 or 16
  constexpr int kRowBytes = kCols * 2;              // 64 or 32 (== TMA store swizzle span)
  constexpr int kChunks = BLOCK_N / kCols;          // column chunks per tile
  const int wo = tc.w0 + er.wi, ho = tc.h0 + er.hi, nn = tc.n0 + er.ni;
  const bool pix_ok = (wo < p.Wo) && (ho < p.Ho) && (nn < p.N);
  const long long pix = (static_cast<long long>(nn) * p.Ho + ho) * p.Wo + wo;
  const int n_base = tc.nblk * BLOCK_N;
  const bool has_res = p.res != nullptr;
  // origin of this warp's 32-row sub-box inside the tile (warp-uniform)
  const int sw0 = tc.w0 + er.sw, sh0 = tc.h0 + er.sh, sn0 = tc.n0 + er.sn;

  uint4 rres[kCols / 8];
  auto load_res = [&](int c) {
    const int c0 = n_base + c * kCols;
    const uint8_t* rp = reinterpret_cast<const uint8_t*>(p.res) + (pix * p.res_pix_stride + c0) * 2;
#pragma unroll
    for (int j = 0; j < kCols / 8; ++j) {
      rres[j] = make_uint4(0u, 0u, 0u, 0u);
      if (has_res && pix_ok && c0 + j * 8 < p.Cout_store) rres[j] = __ldg(reinterpret_cast<const uint4*>(rp) + j);
    }
  };
  // this warp's scale / bias values -> warp-private shared memory: wvec[2*(j*kCols + col) + {0,1}] for its j-th chunk;
  // only when the output-channel block differs from the one staged last (never again on layers with one channel block)
  if (tc.nblk != staged_nblk) {
    staged_nblk = tc.nblk;
    __syncwarp();   // the previous tile's reads of wvec are complete
    if (lane < kCols) {
#pragma unroll
      for (int j = 0; j < kChunks / kSplit; ++j) {
        const int col = n_base + (half + kSplit * j) * kCols + lane;
        wvec[j * kCols + lane] = p.scale ? __ldg(p.scale + col) : 1.f;
        wvec[(kChunks / kSplit) * kCols + j * kCols + lane] = p.bias ? __ldg(p.bias + col) : 0.f;
      }
    }
    __syncwarp();
  }
  const float* wscale = wvec;
  const float* wbias = wvec + (kChunks / kSplit) * kCols;

#pragma unroll 1
  for (int c = half; c < kChunks; c += kSplit) {
    const int cl = c * kCols;                // first column of the chunk within the tile
    const int cg0 = n_base + cl;             // first output channel of the chunk
    const bool beyond = cg0 >= p.Cout_store; // warp-uniform: the chunk lies outside the tensor
    const bool last = (c + kSplit >= kChunks) || (cg0 + kSplit * kCols >= p.Cout_store);
    if (beyond) {   // nothing to read: release the accumulator stage and stop
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      break;
    }
    if (has_res) load_res(c);   // in flight while the accumulator chunk is read from TMEM
    uint8_t* sbase = wstage + sbuf * 2048;
    if (!p.out_f32) {
      // staging[sbuf] may still be read by the store issued two chunks ago
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
    }
    float* yo = reinterpret_cast<float*>(p.y_f32) + pix * p.y_pix_stride + cg0;   // only used when out_f32
    // 16 accumulator columns per TMEM load, 8 channels at a time: keeps the live register set small
    // (16 + 8 outputs + their scale/bias), which the 16-warp variant (112 registers) needs
#pragma unroll
    for (int hh = 0; hh < kCols / 16; ++hh) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_row + cl + hh * 16, v);
      tmem_ld_wait();
      if (last && hh == kCols / 16 - 1) {
        // all TMEM reads of this warp for this accumulator stage are done -> hand it back to the MMA warp
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
      }
#pragma unroll
      for (int c8 = 0; c8 < 2; ++c8) {
        const int ch = hh * 2 + c8;
        float o[8];
        const int vo = (c / kSplit) * kCols + ch * 8;
        const float4 sc0 = *reinterpret_cast<const float4*>(wscale + vo);
        const float4 sc1 = *reinterpret_cast<const float4*>(wscale + vo + 4);
        const float4 bi0 = *reinterpret_cast<const float4*>(wbias + vo);
        const float4 bi1 = *reinterpret_cast<const float4*>(wbias + vo + 4);
        o[0] = act_apply<kAct>(fmaf(__uint_as_float(v[c8 * 8 + 0]), sc0.x, bi0.x));
        o[1] = act_apply<kAct>(fmaf(__uint_as_float(v[c8 * 8 + 1]), sc0.y, bi0.y));
        o[2] = act_apply<kAct>(fmaf(__uint_as_float(v[c8 * 8 + 2]), sc0.z, bi0.z));
        o[3] = act_apply<kAct>(fmaf(__uint_as_float(v[c8 * 8 + 3]), sc0.w, bi0.w));
        o[4] = act_apply<kAct>(fmaf(__uint_as_float(v[c8 * 8 + 4]), sc1.x, bi1.x));
        o[5] = act_apply<kAct>(fmaf(__uint_as_float(v[c8 * 8 + 5]), sc1.y, bi1.y));
        o[6] = act_apply<kAct>(fmaf(__uint_as_float(v[c8 * 8 + 6]), sc1.z, bi1.z));
        o[7] = act_apply<kAct>(fmaf(__uint_as_float(v[c8 * 8 + 7]), sc1.w, bi1.w));
        if (has_res) {
          const uint32_t rr[4] = {rres[ch].x, rres[ch].y, rres[ch].z, rres[ch].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack2<kBf16>(rr[e]);
            o[e * 2 + 0] += f.x;
            o[e * 2 + 1] += f.y;
          }
        }
        if (p.out_f32) {
          if (pix_ok) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (cg0 + ch * 8 + j < p.Cout_store) yo[ch * 8 + j] = o[j];
          }
        } else {
          int phys;
          if constexpr (kRowBytes == 64) phys = ch ^ ((lane >> 1) & 3);
          else phys = ch ^ ((lane >> 2) & 1);
          const uint4 val = make_uint4(pack2<kBf16>(o[0], o[1]), pack2<kBf16>(o[2], o[3]), pack2<kBf16>(o[4], o[5]),
                                       pack2<kBf16>(o[6], o[7]));
          *reinterpret_cast<uint4*>(sbase + lane * kRowBytes + phys * 16) = val;
        }
      }
    }
    if (p.out_f32) {
      if (last) break;
      continue;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (p.upsample2x) {
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) tma_store_4d(&tm.y[pp], sbase, cg0, sw0, sh0, sn0);
      } else {
        tma_store_4d(&tm.y[0], sbase, cg0, sw0, sh0, sn0);
      }
      tma_store_commit();
    }
    sbuf ^= 1;
    if (last) break;
  }
}

// kAcc32 (fp32-accurate mode, out_f32 == 2).  The tensor core's adder aligns the 16 products of an MMA and the fp32
// accumulator to the largest exponent and TRUNCATES every addend there (measured: ~2e-5 relative error after 864
// accumulating MMAs, 5e-6 after 27; sign-symmetric — a negated twin accumulator cancels nothing), i.e. every MMA adds
// ~16 truncations at the ulp of the running sum.  Two measures bring the result to within a small factor of an fp32 FMA chain:
//  * chunks: the MMA warp accumulates only kAccChunkKb k-blocks (8 MMAs) per TMEM stage, and the epilogue warps add every
//    finished chunk into fp32 registers with round-to-nearest (the two TMEM stages double-buffer chunks instead of
//    tiles), so the running sum the truncation is relative to stays a short partial sum;
//  * two accumulators per stage: MMAs over the leading `acc_main_k` K-elements of a filter tap (the x1*w1 products of the
//    split operand, csrc/f32_path.cu) go to M, all others (the five correction terms, 2^-8 .. 2^-16 smaller) go to S,
//    whose truncation unit is smaller by the same factor; the epilogue adds M + S.
constexpr int kAccChunkKb = 2;
// which accumulators the MMAs of k-blocks [kb0, kb1) touch: bit 0 = M, bit 1 = S  (MMA j of k-block kb starts at K-element
// (kb % k_chunks) * BLOCK_K + 16 j of its tap; it goes to S when that is >= main_k)
template <int BLOCK_K>
__device__ __forceinline__ int acc32_touch(int kb0, int kb1, int k_chunks, int main_k) {
  int m = 0;
  for (int kb = kb0; kb < kb1; ++kb) {
    const int k0 = (kb % k_chunks) * BLOCK_K;
    if (k0 < main_k) m |= 1;
    if (k0 + BLOCK_K - 16 >= main_k) m |= 2;
  }
  return m;
}

template <int BLOCK_N, int BLOCK_K, int kSplit, bool kBf16, bool kAcc32 = false>
__global__ void __launch_bounds__(num_threads<kSplit>(), 1)
conv_tc_kernel(const __grid_constant__ ConvTmaps tm, const ConvKArgs p) {
  using S = ConvSmem<BLOCK_N, BLOCK_K, kSplit>;
  constexpr int kNumEpiWarps = S::kNumEpiWarps;
  constexpr int kSwz = BLOCK_K * 2;            // bytes per smem operand row == swizzle span (128 / 64)
  constexpr int kStages = S::kStages;
  // acc32: two accumulators (main products, correction terms) per stage
  constexpr uint32_t kTmemCols = kAcc32 ? 4 * BLOCK_N : (2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N);
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
  static_assert(BLOCK_K == 64 || BLOCK_K == 32, "BLOCK_K");

  griddep_launch_dependents();
  // profile build: wall-clock (globaltimer, ns) of kernel entry / prologue end / role-loop end / exit, min and max over CTAs
  // (slots 8..13 of the counter array, as in conv3x3_halo2_kernel)
  auto stamp = [&](int lo, int hi) {
    if (kProf && p.prof && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (lo >= 0) atomicMin(p.prof + lo, t);
      if (hi >= 0) atomicMax(p.prof + hi, t);
    }
  };
  stamp(8, 9);
  extern __shared__ uint8_t smem_raw[];
  // align by *offset* (not by casting through an integer) so the compiler keeps the shared state space
  // and emits LDS/STS instead of generic LD/ST for the staging buffer and the scale/bias vectors
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kStages * S::kStageBytes;
  float* vecs = reinterpret_cast<float*>(staging + S::kStagingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(vecs) + S::kVecBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;        // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  // warp index through a shuffle (as CUTLASS' canonical_warp_idx_sync): the compiler then knows it is warp-uniform and keeps
  // everything derived from it (role, TMEM lane quarter, staging addresses, TMA-store operands) in uniform registers
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // (same value in every lane: tell the compiler)
  griddep_wait();   // PDL: everything above overlapped the previous kernel's tail; its results are visible from here on
  stamp(-1, 10);

  if (warp_idx == 0) {
    // ------------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0;
      const long long t_begin = clock64();
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(p, tile);
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.kw, s = tap - r * p.kw;
          int dh = r - p.pad, dw = s - p.pad;
          int map_idx = 0;
          int b_k0 = 0, b_tap = tap;
          if (p.pair_w) {
            // rows: parity planes as for any stride-2 conv; columns: pixel pairs (2j, 2j+1) are one 128-byte K block, tap
            // s = 1 of the kw = 2 kernel covers original taps 1, 2 of pair j, tap s = 0 covers tap 0 = second pixel of
            // pair j-1 (its first half meets zero weights: the B box starts 32 elements before the row, OOB = 0)
            const int ph = dh & 1;
            map_idx = ph * 2;
            dh = (dh - ph) >> 1;
            b_k0 = s == 0 ? -32 : 32;
            b_tap = r;
          } else if (p.stride == 2) {
            const int ph = dh & 1, pw = dw & 1;
            map_idx = ph * 2 + pw;
            dh = (dh - ph) >> 1;
            dw = (dw - pw) >> 1;
          }
          const CUtensorMap* amap = &tm.a[map_idx];
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            if (kProf && p.prof) {
              const long long t0 = clock64();
              mbar_wait(&empty_bar[stage], phase ^ 1);
              t_wait += clock64() - t0;
            } else {
              mbar_wait(&empty_bar[stage], phase ^ 1);
            }
            uint8_t* sa = stage_base + stage * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
              tma_load_4d(sa, amap, &full_bar[stage], kc * BLOCK_K, tc.w0 + dw, tc.h0 + dh, tc.n0);
              tma_load_3d(sb, &tm.b, &full_bar[stage], kc * BLOCK_K + b_k0, b_tap, tc.nblk * BLOCK_N + tc.n0 * p.b_img_rows);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (kProf && p.prof && lane == 0) {
        atomicAdd(p.prof + PROF_PRODUCER_WAIT_EMPTY, (unsigned long long)t_wait);
        atomicAdd(p.prof + PROF_PRODUCER_TOTAL, (unsigned long long)(clock64() - t_begin));
        atomicAdd(p.prof + PROF_CTAS, 1ull);
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------------ MMA issuer: the whole warp walks the loop and
    // one elected lane issues each tcgen05 instruction (warp-uniform control flow keeps descriptors / barrier addresses in
    // uniform registers; a lone thread under `if (lane == 0)` needed ~24 SASS instructions per MMA, see conv_halo2.cu)
    {
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N, kBf16 ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int tl = 0;
      long long t_wfull = 0, t_wtempty = 0;
      const long long t_begin = clock64();
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
        if constexpr (kAcc32) {
          for (int kb0 = 0; kb0 < num_kb; kb0 += kAccChunkKb, ++tl) {
            const int as = tl & 1;
            mbar_wait(&tempty_bar[as], ((tl >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + as * 2 * BLOCK_N;
            bool first_m = true, first_s = true;
            const int kb1 = kb0 + kAccChunkKb < num_kb ? kb0 + kAccChunkKb : num_kb;
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after_sync();
              const uint32_t sa = smem_u32(stage_base + stage * S::kStageBytes);
              const uint64_t adesc = umma_desc_kmajor<kSwz>(sa);
              const uint64_t bdesc = umma_desc_kmajor<kSwz>(sa + S::kABytes);
#pragma unroll
              const int k0 = (kb % p.k_chunks) * BLOCK_K;
              const bool el = elect_one();
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k) {
                if (k0 + 16 * k < p.acc_main_k) {
                  if (el) umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, first_m ? 0u : 1u);
                  first_m = false;
                } else {
                  if (el) umma_f16_ss(d_tmem + BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, first_s ? 0u : 1u);
                  first_s = false;
                }
              }
              if (el) umma_commit(&empty_bar[stage]);
              __syncwarp();
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tfull_bar[as]);
            __syncwarp();
          }
          --tl;      // the tile loop's own increment
          continue;
        }
        const int as = tl & 1;
        const uint32_t aphase = (tl >> 1) & 1;
        {
          const long long t0 = (kProf && p.prof) ? clock64() : 0;
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          if (kProf && p.prof) t_wtempty += clock64() - t0;
        }
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          {
            const long long t0 = (kProf && p.prof) ? clock64() : 0;
            mbar_wait(&full_bar[stage], phase);
            if (kProf && p.prof) t_wfull += clock64() - t0;
          }
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(stage_base + stage * S::kStageBytes);
          const uint64_t adesc = umma_desc_kmajor<kSwz>(sa);
          const uint64_t bdesc = umma_desc_kmajor<kSwz>(sa + S::kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the >>4 address field
              umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
            if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (kProf && p.prof && lane == 0) {
        atomicAdd(p.prof + PROF_MMA_WAIT_FULL, (unsigned long long)t_wfull);
        atomicAdd(p.prof + PROF_MMA_WAIT_TEMPTY, (unsigned long long)t_wtempty);
        atomicAdd(p.prof + PROF_MMA_TOTAL, (unsigned long long)(clock64() - t_begin));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp_idx - 2;
    const int q = warp_idx & 3;            // TMEM lane quarter this warp may access
    const int half = ew >> 2;              // which of the kSplit column parts this warp owns
    uint8_t* wstage = staging + ew * 4096;   // this warp's two 2 KB staging buffers
    float* wvec = vecs + ew * S::kVecFloats;
    int tl = 0;
    int sbuf = 0;
    long long t_wtfull = 0;
    const long long t_begin = clock64();
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      if constexpr (kAcc32) {
        static_assert(!kAcc32 || (kSplit == 2 && BLOCK_N == 64), "acc32 variant: 2 x 32 register accumulators per thread");
        constexpr int kMy = BLOCK_N / 64;           // 32-column chunks per thread (chunks half, half + 2)
        const TileCoord tc = tile_coord(p, tile);
        float acc[kMy][32], acc_s[kMy][32];       // main products / correction terms, summed at the end
#pragma unroll
        for (int j = 0; j < kMy; ++j)
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[j][i] = acc_s[j][i] = 0.f;
        for (int kb0 = 0; kb0 < num_kb; kb0 += kAccChunkKb, ++tl) {
          const int as = tl & 1;
          mbar_wait(&tfull_bar[as], (tl >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 2 * BLOCK_N;
          const int kb1 = kb0 + kAccChunkKb < num_kb ? kb0 + kAccChunkKb : num_kb;
          const int touch = acc32_touch<BLOCK_K>(kb0, kb1, p.k_chunks, p.acc_main_k);
#pragma unroll
          for (int j = 0; j < kMy; ++j) {
            uint32_t v[32];
            if (touch & 2) {         // correction terms first: small + small, then the main partial sum on top
              tmem_ld_32x32b_x32(t_row + BLOCK_N + (half + 2 * j) * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) acc_s[j][i] += __uint_as_float(v[i]);
            }
            if (touch & 1) {
              tmem_ld_32x32b_x32(t_row + (half + 2 * j) * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) acc[j][i] += __uint_as_float(v[i]);     // fp32 add, round to nearest
            }
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        --tl;
        const int row = q * 32 + lane;
        const int wo = tc.w0 + (row & (p.tw - 1)), ho = tc.h0 + ((row >> p.tw_log2) & (p.th - 1)),
                  nn = tc.n0 + (row >> (p.tw_log2 + p.th_log2));
        if (wo < p.Wo && ho < p.Ho && nn < p.N) {
          const long long pix = (static_cast<long long>(nn) * p.Ho + ho) * p.Wo + wo;
#pragma unroll
          for (int j = 0; j < kMy; ++j) {
            const int c0 = tc.nblk * BLOCK_N + (half + 2 * j) * 32;
            float* yo = reinterpret_cast<float*>(p.y_f32) + pix * p.y_pix_stride + c0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (c0 + i < p.Cout_store) {
                const float z = fmaf(acc[j][i] + acc_s[j][i], p.scale ? __ldg(p.scale + c0 + i) : 1.f, p.bias ? __ldg(p.bias + c0 + i) : 0.f);
                yo[i] = apply_act(z, p.act);
              }
            }
          }
        }
        continue;
      }
      // 16-bit / fp32-head path: one specialised tile loop per activation (the switch is taken once per kernel, the
      // per-thread row geometry and the staged scale / bias survive from tile to tile)
      const EpiRow er = epi_row(p, q, lane);
      int staged_nblk = -1;
      // residual rows of the tiles this CTA reaches kResPrefetchTiles iterations from now -> L2 (conv_common.cuh)
      auto prefetch_res = [&](int t) {
        if (p.res_pf && p.res != nullptr && half == 0 && t < p.num_tiles) {
          const TileCoord f = tile_coord(p, t);
          const int wo = f.w0 + er.wi, ho = f.h0 + er.hi, nn = f.n0 + er.ni;
          if (wo < p.Wo && ho < p.Ho && nn < p.N) {
            const long long pix = (static_cast<long long>(nn) * p.Ho + ho) * p.Wo + wo;
            const int nb = f.nblk * BLOCK_N;
            const int cols = p.Cout_store - nb < BLOCK_N ? p.Cout_store - nb : BLOCK_N;
            l2_prefetch_row(reinterpret_cast<const uint8_t*>(p.res) + (pix * p.res_pix_stride + nb) * 2, (unsigned)cols * 2u, p.res_pf);
          }
        }
      };
      if (tl == 0) {
        for (int d = 1; d < kResPrefetchTiles; ++d) prefetch_res(tile + d * (int)gridDim.x);
      }
      auto tile_loop = [&](auto act_tag) {
        constexpr int kActC = decltype(act_tag)::value;
        for (; tile < p.num_tiles; tile += gridDim.x, ++tl) {
          prefetch_res(tile + kResPrefetchTiles * (int)gridDim.x);
          const int as = tl & 1;
          const TileCoord tc = tile_coord(p, tile);
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
          {
            const long long t0 = (kProf && p.prof) ? clock64() : 0;
            mbar_wait(&tfull_bar[as], (tl >> 1) & 1);
            if (kProf && p.prof) t_wtfull += clock64() - t0;
          }
          tc_fence_after_sync();
          epilogue_tile<BLOCK_N, BLOCK_K, kSplit, kBf16, kActC>(tm, p, tc, er, t_row, wstage, wvec, sbuf, staged_nblk,
                                                                 &tempty_bar[as], lane, half);
        }
      };
      switch (p.act) {
        case DYK_ACT_LEAKY: tile_loop(std::integral_constant<int, DYK_ACT_LEAKY>{}); break;
        case DYK_ACT_MISH: tile_loop(std::integral_constant<int, DYK_ACT_MISH>{}); break;
        case DYK_ACT_RELU: tile_loop(std::integral_constant<int, DYK_ACT_RELU>{}); break;
        case DYK_ACT_RELU6: tile_loop(std::integral_constant<int, DYK_ACT_RELU6>{}); break;
        case DYK_ACT_HARDSWISH: tile_loop(std::integral_constant<int, DYK_ACT_HARDSWISH>{}); break;
        case DYK_ACT_HARDSIGMOID: tile_loop(std::integral_constant<int, DYK_ACT_HARDSIGMOID>{}); break;
        default: tile_loop(std::integral_constant<int, DYK_ACT_LINEAR>{}); break;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
    if (kProf && p.prof && ew == 0 && lane == 0) {
      atomicAdd(p.prof + PROF_EPI_WAIT_TFULL, (unsigned long long)t_wtfull);
      atomicAdd(p.prof + PROF_EPI_TOTAL, (unsigned long long)(clock64() - t_begin));
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  stamp(12, 11);
  if (warp_idx == 1) tmem_dealloc<kTmemCols>(tmem_base);
  stamp(-1, 13);
}

// ---------------------------------------------------------------------------------------------- host

int encode_map_generic(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes /* rank-1 */, const cuuint32_t* box, int swizzle_bytes,
                       const char* what) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(DYK_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 0  ? CU_TENSOR_MAP_SWIZZLE_NONE
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

static inline int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                             const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes, const char* what) {
  return encode_map_generic(map, base, rank, dims, strides_bytes, box, swizzle_bytes, what);
}

int conv1x1_thin_try(const dyk_conv_params* p, cudaStream_t stream);   // conv_thin.cu
int conv3x3_halo_try(const dyk_conv_params* p, cudaStream_t stream);   // conv_halo.cu
int conv3x3_halo2_try(const dyk_conv_params* p, cudaStream_t stream);  // conv_halo2.cu (CTA pairs)
bool conv3x3_halo2_eligible(const dyk_conv_params* p);

int res_prefetch_mode() {
  static const int mode = getenv("DYK_RES_PF") ? atoi(getenv("DYK_RES_PF")) : 0;
  return mode;
}

static bool dual_source_ok(const dyk_conv_params* p) {
  static const bool no_halo = getenv("DYK_NO_HALO") != nullptr && getenv("DYK_NO_HALO")[0] == '1';
  return !no_halo && p->Cin % 64 == 0 && !p->res && (p->dtype == DYK_F16 || p->dtype == DYK_BF16) && conv3x3_halo2_eligible(p);
}

// Spatial box (tw, th, tn) with tw*th*tn == 128 that wastes the fewest output pixels.
static void pick_tile(int Wo, int Ho, int N, int* tw, int* th, int* tn, bool one_image = false) {
  double best = -1;
  for (int w = 1; w <= 128; w *= 2) {
    for (int h = 1; w * h <= 128; h *= 2) {
      const int n = 128 / (w * h);
      if (one_image && n != 1) continue;      // per-image weights: a tile must not straddle images
      const long long covered = (long long)ceil_div(Wo, w) * w * (long long)ceil_div(Ho, h) * h *
                                (long long)ceil_div(N, n) * n;
      double eff = (double)Wo * Ho * N / (double)covered;
      // prefer wide boxes (longer contiguous runs per TMA row) on ties
      eff += 1e-6 * w - 1e-7 * n;
      if (eff > best) { best = eff; *tw = w; *th = h; *tn = n; }
    }
  }
}

template <int BLOCK_N, int BLOCK_K, int kSplit, bool kBf16, bool kAcc32 = false>
static int launch_conv(const ConvTmaps& tm, const ConvKArgs& ka, cudaStream_t stream) {
  using S = ConvSmem<BLOCK_N, BLOCK_K, kSplit>;
  auto kern = conv_tc_kernel<BLOCK_N, BLOCK_K, kSplit, kBf16, kAcc32>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    DYK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  int grid = ka.num_tiles < ka.grid_cap ? ka.num_tiles : ka.grid_cap;
  DYK_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(num_threads<kSplit>()), S::kTotal, stream, tm, ka));
  DYK_LAUNCH_OK("conv_tc_kernel");
  return DYK_OK;
}

template <int BLOCK_K, bool kBf16>
static int dispatch_n(int block_n, int split, const ConvTmaps& tm, const ConvKArgs& ka, cudaStream_t stream) {
  if (split == 4) {
    switch (block_n) {
      case 64: return launch_conv<64, BLOCK_K, 4, kBf16>(tm, ka, stream);
      case 128: return launch_conv<128, BLOCK_K, 4, kBf16>(tm, ka, stream);
      case 256: return launch_conv<256, BLOCK_K, 4, kBf16>(tm, ka, stream);
    }
  }
  switch (block_n) {
    case 32: return launch_conv<32, BLOCK_K, 2, kBf16>(tm, ka, stream);
    case 64: return launch_conv<64, BLOCK_K, 2, kBf16>(tm, ka, stream);
    case 128: return launch_conv<128, BLOCK_K, 2, kBf16>(tm, ka, stream);
    case 256: return launch_conv<256, BLOCK_K, 2, kBf16>(tm, ka, stream);
  }
  return fail(DYK_EINVAL, "bad BLOCK_N %d", block_n);
}

// Picks the N tile by a small cycle model: a k-block costs max(MMA issue, L2->SM operand fetch at ~42 B/clk/SM),
// a tile adds a fixed epilogue/hand-off cost, and the grid runs in waves over the SMs.
static int pick_block_n(int cout_store, long long m_tiles, int num_kb, int block_k, int sms) {
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

namespace dyk {
unsigned long long* g_conv_prof = nullptr;   // shared with conv_halo.cu
}
using namespace dyk;

extern "C" __attribute__((visibility("default"))) int dyk_conv_set_profile(uint64_t* dev_counters) {
  if (dev_counters != nullptr && !kProf)
    return fail(DYK_EINVAL, "dyk_conv_set_profile: this library was built without -DDYK_CONV_PROFILE");
  g_conv_prof = reinterpret_cast<unsigned long long*>(dev_counters);
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_conv2d_dual_source_supported(const dyk_conv_params* p) {
  return (p != nullptr && dual_source_ok(p)) ? 1 : 0;
}

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
  DYK_REQUIRE(p->out_f32 >= 0 && p->out_f32 <= 2 && !(p->out_f32 == 2 && p->res), "dyk_conv2d_fwd: out_f32 = %d", p->out_f32);
  DYK_REQUIRE(p->y_pix_stride >= p->Cout_store && p->x_pix_stride >= p->Cin, "dyk_conv2d_fwd: stride < channels");
  DYK_REQUIRE((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
              "dyk_conv2d_fwd: x, w, y must be 16-byte aligned");
  if (p->res) {
    DYK_REQUIRE(p->res_pix_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p->res) & 15) == 0,
                "dyk_conv2d_fwd: residual must be 16-byte aligned with stride %% 8 == 0");
    DYK_REQUIRE(!p->upsample2x, "dyk_conv2d_fwd: residual + upsample2x not supported");
  }
  const int Ho = p->out_h > 0 ? p->out_h : (p->H + 2 * p->pad - p->kh) / p->stride + 1;
  const int Wo = p->out_w > 0 ? p->out_w : (p->W + 2 * p->pad - p->kw) / p->stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_conv2d_fwd: empty output");
  DYK_REQUIRE(p->y_plane >= 0 && p->y_plane <= 4, "dyk_conv2d_fwd: y_plane %d", p->y_plane);
  DYK_REQUIRE(!(p->y_plane && (p->res || p->upsample2x || p->out_f32)),
              "dyk_conv2d_fwd: y_plane cannot be combined with residual / upsample2x / out_f32");

  if (p->x2) {
    DYK_REQUIRE(p->x_wts_raw != nullptr, "dyk_conv2d_fwd: x2 without x_wts_raw");
    DYK_REQUIRE(p->x2_pix_stride % 8 == 0 && p->x2_pix_stride >= p->Cin && (reinterpret_cast<uintptr_t>(p->x2) & 15) == 0,
                "dyk_conv2d_fwd: x2 must be 16-byte aligned with stride %% 8 == 0");
    DYK_REQUIRE(dual_source_ok(p), "dyk_conv2d_fwd: the dual-source input is not available for this layer "
                "(see dyk_conv2d_dual_source_supported)");
    return conv3x3_halo2_try(p, stream);
  }
  // 1x1 layers with Cin * Cout <= 2048 (MobileNet expand / project convs): CUDA-core kernel, conv_thin.cu
  {
    const int tr = conv1x1_thin_try(p, stream);
    if (tr <= 0) return tr;
  }
  // 3x3 stride-1 layers: halo kernel (every input pixel loaded once per tile instead of once per tap)
  static const bool no_halo = getenv("DYK_NO_HALO") != nullptr && getenv("DYK_NO_HALO")[0] == '1';
  if (!no_halo) {
    const int h2 = conv3x3_halo2_try(p, stream);
    if (h2 <= 0) return h2;
    const int hr = conv3x3_halo_try(p, stream);
    if (hr <= 0) return hr;   // launched (0) or failed (< 0); 1 = not eligible
  }

  // Stride-2 3x3 layers with 32 input channels (the first downsampling conv of every backbone, 335 MB in at 512x640 x 16)
  // moved 64-byte rows per TMA request and ran at 2.0 TB/s.  With a dense input two W-adjacent pixels are 128 contiguous
  // bytes, so the layer is run as a 3x2 convolution over pixel PAIRS with 64 "channels": 6 K blocks of 64 instead of 9 of
  // 32, 128-byte rows; the weight tile of a tap is a window of the packed [Cout][3][3][32] rows (no re-packing).
  static const bool pair_off = getenv("DYK_PAIR_W") != nullptr && getenv("DYK_PAIR_W")[0] == '0';
  const bool pair_w = !pair_off && p->stride == 2 && p->kh == 3 && p->kw == 3 && p->pad == 1 && p->Cin == 32 &&
                      p->x_pix_stride == 32 && p->W % 2 == 0 && !p->upsample2x && !p->y_plane && p->out_h == 0 && p->out_w == 0;
  const int BK = (p->Cin <= 32 && !pair_w) ? 32 : 64;
  const int swz = BK * 2;
  ConvTmaps tm;
  ConvKArgs ka;
  memset(&tm, 0, sizeof(tm));
  memset(&ka, 0, sizeof(ka));

  // 1x1 stride-1 convolutions are plain GEMMs over the flattened pixel index.
  const bool per_image = p->w_image_stride != 0;
  if (per_image) {
    DYK_REQUIRE(p->kh == 1 && p->kw == 1 && p->stride == 1 && p->pad == 0 && !p->upsample2x && !p->y_plane && p->out_f32 != 2,
                "dyk_conv2d_fwd: per-image weights are available for 1x1 stride-1 convolutions");
    DYK_REQUIRE(p->w_image_stride % p->Cin == 0 && p->w_image_stride / p->Cin >= p->Cout && (p->w_image_stride * 2) % 16 == 0,
                "dyk_conv2d_fwd: w_image_stride must be a whole number (>= Cout) of weight rows");
  }
  const bool flat = (p->kh == 1 && p->kw == 1 && p->stride == 1 && p->pad == 0 && !p->upsample2x && !p->y_plane &&
                     p->out_h == 0 && p->out_w == 0 && !per_image);
  int tw, th, tn;
  int gW = Wo, gH = Ho, gN = p->N;  // logical output grid the tiles run over
  if (flat) {
    gW = (int)((long long)p->N * p->H * p->W);
    DYK_REQUIRE((long long)p->N * p->H * p->W < (1ll << 31), "dyk_conv2d_fwd: too many pixels");
    gH = 1; gN = 1; tw = 128; th = 1; tn = 1;
  } else {
    pick_tile(Wo, Ho, p->N, &tw, &th, &tn, per_image);
  }
  const cuuint32_t abox[4] = {(cuuint32_t)BK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  const long long xs = p->x_pix_stride * 2;  // bytes per pixel step
  int rc;
  if (flat) {
    const cuuint64_t dims[4] = {(cuuint64_t)p->Cin, (cuuint64_t)gW, 1, 1};
    const cuuint64_t str[3] = {(cuuint64_t)xs, (cuuint64_t)xs * gW, (cuuint64_t)xs * gW};
    if ((rc = encode_map(&tm.a[0], p->x, 4, dims, str, abox, swz, "A/flat"))) return rc;
  } else if (pair_w) {
    for (int ph = 0; ph < 2; ++ph) {
      const int Hp = (p->H - ph + 1) / 2;
      if (Hp <= 0) continue;
      const cuuint64_t dims[4] = {64, (cuuint64_t)(p->W / 2), (cuuint64_t)Hp, (cuuint64_t)p->N};
      const cuuint64_t str[3] = {(cuuint64_t)xs * 2, (cuuint64_t)xs * p->W * 2, (cuuint64_t)xs * p->W * p->H};
      const uint8_t* base = reinterpret_cast<const uint8_t*>(p->x) + (long long)ph * p->W * xs;
      if ((rc = encode_map(&tm.a[ph * 2], base, 4, dims, str, abox, swz, "A/pair"))) return rc;
    }
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
  const int taps = pair_w ? 6 : p->kh * p->kw;
  const int cin_eff = pair_w ? 64 : p->Cin;
  const long long m_tiles = (long long)ceil_div(gW, tw) * ceil_div(gH, th) * ceil_div(gN, tn);
  int BN = pick_block_n(p->Cout_store, m_tiles, taps * ceil_div(cin_eff, BK), BK, sm_budget(p->sm_limit));
  const bool acc32 = p->out_f32 == 2;
  if (acc32) BN = 64;      // 2 x 32 register accumulators per epilogue thread
  // 16 epilogue warps when the tile time is the epilogue's: short main loop (<= 16 k-blocks) and an activation that costs
  // MUFU issue slots (Mish).  Measured: dyolov4 (Mish) 8.33 -> 8.09 ms, dyolov3 (leaky) 5.65 -> 5.75 ms, hence the act test.
  static const int force_split = getenv("DYK_EPI_SPLIT") ? atoi(getenv("DYK_EPI_SPLIT")) : 0;
  int split = (taps * ceil_div(cin_eff, BK) <= 16 && BN >= 64 && !p->out_f32 && p->act == DYK_ACT_MISH) ? 4 : 2;
  if (force_split == 2 || (force_split == 4 && BN >= 64)) split = force_split;
  if (acc32) split = 2;
  if (pair_w) {      // one filter row = 96 contiguous elements [s][c]; a tap's tile is a 64-element window of it
    const cuuint64_t dims[3] = {96, 3, (cuuint64_t)p->Cout};
    const cuuint64_t str[2] = {96 * 2, 288 * 2};
    const cuuint32_t box[3] = {64, 1, (cuuint32_t)BN};
    if ((rc = encode_map(&tm.b, p->w, 3, dims, str, box, swz, "B/pair"))) return rc;
  } else {
    const long long img_rows = per_image ? p->w_image_stride / p->Cin : 0;      // (1x1: one tap)
    const cuuint64_t rows = per_image ? (cuuint64_t)(img_rows * (p->N - 1) + p->Cout) : (cuuint64_t)p->Cout;
    const cuuint64_t dims[3] = {(cuuint64_t)p->Cin, (cuuint64_t)taps, rows};
    const cuuint64_t str[2] = {(cuuint64_t)p->Cin * 2, (cuuint64_t)p->Cin * 2 * taps};
    const cuuint32_t box[3] = {(cuuint32_t)BK, 1, (cuuint32_t)BN};
    if ((rc = encode_map(&tm.b, p->w, 3, dims, str, box, swz, "B"))) return rc;
    ka.b_img_rows = (int)img_rows;
  }
  if (!p->out_f32) {
    // every epilogue warp stores 32-row x storeC-channel sub-boxes of the tile (see epilogue_tile)
    const int storeC = BN >= 32 * split ? 32 : BN / split;
    const int sw = tw < 32 ? tw : 32;
    const int sh = th < 32 / sw ? th : 32 / sw;
    const int sn = 32 / (sw * sh);
    const cuuint32_t ybox[4] = {(cuuint32_t)storeC, (cuuint32_t)sw, (cuuint32_t)sh, (cuuint32_t)sn};
    const long long ys = p->y_pix_stride * 2;
    if (p->upsample2x || p->y_plane) {
      const int W2 = 2 * Wo, H2 = 2 * Ho;
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          if (p->y_plane && p->y_plane - 1 != ph * 2 + pw) continue;
          const cuuint64_t dims[4] = {(cuuint64_t)p->Cout_store, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)p->N};
          const cuuint64_t str[3] = {(cuuint64_t)ys * 2, (cuuint64_t)ys * W2 * 2, (cuuint64_t)ys * W2 * H2};
          uint8_t* base = reinterpret_cast<uint8_t*>(p->y) + ((long long)ph * W2 + pw) * ys;
          // a single parity plane (y_plane) is stored through y[0] like a dense output
          if ((rc = encode_map(&tm.y[p->y_plane ? 0 : ph * 2 + pw], base, 4, dims, str, ybox, storeC * 2, "Y/plane"))) return rc;
        }
    } else {
      const cuuint64_t dims[4] = {(cuuint64_t)p->Cout_store, (cuuint64_t)gW, (cuuint64_t)gH, (cuuint64_t)gN};
      const cuuint64_t str[3] = {(cuuint64_t)ys, (cuuint64_t)ys * gW, (cuuint64_t)ys * gW * gH};
      if ((rc = encode_map(&tm.y[0], p->y, 4, dims, str, ybox, storeC * 2, "Y"))) return rc;
    }
  }

  ka.tw = tw; ka.th = th; ka.tn = tn;
  ka.flat = flat ? 1 : 0;
  ka.grid_cap = sm_budget(p->sm_limit);
  ka.tiles_w = ceil_div(gW, tw); ka.tiles_h = ceil_div(gH, th); ka.tiles_b = ceil_div(gN, tn);
  ka.n_blocks = ceil_div(p->Cout_store, BN);
  for (ka.tw_log2 = 0; (1 << ka.tw_log2) < tw; ++ka.tw_log2) {}
  for (ka.th_log2 = 0; (1 << ka.th_log2) < th; ++ka.th_log2) {}
  ka.fd_nblocks = make_fastdiv((unsigned)ka.n_blocks);
  ka.fd_tiles_w = make_fastdiv((unsigned)ka.tiles_w);
  ka.fd_tiles_h = make_fastdiv((unsigned)ka.tiles_h);
  const long long nt = m_tiles * ka.n_blocks;
  DYK_REQUIRE(nt < (1ll << 31), "dyk_conv2d_fwd: too many tiles");
  ka.num_tiles = (int)nt;
  ka.k_chunks = ceil_div(cin_eff, BK);
  ka.kh = p->kh; ka.kw = pair_w ? 2 : p->kw; ka.stride = flat ? 1 : p->stride; ka.pad = p->pad;
  ka.pair_w = pair_w ? 1 : 0;
  ka.Ho = gH; ka.Wo = gW; ka.N = gN;
  ka.Cout_store = p->Cout_store;
  ka.act = p->act;
  ka.upsample2x = p->upsample2x;
  ka.out_f32 = p->out_f32;
  ka.acc_main_k = acc32 ? p->Cin / 6 : 0;       // the operand is [x1|x2|x3|x1|x2|x1]: the first Cin/6 channels are x1 (* w1)
  ka.y_f32 = p->y;
  ka.y_pix_stride = p->y_pix_stride;
  ka.scale = p->scale; ka.bias = p->bias;
  ka.res = p->res; ka.res_pix_stride = p->res_pix_stride;
  ka.prof = g_conv_prof;
  ka.res_pf = res_prefetch_mode();

  const bool bf = p->dtype == DYK_BF16;
  if (acc32) {
    DYK_REQUIRE(BK == 64 && bf, "dyk_conv2d_fwd: out_f32 = 2 (fp32-accurate accumulation) needs bf16 operands and Cin > 32");
    DYK_REQUIRE(p->Cin % 6 == 0, "dyk_conv2d_fwd: out_f32 = 2 expects the 6-way split operand (Cin %% 6 == 0)");
    return launch_conv<64, 64, 2, true, true>(tm, ka, stream);
  }
  if (BK == 64) return bf ? dispatch_n<64, true>(BN, split, tm, ka, stream) : dispatch_n<64, false>(BN, split, tm, ka, stream);
  return bf ? dispatch_n<32, true>(BN, split, tm, ka, stream) : dispatch_n<32, false>(BN, split, tm, ka, stream);
}
