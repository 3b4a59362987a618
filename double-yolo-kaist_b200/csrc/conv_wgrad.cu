// Weight gradient of a dense convolution as a split-K tcgen05 GEMM with both operands MN-major.
//
// Replaces autograd's conv2d weight gradient for nn.Conv2d (reference models.py:35; backward driven by
// train_utils/kaist_train_eval_utils.py:103 `scaler.scale(losses).backward()`):
//     dW[co][tap][ci] = sum over output pixels p of dz[p][co] * x[p (+) tap][ci]
// GEMM view per filter tap:  D[M = 128 out-channels][N = BLOCK_N in-channels] += A[M][K] * B[N][K]^T with K = pixels.
// In NHWC the pixel index is the slow dimension of both tensors, so a TMA box {64 channels, 8 x 8 pixels} lands in
// shared memory as 64 pixel rows of 128 bytes (SWIZZLE_128B) = the canonical **MN-major** operand layout of
// tcgen05.mma (8 K-rows x 64 MN-elements per swizzle atom; SBO = 1024 B between K groups, LBO = one box between
// 64-channel blocks).  No transpose is ever materialised.
//   * A = dz box at output pixel (n, h0, w0);  B = x box of the same pixels shifted by the tap (conv padding = TMA
//     out-of-bounds zero fill; stride-2 convolutions read the four parity planes of x like the forward kernel).
//   * one CTA owns one (128 out-channel block, BLOCK_N in-channel block, group of kTaps taps) and one contiguous range
//     of the pixel blocks (split-K); accumulators (kTaps x BLOCK_N fp32 columns) stay in TMEM for the whole range.
//   * the CTAs of one thread-block cluster (2 by default) are consecutive splits of the same tile: after the main loop
//     every CTA parks its fp32 accumulator tile in its own shared memory (the drained operand pipeline), and each CTA sums
//     one row slab over all CTAs of the cluster through distributed shared memory, in rank order.  Only that sum goes to
//     the workspace [split / cluster][tap][Cout_pad][Cin_pad]; with one wave of CTAs and pairs the partial volume of a
//     128x128x9 layer drops from 58 MB to 14.5 MB.  A second kernel sums the remaining partials in a fixed order and
//     writes / accumulates the OIHW gradient -> deterministic, no float atomics.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue.
#include "common.h"
#include "ptx.cuh"
#include "conv_common.cuh"
#include "vec.cuh"
#include <cstdlib>
#include <cstring>

namespace dyk {

int encode_map_generic(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes, const char* what);

struct WgradTmaps {
  CUtensorMap dz;     // (Cout, Wo, Ho, N), box {64, 8, 8, 1}
  CUtensorMap x[4];   // (Cin, W, H, N) or its four parity planes for stride 2, box {64, 8, 8, 1}
};

struct WgradKArgs {
  int tiles_w, tiles_h;          // 8 x 8 pixel blocks over (Wo, Ho)
  int n_kb;                      // tiles_w * tiles_h * N
  int kb_per_split, splits;
  int co_blocks, ci_blocks, tap_groups;
  int kw, stride, pad;
  int Cout_pad, Cin_pad, taps_total;
  FastDiv fd_tiles_w, fd_tiles_h, fd_ci_blocks, fd_tap_groups;
  float* part;
};

constexpr int kWgThreads = 192;
constexpr int kBoxBytes = 64 * 128;    // 64 pixels x 64 channels x 2 B

template <int BLOCK_N, int kTaps>
struct WgradSmem {
  static constexpr int kABytes = 2 * kBoxBytes;                       // 128 out-channels
  static constexpr int kBBytes = (BLOCK_N / 64) * kBoxBytes;          // per tap
  static constexpr int kStageBytes = kABytes + kTaps * kBBytes;
  static constexpr int kStagesRaw = (227 * 1024 - 2048) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 6 ? 6 : kStagesRaw;
  static constexpr int kTotal = kStages * kStageBytes + 2048;
  static_assert(kStages >= 2, "wgrad pipeline too shallow");
};

// MN-major SWIZZLE_128B operand descriptor: LBO = byte offset between 64-element blocks along M/N, SBO = 1024 B
// between groups of 8 K-rows.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int BLOCK_N, int kTaps, bool kBf16>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgradTmaps tm, const WgradKArgs p) {
  using S = WgradSmem<BLOCK_N, kTaps>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kAccCols = kTaps * BLOCK_N;
  constexpr uint32_t kTmemCols = kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512)));
  static_assert(kAccCols <= 512, "accumulators do not fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* done_bar = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  // warp index through a shuffle (as CUTLASS' canonical_warp_idx_sync): the compiler then knows it is warp-uniform and keeps
  // everything derived from it (role, TMEM lane quarter, staging addresses, TMA-store operands) in uniform registers
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  // work item: (co block, ci block, tap group, split)
  unsigned w = blockIdx.x;
  const unsigned split = w % (unsigned)p.splits;  w /= (unsigned)p.splits;
  const unsigned tg = w - fd_div(w, p.fd_tap_groups) * p.fd_tap_groups.div;  w = fd_div(w, p.fd_tap_groups);
  const unsigned cib = w - fd_div(w, p.fd_ci_blocks) * p.fd_ci_blocks.div;
  const unsigned cob = fd_div(w, p.fd_ci_blocks);
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.n_kb, kb0 + p.kb_per_split);
  const int nkb = kb1 - kb0;   // may be <= 0 for the last splits: the CTA then writes zeros

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tm.dz);
    tma_prefetch_desc(&tm.x[0]);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // (same value in every lane: tell the compiler)

  if (warp_idx == 0) {
    {   // TMA producer: whole warp in the loop, one elected lane issues (uniform-register operands)
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const unsigned rowt = fd_div((unsigned)kb, p.fd_tiles_w);
        const int w0 = (kb - rowt * p.fd_tiles_w.div) * 8;
        const unsigned n = fd_div(rowt, p.fd_tiles_h);
        const int h0 = (rowt - n * p.fd_tiles_h.div) * 8;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * S::kStageBytes;
        if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
        tma_load_4d(sa, &tm.dz, &full_bar[stage], cob * 128, w0, h0, n);
        tma_load_4d(sa + kBoxBytes, &tm.dz, &full_bar[stage], cob * 128 + 64, w0, h0, n);
#pragma unroll
        for (int t = 0; t < kTaps; ++t) {
          const int tap = tg * kTaps + t;
          const int r = tap / p.kw, s = tap - r * p.kw;
          int dh = r - p.pad, dw = s - p.pad, map_idx = 0;
          if (p.stride == 2) {
            const int ph = dh & 1, pw = dw & 1;
            map_idx = ph * 2 + pw;
            dh = (dh - ph) >> 1;
            dw = (dw - pw) >> 1;
          }
          uint8_t* sb = sa + S::kABytes + t * S::kBBytes;
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_4d(sb + j * kBoxBytes, &tm.x[map_idx], &full_bar[stage], cib * BLOCK_N + j * 64, w0 + dw, h0 + dh, n);
        }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    {   // whole warp in the loop, one elected lane per tcgen05 instruction (uniform-register operands; see conv_halo2.cu)
      // both operands MN-major: bits 15 / 16 of the instruction descriptor
      constexpr uint32_t idesc = umma_idesc_f16(128, BLOCK_N, kBf16 ? 1 : 0) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
        if (elect_one()) {
#pragma unroll
          for (int t = 0; t < kTaps; ++t) {
            const uint32_t sb = sa + S::kABytes + t * S::kBBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 16 pixels (= 2 groups of 8 K-rows = 2048 B) per MMA
              const uint64_t adesc = umma_desc_mnmajor(sa + k * 2048, kBoxBytes);
              const uint64_t bdesc = umma_desc_mnmajor(sb + k * 2048, kBoxBytes);
              umma_f16_ss(tmem_base + t * BLOCK_N, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (i == nkb - 1) umma_commit(done_bar);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (nkb <= 0 && elect_one()) umma_commit(done_bar);
      __syncwarp();
    }
  }
  // ------------------------------------------------------------------ epilogue: in-cluster reduction of the split-K tiles
  // Per tap: (1) the four epilogue warps copy the CTA's 128 x BLOCK_N fp32 accumulator tile from TMEM (lane = out channel)
  // into shared memory (float4 index XOR-swizzled by the row so that the 32 rows of a warp do not collide on banks);
  // (2) cluster barrier; (3) CTA `rank` sums rows [rank*128/CS, (rank+1)*128/CS) over the CS tiles in rank order (DSMEM
  // loads) and writes them to the workspace; (4) cluster barrier before the tile buffer is reused for the next tap.
  __syncwarp();   // the single-thread producer / MMA roles rejoin their warps before the warp-aligned cluster barriers
  const uint32_t cs = cluster_nctarank(), crank = cluster_ctarank();
  constexpr int kRowF4 = BLOCK_N / 4;
  float4* tile = reinterpret_cast<float4*>(smem);
  if (warp_idx >= 2) {
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
  }
  const int q = warp_idx & 3;
  const unsigned csplit = split / cs;
#pragma unroll 1
  for (int t = 0; t < kTaps; ++t) {
    if (warp_idx >= 2) {
      const int row = q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        if (nkb > 0) {
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * BLOCK_N + c, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          tile[row * kRowF4 + (((c >> 2) + j) ^ (row & 7))] =
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                          __uint_as_float(v[4 * j + 3]));
      }
    }
    cluster_sync_all();
    {
      const int tap = tg * kTaps + t;
      const int rows_per = 128 / (int)cs;
      const int row0 = (int)crank * rows_per;
      float* dst = p.part + (((long long)csplit * p.taps_total + tap) * p.Cout_pad + cob * 128) * p.Cin_pad + cib * BLOCK_N;
      const uint32_t tile_addr = smem_u32(tile);
      for (int e = threadIdx.x; e < rows_per * kRowF4; e += kWgThreads) {
        const int row = row0 + e / kRowF4, c4 = e % kRowF4;
        const uint32_t off = (uint32_t)(row * kRowF4 + (c4 ^ (row & 7))) * 16u;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t r = 0; r < cs; ++r) {
          const float4 v = ld_shared_cluster_f4(cluster_map_shared(tile_addr + off, r));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(dst + (long long)row * p.Cin_pad + c4 * 4) = acc;
      }
    }
    cluster_sync_all();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// grad_w[co][ci][r][s] (+)= sum over splits (fixed order) of part[split][tap][co][ci]
// One thread per output, coalesced over ci.  (A variant with 8 split lanes per output was not faster: the cost is the
// volume of the partials — e.g. 58 MB for 128x128x9 weights at 98 splits — not the length of the serial sum.)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int taps, int Cout_pad, int Cin_pad, int Cout,
                                    int Cin, float* __restrict__ grad, int accumulate) {   // Cin = real input channels
  const long long total = (long long)Cout * taps * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    long long t = i / Cin;
    const int tap = (int)(t % taps);
    const int co = (int)(t / taps);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp)
      s += part[(((long long)sp * taps + tap) * Cout_pad + co) * Cin_pad + ci];
    const long long gi = ((long long)co * Cin + ci) * taps + tap;
    grad[gi] = accumulate ? grad[gi] + s : s;
  }
}
// Stem convolutions (Cin <= 4): the frames are NCHW fp32 / uint8 and K = 27 is far too small for a GEMM tile, so
// this is a CUDA-core reduction: block = 32 pixel lanes x 8 threads, each thread owns 4 out channels ... kept simple:
// thread t of 256 handles output column (co, tap*Cin+ci) pairs round-robin over a strip of pixels; strips are reduced
// through the same two-stage fixed-order scheme as the BN statistics.
template <bool kBf16>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const void* __restrict__ x, int x_kind, const uint8_t* __restrict__ dz, long long dzs, int N, int H, int W,
                  int Cin, int Cout, int k, int stride, int pad, int Ho, int Wo, int strips, float* __restrict__ part) {
  extern __shared__ float sm[];
  const int KK = k * k * Cin;                    // <= 4 * 49
  float* sx = sm;                                // [64 pixels][KK]
  float* sdz = sm + 64 * KK;                     // [64 pixels][Cout]
  const long long npix = (long long)N * Ho * Wo;
  const long long per = (npix + strips - 1) / strips;
  const long long p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  const int nout = Cout * KK;
  float acc[8];                                  // outputs tid, tid+256, ... (nout <= 2048)
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long base = p0; base < p1; base += 64) {
    const int cnt = (int)min((long long)64, p1 - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * KK; i += 256) {
      const int pl = i / KK, kk = i - pl * KK;
      const int ci = kk % Cin, tap = kk / Cin, r = tap / k, s = tap - r * k;
      const long long pix = base + pl;
      const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
      const int hi = ho * stride - pad + r, wi = wo * stride - pad + s;
      float v = 0.f;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
        const long long xi = (((long long)n * Cin + ci) * H + hi) * W + wi;
        v = x_kind == 1 ? (float)reinterpret_cast<const uint8_t*>(x)[xi] / 255.0f : reinterpret_cast<const float*>(x)[xi];
      }
      sx[pl * KK + kk] = v;
    }
    for (int i = threadIdx.x; i < cnt * Cout; i += 256) {
      const int pl = i / Cout, co = i - pl * Cout;
      sdz[pl * Cout + co] = load1<kBf16>(dz, (base + pl) * dzs + co);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = threadIdx.x + j * 256;
      if (o < nout) {
        const int co = o / KK, kk = o - co * KK;
        float a = acc[j];
        for (int pl = 0; pl < cnt; ++pl) a = fmaf(sdz[pl * Cout + co], sx[pl * KK + kk], a);
        acc[j] = a;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int o = threadIdx.x + j * 256;
    if (o < nout) part[(long long)blockIdx.x * nout + o] = acc[j];
  }
}
// grad[co][ci][r][s] (+)= sum over strips; part is [strip][co][tap*Cin + ci]
__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ part, int strips, int Cout, int Cin, int taps,
                                         float* __restrict__ grad, int accumulate) {
  const int nout = Cout * taps * Cin;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nout) return;
  const int KK = taps * Cin;
  const int co = o / KK, kk = o - co * KK, tap = kk / Cin, ci = kk - tap * Cin;
  float s = 0.f;
  for (int st = 0; st < strips; ++st) s += part[(long long)st * nout + o];
  const int gi = (co * Cin + ci) * taps + tap;
  grad[gi] = accumulate ? grad[gi] + s : s;
}

struct WgradPlan {
  int BN, kTaps, co_blocks, ci_blocks, tap_groups, G;
};
static WgradPlan wgrad_plan(int Cin, int Cout, int k) {
  WgradPlan w;
  const int taps = k * k;
  w.kTaps = (taps % 3 == 0) ? 3 : 1;
  if (w.kTaps == 3) w.BN = Cin <= 64 ? 64 : 128;
  else w.BN = Cin <= 64 ? 64 : (Cin <= 128 ? 128 : 256);
  w.co_blocks = ceil_div(Cout, 128);
  w.ci_blocks = ceil_div(Cin, w.BN);
  w.tap_groups = taps / w.kTaps;
  w.G = w.co_blocks * w.ci_blocks * w.tap_groups;
  return w;
}
// Split-K factor and cluster size.  Measured on the dyolov4_fshare bs-16 training step (profiles/r02_wgrad_sweep.txt;
// DYK_WG_CLUSTER / DYK_WG_WAVES override):  one CTA per SM over the whole grid (one wave) beats two (52.9 -> 49.8 ms per
// step: half the partial tiles to write and to reduce, one prologue per SM); pairs of splits reduced through DSMEM
// (cluster of 2) gain a little more (49.5 ms); clusters of 4 / 8 LOSE (53.5 / 54.1 ms): a cluster needs all its SMs free in
// one GPC at once, the grid then runs in more, emptier waves, and every CTA waits for the slowest of its cluster.
static void wgrad_splits(const WgradPlan& w, int n_kb, int* splits_out, int* cs_out) {
  static const int max_cs = getenv("DYK_WG_CLUSTER") ? atoi(getenv("DYK_WG_CLUSTER")) : 2;
  static const int waves = getenv("DYK_WG_WAVES") ? atoi(getenv("DYK_WG_WAVES")) : 1;
  int want = (waves * num_sms()) / w.G;
  if (want > n_kb) want = n_kb;
  if (want < 1) want = 1;
  int cs = 1;
  while (cs * 2 <= want && cs * 2 <= max_cs) cs *= 2;
  *cs_out = cs;
  *splits_out = want / cs * cs;
}
static int wgrad_max_partials(const WgradPlan& w) {     // upper bound of splits / cluster size over all n_kb
  int s = (2 * num_sms()) / w.G;
  return s < 1 ? 1 : s;
}

template <int BLOCK_N, int kTaps, bool kBf16>
static int launch_wgrad(const WgradTmaps& tm, const WgradKArgs& ka, int grid, int cs, cudaStream_t stream) {
  using S = WgradSmem<BLOCK_N, kTaps>;
  auto kern = conv_wgrad_kernel<BLOCK_N, kTaps, kBf16>;
  static bool configured = false;
  if (!configured) {
    DYK_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kWgThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DYK_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tm, ka));
  DYK_LAUNCH_OK("conv_wgrad_kernel");
  return DYK_OK;
}

}  // namespace dyk

using namespace dyk;

extern "C" __attribute__((visibility("default"))) int64_t dyk_conv2d_wgrad_workspace_bytes(int32_t Cin, int32_t Cout, int32_t k) {
  if (Cin <= 0 || Cout <= 0 || k <= 0) return 0;
  const WgradPlan w = wgrad_plan(Cin, Cout, k);
  return (int64_t)wgrad_max_partials(w) * k * k * (w.co_blocks * 128) * (int64_t)(w.ci_blocks * w.BN) * 4;
}

extern "C" __attribute__((visibility("default"))) int dyk_conv2d_wgrad(
    const void* x, int64_t xs, const void* dz, int64_t dzs, float* grad, int32_t N, int32_t H, int32_t W, int32_t Cin,
    int32_t Cin_real, int32_t Cout, int32_t Cout_real, int32_t k, int32_t stride, int32_t pad, int32_t accumulate,
    int32_t dtype, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x && dz && grad && workspace, "dyk_conv2d_wgrad: null pointer");
  DYK_REQUIRE(dtype == DYK_F16 || dtype == DYK_BF16, "dyk_conv2d_wgrad: dtype %d", dtype);
  DYK_REQUIRE(stride == 1 || stride == 2, "dyk_conv2d_wgrad: stride %d", stride);
  DYK_REQUIRE(k >= 1 && k <= 7, "dyk_conv2d_wgrad: kernel %d", k);
  DYK_REQUIRE(Cin > 0 && Cin % 8 == 0 && Cout > 0 && Cout % 8 == 0 && Cout_real > 0 && Cout_real <= Cout &&
                  Cin_real > 0 && Cin_real <= Cin,
              "dyk_conv2d_wgrad: Cin=%d Cout=%d must be positive multiples of 8", Cin, Cout);
  DYK_REQUIRE(xs % 8 == 0 && dzs % 8 == 0 && xs >= Cin && dzs >= Cout, "dyk_conv2d_wgrad: pixel strides");
  DYK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0,
              "dyk_conv2d_wgrad: 16-byte alignment");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_conv2d_wgrad: empty output");
  const WgradPlan wp = wgrad_plan(Cin, Cout, k);

  WgradTmaps tm;
  WgradKArgs ka;
  memset(&tm, 0, sizeof(tm));
  memset(&ka, 0, sizeof(ka));
  int rc;
  const cuuint32_t box[4] = {64, 8, 8, 1};
  {
    const long long s = dzs * 2;
    const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
    const cuuint64_t str[3] = {(cuuint64_t)s, (cuuint64_t)s * Wo, (cuuint64_t)s * Wo * Ho};
    if ((rc = encode_map_generic(&tm.dz, dz, 4, dims, str, box, 128, "wgrad dz"))) return rc;
  }
  {
    const long long s = xs * 2;
    if (stride == 1) {
      const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t str[3] = {(cuuint64_t)s, (cuuint64_t)s * W, (cuuint64_t)s * W * H};
      if ((rc = encode_map_generic(&tm.x[0], x, 4, dims, str, box, 128, "wgrad x"))) return rc;
    } else {
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const int Hp = (H - ph + 1) / 2, Wp = (W - pw + 1) / 2;
          if (Hp <= 0 || Wp <= 0) continue;
          const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)N};
          const cuuint64_t str[3] = {(cuuint64_t)s * 2, (cuuint64_t)s * W * 2, (cuuint64_t)s * W * H};
          const uint8_t* base = reinterpret_cast<const uint8_t*>(x) + ((long long)ph * W + pw) * s;
          if ((rc = encode_map_generic(&tm.x[ph * 2 + pw], base, 4, dims, str, box, 128, "wgrad x/parity"))) return rc;
        }
    }
  }
  ka.tiles_w = ceil_div(Wo, 8);
  ka.tiles_h = ceil_div(Ho, 8);
  const long long nkb = (long long)ka.tiles_w * ka.tiles_h * N;
  DYK_REQUIRE(nkb < (1ll << 30), "dyk_conv2d_wgrad: too many pixel blocks");
  ka.n_kb = (int)nkb;
  int splits, cs;
  wgrad_splits(wp, ka.n_kb, &splits, &cs);
  ka.kb_per_split = ceil_div(ka.n_kb, splits);     // trailing splits may get no pixel block: they contribute zeros
  ka.splits = splits;
  const int parts = splits / cs;
  ka.co_blocks = wp.co_blocks; ka.ci_blocks = wp.ci_blocks; ka.tap_groups = wp.tap_groups;
  ka.kw = k; ka.stride = stride; ka.pad = pad;
  ka.Cout_pad = wp.co_blocks * 128;
  ka.Cin_pad = wp.ci_blocks * wp.BN;
  ka.taps_total = k * k;
  ka.fd_tiles_w = make_fastdiv((unsigned)ka.tiles_w);
  ka.fd_tiles_h = make_fastdiv((unsigned)ka.tiles_h);
  ka.fd_ci_blocks = make_fastdiv((unsigned)wp.ci_blocks);
  ka.fd_tap_groups = make_fastdiv((unsigned)wp.tap_groups);
  ka.part = reinterpret_cast<float*>(workspace);
  const int64_t need = (int64_t)parts * ka.taps_total * ka.Cout_pad * (int64_t)ka.Cin_pad * 4;
  DYK_REQUIRE(workspace_bytes >= need, "dyk_conv2d_wgrad: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
              (long long)need);
  const int grid = wp.G * splits;
  const bool bf = dtype == DYK_BF16;
#define DYK_WG(BN_, T_) (bf ? launch_wgrad<BN_, T_, true>(tm, ka, grid, cs, stream) : launch_wgrad<BN_, T_, false>(tm, ka, grid, cs, stream))
  if (wp.kTaps == 3) rc = wp.BN == 64 ? DYK_WG(64, 3) : DYK_WG(128, 3);
  else rc = wp.BN == 64 ? DYK_WG(64, 1) : (wp.BN == 128 ? DYK_WG(128, 1) : DYK_WG(256, 1));
#undef DYK_WG
  if (rc) return rc;
  const long long total = (long long)Cout_real * ka.taps_total * Cin_real;
  long long g = (total + 255) / 256;
  if (g > num_sms() * 16) g = num_sms() * 16;
  wgrad_reduce_kernel<<<(int)g, 256, 0, stream>>>(ka.part, parts, ka.taps_total, ka.Cout_pad, ka.Cin_pad, Cout_real, Cin_real,
                                                  grad, accumulate);
  DYK_LAUNCH_OK("wgrad_reduce_kernel");
  return DYK_OK;
}

extern "C" __attribute__((visibility("default"))) int dyk_conv2d_stem_wgrad(
    const void* x_nchw, const void* dz, int64_t dzs, float* grad, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
    int32_t k, int32_t stride, int32_t pad, int32_t accumulate, int32_t dtype, int32_t x_kind, float* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DYK_REQUIRE(x_nchw && dz && grad && workspace, "dyk_conv2d_stem_wgrad: null pointer");
  DYK_REQUIRE(Cin >= 1 && Cin <= 4 && k >= 1 && k <= 7 && Cout >= 1, "dyk_conv2d_stem_wgrad: Cin=%d k=%d", Cin, k);
  const int KK = k * k * Cin;
  DYK_REQUIRE(Cout * KK <= 2048, "dyk_conv2d_stem_wgrad: Cout*k*k*Cin = %d > 2048", Cout * KK);
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  DYK_REQUIRE(Ho > 0 && Wo > 0, "dyk_conv2d_stem_wgrad: empty output");
  const int strips = DYK_STEM_WGRAD_STRIPS;
  const size_t smem = (size_t)64 * (KK + Cout) * sizeof(float);
  DYK_REQUIRE(smem <= 48 * 1024, "dyk_conv2d_stem_wgrad: tile does not fit shared memory");
  if (dtype == DYK_BF16)
    stem_wgrad_kernel<true><<<strips, 256, smem, stream>>>(x_nchw, x_kind, (const uint8_t*)dz, dzs, N, H, W, Cin, Cout, k, stride,
                                                           pad, Ho, Wo, strips, workspace);
  else
    stem_wgrad_kernel<false><<<strips, 256, smem, stream>>>(x_nchw, x_kind, (const uint8_t*)dz, dzs, N, H, W, Cin, Cout, k, stride,
                                                            pad, Ho, Wo, strips, workspace);
  DYK_LAUNCH_OK("stem_wgrad_kernel");
  const int nout = Cout * KK;
  stem_wgrad_reduce_kernel<<<(nout + 255) / 256, 256, 0, stream>>>(workspace, strips, Cout, Cin, k * k, grad, accumulate);
  DYK_LAUNCH_OK("stem_wgrad_reduce_kernel");
  return DYK_OK;
}
