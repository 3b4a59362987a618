/* dyk_b200.h — C-ABI of libdyk_b200.so: the B200 (sm_100a) implementation of the Double-YOLO-Kaist
 * forward hot path (SURVEY.md §8a rows a4–a12).
 *
 * The reference has no native layer: every op below is a PyTorch / torchvision library call made from
 * the reference's Python modules.  Each entry point cites the reference call site it replaces
 * (paths relative to the reference repo).  The Python host in double-yolo-kaist_b200/ binds these with
 * ctypes; nothing in the signatures depends on torch.
 *
 * Conventions
 *  - All device tensors are channels-last ("NHWC") unless a name says nchw.  A tensor argument is a
 *    base pointer to channel 0 of pixel (0,0,0) of the *slice* being addressed plus a pixel stride in
 *    elements, so a contiguous channel slice of a wider (concat) buffer is addressed without copies.
 *  - dtype: DYK_F16 or DYK_BF16 activations / packed weights; accumulation is always fp32.
 *  - Every function is asynchronous on `stream` (a cudaStream_t passed as void*) and returns 0 on
 *    success or a negative DYK_E* code; dyk_last_error() then returns a thread-local message.
 *    No function aborts the process.
 */
#ifndef DYK_B200_H_
#define DYK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYK_ABI_VERSION 7

enum { DYK_F16 = 0, DYK_BF16 = 1 };

/* activation ids — reference: models.py:51-64 (create_modules) */
enum {
  DYK_ACT_LINEAR = 0,
  DYK_ACT_LEAKY = 1,       /* nn.LeakyReLU(0.1)  */
  DYK_ACT_MISH = 2,        /* nn.Mish            */
  DYK_ACT_RELU = 3,        /* nn.ReLU            */
  DYK_ACT_RELU6 = 4,       /* nn.ReLU6           */
  DYK_ACT_HARDSWISH = 5,   /* nn.Hardswish       */
  DYK_ACT_HARDSIGMOID = 6  /* nn.Hardsigmoid     */
};

enum {
  DYK_OK = 0,
  DYK_EINVAL = -1,   /* bad argument / unsupported shape                 */
  DYK_ECUDA = -2,    /* CUDA runtime / driver error (message has detail) */
  DYK_EARCH = -3     /* device is not sm_100                             */
};

int dyk_abi_version(void);
/* sizeof(dyk_conv_params) as this library was compiled: lets a binding (ctypes / cgo / JNI struct mirror) verify its layout. */
int dyk_conv_params_size(void);
const char* dyk_last_error(void);
/* 0 when the current device is a compute-capability-10.x GPU and the driver entry points resolve. */
int dyk_check_device(void);

/* ---- dense convolution + folded BN + activation (+ residual) --------------------------------
 * Replaces nn.Conv2d -> nn.BatchNorm2d -> activation of a [convolutional] block
 * (models.py:28-64) and, with `res`, the following [shortcut] add (build_utils/layers.py:63-85,
 * unweighted case).  tcgen05 implicit GEMM, TMA-staged NHWC tiles, fp32 accumulation in TMEM.
 *   y[n,ho,wo,co] = act(scale[co] * sum_{r,s,ci} x[n, ho*stride+r-pad, wo*stride+s-pad, ci] * w[co,r,s,ci]
 *                       + bias[co]) (+ res[n,ho,wo,co])
 * Requirements: groups == 1, Cin % 8 == 0, Cout_stride/x strides % 8 == 0, stride in {1,2}.
 * 1x1 layers with fewer than 64 input channels (MobileNet expand / project convs) are served by a warp-MMA kernel
 * (csrc/conv_thin.cu) with the same epilogue arithmetic; the dispatch is internal.
 */
typedef struct dyk_conv_params {
  const void* x;          /* input slice, NHWC                                   */
  int64_t x_pix_stride;   /* elements between consecutive pixels of x            */
  const void* w;          /* packed weights [Cout][kh][kw][Cin], dtype           */
  const float* scale;     /* [Cout_pad] per-channel multiplier, or NULL (= 1)    */
  const float* bias;      /* [Cout_pad] per-channel addend, or NULL (= 0)        */
  void* y;                /* output slice, NHWC                                  */
  int64_t y_pix_stride;
  const void* res;        /* optional residual (same N,Ho,Wo,Cout), or NULL      */
  int64_t res_pix_stride;
  int32_t N, H, W, Cin;
  int32_t Cout;           /* real output channels (weights rows)                 */
  int32_t Cout_store;     /* channels written to y (>= Cout; extra are act(bias))*/
  int32_t kh, kw, stride, pad;
  int32_t act;            /* DYK_ACT_*                                           */
  int32_t dtype;          /* DYK_F16 / DYK_BF16                                  */
  int32_t upsample2x;     /* 1: y is the 2x nearest-upsampled map (models.py:100-101 fused) */
  int32_t out_f32;        /* 1: y is fp32 (y_pix_stride in floats); used by the head convs;
                             2: fp32 y AND fp32-accurate accumulation (fp32 mode, see below)  */
  /* ---- ABI v2 additions (zero = previous behaviour) ---- */
  int32_t out_h, out_w;   /* explicit output size (0 = derive from H, W, pad, k, stride); with pad = 0 this gives
                             "pad only at the bottom / right", which the strided-conv data gradient needs          */
  int32_t y_plane;        /* 1 + (row parity*2 + col parity): y is that parity plane of a tensor of size
                             (2*out_h, 2*out_w) and pixel stride y_pix_stride (0 = dense output)                   */
  /* ---- ABI v5 additions (zero = previous behaviour): dual-source input ----
   * WeightedFeatureFusion (build_utils/layers.py:63-85, the modality-fusion add `x * w[0] + a * w[1]` with
   * w = sigmoid(self.w) * (2 / n), n = 2) fused into the convolution that consumes it: the input of the convolution is
   *   w0 * x + w1 * x2   (rounded once to the storage type, exactly what dyk_fused_add would have stored)
   * formed in shared memory while the operand is staged, so neither modality tensor is read twice and the sum is never
   * written.  x2 has the geometry of x; x_wts_raw points to the module's raw parameter `w` (2 floats, device memory).
   * Available where dyk_conv2d_dual_source_supported() says so; dyk_conv2d_fwd fails with DYK_EINVAL elsewhere.   */
  const void* x2;
  int64_t x2_pix_stride;
  const float* x_wts_raw;
  /* ---- ABI v6 addition (zero = previous behaviour): per-image weights ----
   * SqueezeExcitation.forward's `scale * x` (build_utils/layers.py:184-190) folded into the 1x1 convolution that consumes
   * it: conv(x * gate[n]) = x . (W * gate[n]) — image n is convolved with its own weight tensor at w + n * w_image_stride
   * (elements; a whole number >= Cout of [Cin]-rows), written by dyk_scale_weights_per_image.  The gated activation tensor
   * is never written or read; tiles are kept inside one image.  1x1, stride 1, pad 0 only. */
  int64_t w_image_stride;
  /* ---- ABI v7 addition (zero = previous behaviour): SM budget ----
   * Upper bound on the SMs (= persistent CTAs) this launch occupies; 0 = all.  The two modality backbones of a dual-stream
   * model are independent between fusion points and run on two CUDA streams: with half the machine each, both kernels are
   * resident at once, a kernel boundary of one stream idles half the SMs instead of all, and the two launches' tiles are
   * rounded up to whole rounds over 74 SMs each instead of one after the other over 148. */
  int32_t sm_limit;
} dyk_conv_params;
int dyk_conv2d_fwd(const dyk_conv_params* p, void* stream);
/* 1 when dyk_conv2d_fwd accepts x2 != NULL for this layer (3x3, stride 1, pad 1, >= 256 output channels, Cin % 64 == 0:
 * the CTA-pair halo kernel), else 0.  Only the shape / flag fields of p are read. */
int dyk_conv2d_dual_source_supported(const dyk_conv_params* p);
/* out[n][co][ci] = round_dtype(w_packed[co][ci] * gate[n * gate_stride + ci]) for a 1x1 convolution's packed weights
 * ([Cout][Cin], dtype): the per-image weights of dyk_conv_params.w_image_stride (= Cout * Cin here). */
int dyk_scale_weights_per_image(const void* w_packed, const float* gate, int64_t gate_stride, void* out, int32_t N, int32_t Cout,
                                int32_t Cin, int32_t dtype, void* stream);

/* Diagnostics (no reference counterpart): when dev_counters != NULL, every later dyk_conv2d_fwd launch adds
 * its role-cycle counters into dev_counters[0..7] (device memory, 8 x uint64, caller zeroes them):
 *   0 producer cycles waiting for a free smem stage   1 producer total
 *   2 MMA-issuer cycles waiting for TMA data          3 MMA-issuer cycles waiting for a free accumulator
 *   4 MMA-issuer total                                5 epilogue (warp 0) cycles waiting for an accumulator
 *   6 epilogue total                                  7 number of CTAs
 * Pass NULL to switch the counters off (the default). */
int dyk_conv_set_profile(uint64_t* dev_counters);

/* ---- first-layer convolution reading the caller's NCHW frames ----------------------------------
 * Replaces the stem nn.Conv2d+BN+act (models.py:35-36: in_channels=3, also at second_index) fused
 * with the NCHW->NHWC / fp32->dtype conversion.  Direct (CUDA-core) kernel, Cin <= 4.
 * w is fp32 [Cout][kh][kw][Cin].  x_kind 0: x is fp32; x_kind 1: x is uint8 and is normalised as the
 * reference's callers do (`imgs.float() / 255.0`, train_utils/kaist_train_eval_utils.py:54-55,
 * evaluate.py:67-68) inside the kernel.
 */
int dyk_conv2d_stem_nchw_fwd(const void* x_nchw, const float* w, const float* scale, const float* bias,
                             void* y, int64_t y_pix_stride, int32_t N, int32_t H, int32_t W, int32_t Cin,
                             int32_t Cout, int32_t k, int32_t stride, int32_t pad, int32_t act,
                             int32_t dtype, int32_t x_kind, void* stream);
/* Multi-scale training (train_utils/kaist_train_eval_utils.py:59-71): every `accumulate` batches the reference picks a new
 * size and runs `imgs = F.interpolate(imgs, size=ns, mode='bilinear', align_corners=False)` before the model, i.e. it
 * reads the batch, writes a resized fp32 copy and the stem reads that again.  Here the stem convolution samples the resized
 * frame while it gathers its 3x3 neighbourhood: x holds the ORIGINAL Hs x Ws frames (fp32, or uint8 as above), the
 * convolution runs over their bilinear resize to H x W (PyTorch's align_corners = False arithmetic) and y is N x H x W.
 * Available for the 3x3 / stride 1 / pad 1 / 3 -> 32 channel stem of the shipped cfgs (tensor-core stem kernel). */
int dyk_conv2d_stem_nchw_resize_fwd(const void* x_nchw, const float* w_ohwi, const float* scale, const float* bias, void* y,
                                    int64_t y_pix_stride, int32_t N, int32_t Hs, int32_t Ws, int32_t H, int32_t W, int32_t Cin,
                                    int32_t Cout, int32_t k, int32_t stride, int32_t pad, int32_t act, int32_t dtype,
                                    int32_t x_kind, void* stream);

/* ---- depthwise convolution + BN + activation ---------------------------------------------------
 * Replaces nn.Conv2d(groups=C) (+BN+act): models.py:41 (groups key) and
 * build_utils/layers.py:224-226 (DepthwiseSeparableConv2d).  w is fp32 [k][k][C].
 * k in {3, 5} with stride in {1, 2}: TMA-staged tile kernel (csrc/dwconv_tile.cu; the zero padding is the TMA out-of-bounds
 * fill); other shapes: CUDA-core strip / generic kernels.  All paths accumulate in fp32 in (r, s) order: identical results.
 */
int dyk_dwconv2d_fwd(const void* x, int64_t x_pix_stride, const float* w, const float* scale,
                     const float* bias, void* y, int64_t y_pix_stride, int32_t N, int32_t H, int32_t W,
                     int32_t C, int32_t k, int32_t stride, int32_t pad, int32_t act, int32_t dtype,
                     void* stream);

/* ---- WeightedFeatureFusion (build_utils/layers.py:47-85) ---------------------------------------
 * y = w0*a + w1*b over the first C channels (w0 = w1 = 1 for the unweighted shortcut).
 * The weights are read on device: wts -> float[2] (already sigmoid(w)*2/n), or NULL.
 */
int dyk_fused_add(const void* a, int64_t a_pix_stride, const void* b, int64_t b_pix_stride, void* y,
                  int64_t y_pix_stride, int64_t npix, int32_t C, const float* wts, int32_t dtype,
                  void* stream);
/* the same with operand a seen through a SqueezeExcitation gate that was not materialised (layers.py:184-190 followed by
 * layers.py:63-85): a[n,h,w,c] * gate[n * gate_stride + c], rounded to the storage type, enters the sum in place of a. */
int dyk_fused_add_gated(const void* a, int64_t a_pix_stride, const float* gate, int64_t gate_stride, int32_t HW, const void* b,
                        int64_t b_pix_stride, void* y, int64_t y_pix_stride, int64_t npix, int32_t C, const float* wts,
                        int32_t dtype, void* stream);
/* w_out[i] = sigmoid(w_raw[i]) * 2 / n   (layers.py:66) */
int dyk_fusion_weights(const float* w_raw, float* w_out, int32_t n, void* stream);

/* ---- FeatureConcat fallback (build_utils/layers.py:32-44): copy one source into a channel slice */
int dyk_copy_slice(const void* src, int64_t src_pix_stride, void* dst, int64_t dst_pix_stride,
                   int64_t npix, int32_t C, int32_t dtype, void* stream);

/* ---- nn.MaxPool2d(k, stride, (k-1)//2) (models.py:91-94) ---------------------------------------- */
int dyk_maxpool2d(const void* x, int64_t x_pix_stride, void* y, int64_t y_pix_stride, int32_t N,
                  int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride, int32_t dtype, void* stream);

/* ---- nn.Upsample(scale_factor=s) nearest (models.py:100-101) ------------------------------------ */
int dyk_upsample_nearest(const void* x, int64_t x_pix_stride, void* y, int64_t y_pix_stride, int32_t N,
                         int32_t H, int32_t W, int32_t C, int32_t s, int32_t dtype, void* stream);

/* ---- SqueezeExcitation (build_utils/layers.py:175-190) ------------------------------------------
 * dyk_se_gate:  gate[n,c] = hardsigmoid(W2 relu(W1 mean_hw(x[n]) + b1) + b2)  (fp32 [N][C]);
 *               pooled_scratch: N * DYK_SE_MAX_SLABS * C floats (per-slab partial sums, reduced in a fixed order
 *               so the result is deterministic).
 * dyk_scale_channels: y = x * gate (broadcast over H,W).
 * w1 fp32 [Csq][C], w2 fp32 [C][Csq].
 */
#define DYK_SE_MAX_SLABS 32
int dyk_se_gate(const void* x, int64_t x_pix_stride, int32_t N, int32_t HW, int32_t C, const float* w1,
                const float* b1, const float* w2, const float* b2, int32_t Csq, float* pooled_scratch,
                float* gate, int32_t dtype, void* stream);
int dyk_scale_channels(const void* x, int64_t x_pix_stride, const float* gate, void* y,
                       int64_t y_pix_stride, int32_t N, int32_t HW, int32_t C, int32_t dtype, void* stream);
/* the two 1x1 "convolutions" of the block on the pooled vector (fc1 + ReLU, fc2 + hardsigmoid; layers.py:186-189), from
 * per-slab partial sums pooled[N][slabs][C] (slabs <= 31; the hidden activations are kept behind them in the same scratch) */
int dyk_se_mlp(float* pooled_scratch, int32_t slabs, int32_t N, int32_t HW, int32_t C, const float* w1, const float* b1,
               const float* w2, const float* b2, int32_t Csq, float* gate, void* stream);

/* ---- fp32-accurate mode (model.compute_dtype = torch.float32) -------------------------------------
 * The north-star asks for box / conf / class within 1e-3 of the reference's fp32 path (models.py:279-315 on fp32
 * tensors).  In this mode activations are fp32 NHWC; dense convolutions still run on dyk_conv2d_fwd (tcgen05, bf16
 * operands, out_f32 = 2: every 32 MMAs the TMEM accumulator is added into fp32 registers with round-to-nearest, because
 * the tensor core's own accumulation truncates) as a 3-way bf16 split product with K = 6*Cin:
 *   dyk_f32_split6:      y[pix][6C]  = [x1|x2|x3|x1|x2|x1] (bf16),  x = x1 + x2 + x3 exactly
 *   dyk_f32_pack_split6: out[O][kh*kw][6I] = [w1|w1|w1|w2|w2|w3] (bf16) from OIHW fp32
 * so that x*w = x1w1 + x2w1 + x3w1 + x1w2 + x2w2 + x1w3 up to 2^-24.  The other ops of the graph have fp32 kernels with
 * the signatures of their 16-bit counterparts (strides in fp32 elements, C a multiple of 4, 16-byte aligned pointers):
 * stem convolution from NCHW frames (nn.Conv2d + BatchNorm2d + activation, models.py:28-64), depthwise convolution
 * (layers.py:224-226), WeightedFeatureFusion / concat copy (b = NULL) (layers.py:32-85), nn.MaxPool2d, nn.Upsample,
 * SqueezeExcitation (pooled_scratch: N * DYK_SE_MAX_SLABS * C floats). */
int dyk_f32_split6(const float* x, int64_t x_pix_stride, void* y_bf16, int64_t npix, int32_t C, void* stream);
int dyk_f32_pack_split6(const float* w_oihw, void* out_bf16, int32_t O, int32_t I, int32_t kh, int32_t kw, void* stream);
int dyk_f32_stem_nchw_fwd(const void* x_nchw, const float* w_ohwi, const float* scale, const float* bias, float* y,
                          int64_t y_pix_stride, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k,
                          int32_t stride, int32_t pad, int32_t act, int32_t x_kind, void* stream);
int dyk_f32_dwconv2d_fwd(const float* x, int64_t x_pix_stride, const float* w_kkc, const float* scale, const float* bias,
                         float* y, int64_t y_pix_stride, int32_t N, int32_t H, int32_t W, int32_t C, int32_t k,
                         int32_t stride, int32_t pad, int32_t act, void* stream);
int dyk_f32_fused_add(const float* a, int64_t a_pix_stride, const float* b, int64_t b_pix_stride, float* y,
                      int64_t y_pix_stride, int64_t npix, int32_t C, const float* weights, void* stream);
int dyk_f32_maxpool2d(const float* x, int64_t x_pix_stride, float* y, int64_t y_pix_stride, int32_t N, int32_t H, int32_t W,
                      int32_t C, int32_t k, int32_t stride, void* stream);
int dyk_f32_upsample_nearest(const float* x, int64_t x_pix_stride, float* y, int64_t y_pix_stride, int32_t N, int32_t H,
                             int32_t W, int32_t C, int32_t s, void* stream);
int dyk_f32_se(const float* x, int64_t x_pix_stride, float* y, int64_t y_pix_stride, int32_t N, int32_t HW, int32_t C,
               const float* w1, const float* b1, const float* w2, const float* b2, int32_t Csq, float* pooled_scratch,
               float* gate, void* stream);

/* ---- YOLOLayer.forward, eval + train views (models.py:218-258) ------------------------------------
 * p_nhwc: head conv output [N][ny][nx][p_pix_stride] (channel a*no + o); in_kind 0 = fp16, 1 = bf16,
 *         2 = fp32 (the head convs are run with out_f32 so logits are not rounded to 16 bits).
 * p_out:  fp32 [N][na][ny][nx][no]           (the reference's permuted `p`)
 * io_out: fp32 rows of the concatenated prediction: io_out[(n*rows_total + row_off + (a*ny+gy)*nx+gx)*no + o]
 *         (models.py:315 torch.cat(x, 1)); pass NULL in training mode.
 * anchor_vec: fp32 [na][2] anchors in grid units (models.py:182 anchors/stride); v4 != 0 selects
 *         the decode of models.py:249-252, else models.py:243-246.
 */
int dyk_yolo_decode(const void* p_nhwc, int64_t p_pix_stride, float* p_out, float* io_out, int32_t N,
                    int32_t ny, int32_t nx, int32_t na, int32_t no, const float* anchor_vec, float stride,
                    int32_t v4, int64_t rows_total, int64_t row_off, int32_t in_kind, void* stream);

/* ---- non_max_suppression (build_utils/utils.py:387-464, incl. torchvision.ops.nms at :448) ---------
 * pred: fp32 [B][rows][5+nc] (x,y,w,h,obj,cls...).  All images in one launch sequence.
 * out:  fp32 [B][max_num][6] (x1,y1,x2,y2,score,label), out_count: int32 [B] (0 => reference None).
 * classes_mask: bit k set => class k allowed (0 = no filter; nc <= 64 when used).
 * workspace: device scratch of dyk_nms_workspace_bytes(B, rows, nc, multi_label) bytes.
 * conf_thres is compared in fp32 (torch compares an fp32 tensor with a Python scalar in fp32); iou_thres
 * is a double because torchvision's CPU nms compares the fp32 IoU against the double threshold.
 */
int64_t dyk_nms_workspace_bytes(int32_t B, int32_t rows, int32_t nc, int32_t multi_label);
int dyk_nms_batched(const float* pred, int32_t B, int32_t rows, int32_t nc, float conf_thres,
                    double iou_thres, int32_t multi_label, uint64_t classes_mask, int32_t agnostic,
                    int32_t max_num, float* out, int32_t* out_count, void* workspace,
                    int64_t workspace_bytes, void* stream);


/* =============================== training path ===============================
 * The reference trains through PyTorch autograd (train_utils/kaist_train_eval_utils.py:74-108:
 * `pred = model(v, l)` under autocast, `scaler.scale(loss).backward()`).  The entry points below are the native
 * forward-in-training-mode and backward kernels the Python plan (dyk/train_plan.py) strings together inside one
 * torch.autograd.Function.  All per-channel reductions use a fixed two-stage order (bit-reproducible, no float atomics).
 * `workspace` arguments are caller-provided device scratch, sizes in floats given per function.
 */
#define DYK_TRAIN_MAX_SLABS 1024

/* nn.BatchNorm2d in training mode (models.py:47), statistics part: per-channel batch mean / biased variance of the
 * 16-bit conv output z -> scale = gamma*invstd, shift = beta - mean*scale (so y = z*scale + shift), saved mean /
 * invstd for backward, and the in-place running_mean / running_var update (momentum, unbiased variance).
 * workspace: DYK_TRAIN_MAX_SLABS * 2 * C floats.
 * counters: NULL, or C/8 (at least) uint32 that are ZERO on entry and are left zero: the per-channel finalisation then runs
 * inside the reduction kernel (the block that finishes last does it) instead of as a second launch.  One counter array
 * may be shared by consecutive calls on one stream, never by concurrent ones. */
int dyk_bn_train_stats(const void* z, int64_t z_pix_stride, int64_t npix, int32_t C, int32_t dtype, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                       float* scale, float* shift, float* mean, float* invstd, float* workspace, uint32_t* counters,
                       void* stream);
/* y = act(z*scale[c] + shift[c])  — BatchNorm normalisation + activation (models.py:47-64) */
int dyk_bn_act_apply(const void* z, int64_t z_pix_stride, const float* scale, const float* shift, int32_t act, void* y,
                     int64_t y_pix_stride, int64_t npix, int32_t C, int32_t dtype, void* stream);
/* backward of activation + train-mode BatchNorm: dz (16-bit) from dy and the saved z / statistics; dgamma, dbeta are
 * ACCUMULATED into fp32 vectors (may be NULL).  workspace: (DYK_TRAIN_MAX_SLABS * 2 + 3) * C floats; counters as for
 * dyk_bn_train_stats (NULL = separate finalisation launch). */
int dyk_bn_act_bwd(const void* dy, int64_t dy_pix_stride, const void* z, int64_t z_pix_stride, const float* scale,
                   const float* shift, const float* mean, const float* invstd, const float* gamma, int32_t act,
                   int64_t npix, int32_t C, int32_t dtype, void* dz, int64_t dz_pix_stride, float* dgamma, float* dbeta,
                   float* workspace, uint32_t* counters, void* stream);
/* out[c] (+)= sum over pixels of x[pix][c]  (bias gradient of the head convs).  workspace: DYK_TRAIN_MAX_SLABS*2*C floats */
int dyk_chan_sum(const void* x, int64_t x_pix_stride, int64_t npix, int32_t C, int32_t dtype, float* out,
                 int32_t accumulate, float* workspace, void* stream);
/* dst = alpha*src (+ dst when accumulate); alpha read on device (NULL = 1).  Gradient routing of [shortcut] / [route]. */
int dyk_axpby(const void* src, int64_t src_pix_stride, const float* alpha, void* dst, int64_t dst_pix_stride,
              int64_t npix, int32_t C, int32_t accumulate, int32_t dtype, void* stream);
/* WeightedFeatureFusion.w gradient (layers.py:65-73): grad_w[i] += sum(dy*operand_i) * (2/n) * sigmoid'(w_i), n == 2.
 * workspace: 2 * DYK_TRAIN_MAX_SLABS * 2 * C floats. */
int dyk_fusion_weights_bwd(const void* dy, int64_t dy_pix_stride, const void* a, int64_t a_pix_stride, const void* b,
                           int64_t b_pix_stride, int64_t npix, int32_t C, int32_t dtype, const float* w_raw, int32_t n,
                           float* grad_w, float* workspace, void* stream);
/* nn.MaxPool2d backward (gradient goes to the first maximum of each window, like PyTorch).
 * idx_workspace: N*Ho*Wo*C int32. */
int dyk_maxpool2d_bwd(const void* x, int64_t x_pix_stride, const void* dy, int64_t dy_pix_stride, void* dx,
                      int64_t dx_pix_stride, int32_t N, int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride,
                      int32_t accumulate, int32_t dtype, int32_t* idx_workspace, void* stream);
/* nn.Upsample(nearest, s) backward: dx[h][w] (+)= sum of the s x s block of dy.  H, W are the INPUT (small) sizes. */
int dyk_upsample_nearest_bwd(const void* dy, int64_t dy_pix_stride, void* dx, int64_t dx_pix_stride, int32_t N, int32_t H,
                             int32_t W, int32_t C, int32_t s, int32_t accumulate, int32_t dtype, void* stream);
/* SqueezeExcitation backward (layers.py:184-190): dx (+)= dy*gate + dmean/HW, and the fc1/fc2 weight / bias gradients
 * ACCUMULATED into gw1 [Csq][C], gb1 [Csq], gw2 [C][Csq], gb2 [C].  pooled / gate: as written by dyk_se_gate in the
 * forward.  workspace: 72 * N * C floats. */
int dyk_se_bwd(const void* x, int64_t x_pix_stride, const void* dy, int64_t dy_pix_stride, void* dx, int64_t dx_pix_stride,
               int32_t N, int32_t HW, int32_t C, const float* w1, const float* b1, const float* w2, const float* b2,
               int32_t Csq, const float* pooled, const float* gate, float* gw1, float* gb1, float* gw2, float* gb2,
               int32_t accumulate, int32_t dtype, float* workspace, void* stream);
/* inverse of YOLOLayer's view+permute (models.py:229) for the gradient: dp fp32 [N][na][ny][nx][no] -> 16-bit NHWC
 * [N][ny][nx][Cpad], channels >= na*no zero-filled. */
int dyk_yolo_train_bwd(const float* dp, int32_t N, int32_t na, int32_t ny, int32_t nx, int32_t no, void* dz, int32_t Cpad,
                       int32_t dtype, void* stream);
/* data-gradient weights: OIHW fp32 -> [I][kh][kw][Opad] 16-bit, taps rotated 180 degrees, so that
 * dx = dyk_conv2d_fwd(dz, W', pad' = k-1-pad) for stride-1 convolutions. */
int dyk_pack_weights_dgrad(const float* w_oihw, void* w_packed, int32_t O, int32_t I, int32_t kh, int32_t kw, int32_t Opad,
                           int32_t dtype, void* stream);
/* weight gradient of a dense convolution (tcgen05, split over pixel ranges, deterministic two-pass reduction):
 *   grad_w[co][ci][r][s] (+)= sum_{n,ho,wo} dz[n,ho,wo,co] * x[n, ho*stride+r-pad, wo*stride+s-pad, ci]
 * x: NHWC 16-bit input of the convolution, dz: NHWC 16-bit gradient of its (pre-BN) output, grad_w: fp32 OIHW
 * (the nn.Conv2d weight layout).  Requirements as dyk_conv2d_fwd; Cout % 8 == 0 (pad dz for the 18-channel heads).
 * workspace: dyk_conv2d_wgrad_workspace_bytes(...) bytes. */
int64_t dyk_conv2d_wgrad_workspace_bytes(int32_t Cin, int32_t Cout, int32_t k);
int dyk_conv2d_wgrad(const void* x, int64_t x_pix_stride, const void* dz, int64_t dz_pix_stride, float* grad_w_oihw,
                     int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cin_real, int32_t Cout, int32_t Cout_real,
                     int32_t k, int32_t stride, int32_t pad, int32_t accumulate, int32_t dtype, void* workspace,
                     int64_t workspace_bytes, void* stream);
/* Cin_real / Cout_real: channels of the OIHW gradient actually written when x / dz carry zero-padded channels
 * (stem frames padded 3 -> 8, head logits padded 18 -> 32). */
/* every convolution weight of a model in one launch (the weights change each optimizer step): descs is a DEVICE array
 * int64 [n][8] = (w OIHW fp32 ptr, forward-layout ptr [O][kh][kw][I], dgrad-layout ptr [I][kh][kw][Opad] or 0, O, I,
 * taps = kh*kw <= 9, Opad, index of this layer's first tile), tiles = ceil(O/32)*ceil(I/32) per layer in order,
 * total_tiles their sum.  Rows O..Opad of the dgrad layout are not written (zero them once). */
int dyk_pack_weights_multi(const int64_t* descs, int32_t n, int32_t total_tiles, int32_t dtype, void* stream);
/* caller's NCHW fp32 / uint8 frames (as dyk_conv2d_stem_nchw_fwd, uint8 divided by 255) -> NHWC 16-bit with the channel
 * dimension zero-padded to 8, the operand layout the tensor-core wgrad kernel needs for the stem convolutions. */
int dyk_frames_to_nhwc8(const void* x_nchw, void* y, int32_t N, int32_t Cin, int32_t H, int32_t W, int32_t dtype,
                        int32_t x_kind, void* stream);
/* 3-channel frames -> NHWC 16-bit im2col rows of the 3x3 / pad 1 neighbourhood: y[n][h][w][ci*9 + r*3 + s], 32 channels
 * (27..31 zero).  dyk_conv2d_wgrad on it with k = 1, Cin = 32, Cin_real = 27 yields the stem's OIHW weight gradient. */
int dyk_frames_to_im2col32(const void* x_nchw, void* y, int32_t N, int32_t H, int32_t W, int32_t dtype, int32_t x_kind,
                           void* stream);
/* the same rows taken from the bilinear resize of Hs x Ws frames to H x W (see dyk_conv2d_stem_nchw_resize_fwd): the stem's
 * weight gradient in a multi-scale training step without a resized copy of the batch. */
int dyk_frames_to_im2col32_resize(const void* x_nchw, void* y, int32_t N, int32_t Hs, int32_t Ws, int32_t H, int32_t W,
                                  int32_t dtype, int32_t x_kind, void* stream);
/* weight gradient of the stem convolutions (Cin <= 4, NCHW fp32 / uint8 frames as in dyk_conv2d_stem_nchw_fwd).
 * workspace: DYK_STEM_WGRAD_STRIPS * Cout * k*k*Cin floats. */
#define DYK_STEM_WGRAD_STRIPS 592
int dyk_conv2d_stem_wgrad(const void* x_nchw, const void* dz, int64_t dz_pix_stride, float* grad_w_oihw, int32_t N,
                          int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t k, int32_t stride, int32_t pad,
                          int32_t accumulate, int32_t dtype, int32_t x_kind, float* workspace, void* stream);

/* ---- depthwise convolution backward (nn.Conv2d(groups = C): models.py:41, build_utils/layers.py:224; autograd in the
 * reference).  w and the forward layout are as in dyk_dwconv2d_fwd (fp32 [k][k][C]); H, W are the INPUT sizes.
 * dgrad: dx (+)= sum over taps of dz[(h + pad - r)/stride, (w + pad - s)/stride] * w[r][s]   (gather form, no atomics).
 * wgrad: grad_w is the state_dict layout [C][1][k][k] fp32; k = 3 or 5; workspace: DYK_DW_WGRAD_SLABS * k*k*C floats,
 * reduced in a fixed order (bit-reproducible). */
#define DYK_DW_WGRAD_SLABS 256
int dyk_dwconv2d_dgrad(const void* dz, int64_t dz_pix_stride, const float* w, void* dx, int64_t dx_pix_stride, int32_t N,
                       int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride, int32_t pad, int32_t accumulate,
                       int32_t dtype, void* stream);
int dyk_dwconv2d_wgrad(const void* x, int64_t x_pix_stride, const void* dz, int64_t dz_pix_stride, float* grad_w, int32_t N,
                       int32_t H, int32_t W, int32_t C, int32_t k, int32_t stride, int32_t pad, int32_t accumulate,
                       int32_t dtype, float* workspace, void* stream);

/* ---- scale_coords + clip_coords (build_utils/utils.py:60-92): xyxy boxes from the network input size back to the
 * original image, in place: x = clamp((x - pad_x) / gain, 0, img_w), y = clamp((y - pad_y) / gain, 0, img_h) for the first
 * four columns of n rows (row_stride floats apart).  clip_coords alone = pad 0, gain 1. */
int dyk_scale_coords(float* boxes, int64_t row_stride, int32_t n, float pad_x, float pad_y, float gain, float img_w,
                     float img_h, void* stream);

/* ---- training loss of the YOLO heads (build_utils/utils.py:209-384: compute_loss + build_targets; bbox_iou :95-138,
 * wh_iou :166-172), forward and gradient fused, one head per call, no host synchronisation.
 *
 * dyk_yolo_build_targets: targets [nt][6] fp32 (image, class, x, y, w, h normalised) against the na anchors of one head
 *   (anchor_vec = anchors / stride) on an ny x nx grid; positives in the reference's anchor-major order:
 *   *count (device) = nb, idx[k] = (b, a, gj, gi), tbox[k] = (gx - gi, gy - gj, gw, gh), tcls[k].  Arrays hold na*nt entries.
 * dyk_yolo_loss_head: p, dp fp32 [B][na][ny][nx][no].  dp is fully written with d(w_box*lbox + w_obj*lobj + w_cls*lcls)/dp of
 *   this head (box channels 0-3, objectness 4, classes 5+; class loss only when no > 6 as in the reference's nc > 1).
 *   acc[3] (device, zeroed by the caller before the first head) accumulates the per-head means; with finalize != 0 the
 *   call also writes out[3] = acc * (w_box, w_obj, w_cls) (utils.py:296-298).  ciou selects CIoU (hyp has the key 'ciou')
 *   or GIoU; gr = model.gr; focal loss (fl_gamma > 0) is not implemented.  err (device int) is set to 1 when a target
 *   indexes outside the batch / grid, 2 when a class id is >= nc (the reference raises for both).
 *   workspace: dyk_yolo_loss_workspace_floats(max_pos = na*nt, no, cells = B*na*ny*nx) floats.
 * dyk_yolo_loss_scale_grad: out = dp * upstream[group of the channel] (upstream = 3 device floats: grads of the box / obj /
 *   class loss outputs), the backward of the autograd node. */
int dyk_yolo_build_targets(const float* targets, int32_t nt, const float* anchor_vec, int32_t na, int32_t ny, int32_t nx,
                           float iou_t, int32_t* count, int32_t* idx, float* tbox, int32_t* tcls, void* stream);
int64_t dyk_yolo_loss_workspace_floats(int32_t max_pos, int32_t no, int64_t cells);
int dyk_yolo_loss_head(const float* p, float* dp, int32_t B, int32_t na, int32_t ny, int32_t nx, int32_t no,
                       const int32_t* count, const int32_t* idx, const float* tbox, const int32_t* tcls,
                       const float* anchor_vec, int32_t max_pos, int32_t v4, int32_t ciou, float gr, float obj_pw, float cls_pw,
                       float w_box, float w_obj, float w_cls, float* acc, int32_t finalize, float* out, int32_t* err,
                       float* workspace, void* stream);
int dyk_yolo_loss_scale_grad(const float* dp, float* out, int64_t n, int32_t no, const float* upstream, void* stream);

/* ---- layout / packing helpers ---------------------------------------------------------------------
 * pack: OIHW fp32 (state_dict layout, models.py:35) -> [O][kh][kw][I] dtype.
 */
int dyk_pack_weights_ohwi(const float* w_oihw, void* w_packed, int32_t O, int32_t I, int32_t kh, int32_t kw,
                          int32_t dtype, void* stream);
/* eval-mode nn.BatchNorm2d folded into the convolution epilogue's per-channel (scale, bias) (models.py:44-47: the
 * reference runs conv -> BatchNorm2d -> activation as three library calls), for every layer of a model in one launch:
 * scale = gamma * rsqrt(running_var + eps), bias = beta - running_mean * scale; a layer without BN gets its conv bias.
 * descs: DEVICE int64 [n][10] = (gamma or 0, beta or 0, running_mean, running_var or 0 = no BN, conv bias or 0,
 * scale_out or 0, bias_out or 0, C, Cpad, float bits of eps); outputs are Cpad floats, zero beyond C. */
int dyk_fold_bn_multi(const int64_t* descs, int32_t n, void* stream);
int dyk_nchw_f32_to_nhwc(const float* x, void* y, int64_t y_pix_stride, int32_t N, int32_t C, int32_t H,
                         int32_t W, int32_t dtype, void* stream);
int dyk_nhwc_to_nchw_f32(const void* x, int64_t x_pix_stride, float* y, int32_t N, int32_t C, int32_t H,
                         int32_t W, int32_t dtype, void* stream);

/* ---- fused multi-tensor optimizer step (train.py:85-91: optim.SGD(momentum, nesterov=True, weight_decay) or
 * optim.Adam(betas=(momentum, 0.999), weight_decay); stepped by GradScaler in kaist_train_eval_utils.py:103-108) --------
 * descs: device int64 [n][6] = (param ptr, grad ptr, state1 ptr, state2 ptr or 0, numel, first block of this tensor);
 * a tensor occupies ceil(numel / dyk_optim_block_elems()) consecutive blocks; total_blocks = their sum.  All tensors fp32.
 * grad_scale / found_inf: optional device scalars of torch.amp.GradScaler (gradients are divided by *grad_scale; the
 * whole step is skipped when *found_inf != 0) — no host synchronisation.
 * SGD : g += wd*p; buf = first_step ? g : momentum*buf + (1-dampening)*g; g = nesterov ? g + momentum*buf : buf; p -= lr*g
 * Adam: g += wd*p; m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g^2; p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps) */
int32_t dyk_optim_block_elems(void);
int dyk_optim_sgd_multi(const int64_t* descs, int32_t n, int64_t total_blocks, float lr, float momentum, float dampening,
                        float weight_decay, int32_t nesterov, int32_t first_step, const float* grad_scale,
                        const float* found_inf, void* stream);
int dyk_optim_adam_multi(const int64_t* descs, int32_t n, int64_t total_blocks, float lr, float beta1, float beta2, float eps,
                         float weight_decay, int64_t step, const float* grad_scale, const float* found_inf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYK_B200_H_ */
