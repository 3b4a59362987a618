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

#define DYK_ABI_VERSION 1

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
  int32_t out_f32;        /* 1: y is fp32 (y_pix_stride in floats); used by the head convs   */
} dyk_conv_params;
int dyk_conv2d_fwd(const dyk_conv_params* p, void* stream);

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

/* ---- depthwise convolution + BN + activation ---------------------------------------------------
 * Replaces nn.Conv2d(groups=C) (+BN+act): models.py:41 (groups key) and
 * build_utils/layers.py:224-226 (DepthwiseSeparableConv2d).  w is fp32 [k][k][C].
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

/* ---- layout / packing helpers ---------------------------------------------------------------------
 * pack: OIHW fp32 (state_dict layout, models.py:35) -> [O][kh][kw][I] dtype.
 */
int dyk_pack_weights_ohwi(const float* w_oihw, void* w_packed, int32_t O, int32_t I, int32_t kh, int32_t kw,
                          int32_t dtype, void* stream);
int dyk_nchw_f32_to_nhwc(const float* x, void* y, int64_t y_pix_stride, int32_t N, int32_t C, int32_t H,
                         int32_t W, int32_t dtype, void* stream);
int dyk_nhwc_to_nchw_f32(const void* x, int64_t x_pix_stride, float* y, int32_t N, int32_t C, int32_t H,
                         int32_t W, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYK_B200_H_ */
