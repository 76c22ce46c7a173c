/*
 * asn_b200.h -- C ABI of libasn_b200.so: the B200 (sm_100a) implementation of the
 * AdaptSegNet output-space-adaptation hot path.
 *
 * The reference (sahngmin/AdaptSegNet) has no FFI of its own: its hot path is a set
 * of torch.nn / numpy calls made from three Python files plus one free function
 * (SURVEY.md section 8b).  Each entry point below replaces the arithmetic behind one of
 * those call sites; the citation after "replaces:" is the reference file:line.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless it says host.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises, nothing allocates: scratch comes in as (workspace, workspace_bytes),
 *     sized by the matching asn_*_workspace_bytes().
 *   - return 0 on success, a negative ASN_E* code on failure; asn_last_error() returns a
 *     thread-local message for the last failure.
 *   - activations fp32 NCHW contiguous at the boundary (what the reference's modules
 *     exchange); bf16 / NHWC staging layouts are internal.
 */
#ifndef ASN_B200_H_
#define ASN_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ASN_API __attribute__((visibility("default")))
#else
#define ASN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ASN_ABI_VERSION 2

enum {
  ASN_OK = 0,
  ASN_EINVAL = -1,      /* bad argument (shape, alignment, null pointer) */
  ASN_EWORKSPACE = -2,  /* workspace too small */
  ASN_ECUDA = -3,       /* a CUDA runtime / driver call failed */
  ASN_EUNSUPPORTED = -4 /* configuration outside what the kernels cover */
};

/* precision of the convolution paths */
enum {
  ASN_PREC_BF16 = 0, /* bf16 operands, fp32 accumulate on tcgen05 tensor cores (default) */
  ASN_PREC_FP32 = 1  /* fp32 FFMA on CUDA cores: the 1e-4 "fp32 mode" of the north star */
};

enum { ASN_LABEL_U8 = 0, ASN_LABEL_I32 = 1, ASN_LABEL_I64 = 2 };
enum { ASN_GAN_BCE = 0, ASN_GAN_MSE = 1 };

ASN_API int asn_abi_version(void);
/* sha256 prefix over the sources (csrc, this header) and compiler flags the library was built from (adaptsegnet_b200/
 * _build.py::source_id): lets a run prove which tree produced the binary it loaded */
ASN_API const char* asn_build_id(void);
ASN_API const char* asn_last_error(void);
/* number of SMs of the current device (grids are sized from it) */
ASN_API int asn_sm_count(int* out_host);

/* Optional per-kernel timing (CUDA events recorded on the launching stream around every kernel of
 * this library).  asn_prof_enable(1) clears and starts recording, asn_prof_enable(0) stops;
 * asn_prof_report synchronises the recorded events and writes a JSON object
 * {"kernel": {"launches", "ms", "flops", "bytes"}} (algorithmic flops / bytes); returns the size needed. */
/* number of kernels this library has launched in this process (every launch site counts itself) */
ASN_API int64_t asn_launch_count(void);
ASN_API int asn_prof_enable(int on);
ASN_API int64_t asn_prof_report(char* buf_host, int64_t capacity);
/* the recorded scope names in launch order, one per line (same buffer convention as asn_prof_report) */
ASN_API int64_t asn_prof_sequence(char* buf_host, int64_t capacity);

/* ------------------------------------------------------------------------------------
 * K7  confusion matrix.  replaces: compute_iou.py:15-17 (fast_hist), accumulate :57
 *   hist[n_cls*a+b] += 1 for every pixel with 0 <= a < n_cls.  `hist` (n_cls*n_cls int64)
 *   is ACCUMULATED into (zero it for a fresh matrix).  Like np.bincount in the reference, a
 *   prediction b >= n_cls under a valid label lands at the flat index n_cls*a+b; indices
 *   >= n_cls^2 (where the reference's reshape raises ValueError) are counted in
 *   *overflow instead.  Integer arithmetic, bit-exact, order independent.
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_fast_hist(const void* label, int label_dtype, const uint8_t* pred, int64_t n_px,
                  int n_cls, int64_t* hist, int64_t* overflow, void* stream);
/* same with compute_iou.py:24-28 (label_mapping) fused in: labels in [0,256) are replaced by lut256[label]
 * (device, 256 x uint8; identity for ids the mapping does not mention) before counting; other labels are left as
 * they are, like the reference's `output[input == src] = dst` passes.  Replaces label_mapping + fast_hist at
 * compute_iou.py:55-57. */
ASN_API int asn_fast_hist_lut(const void* label, int label_dtype, const uint8_t* lut256, const uint8_t* pred,
                      int64_t n_px, int n_cls, int64_t* hist, int64_t* overflow, void* stream);

/* per_class_iu + nanmean on the device (compute_iou.py:20-21,61-64): iu[c] = hist[c][c] / (row + col - diag) in float64
 * (0/0 -> nan), *miou = nanmean(iu).  One tiny launch; the matrix never visits the host. */
ASN_API int asn_per_class_iu(const int64_t* hist, int n_cls, double* iu, double* miou, void* stream);

/* ------------------------------------------------------------------------------------
 * Input pipeline, the tensor-forming tail of dataset/gta5_dataset.py:58-71 on the raw 8-bit
 * buffers: RGB HWC uint8 -> fp32 CHW in BGR order with the per-channel mean subtracted (:66-69; mean_b/g/r = the script's
 * IMG_MEAN), and label ids -> train ids through a 256-entry table (:61-64: the 19 pairs, 255 elsewhere) as int64 (the dtype
 * the loss wants, train...:595).  The PIL resize in front (:51-52) stays on the host.
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_image_u8_to_bgr_f32(const uint8_t* rgb_hwc, float* out_chw, int N, int H, int W, float mean_b, float mean_g,
                            float mean_r, void* stream);
ASN_API int asn_label_u8_to_trainid_i64(const uint8_t* ids, const uint8_t* lut256, int64_t* out, int64_t n_px, void* stream);

/* ------------------------------------------------------------------------------------
 * K2 / K2b / K9  bilinear resize, align_corners=True.
 *   replaces: model/deeplab_multi.py:188-189, evaluate_cityscapes.py:153 (nn.Upsample)
 *   and its autograd; asn_upsample_argmax_u8 replaces evaluate_cityscapes.py:163,168-169
 *   (interp -> cpu -> transpose -> np.argmax -> uint8; first maximum wins).
 *   bwd is the gather-form adjoint: no atomics, deterministic.  workspace for bwd:
 *   N*C*H*w floats (asn_upsample_bwd_workspace_bytes).
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_upsample_bilinear_fwd(const float* x, int N, int C, int h, int w, float* y, int H, int W,
                              void* stream);
ASN_API size_t asn_upsample_bwd_workspace_bytes(int N, int C, int H, int W, int h, int w);
ASN_API int asn_upsample_bilinear_bwd(const float* dy, int N, int C, int H, int W, float* dx, int h, int w,
                              void* workspace, size_t workspace_bytes, void* stream);
ASN_API int asn_upsample_argmax_u8(const float* x, int N, int C, int h, int w, uint8_t* pred, int H, int W,
                           void* stream);
/* the fork's evaluation chain as it really runs: x (h x w low-res logits) -> bilinear to (Hm x Wm), the upsample to the
 * input size inside ResNetMulti.forward (model/deeplab_multi.py:188-189) -> bilinear to (H x W), the script's `interp`
 * (evaluate_cityscapes.py:153,163) -> argmax -> uint8 (:168-169).  The two stages do not compose into one unless
 * (out-1) % (in-1) == 0; both are evaluated with ATen's float op order (bit-identical intermediate), the
 * intermediate lives in shared memory only.  ASN_EUNSUPPORTED when the second stage minifies so strongly that a
 * tile's intermediate pixels do not fit in shared memory. */
ASN_API int asn_upsample2_argmax_u8(const float* x, int N, int C, int h, int w, int Hm, int Wm, uint8_t* pred, int H,
                            int W, void* stream);
/* the same with the confusion matrix of compute_iou.py:55-57 accumulated in the SAME kernel (SURVEY.md section 8f row 3:
 * no PNG round trip, no second pass over the prediction): hist[a][b] += 1 for every pixel whose (optionally lut256-
 * mapped, compute_iou.py:24-28) label a is in [0, n_cls); `pred` may be NULL when only the matrix is wanted.
 * label: [N][H][W] of label_dtype (ASN_LABEL_*); n_cls <= 64; hist / overflow as asn_fast_hist (accumulated). */
ASN_API int asn_upsample2_argmax_hist(const float* x, int N, int C, int h, int w, int Hm, int Wm, uint8_t* pred, int H,
                              int W, const void* label, int label_dtype, const uint8_t* lut256, int n_cls,
                              int64_t* hist, int64_t* overflow, void* stream);

/* ------------------------------------------------------------------------------------
 * K3  softmax + cross entropy with ignore label over (N,C,H,W) logits.
 *   replaces: torch.nn.CrossEntropyLoss(ignore_index=255) train_gta2cityscapes_multi.py:248,
 *   359,546 applied :282,407,599-600, and utils/loss.py:7-36 (CrossEntropy2d; mask_negative=1
 *   adds its `target >= 0` mask, class_weight its `weight`, size_average its flag).
 *   stats (device, 4 x 8 bytes, written by fwd):
 *     [0] double  sum_i w_i * nll_i          [1] double  sum_i w_i   (valid pixels)
 *     [2] int64   n_valid (exact)            [3] int64   n_bad: labels neither ignored nor in
 *                                                         [0,C) -- torch raises for those
 *   loss (device float): size_average ? stats0/stats1 : stats0   (0/0 = nan as the reference).
 *   bwd: dz = gscale * (softmax(z) - onehot(y)) * w_y * valid / (size_average ? stats1 : 1);
 *   gscale is a device float (upstream gradient), may be NULL (= 1).
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_softmax_ce_fwd(const float* z, const int64_t* y, int N, int C, int H, int W, int ignore_label,
                       int mask_negative, const float* class_weight, int size_average, void* stats,
                       float* loss, void* stream);
ASN_API int asn_softmax_ce_bwd(const float* z, const int64_t* y, int N, int C, int H, int W, int ignore_label,
                       int mask_negative, const float* class_weight, int size_average,
                       const void* stats, const float* gscale, float* dz, void* stream);

/* ------------------------------------------------------------------------------------
 * K2+K3 fused ("Tier-B"): bilinear upsample (align_corners=True) of LOW-RES logits to
 * (H,W) -> softmax -> cross entropy with ignore label, forward AND the gradient w.r.t. the low-res
 * logits in one pass over the labels; the full-resolution logits are never written.
 *   replaces: interp(x) model/deeplab_multi.py:188-189 followed by CrossEntropyLoss(ignore_index=255)
 *   train_gta2cityscapes_multi.py:599-600 (or utils/loss.py:7-36) and their autograd.
 *   stats / loss / flags as asn_softmax_ce_fwd.  dz_low (N,C,h,w) = d loss / d z_low for an upstream
 *   gradient of 1 (scale it by the upstream scalar).  Deterministic (fixed summation order) except
 *   for the last bits of the double statistics.  asn_upsample_ce_supported: C == 19, H >= h, W >= w;
 *   other cases return ASN_EUNSUPPORTED (compose asn_upsample_bilinear_fwd + asn_softmax_ce_*).
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_upsample_ce_supported(int C, int h, int w, int H, int W);
ASN_API size_t asn_upsample_ce_workspace_bytes(int N, int C, int h, int w, int H, int W);
ASN_API int asn_upsample_ce_fwd_bwd(const float* z_low, const int64_t* y, int N, int C, int h, int w, int H, int W,
                            int ignore_label, int mask_negative, const float* class_weight, int size_average,
                            void* stats, float* loss, float* dz_low, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ------------------------------------------------------------------------------------
 * K4  softmax over dim 1 of (N,C,H,W).  replaces: F.softmax(pred) (implicit dim = 1)
 *   train_gta2cityscapes_multi.py:423,442,454,617-618,645-646,665-666 and its autograd.
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_softmax_fwd(const float* z, int N, int C, int H, int W, float* p, void* stream);
ASN_API int asn_softmax_bwd(const float* p, const float* dp, int N, int C, int H, int W, float* dz,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * K6  adversarial loss against a constant target (the reference builds the target tensor on
 *   the CPU each call).  replaces: BCEWithLogitsLoss / MSELoss
 *   train_gta2cityscapes_multi.py:356,358,543,545 applied :425,444,456,620-624,648-650,668-670.
 *   loss (device float) = mean(...); dx (nullable) = grad_scale * dloss/dx.
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_gan_loss_fwd_bwd(const float* x, int64_t n, float target, int kind, float grad_scale,
                         float* loss, float* dx, void* stream);

/* ------------------------------------------------------------------------------------
 * generic fp32 convolution on CUDA cores (ASN_PREC_FP32 path of K1 / K5 and the on-device
 * cross-check of the tensor-core kernels).  NCHW fp32, OIHW weights, zero padding.
 *   fwd : y = conv(x, w) + bias, optional LeakyReLU (slope; pass 1.0f for none);
 *         accumulate != 0 adds into y (used to sum the ASPP branches).
 *   dgrad: dx (+)= conv_transpose(dy, w);  wgrad: dw = corr(x, dy), db = sum dy (db nullable).
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_conv2d_fwd_f32(const float* x, const float* w, const float* bias, float* y, int N, int C,
                       int H, int W, int O, int KH, int KW, int stride, int pad, int dil,
                       float lrelu_slope, int accumulate, void* stream);
ASN_API int asn_conv2d_dgrad_f32(const float* dy, const float* w, float* dx, int N, int C, int H, int W,
                         int O, int KH, int KW, int stride, int pad, int dil, int accumulate,
                         void* stream);
ASN_API int asn_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int N, int C, int H,
                         int W, int O, int KH, int KW, int stride, int pad, int dil, void* stream);
/* dx = dy * (post > 0 ? 1 : slope): LeakyReLU backward from the stored post-activation
 * (replaces the autograd of model/discriminator.py:23,25,27,29). */
ASN_API int asn_lrelu_bwd_f32(const float* dy, const float* post, float* dx, int64_t n, float slope,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * K1 / K1b  ASPP classifier head on tcgen05 (bf16 operands, fp32 accumulate).
 *   replaces: Classifier_Module.forward model/deeplab_multi.py:117-121 (n_active = 4) and the
 *   early-return variants model/deeplab.py:112-116, model/deeplab_vgg.py:17-21 (n_active = 2),
 *   plus their autograd.
 *   Formulation: with T = 9*n_active taps and NP = round_up(T*n_cls, 16)
 *     fwd  : Z[px, NP]   = Xnhwc[px, Cin] . Wp[NP, Cin]^T      (dense GEMM, no im2col of X)
 *            y[c, p]     = bias_sum[c] + sum_t Z[p + shift_t, t*n_cls + c]   (gather-sum)
 *     dgrad: dX[Cin, px] = WpT[Cin, NP] . dYcol[px, NP]^T ,  dYcol[q, t*n_cls+c] = dy[c, q - shift_t]
 *     wgrad: dWp[Cin,NP] = Xnhwc[px, Cin]^T . dYcol[px, NP]     (MN-major operands, split-K over pixels)
 *   asn_aspp_pack_weights builds Wp (bf16 [NP][Cin]) and WpT (bf16 [Cin][NP]) from the four
 *   fp32 OIHW weights; call it after every optimizer step.
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_aspp_np(int n_cls, int n_active); /* NP */
ASN_API int asn_aspp_pack_weights(const float* const* w_oihw /* host array of n_branches device ptrs */,
                          int n_active, int n_cls, int Cin, void* wp_bf16, void* wpt_bf16,
                          void* stream);
ASN_API size_t asn_aspp_workspace_bytes(int N, int Cin, int H, int W, int n_cls, int n_active);
/* x_channels_last != 0: x (and dx) are (N,Cin,H,W) tensors stored channels_last (NHWC in memory), as a
 * channels_last trunk hands them over; y / dy stay NCHW.
 * x_bf16_keep (nullable): caller-owned N*H*W*Cin bf16 buffer that receives the NHWC bf16 copy of x the forward
 * makes anyway; asn_aspp_bwd takes it as `x_bf16` for the weight gradient (no second pass over the fp32 x). */
ASN_API int asn_aspp_fwd(const float* x_nchw, int x_channels_last, void* x_bf16_keep, const void* wp_bf16, const float* bias_sum, float* y_nchw,
                 int N, int Cin, int H, int W, int n_cls, const int* dil_host, int n_active,
                 void* workspace, size_t workspace_bytes, void* stream);
/* any of dx / dw / db may be NULL (skipped).  dw: host array of n_active device ptrs (OIHW fp32,
 * overwritten); db: device [n_cls] = sum over pixels of dy (identical for every active branch). */
ASN_API int asn_aspp_bwd(const void* x_bf16, int dx_channels_last, const void* wpt_bf16, const float* dy_nchw, float* dx_nchw,
                 float* const* dw_oihw, float* db, int N, int Cin, int H, int W, int n_cls,
                 const int* dil_host, int n_active, void* workspace, size_t workspace_bytes,
                 void* stream);

/* ------------------------------------------------------------------------------------
 * K5 / K5b / K8  FCDiscriminator on tcgen05 (bf16 NHWC activations internally).
 *   replaces: FCDiscriminator.forward model/discriminator.py:21-34 and its autograd:
 *   conv1..conv4 (k4 s2 p1 + bias + LeakyReLU 0.2) as implicit GEMMs whose im2col tiles are
 *   TMA boxes over stride-2 "parity views" of the NHWC activation; classifier (512 -> 1) is a
 *   CUDA-core reduction.  Requires ndf % 64 == 0 (reference default ndf = 64), C_in <= 32.
 *   Weights are packed once per optimizer step by asn_fcd_pack_weights into `wpack`.
 *   The forward keeps its bf16 activations in `acts` (asn_fcd_acts_bytes) for the backward.
 *   bwd: x_logits (nullable): when the forward ran with x_is_logits, pass the same logits and
 *   dx is the gradient w.r.t. them (softmax backward fused).
 *   dx (nullable: D-step, input detached) and/or dparams (nullable: G-step, parameters
 *   frozen) -- host array of 10 device pointers {conv1.w, conv1.b, ..., classifier.w,
 *   classifier.b}, fp32, same shapes as the parameters, overwritten.
 * ---------------------------------------------------------------------------------- */
ASN_API size_t asn_fcd_wpack_bytes(int n_cls, int ndf);
ASN_API int asn_fcd_pack_weights(const float* const* params_host /* 10 device ptrs, w/b per layer */,
                         int n_cls, int ndf, void* wpack, void* stream);
ASN_API size_t asn_fcd_acts_bytes(int N, int n_cls, int ndf, int H, int W);
ASN_API size_t asn_fcd_workspace_bytes(int N, int n_cls, int ndf, int H, int W);
/* layout of `acts` for inspection / tests: out_host[4*l + {0,1,2,3}] = {byte offset, H_l, W_l, C_l} of
 * the bf16 NHWC activation of level l = 0..4 (level 0 = packed input [N][H][W0p][32], W_0 = W0p). */
ASN_API int asn_fcd_act_layout(int N, int n_cls, int ndf, int H, int W, int64_t* out_host);
/* x_is_logits != 0 fuses the channel softmax (K4) into the input pack. */
ASN_API int asn_fcd_fwd(const float* x_nchw, int x_is_logits, const void* wpack, void* acts, float* out,
                int N, int n_cls, int ndf, int H, int W, void* workspace, size_t workspace_bytes,
                void* stream);
ASN_API int asn_fcd_bwd(const float* dout, const float* x_logits, const void* wpack, const void* acts, float* dx_nchw,
                float* const* dparams_host, int N, int n_cls, int ndf, int H, int W,
                void* workspace, size_t workspace_bytes, void* stream);
/* Tier-B flavour: the discriminator sees softmax(interp(z_low)) at H x W, computed inside the input pack from the
 * LOW-RES logits z_low (N,n_cls,x_h,x_w) -- replaces interp + F.softmax + D(...) (train...:617-618,645-646,665-666)
 * without materialising the (N,C,H,W) tensors; bwd returns dz_low (N,n_cls,x_h,x_w).  Same acts / wpack as above;
 * the workspace additionally holds the partial sums of the transposed interpolation. */
ASN_API size_t asn_fcd_workspace_bytes_lowres(int N, int n_cls, int ndf, int H, int W, int x_h, int x_w);
ASN_API int asn_fcd_fwd_lowres(const float* z_low, int x_h, int x_w, const void* wpack, void* acts, float* out,
                       int N, int n_cls, int ndf, int H, int W, void* workspace, size_t workspace_bytes,
                       void* stream);
ASN_API int asn_fcd_bwd_lowres(const float* dout, const float* z_low, int x_h, int x_w, const void* wpack,
                       const void* acts, float* dz_low, float* const* dparams_host, int N, int n_cls, int ndf,
                       int H, int W, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused optimizer steps over flat fp32 buffers, one launch each.
 *   asn_sgd_step: torch.optim.SGD(momentum, weight_decay) train_gta2cityscapes_multi.py:244,347,532,
 *   stepped at :681.  The flat buffers hold n elements (n % 4 == 0) cut into n_seg segments
 *   [seg_begin[s], seg_begin[s+1]) (device arrays; begins are multiples of 4); segment s uses the
 *   learning rate group_lr_host[seg_group[s]] (host array, passed by value) and receives the update
 *   seg_repeat[s] times in sequence -- the reference's parameter groups name most trunk parameters
 *   several times (model/deeplab_multi.py:196-218) and a sequential optimizer then
 *   steps them that often.  first_step != 0: momentum buffers do not exist yet -- every mention
 *   starts from buf = d_p and the last one is kept (what torch.optim.SGD's sequential path does on
 *   its first step()); afterwards buf = momentum * buf + d_p per mention.
 *   asn_adam_step: torch.optim.Adam(lr, betas, eps) train...:351-355,538-540, stepped at :682-683;
 *   `step` is the 1-based step count (bias corrections).
 *   grad_scale (both): the gradient is multiplied by it first (one rounded multiply, exactly `grad.mul_(scale)`):
 *   1 / world size after a SUM all-reduce of the flat gradient buffer, 1 otherwise.
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n,
                 const int64_t* seg_begin, const int* seg_group, const int* seg_repeat, int n_seg,
                 const float* group_lr_host, int n_groups, float momentum, float weight_decay,
                 int first_step, float grad_scale, void* stream);
ASN_API int asn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------
 * raw tcgen05 GEMM (exposed for tests / benchmarking of the tensor-core core):
 *   C[M,N] (fp32, row major, ldc) = A[M,K] . B[N,K]^T, A and B bf16 row major (K contiguous),
 *   lda/ldb in elements and multiples of 8.  split_k > 1 writes split_k partial matrices
 *   C + s*M*ldc.
 * ---------------------------------------------------------------------------------- */
ASN_API int asn_gemm_bf16_tn(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb,
                     int ldc, int split_k, void* stream);
/* same result from "MN-major" operands: C[M,N] = A[K,M]^T . B[K,N], A and B bf16 row major with the
 * reduction dimension as the ROW index (M / N contiguous) - the weight-gradient form X^T . dY read
 * straight from NHWC tensors.  lda/ldb multiples of 8. */
ASN_API int asn_gemm_bf16_nt_mn(const void* A, const void* B, float* C, int M, int N, int K, int lda, int ldb,
                        int ldc, int split_k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ASN_B200_H_ */
