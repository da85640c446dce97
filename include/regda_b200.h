/*
 * regda_b200 -- C ABI of the B200-native RegDA self-training hot path.
 *
 * The reference (StuLiu/RegDA) is pure Python/PyTorch and has no FFI of its own; the
 * "operator interface" it calls on this path is torch / torch_scatter.  Every entry point
 * below names the reference call site (file:line under the reference tree) it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the caller owns every buffer (inputs, outputs, workspace); the library never
 *    allocates, frees or keeps a pointer past return; inputs are const and never written;
 *  - `stream` is a cudaStream_t passed as void*; kernels are launched on it and nothing
 *    synchronises the host (safe inside CUDA-graph capture);
 *  - return value: REGDA_OK or a REGDA_ERR_* code; regda_last_error() gives the text for
 *    the calling thread;
 *  - data-dependent domain errors (label / region id / probability out of range -- the
 *    cases where the reference raises from one_hot / scatter / assert) cannot be known
 *    without a sync, so kernels OR REGDA_FLAG_* bits into the caller's int32 `flags`
 *    word (may be NULL); the output at offending pixels is unspecified.
 *  - tensors are dense row-major with the shapes given; int64 labels / regions exactly as
 *    the reference's LongTensors.
 */
#ifndef REGDA_B200_H
#define REGDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REGDA_ABI_VERSION 3   /* 2: bn_forward relu_mask, bn_backward beta + dz_ready; new entry points (dgrad_bnred, stem, inference, pcl)
                               * 3: stem patch rows padded per filter row, G_k in (o, tap) column order, ppm gather / scatter removed (weights in place) */

#define REGDA_OK 0
#define REGDA_ERR_INVALID_ARG 1
#define REGDA_ERR_WORKSPACE 2
#define REGDA_ERR_CUDA 3
#define REGDA_ERR_UNSUPPORTED 4

#define REGDA_FLAG_LABEL_RANGE 1  /* reference: RuntimeError from one_hot (local_region_homog.py:121) */
#define REGDA_FLAG_REGION_RANGE 2 /* reference: RuntimeError from scatter index (local_region_homog.py:140) */
#define REGDA_FLAG_PROB_RANGE 4   /* reference: AssertionError (pseudo_generation.py:71) */

int regda_abi_version(void);
const char *regda_last_error(void);
/* 0: LRH auto, 1: force the generic (global-bin) path, 2: force the cluster path (error if it does not
 * fit), 3: cluster path with the loop-over-distinct-regions histogram (A/B switch for profiling) */
int regda_set_lrh_path(int mode);
/* which path the last regda_lrh_forward on this thread took: 1 generic, 2 cluster; cluster size in *cluster */
int regda_lrh_last_path(int *cluster);

/* ---- Local Region Homogenizing ------------------------------------------------------
 * Replaces Homogenizer.forward, regda/utils/local_region_homog.py:125-152 (incl. the
 * torch_scatter.scatter(reduce='sum') at :140 and _index2onehot at :107-123).
 * labels, regions, out: int64 [b][h*w].  region ids must lie in [0, region_bound);
 * region_bound plays the role of the reference's batch-global `regions.max()+1`.
 * percent is the Python double of the reference; it is compared as float32, as there. */
size_t regda_lrh_workspace_bytes(int b, int64_t hw, int class_num, int64_t region_bound);
int regda_lrh_forward(const int64_t *labels, const int64_t *regions, int64_t *out,
                      int b, int64_t hw, int class_num, int64_t ignore_label, double percent,
                      int64_t region_bound, int32_t *flags,
                      void *workspace, size_t workspace_bytes, void *stream);
/* max(regions)+1 (and a REGION_RANGE flag for negative ids) without leaving the device:
 * bound_out is one int64.  Replaces the implicit index.max() inside scatter (:140). */
int regda_region_bound(const int64_t *regions, int64_t n, int64_t *bound_out, int32_t *flags, void *stream);

/* ---- pseudo_selection ---------------------------------------------------------------
 * Replaces regda/gast/pseudo_generation.py:59-93.  soft: float32 [b][c][hw];
 * out: int64 [b][hw].  workspace: regda_select_workspace_bytes(b, c). */
size_t regda_select_workspace_bytes(int b, int c);
int regda_pseudo_select(const float *soft, int64_t *out, int b, int c, int64_t hw,
                        double cutoff_top, double cutoff_low, int64_t ignore_label,
                        int32_t *flags, void *workspace, size_t workspace_bytes, void *stream);

/* ---- Aligner.label_refine (label_t_sup=None, mode='all') -----------------------------
 * Replaces regda/gast/alignment.py:194-265 with _pearson_dist :396-423, _softmax_T
 * :283-286 and _logits_norm :288-298.
 *   feat_nhwc   float32 [b][h][w][k]   (channels-last view of the reference's [b,k,h,w])
 *   prototypes  float32 [c][k]
 *   pred1/pred2 float32 [b][c][h][w]   (pred2 may be NULL: single-head form of :232-234)
 *   soft_in/out float32 [b][c][H][W]
 * regda_refine_select is the fused form used by the training step: label_refine followed
 * by pseudo_selection (:218 of tools/train_ssl_reg.py) without materialising the refined
 * [b,c,H,W] tensor; hard_out int64 [b][H][W].
 * workspace: regda_refine_workspace_bytes(b, c, k, h, w) for both. */
size_t regda_refine_workspace_bytes(int b, int c, int k, int h, int w);
/* regda_refine_select with this (larger) workspace keeps the refined probabilities between its two passes instead of
 * recomputing them (same result; +b*c*H*W floats) */
size_t regda_refine_select_workspace_bytes(int b, int c, int k, int h, int w, int H, int W);
int regda_label_refine(const float *feat_nhwc, const float *prototypes,
                       const float *pred1, const float *pred2,
                       const float *soft_in, float *soft_out,
                       int b, int c, int k, int h, int w, int H, int W, double temp,
                       void *workspace, size_t workspace_bytes, void *stream);
/* regda_refine_select with the feature rows in bf16 (the model's InstanceNorm kernel writes bf16; float32 arithmetic inside) */
int regda_refine_select_bf16feat(const void *feat_nhwc_bf16, const float *prototypes, const float *pred1, const float *pred2,
                                 const float *soft_in, int64_t *hard_out, int b, int c, int k, int h, int w, int H, int W,
                                 double temp, double cutoff_top, double cutoff_low, int64_t ignore_label,
                                 void *workspace, size_t workspace_bytes, void *stream);
int regda_refine_select(const float *feat_nhwc, const float *prototypes,
                        const float *pred1, const float *pred2,
                        const float *soft_in, int64_t *hard_out,
                        int b, int c, int k, int h, int w, int H, int W, double temp,
                        double cutoff_top, double cutoff_low, int64_t ignore_label,
                        void *workspace, size_t workspace_bytes, void *stream);
/* _pearson_dist alone (alignment.py:396-423): rows float32 [n][k], prototypes [c][k] ->
 * dist float32 [n][c].  workspace: regda_pearson_workspace_bytes(c, k). */
size_t regda_pearson_workspace_bytes(int c, int k);
int regda_pearson_dist(const float *rows, const float *prototypes, float *dist,
                       int64_t n, int c, int k, void *workspace, size_t workspace_bytes, void *stream);

/* ---- DownscaleLabel + prototype update ------------------------------------------------
 * regda_downscale_label replaces DownscaleLabel.forward, alignment.py:466-481:
 *   label int64 [b][H][W] -> out int64 [b][H/scale][W/scale].
 * regda_class_sums replaces the one_hot x feature broadcast-sum of
 *   _compute_local_prototypes (:313-320) and update_avg (:107-119):
 *   feat_nhwc float32 [n][k], label_ds int64 [n] -> sums float32 [c][k], counts float32 [c]
 *   (ACCUMULATED into sums/counts when accumulate != 0, which is update_avg's running sum).
 * regda_prototype_ema replaces :319-325 + _ema :435-438 in place on `prototypes`;
 * regda_prototype_init_avg replaces init_avg :121-122. */
int regda_downscale_label(const int64_t *label, int64_t *out, int b, int H, int W, int scale,
                          int n_classes, int64_t ignore_label, double min_ratio,
                          int32_t *flags, void *stream);
size_t regda_class_sums_workspace_bytes(int64_t n, int c, int k);
int regda_class_sums_bf16feat(const void *feat_nhwc_bf16, const int64_t *label_ds, float *sums, float *counts,
                              int64_t n, int c, int k, int64_t ignore_label, int accumulate,
                              void *workspace, size_t workspace_bytes, void *stream);
int regda_class_sums(const float *feat_nhwc, const int64_t *label_ds, float *sums, float *counts,
                     int64_t n, int c, int k, int64_t ignore_label, int accumulate,
                     void *workspace, size_t workspace_bytes, void *stream);
int regda_prototype_ema(float *prototypes, const float *sums, const float *counts,
                        int c, int k, double decay, void *stream);
int regda_prototype_init_avg(float *prototypes, const float *sums, const float *counts,
                             int c, int k, void *stream);

/* ---- loss_calc + CrossEntropy --------------------------------------------------------
 * Replaces regda/utils/tools.py:240-252 (bilinear align_corners=True upsample of each
 * head) + regda/gast/balance.py:88-101 (per-pixel CE, ignore_index, mean over ALL pixels).
 *   pred     float32 [b][c][h][w] low-resolution logits of ONE head
 *   label    int64   [b][H][W]
 *   loss     float32 [1]  : mean CE of this head (written, not accumulated)
 *   dpred    float32 [b][c][h][w] : d(loss)/d(pred) * grad_scale  (may be NULL: forward only)
 *   class_weight float32 [c] or NULL: ClassBalance per-class pixel weight (balance.py:30-33)
 * workspace: regda_ce_workspace_bytes(b, h, H). */
size_t regda_ce_workspace_bytes(int b, int h, int H);
int regda_ce_bilinear(const float *pred, const int64_t *label, float *loss, float *dpred,
                      int b, int c, int h, int w, int H, int W, int64_t ignore_label,
                      double grad_scale, const float *class_weight,
                      int32_t *flags, void *workspace, size_t workspace_bytes, void *stream);

/* ---- ClassBalance counts (flag-gated --bcs/--bct, balance.py:43-66) ---------------------
 * counts_out int64 [c+1]: per-class pixel counts, then the number of non-ignored pixels. */
int regda_class_count(const int64_t *label, int64_t n, int c, int64_t ignore_label,
                      int64_t *counts_out, int32_t *flags, void *stream);

/* ---- clip_grad_norm_ + SGD(momentum, weight_decay) -------------------------------------
 * Replaces tools/train_ssl_reg.py:239-241 over ONE flat fp32 parameter arena.
 * regda_sumsq: sum of squares -> sumsq_out float32 [1] (fixed-order two-stage; accumulate != 0
 *   adds to the value already there, for arenas split over several calls).
 * regda_sgd_step: g = grad*grad_scale; coef = min(1, max_norm/(sqrt(sumsq)*grad_scale + 1e-6));
 *   g = coef*g + wd*p;  buf = first_step ? g : momentum*buf + g;  p -= lr*buf;
 *   optional bf16 shadow copy of p (param_bf16, may be NULL); sumsq NULL = no clipping;
 *   lr_device (may be NULL) is a float32 device scalar that overrides `lr`, so a captured CUDA
 *   graph can follow the learning-rate schedule (tools.py:199-207) without re-capture. */
size_t regda_sumsq_workspace_bytes(int64_t n);
int regda_sumsq(const float *x, int64_t n, float *sumsq_out, int accumulate,
                void *workspace, size_t workspace_bytes, void *stream);
int regda_sgd_step(float *param, const float *grad, float *momentum_buf, void *param_bf16,
                   int64_t n, const float *sumsq, double max_norm, double grad_scale, double lr,
                   const float *lr_device, double momentum, double weight_decay, int first_step,
                   void *stream);
/* ExponentialMovingAverage.update, regda/utils/ema.py:46-51 (opt-in; unused by the reference loop) */
int regda_ema_update(float *shadow, const float *param, int64_t n, double decay, void *stream);


/* ---- convolution on the tcgen05 tensor cores -------------------------------------------
 * Replaces the cuDNN convolutions behind nn.Conv2d in regda/_resnets.py:92-112 (Bottleneck) and
 * regda/models/Encoder.py:17-23,33-40 (PPM branch convs, fuse conv).  Implicit GEMM, any dilation / padding:
 *   x   bf16 [n][h][w][cin]        (channels-last)
 *   wgt bf16 [cout][r][s][cin]     (OHWI = the channels-last memory of the reference's weight)
 *   y   bf16 [n][oh][ow][cout],  oh = (h + 2*pad - dil*(r-1) - 1)/stride + 1
 * Stride 2 is handled through TMA element strides (every other input pixel is fetched).  Feature maps smaller than
 * 128 pixels (pyramid-pooling branches: 1x1 .. 6x6) put several images into one 128-row M tile.
 * regda_conv_fprop_supported returns 1 when the shape is covered: cin, cout multiples of 64, stride 1 or 2. */
int regda_conv_fprop_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil);
/* tile-policy knob for sweeps (0 = default): minimum count of 128x256 tiles for which the 256-wide tile is chosen over 128 */
int regda_conv_tune(int min_tiles_256_value);
/* Hint for the next forward / data-gradient convolution launched by the calling thread: its WEIGHT tensor was last written long
 * before the kernel preceding the launch in the stream (the optimizer writes the bf16 weight copies once per step), so the kernel
 * may request its first weight tiles before its programmatic-dependent-launch wait.  Consumed by that launch. */
int regda_conv_hint_static_weights(void);
/* sweep knob: whether the BatchNorm kernels release their programmatic-dependent-launch successors at their start (1, default) or at exit (0) */
int regda_bn_tune(int early_trigger);
int regda_conv_fprop_bf16(const void *x, const void *wgt, void *y, int n, int h, int w, int cin, int cout,
                          int r, int s, int stride, int pad, int dil, void *stream);
/* same convolution, raw float32 accumulators out: y float32 [n][oh][ow][cout].  With the operands of a float32 convolution
 * split into bf16 (hi, hi, lo) x (hi, lo, hi) channel triples this is the float32 PARITY path (error ~2^-17 per product). */
int regda_conv_fprop_bf16_f32out(const void *x, const void *wgt, float *y, int n, int h, int w, int cin, int cout,
                                 int r, int s, int stride, int pad, int dil, void *stream);
/* same, and the epilogue also accumulates the train-mode BatchNorm statistics of y: bn_stats float32
 * [groups][2][cout] (per-group per-channel sum and sum of squares of the bf16 outputs; zeroed inside unless
 * stats_zeroed != 0, i.e. the caller hands out slices of a pool it zeroes once per step).  Needs every M tile inside one
 * statistics group (regda_conv_fprop_stats_supported; always true for maps of >= 128 pixels). */
int regda_conv_fprop_stats_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil, int groups);
int regda_conv_fprop_stats_bf16(const void *x, const void *wgt, void *y, int n, int h, int w, int cin, int cout,
                                int r, int s, int stride, int pad, int dil, float *bn_stats, int groups, int stats_zeroed, void *stream);

/* Data gradient of a stride-1 convolution, reading the forward OHWI weights in place (MN-major B operand):
 *   dy bf16 [n][oh][ow][cout], wgt bf16 [cout][r][s][cin] -> dx bf16 [n][h][w][cin]   (h, w, cin, cout, ... are the
 *   FORWARD convolution's geometry).  addend (may be NULL): bf16 [n][h][w][cin] added in the epilogue, dx = dgrad + addend --
 *   the gradient of a residual branch that also reads x (`out += identity`, regda/_resnets.py:108) joins here.
 * The data gradient of a stride-2 convolution is this call on regda_zero_insert2_bf16(dy) with stride = 1. */
int regda_conv_dgrad_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil);
int regda_conv_dgrad_bf16(const void *dy, const void *wgt, void *dx, int n, int h, int w, int cin, int cout,
                          int r, int s, int stride, int pad, int dil, const void *addend, void *stream);
/* float32 accumulators out (dx float32 [n][h][w][cin]; addend float32 or NULL): data gradient of the float32 parity path */
int regda_conv_dgrad_bf16_f32out(const void *dy, const void *wgt, float *dx, int n, int h, int w, int cin, int cout,
                                 int r, int s, int stride, int pad, int dil, const float *addend, void *stream);
/* Data gradient fused with the reductions of the BatchNorm(+ReLU) backward that consumes it (the BatchNorm whose OUTPUT is
 * this convolution's input; regda/_resnets.py:92-112 chains conv -> bn -> relu -> conv): dx receives
 * dz = (dgrad + addend) * [bn output > 0] (mask bits written by regda_bn_forward_bf16), red float32 [groups][2][cin]
 * ACCUMULATES sum(dz), sum(dz * bn_y).  Follow with regda_bn_backward_bf16(..., dz_ready = 1). */
int regda_conv_dgrad_bnred_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil, int groups);
int regda_conv_dgrad_bnred_bf16(const void *dy, const void *wgt, void *dx, int n, int h, int w, int cin, int cout,
                                int r, int s, int stride, int pad, int dil, const void *addend, const void *bn_y,
                                const void *relu_mask, float *red, int groups, void *stream);
/* Weight gradient (stride 1 or 2), ACCUMULATED into dw fp32 [cout][r][s][cin] with TMA reduce-stores:
 *   dy bf16 [n][oh][ow][cout], x bf16 [n][h][w][cin]. */
int regda_conv_wgrad_supported(int n, int h, int w, int cin, int cout, int r, int s, int stride, int pad, int dil);
int regda_conv_wgrad_bf16(const void *dy, const void *x, float *dw, int n, int h, int w, int cin, int cout,
                          int r, int s, int stride, int pad, int dil, void *stream);
/* dst bf16 [n][OH][OW][c] = zero-inserted src bf16 [n][oh][ow][c]: dst[n][2i][2j] = src[n][i][j], zero elsewhere
 * (the dY operand of a stride-2 convolution's data gradient: regda/_resnets.py:92-112, layer2.0 / layer3.0). */
int regda_zero_insert2_bf16(const void *src, void *dst, int n, int oh, int ow, int OH, int OW, int c, void *stream);

/* ---- folded PPM fuse convolution: layout glue (regda/models/Encoder.py:43-52; see csrc/ppm.cu) ------------------
 * The upsampled pyramid branches enter the 3x3 fuse convolution through two small GEMMs (run as 1x1 convolutions on the
 * tcgen05 kernels) instead of 2048 materialised channels; every GEMM uses its channel block of the OHWI weight in place (weight
 * channel stride of the convolution entry points), these kernels re-lay the intermediate G between the two GEMMs. */
int regda_ppm_g_pack(void *g0, void *g1, void *g2, void *g3, void *gt, int b, int O, int T, int kp, const int *scales_host,
                     int nscales, int pack, void *stream);
int regda_transpose_bf16(const void *src, void *dst, int batch, int rows, int cols, void *stream);
/* forward convolution + BatchNorm statistics (bn_stats may be NULL: none) with an epilogue addend bf16 [n][oh][ow][cout] */
int regda_conv_fprop_addend_bf16(const void *x, const void *wgt, int wct, void *y, int n, int h, int w, int cin, int cout,
                                 int r, int s, int stride, int pad, int dil, const void *addend, float *bn_stats, int groups,
                                 int stats_zeroed, void *stream);
/* `wct` >= cin: the weight has wct channels per tap ([cout][r][s][wct]) and the convolution runs over its FIRST cin channels, in
 * place -- forward above, data gradient and weight gradient (dw float32 [cout][r][s][wct], columns [0, cin) of every tap) here. */
int regda_conv_dgrad_wslice_bf16(const void *dy, const void *wgt, int wct, void *dx, int n, int h, int w, int cin, int cout,
                                 int r, int s, int stride, int pad, int dil, const void *addend, void *stream);
int regda_conv_wgrad_wslice_bf16(const void *dy, const void *x, float *dw, int wct, int n, int h, int w, int cin, int cout,
                                 int r, int s, int stride, int pad, int dil, void *stream);

/* ---- PPM head tail: Dropout2d + classifier (regda/models/Encoder.py:39-40) ----------------------------
 * regda_dropout2d_mask: keep_scale float32 [n] = 0 with probability p, else 1/(1-p), n = images * channels; the generator
 *   state (uint64 {seed, draw counter}) lives in DEVICE memory and is advanced by the kernel, so a replayed CUDA graph
 *   draws a fresh mask every step (torch's Dropout2d stream is not reproduced: parity tests run with p = 0).
 * regda_classifier_fwd: out float32 [b][ncls][hw] = bias + W (keep * y); y [b][hw][cin] bf16 (or float32 if y_is_f32),
 *   w float32 [ncls][cin], bias [ncls] or NULL, keep float32 [b][cin] or NULL.
 * regda_classifier_bwd: dout float32 [b][ncls][hw] -> dy (written, same type as y), dw / dbias (ACCUMULATED). */
int regda_dropout2d_mask(void *state_u64x2, double p, float *keep_scale, int n, void *stream);
int regda_classifier_fwd(const void *y, int y_is_f32, const float *w, const float *bias, const float *keep, float *out, int b,
                         int hw, int cin, int ncls, void *stream);
int regda_classifier_bwd(const void *y, int y_is_f32, const float *w, const float *keep, const float *dout, void *dy, float *dw,
                         float *dbias, int b, int hw, int cin, int ncls, void *stream);

/* ---- train-mode BatchNorm2d (+ residual add + ReLU) over channels-last bf16 -------------------
 * Replaces nn.BatchNorm2d / F.relu / `out += identity` in regda/_resnets.py:92-112 and
 * regda/models/Encoder.py:24-40.  y, residual (may be NULL), out: bf16 [npix][c];
 * gamma/beta/running_*: float32 [c]; num_batches_tracked: int64 [1] (may be NULL).
 *   out = relu?( gamma*(y-mean)*rstd + beta + residual ),  batch statistics in fp32 (biased variance for the
 *   normalisation, unbiased for running_var, as torch).
 * `groups` > 1 = that many independent statistics groups over equal contiguous pixel ranges: the source and target
 * batches of one training step run through every layer as ONE tensor while BatchNorm keeps the reference's
 * per-forward-call (per-domain) batch statistics; running statistics are updated once per group, in order.
 * No separate "finalize" launches: the apply kernels derive mean / rstd / scale / shift of their channels from the sums.
 * Backward: dout, out (needed when relu), y -> dy (and dres = masked dout when dres != NULL);
 * dgamma / dbeta float32 [c] are ACCUMULATED (+=). */
int regda_bn_supported(int64_t npix, int c);
/* stats float32 [groups][2][c]: per-group per-channel sum and sum of squares of y.  have_stats != 0: already produced by
 * regda_conv_fprop_stats_bf16 (the statistics pass over y is skipped); otherwise computed here into `stats` (zeroed first
 * unless stats_zeroed != 0).  Keep `stats` for the backward call. */
int regda_bn_forward_bf16(const void *y, const void *residual, void *out, int64_t npix, int c, int groups,
                          const float *gamma, const float *beta, float *running_mean, float *running_var,
                          int64_t *num_batches_tracked, double eps, double momentum, int relu,
                          float *stats, int have_stats, int stats_zeroed, void *relu_mask, void *stream);
/* red float32 [groups][2][c]: scratch for the two backward reductions (zeroed here unless red_zeroed != 0). */
/* relu_mask (forward, may be NULL; needs relu != 0): uint8 [npix*c/8], bit e%8 of byte e/8 = [out element e > 0], for
 * regda_conv_dgrad_bnred_bf16.  dz_ready != 0 (backward): dout is already dz and red already holds the two sums (both made
 * by regda_conv_dgrad_bnred_bf16): only the apply pass runs, out / dres are not used (the residual gradient IS dout).
 * ReLU mask: from `out` (out > 0) when it is given; with relu != 0, out == NULL and dres == NULL (no residual) the mask is
 * recomputed from y, gamma, beta and stats exactly as the forward computed the output, so `out` is not read at all. */
int regda_bn_backward_bf16(const void *dout, const void *out, const void *y, void *dy, void *dres, int64_t npix, int c,
                           int groups, const float *gamma, const float *beta, const float *stats, double eps, float *dgamma,
                           float *dbeta, int relu, float *red, int red_zeroed, int dz_ready, void *stream);

/* PrototypeContrastiveLoss (regda/loss.py:18-47; the alignment loss of the stage-2 step, tools/train_align_reg.py:186-189):
 * loss = mean over rows with label != ignore_label of CE(normalize(feat_row) . normalize(proto_c) / temperature, label).
 * feat float32 [n][k] (k % 4 == 0), label int64 [n], proto float32 [c][k] (c <= 8); stats float32 [2] receives (loss, valid
 * rows).  backward: dfeat float32 [n][k] = upstream[0] * dloss/dfeat (upstream = device scalar or NULL for 1), zeros for
 * ignored rows; pass the forward call's workspace and stats.  A label outside [0, c) that is not ignore_label sets
 * REGDA_FLAG_LABEL_RANGE (nn.CrossEntropyLoss raises there). */
size_t regda_pcl_workspace_bytes(int64_t n);
int regda_pcl_forward(const float *feat, const int64_t *label, const float *proto, float *stats, int64_t n, int k, int c,
                      int64_t ignore_label, double temperature, int32_t *flags, void *workspace, size_t workspace_bytes,
                      void *stream);
int regda_pcl_backward(const float *feat, const int64_t *label, const float *proto, const float *stats, const float *upstream,
                       float *dfeat, int64_t n, int k, int c, int64_t ignore_label, double temperature, void *workspace,
                       size_t workspace_bytes, void *stream);

/* Eval-mode BatchNorm2d (+ residual + ReLU) over channels-last bf16 with the running statistics (regda/_resnets.py:92-112 under
 * model.eval(): the offline teacher pass pseudo_generation.py:96-141 and evaluate() eval.py:14-56). */
int regda_bn_inference_bf16(const void *y, const void *residual, void *out, int64_t npix, int c, const float *gamma,
                            const float *beta, const float *running_mean, const float *running_var, double eps, int relu,
                            void *stream);
/* Eval tail of Deeplabv2.forward (regda/models/Encoder.py:152-155): out[b][c][H][W] = mean over the heads of
 * softmax_c(bilinear_align_corners(x_head [b][c][h][w])), float32; x2 may be NULL (one head). */
int regda_upsample_softmax_mean(const float *x1, const float *x2, float *out, int b, int c, int h, int w, int H, int W,
                                void *stream);

/* Patch matrix of the stem convolution (regda/_resnets.py:150: Conv2d(3, 64, 7, stride 2, padding 3)): x bf16 [n][h][w][3]
 * -> a bf16 [n][oh][ow][192], k = r*24 + s*3 + c (filter row r: 21 taps + 3 zeros; k >= 168 zero); the stem then runs on the tcgen05 kernels
 * as a 1x1 convolution over 192 channels (forward + weight gradient). */
int regda_stem_im2col_bf16(const void *x, void *a, int n, int h, int w, void *stream);
/* same patch matrix straight from the loader's float32 [n][3][h][w] image (rounded to bf16 while staged) */
int regda_stem_im2col_f32nchw(const float *x, void *a, int n, int h, int w, void *stream);

/* MaxPool2d(3, stride 2, padding 1) over channels-last bf16 (regda/_resnets.py:153): x [n][h][w][c] -> y [n][oh][ow][c],
 * oh = (h-1)/2+1; argmax_u8 (may be NULL for inference) [n][oh][ow][c] receives the position 0..8 of the first maximum
 * inside each window, which is all the backward needs. */
int regda_maxpool3s2_fwd_bf16(const void *x, void *y, void *argmax_u8, int n, int h, int w, int c, void *stream);
int regda_maxpool3s2_bwd_bf16(const void *argmax_u8, const void *dy, void *dx, int n, int h, int w, int c, void *stream);

/* ---- pyramid pooling front end of the PPM heads (regda/models/Encoder.py:43-52) -----------------
 * scales_host: HOST array of nscales (<= 4) pool sizes, e.g. {1,2,3,6}; ncell = sum s^2; cells of scale k start
 * at sum_{j<k} s_j^2, row-major.
 *   regda_ppm_pool_fwd : AdaptiveAvgPool2d(s) for every scale in one pass: feat bf16 [b][h][w][c] -> pooled f32 [b][ncell][c]
 *   regda_ppm_pool_bwd : dpooled f32 [b][ncell][c] -> dfeat bf16 [b][h][w][c]
 *   regda_ppm_upcat_fwd: cat bf16 [b][h][w][c + nscales*cb] = concat(feat, bilinear(align_corners=False) of
 *                        branch_k bf16 [b][s_k][s_k][cb])  (F.interpolate + torch.cat of Encoder.py:47-52)
 *   regda_ppm_upcat_bwd: dcat -> dbranch_k f32 [b][s_k][s_k][cb] (written); d feat is dcat[..., :c] itself. */
int regda_ppm_pool_fwd(const void *feat, float *pooled, int b, int h, int w, int c, const int *scales_host, int nscales,
                       void *stream);
int regda_ppm_pool_bwd(const float *dpooled, void *dfeat, int b, int h, int w, int c, const int *scales_host, int nscales,
                       void *stream);
/* dfeat = addend + pool_backward(dpooled); addend bf16 [b][h][w][c] or NULL: the gradient the feature map's other readers sent
 * (autograd's AccumulateGrad add after regda/models/Encoder.py:43-52's AdaptiveAvgPool2d backward, fused) */
int regda_ppm_pool_bwd_add(const float *dpooled, const void *addend, void *dfeat, int b, int h, int w, int c, const int *scales_host,
                           int nscales, void *stream);
/* the pooled cells as the branch convolutions read them (Encoder.py:45-47: ppm[k](AdaptiveAvgPool2d(s_k)(x)) takes the s_k x s_k map):
 * split != 0: pooled float32 [b][ncell][c] -> p_k bf16 [b][s_k][s_k][c]; split == 0: the gradient's way back (float32 result) */
int regda_ppm_cells(float *pooled, void *p0, void *p1, void *p2, void *p3, int b, int c, const int *scales_host, int nscales, int split,
                    void *stream);
int regda_ppm_upcat_fwd(const void *feat, const void *br0, const void *br1, const void *br2, const void *br3, void *cat,
                        int b, int h, int w, int c, int cb, const int *scales_host, int nscales, void *stream);
int regda_ppm_upcat_bwd(const void *dcat, float *dbr0, float *dbr1, float *dbr2, float *dbr3, int b, int h, int w, int c,
                        int cb, const int *scales_host, int nscales, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* REGDA_B200_H */
