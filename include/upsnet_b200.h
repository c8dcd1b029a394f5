/*
 * upsnet_b200.h -- C ABI of libupsnet_b200.so: hand-written sm_100a CUDA for the UPSNet
 * per-image inference hot path.  Plain pointers and sizes only; no
 * torch types.  Every entry point
 *   - takes DEVICE pointers unless the name ends in _host,
 *   - enqueues on `stream` (a cudaStream_t passed as void*) and does not synchronise,
 *   - never allocates (caller-owned outputs and workspaces; sizes from *_workspace_bytes),
 *   - is re-entrant (no global state) and returns 0 on success, a positive cudaError_t
 *     value on a CUDA failure, or a negative UPSNET_E_* code on an argument error.
 * Paths in "replaces:" comments are relative to /root/reference/upsnet/.
 */
#ifndef UPSNET_B200_H_
#define UPSNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UPSNET_E_BADARG (-1)
#define UPSNET_E_UNSUPPORTED (-2)
#define UPSNET_E_WORKSPACE (-3)

#define UPSNET_LAYOUT_NCHW 0
#define UPSNET_LAYOUT_NHWC 1
/* upsnet_roi_align*_forward with UPSNET_DTYPE_PAIR only: out = [R][2][PH*PW*C], the hi plane of the flattened (ph,pw,c)
 * roi feature followed by its lo plane -- i.e. a pair tensor of R 1x1 'images' with PH*PW*C channels (the RCNN fc6 input) */
#define UPSNET_LAYOUT_FLAT_PAIR 2

#define UPSNET_DTYPE_F32 0
#define UPSNET_DTYPE_BF16 1
/* hi/lo bf16 PAIR: an NHWC tensor with 2*C bf16 channels per pixel, [0,C) = bf16(v), [C,2C) = bf16(v - hi) -- the 16-bit
 * storage of precision UPSNET_PREC_BF16X3 (same bytes as fp32, ~16 mantissa bits); v = hi + lo is exact in fp32 */
#define UPSNET_DTYPE_PAIR 2

/* epilogue flags for the convolution entry points */
#define UPSNET_EPI_RELU 1
#define UPSNET_EPI_RES_UP2 2 /* upsnet_igemm_forward only: residual is [N,Ho/2,Wo/2,Cout], read with nearest 2x upsampling */
#define UPSNET_EPI_STEM_PAIR 8 /* upsnet_stem_forward only: y is a hi/lo pair tensor [N,Ho,Wo,2*Cout] (precision bf16x3) */
#define UPSNET_EPI_NO_TMA 4  /* upsnet_igemm_forward only: use the cp.async gather kernel even where the TMA-fed one qualifies */
/* upsnet_igemm_forward, y_dtype PAIR only: store the output channels as [hi G][lo G] per group of G channels instead of
 * [hi Cout][lo Cout] (G % 64 == 0, Cout % G == 0; 0 = Cout).  Lets a 1x1 conv that emulates a 2x2 deconvolution write
 * its four (a,b) sub-pixel groups as four pair pixels (models/rcnn.py:62 mask_deconv1). */
#define UPSNET_EPI_PAIR_GROUP(G) ((((G) / 64) & 0xfff) << 8)
/* upsnet_igemm_forward, fp32 plane-wise (NCHW) or small-Cout outputs only: output channels >= c get a logistic sigmoid
 * 1 / (1 + expf(-v)) after the bias (models/rpn.py:55 cls_prob = sigmoid(cls_score): the RPN head writes the logits and,
 * from duplicated weight rows, their probabilities in one launch). */
#define UPSNET_EPI_SIGMOID_FROM(c) ((((c) + 1) & 0x3ff) << 20)

/* precision of the tensor-core convolution path */
#define UPSNET_PREC_FP32_SIMT 0 /* fp32 FFMA tiles (exact-order-free fp32)            */
#define UPSNET_PREC_BF16X3 1    /* tcgen05 kind::f16, 3-term bf16 split (~fp32 result) */
#define UPSNET_PREC_BF16 2      /* tcgen05 kind::f16, single bf16 pass                 */

/* library / build identification: returns e.g. 100 for sm_100a, fills `n_sm` if non-NULL */
int upsnet_version(int *n_sm);

/* ---------------------------------------------------------------------------------------
 * ROIAlign forward (sampling grid sr x sr, no half-pixel shift).
 * replaces: operators/src/roi_align_cuda.cpp:39-75 roi_align_forward_cuda
 *           -> operators/src/roi_align_kernel.cu:351 roi_align_forward_gpu_kernel_launcher
 * feat [B,C,H,W] (NCHW) or [B,H,W,C] (NHWC) fp32; rois [R,5] = (batch,x1,y1,x2,y2);
 * out [R,C,PH,PW] (NCHW) or [R,PH,PW,C] (NHWC) -- same layout flag as feat.
 * dtype: UPSNET_DTYPE_F32, or UPSNET_DTYPE_BF16 (NHWC only: bf16 features in, bf16 out, fp32 accumulation), or
 * UPSNET_DTYPE_PAIR (NHWC: hi/lo pair features [B,H,W,2C] in, pair out [R,PH,PW,2C] or UPSNET_LAYOUT_FLAT_PAIR;
 * sampling_ratio > 0, C % 8 == 0).
 */
int upsnet_roi_align_forward(const void *feat, int B, int C, int H, int W, int layout, int dtype,
                             const float *rois, int R, int PH, int PW, int sampling_ratio,
                             float spatial_scale, void *out, void *stream);

/* FPN ROIAlign: level assignment + 4 pyramid levels + un-permute in ONE launch.
 * replaces: operators/modules/fpn_roi_align.py:32-62 FPNRoIAlign.forward (host bucketing,
 *           4 launches, cat, index_select).  feats[l] has spatial size (Hs[l],Ws[l]) and
 *           scale scales[l]; level(roi) = clip(floor(2+log2(sqrt(w*h)/224+1e-6)),0,3).
 * levels_out (optional, may be NULL): int32 [R] chosen level per roi. */
int upsnet_roi_align_fpn_forward(const void *const feats[4], const int Hs[4], const int Ws[4],
                                 const float scales[4], int B, int C, int layout, int dtype,
                                 const float *rois, int R, int PH, int PW, int sampling_ratio,
                                 void *out, int *levels_out, void *stream);

/* ---------------------------------------------------------------------------------------
 * NMS (IoU with the legacy +1 box area, suppress when IoU > thresh).
 * replaces: nms/gpu_nms.hpp:14 _nms  (nms/nms_kernel.cu:40-84 nms_kernel + :97-150 host sweep)
 *
 * Device-resident, segmented: S independent problems in one launch pair.  boxes [total,4]
 * fp32 (x1,y1,x2,y2), already sorted by descending score inside each segment;
 * seg_offsets int32 [S+1] on the DEVICE; max_seg_len = host-known upper bound of a segment
 * length (<= 65536).  keep_out int32 [S, max_seg_len]: positions (relative to the segment
 * start) of kept boxes in ascending order; keep_cnt int32 [S].  Nothing returns to the host.
 */
int upsnet_nms_workspace_bytes(int S, int max_seg_len, size_t *bytes);
int upsnet_nms_segmented(const float *boxes, const int *seg_offsets, int S, int max_seg_len,
                         float thresh, int *keep_out, int *keep_cnt, void *workspace,
                         size_t workspace_bytes, void *stream);
/* Drop-in for the reference's host-pointer entry (same arguments as _nms): boxes_host
 * [n,boxes_dim>=4] sorted by score desc; keep_out/num_out on the host; synchronises. */
int upsnet_nms_host(int *keep_out, int *num_out, const float *boxes_host, int boxes_num,
                    int boxes_dim, float thresh, int device_id);

/* ---------------------------------------------------------------------------------------
 * Deformable convolution v1 / v2 forward, fused (no column buffer in HBM).
 * replaces: operators/functions/deform_conv.py:26-57 DeformConvFunction.forward
 *           (= deform_conv_cuda.deform_im2col, operators/src/deform_conv_cuda.cpp:49-68,
 *              kernel operators/src/deform_conv_kernel.cu:194-242, + torch.mm + bias)
 *           operators/functions/mod_deform_conv.py:25-59 for mask != NULL
 *           (kernel operators/src/mod_deform_conv_kernel.cu:187-249).
 * x [N,Cin,H,W]; offset [N,dg*2*kh*kw,Ho,Wo] (pairs (dh,dw) per tap); mask [N,dg*kh*kw,Ho,Wo]
 * or NULL (already activated: 2*sigmoid); weight [Cout,Cin,kh,kw]; bias [Cout] or NULL;
 * y [N,Cout,Ho,Wo].  All fp32 NCHW contiguous.  groups must be 1 (the reference ignores it).
 */
int upsnet_dcn_forward(const float *x, const float *offset, const float *mask,
                       const float *weight, const float *bias, float *y, int N, int Cin, int H,
                       int W, int Cout, int kh, int kw, int stride_h, int stride_w, int pad_h,
                       int pad_w, int dil_h, int dil_w, int deformable_groups, int epi_flags,
                       int precision, void *stream);

/* Dense convolution forward with fused bias / residual-add / ReLU epilogue (frozen BN is
 * folded into weight+bias by the caller).
 * replaces: the cuDNN conv + BN + ReLU + add chains of models/resnet.py:80-100,
 *           models/fpn.py:78-104, models/rpn.py:52-56, models/rcnn.py:79-87,132-141.
 * x [N,Cin,H,W]; weight [Cout,Cin,kh,kw]; bias/residual may be NULL; residual, y [N,Cout,Ho,Wo]. */
int upsnet_conv2d_forward(const float *x, const float *weight, const float *bias,
                          const float *residual, float *y, int N, int Cin, int H, int W,
                          int Cout, int kh, int kw, int stride_h, int stride_w, int pad_h,
                          int pad_w, int dil_h, int dil_w, int epi_flags, int precision,
                          void *stream);

/* ---------------------------------------------------------------------------------------
 * tcgen05 implicit-GEMM convolution / deformable convolution (engine entry point).
 * Same arithmetic contract as upsnet_conv2d_forward / upsnet_dcn_forward, but
 *   - x is NHWC [N,H,W,Cin] (Cin % 64 == 0) stored as fp32 or bf16 (x_dtype) -- or, for a tiny Cin <= 8 (the
 *     RGB stem), the fp32 NCHW image itself: K = kh*kw*Cin is flattened and zero-padded to a multiple of 64; y and residual are NHWC or
 *     NCHW (out_layout) stored as fp32 or bf16 (y_dtype); bf16 activations are copied by TMA / cp.async straight
 *     into the tensor-core layout (UPSNET_PREC_BF16); UPSNET_PREC_BF16X3 takes fp32 activations (split in the
 *     gather) or UPSNET_DTYPE_PAIR activations (x [N,H,W,2*Cin], y / residual [N,Ho,Wo,2*Cout]: TMA-fed, three MMAs
 *     per k-slice, fp32 result re-split in the epilogue),
 *   - weights are pre-packed once with upsnet_igemm_pack_weight (bf16 hi/lo planes,
 *     [Cout_pad][kh*kw][Cin]); `packed` must hold upsnet_igemm_packed_weight_bytes bytes,
 *   - offset [N,2*kh*kw,Ho,Wo] / mask [N,kh*kw,Ho,Wo] stay NCHW (reference layout), NULL for a
 *     dense convolution; deformable_groups must be 1,
 *   - precision is UPSNET_PREC_BF16X3 (fp32-grade result) or UPSNET_PREC_BF16.
 * replaces: the same reference call sites as upsnet_conv2d_forward / upsnet_dcn_forward.
 */
int upsnet_igemm_packed_weight_bytes(int Cout, int Cin, int kh, int kw, size_t *bytes);
int upsnet_igemm_pack_weight(const float *weight, int Cout, int Cin, int kh, int kw, void *packed,
                             void *stream);
int upsnet_igemm_forward(const void *x_nhwc, const float *offset, const float *mask,
                         const void *packed, const float *bias, const void *residual, void *y,
                         int N, int H, int W, int Cin, int Cout, int kh, int kw, int stride_h,
                         int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int out_layout,
                         int x_dtype, int y_dtype, int epi_flags, int precision, void *stream);

/* ---------------------------------------------------------------------------------------
 * Deformable convolution v1 / v2 on hi/lo PAIR activations with the bilinear corners gathered from a shared-memory
 * window (csrc/dcn_win.cu): 3x3, stride 1, deformable_groups 1, Cin % 64 == 0, Cout % 16 == 0, precision
 * UPSNET_PREC_BF16X3.  x [N,H,W,2*Cin] pair NHWC -> y [N,Ho,Wo,2*Cout] pair NHWC; offset / mask as in
 * upsnet_igemm_forward; epi_flags: UPSNET_EPI_RELU.  Same arithmetic contract as upsnet_igemm_forward with an offset
 * (samples outside the staged window are gathered from global memory: the result does not depend on the window size).
 * `packed` comes from upsnet_dcn_pack_weight (K order: 16-channel sub-chunk, tap, channel) and holds
 * upsnet_dcn_packed_weight_bytes bytes.  All three return UPSNET_E_UNSUPPORTED for other layer shapes: callers then use
 * upsnet_igemm_forward.
 * replaces: operators/functions/deform_conv.py:26-57 (deformable_im2col + torch.mm),
 *           operators/src/deform_conv_kernel.cu:89-118,194-242, mod_deform_conv_kernel.cu (v2 mask). */
int upsnet_dcn_packed_weight_bytes(int Cout, int Cin, int kh, int kw, size_t *bytes);
int upsnet_dcn_pack_weight(const float *weight, int Cout, int Cin, int kh, int kw, void *packed, void *stream);
int upsnet_dcn_pair_forward(const void *x_pair, const float *offset, const float *mask, const void *packed,
                            const float *bias, void *y_pair, int N, int H, int W, int Cin, int Cout, int kh, int kw,
                            int pad_h, int pad_w, int dil_h, int dil_w, int epi_flags, void *stream);
/* Dense 3x3 / stride-1 convolution on hi/lo PAIR activations through the same window pipeline (csrc/dcn_win.cu, DENSE mode):
 * the input window of a 16x8-pixel tile is staged once per 16-channel sub-chunk by TMA and feeds all nine taps, the A operand
 * is copied window -> TMEM.  Meant for the small-N layers (18-channel offset convs of the semantic head, 64->64 bottleneck
 * convs) whose per-tap TMA boxes make upsnet_igemm_forward L2->SM-bandwidth-bound.  x [N,H,W,2*Cin] pair NHWC; `packed`
 * from upsnet_dcn_pack_weight; y = fp32 NCHW [N,Cout,Ho,Wo] (UPSNET_LAYOUT_NCHW, any Cout) or pair NHWC [N,Ho,Wo,2*Cout]
 * (UPSNET_LAYOUT_NHWC, Cout % 16 == 0); epi_flags: UPSNET_EPI_RELU.  Same arithmetic contract as upsnet_igemm_forward
 * (precision UPSNET_PREC_BF16X3).  UPSNET_E_UNSUPPORTED for other shapes (dilation > 7, Cin % 64 != 0).
 * replaces: the cuDNN 3x3 convs of models/fcn.py:40-55 (conv_offset) and models/resnet.py:80-100 (conv2). */
int upsnet_conv3x3_pair_forward(const void *x_pair, const void *packed, const float *bias, void *y, int N, int H, int W,
                                int Cin, int Cout, int pad_h, int pad_w, int dil_h, int dil_w, int out_layout,
                                int epi_flags, void *stream);

/* ---------------------------------------------------------------------------------------
 * Parameter-free panoptic head, fused: MaskRemoval + SegTerm + void/concat/argmax.
 * replaces: models/resnet_upsnet.py:223-240 with operators/modules/mask_removal.py:29-93,
 *           operators/modules/unary_logits.py:78-105.  Never materialises the [1,k,H,W] planes.
 * fcn [S,H,W] fp32 (fcn_output of one image); boxes [n,4] (mask_rois[:,1:]); cls_prob [n];
 * mask_logit [n,28,28] (logit of the predicted class); cls_idx int64 [n] (1-based thing class,
 * <= num_thing); num_stuff = S - num_thing.
 * n is the (maximum) instance count known to the host; n_dev (optional, may be NULL) is a DEVICE int32 with
 * the actual count <= n, so the call can be enqueued without knowing it (static-shape engine / CUDA graphs).
 * keep_out int64 [max(n,1)] original indices of kept instances in score order, k_out int32[1];
 * labels int64 [H,W] (255 = void); sem_labels int64 [H,W] or NULL (argmax_c fcn).
 * Workspace: upsnet_panoptic_workspace_bytes is the preferred size (the 1-bit mask windows of all n instances resident,
 * capped at 64 MB); any size >= upsnet_panoptic_workspace_min_bytes is accepted -- the windows are then built and
 * consumed in rounds of consecutive score ranks (at most 64), with identical results.
 */
int upsnet_panoptic_workspace_bytes(int n, int H, int W, int num_thing, size_t *bytes);
int upsnet_panoptic_workspace_min_bytes(int n, int H, int W, int num_thing, size_t *bytes);
int upsnet_panoptic_head(const float *fcn, int S, int H, int W, const float *boxes,
                         const float *cls_prob, const float *mask_logit, const int64_t *cls_idx,
                         int n, const int *n_dev, int num_stuff, double fraction_threshold, int64_t *keep_out,
                         int *k_out, int64_t *labels, int64_t *sem_labels, void *workspace,
                         size_t workspace_bytes, void *stream);
/* Same head on the QUARTER-resolution score map: score [S,Hs,Ws] (models/fcn.py:94-101 `score` before the final
 * nn.Upsample(scale_factor=4, mode='bilinear')), labels / sem_labels [4*Hs,4*Ws].  The up-sampling is evaluated inside the
 * fusion kernel with the arithmetic of upsnet_upsample_bilinear_nchw (factor 4), so the results are bit-identical to
 * upsnet_panoptic_head on the materialised [S,4*Hs,4*Ws] logits, which are neither written nor read (159 MB each way at
 * 19 x 1024 x 2048).  Workspace sizes: upsnet_panoptic_workspace_bytes(n, 4*Hs, 4*Ws, num_thing). */
int upsnet_panoptic_head_up4(const float *score, int S, int Hs, int Ws, const float *boxes,
                             const float *cls_prob, const float *mask_logit, const int64_t *cls_idx,
                             int n, const int *n_dev, int num_stuff, double fraction_threshold, int64_t *keep_out,
                             int *k_out, int64_t *labels, int64_t *sem_labels, void *workspace,
                             size_t workspace_bytes, void *stream);

/* MaskRemoval alone (API parity with operators/modules/mask_removal.py:29-93): score-ordered overlap
 * pruning; keep_out / k_out as above; mask_energy (optional, may be NULL) float [n,H,W]: planes 0..k-1
 * receive the pasted, resized logits of the kept instances (zeros elsewhere), as the reference returns
 * them.  Workspace size: upsnet_panoptic_workspace_bytes. */
int upsnet_mask_removal(const float *boxes, const float *cls_prob, const float *mask_logit,
                        const int64_t *cls_idx, int n, const int *n_dev, int H, int W, int num_thing,
                        double fraction_threshold, int64_t *keep_out, int *k_out, float *mask_energy,
                        void *workspace, size_t workspace_bytes, void *stream);

/* RGB stem on the TMA kernel: k x k (kw <= 8) / stride 2 / pad `pad` convolution of a tiny-Cin (<= 8) fp32 NCHW
 * image, bf16 NHWC output [N,Ho,Wo,Cout] (Cout % 64 == 0) -- or, with UPSNET_EPI_STEM_PAIR, the hi/lo pair tensor
 * [N,Ho,Wo,2*Cout] computed with the three-pass split from hi/lo copies of the image -- fused bias + ReLU (UPSNET_EPI_RELU).
 * replaces: models/resnet.py:155-162 conv1 + bn1 (folded) + relu.
 * The call first packs the image to a zero-padded bf16 NHWC8 copy in `workspace` (upsnet_stem_workspace_bytes),
 * then runs the tcgen05 kernel whose A tiles are boxes of a 5-D tensor map over that copy; weights are packed
 * once with upsnet_stem_pack_weight ([Cout][kh][8][8] bf16, upsnet_stem_packed_weight_bytes).
 * Returns UPSNET_E_UNSUPPORTED if the driver rejects the tensor map (callers fall back to upsnet_igemm_forward). */
int upsnet_stem_workspace_bytes(int N, int H, int W, int kh, int kw, int pad, size_t *bytes);
int upsnet_stem_packed_weight_bytes(int Cout, int kh, size_t *bytes);
int upsnet_stem_pack_weight(const float *weight, int Cout, int Cin, int kh, int kw, void *packed,
                            void *stream);
int upsnet_stem_forward(const float *x, const void *packed_w, const float *bias, void *y, int N, int Cin,
                        int H, int W, int Cout, int kh, int kw, int pad, int epi_flags, void *workspace,
                        size_t workspace_bytes, void *stream);

/* Max-pooling on NHWC activations (bf16, hi/lo pair or fp32 storage; C % 8 == 0 resp. C % 4 == 0), floor output size.
 * UPSNET_DTYPE_PAIR: x [N,H,W,2C] -> y [N,Ho,Wo,2C], the window element with the largest hi + lo is copied.
 * replaces: models/resnet.py:163 nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of the stem.
 * x [N,H,W,C] -> y [N,Ho,Wo,C], Ho = (H + 2*pad - k)/stride + 1; padding never wins the max. */
int upsnet_maxpool2d_nhwc(const void *x, void *y, int N, int H, int W, int C, int k, int stride,
                          int pad, int dtype, void *stream);

/* Bilinear up-sampling by an integer factor, NCHW fp32 planes, align_corners = False.
 * replaces: models/fcn.py:88-101 nn.Upsample(scale_factor=4, mode='bilinear') of the semantic logits
 *           (source index (dst + 0.5)/factor - 0.5 clamped at 0, as ATen's upsample_bilinear2d).
 * x [planes,H,W] -> y [planes,H*factor,W*factor]; (W*factor) % 4 == 0. */
int upsnet_upsample_bilinear_nchw(const float *x, float *y, int planes, int H, int W, int factor,
                                  void *stream);

/* Semantic-head score assembly: out = s2 + up2(s3) + up4(s4) + up8(s5), bilinear, align_corners = False.
 * replaces: models/fcn.py:94-101 (three F.interpolate of 128-channel maps + cat + score conv; the engine scores every level
 *           at its own resolution first -- the 1x1 conv and the up-sampling commute -- and sums the 19-plane maps here).
 * s2 [planes,H,W], s3 [planes,H/2,W/2], s4 [planes,H/4,W/4], s5 [planes,H/8,W/8] fp32; H % 8 == W % 8 == 0. */
int upsnet_fcn_score_fuse(const float *s2, const float *s3, const float *s4, const float *s5, float *out,
                          int planes, int H, int W, void *stream);

/* ---------------------------------------------------------------------------------------
 * Detection glue, fused (device-resident; nothing returns to the host).
 *
 * upsnet_rpn_decode: anchors + deltas -> clipped proposal boxes for the pre-NMS top-k of every pyramid level.
 * replaces: operators/functions/pyramid_proposal.py:83-131 (shifted anchors, bbox_transform, clip_boxes;
 *           bbox/bbox_transform.py:290-330,45-60) for the selected indices.
 * deltas[l] fp32 [4A,h_l,w_l] (RPN head output, channel a*4+c); top_idx[l] int64 [k_l] flat (y,x,a) indices
 * (device); k/hs/ws/strides host arrays [L]; base_anchors float64 [L,A,4] on the device (generate_anchors);
 * boxes_out fp32 [sum k_l, 4] in level order. */
int upsnet_rpn_decode(const float *const *deltas, const long long *const *top_idx, const int *k,
                      const int *hs, const int *ws, const int *strides, const double *base_anchors,
                      int L, int A, float im_h, float im_w, float *boxes_out, void *stream);

/* upsnet_rpn_topk: the pre_nms_top_n best anchors of every pyramid level, sorted by descending score, one call.
 * replaces: operators/functions/pyramid_proposal.py:104-118 (per-level argsort(-scores)[:pre_nms_top_n]).
 * probs[l] fp32 [A,h_l,w_l] (device); k_l = min(pre_nms_top_n, A*h_l*w_l) <= 2048; out_scores fp32 / out_idx int64
 * [sum k_l]: level after level, flat (y,x,a) indices as the reference's transposed score vector uses; equal
 * scores are ordered by ascending index.  Workspace: upsnet_rpn_topk_workspace_bytes(L). */
int upsnet_rpn_topk_workspace_bytes(int L, size_t *bytes);
int upsnet_rpn_topk(const float *const *probs, const int *hs, const int *ws, int L, int A,
                    int pre_nms_top_n, float *out_scores, long long *out_idx, void *workspace,
                    size_t workspace_bytes, void *stream);

/* upsnet_rpn_collect: per level the first min(keep_cnt, post_nms_top_n) NMS survivors, then the post_nms_top_n
 * best of their union by score (descending), as fixed-size outputs.
 * replaces: operators/functions/pyramid_proposal.py:196-222 + modules/pyramid_proposal.py:61-67.
 * keep/keep_cnt/seg_offsets from upsnet_nms_segmented over (boxes, scores) [total]; rois fp32 [post,5] =
 * (0,x1,y1,x2,y2), rows past the live count are zero; out_scores [post]; valid uint8 [post]. */
int upsnet_rpn_collect(const int *keep, const int *keep_cnt, const int *seg_offsets, const float *boxes,
                       const float *scores, int S, int max_seg_len, int post_nms_top_n, float *rois,
                       float *out_scores, unsigned char *valid, void *stream);

/* upsnet_maskroi_prepare: candidate selection (prob > score_thresh, roi valid), ordering (class segment
 * ascending -- one segment when class_agnostic --, score descending, roi-major index ascending) and box decode
 * (weights, clip) in one launch.  replaces: operators/modules/mask_roi.py:36-95 up to the per-class NMS.
 * rois [R,5]; roi_valid uint8 [R]; bbox_delta [R,4C]; cls_prob [R,C]; R*(C-1) <= 8192.
 * sc_out/cls_out/bx_out [R*(C-1)] (/[.,4]): candidates first, in NMS input order (others: score -1);
 * offs_out int32 [nseg+1]: segment offsets for upsnet_nms_segmented. */
int upsnet_maskroi_prepare(const float *rois, const unsigned char *roi_valid, const float *bbox_delta,
                           const float *cls_prob, int R, int C, int class_agnostic, float score_thresh,
                           const float weights[4], float im_h, float im_w, float *sc_out, int *cls_out,
                           float *bx_out, int *offs_out, void *stream);

/* upsnet_maskroi_finish: NMS survivors (class-major) -> keep scores >= the top_n-th largest -> `cap` output
 * slots (score, (0,x1,y1,x2,y2), class) + device count; no survivor -> one dummy detection (score 1, zero box,
 * class 0).  replaces: operators/modules/mask_roi.py:96-139.  keep/keep_cnt/seg_offsets as produced by
 * upsnet_nms_segmented on bx; nseg <= 128.  n_out int32 [2]: [0] = number of detections, [1] = truncation flags (the
 * reference keeps every survivor and every box tied at the top-n threshold): bit 0 = more NMS survivors than the 4096
 * candidate slots, bit 1 = more boxes at / above the threshold than `cap` output slots. */
int upsnet_maskroi_finish(const int *keep, const int *keep_cnt, const int *seg_offsets, const float *sc,
                          const int *cls, const float *bx, int nseg, int max_seg_len, int top_n, int cap,
                          float *out_sc, float *out_bx, long long *out_cls, int *n_out, void *stream);

/* ---------------------------------------------------------------------------------------
 * Callers either side of the per-image forward, device resident.
 *
 * upsnet_unified_pan_result: the 2-channel panoptic result the PQ evaluation consumes.
 * replaces: dataset/base_dataset.py:332-371 get_unified_pan_result (numpy on the host: np.unique per segment).
 * seg int64 [H,W] (semantic argmax, 'fcn_outputs'), pan int64 [H,W] ('panoptic_outputs': 0..id_last_stuff stuff,
 * id_last_stuff + 1 + j = j-th kept instance, 255 void; id_last_stuff = num_seg_classes - num_classes), cls_inds int64
 * [k] ('panoptic_cls_inds', 1-based thing class of kept instance j); k_dev optional DEVICE count <= k.
 * pan_2ch uint8 [H,W,3]: channel 0 = semantic class (255 void), 1 = instance number (rank among the instance ids present,
 * from 1; 0 = stuff), 2 = 0.  A segment takes its semantic majority class when that is a stuff class holding >= half of
 * it; stuff classes smaller than stuff_area_limit pixels become void.  err_out (optional DEVICE int): bit 0 = label out of
 * range, bit 1 = an instance id without an entry in cls_inds (the reference raises IndexError).
 */
int upsnet_unified_pan_workspace_bytes(int num_seg_classes, size_t *bytes);
int upsnet_unified_pan_result(const long long *seg, const long long *pan, const long long *cls_inds, int k,
                              const int *k_dev, int H, int W, int num_seg_classes, int num_classes,
                              int stuff_area_limit, unsigned char *pan_2ch, int *err_out, void *workspace,
                              size_t workspace_bytes, void *stream);

/* upsnet_prep_image: raw uint8 HWC (BGR) image -> the network input blob.
 * replaces: dataset/base_dataset.py:143-174 prep_im_for_blob (mean subtraction, cv2.resize INTER_LINEAR) and :898-923
 *           im_list_to_blob (zero padding to a multiple of the FPN stride, HWC -> CHW), plus the 4x larger fp32 H2D copy.
 * image_hwc uint8 [h,w,3] on the DEVICE; scale = the fx = fy factor handed to cv2.resize (the source step is 1/scale,
 * as OpenCV does when factors are given); (out_h,out_w) = the resized size (cvRound(h*scale), cvRound(w*scale));
 * (pad_h,pad_w) >= it; blob fp32 [3,pad_h,pad_w]: resized (image - pixel_means), zeros in the padding.  The means are
 * float64 like config.network.pixel_means: numpy subtracts in double and stores float32 (base_dataset.py:154).
 */
int upsnet_prep_image(const unsigned char *image_hwc, int h, int w, double scale, int out_h, int out_w, int pad_h,
                      int pad_w, const double pixel_means[3], float *blob, void *stream);

/* ---------------------------------------------------------------------------------------
 * Instance-mask post-processing of the test loop on the device.
 * replaces: upsnet_end2end_test.py:95-152 `im_post` (expand_boxes bbox/bbox_transform.py:365-381 -> int32 boxes,
 *           (M+2)x(M+2) zero-padded mask -> cv2.resize -> > 0.5 -> paste into [H,W] -> pycocotools.mask.encode), run by the
 *           reference on the host with numpy + cv2 + pycocotools for every detection.
 * mask_probs [n,C,M,M] fp32 (M <= 28; the plane of cls_inds[d] is used when C > 1, plane 0 otherwise), boxes [n,4]
 * (x1,y1,x2,y2 = pred_boxes[:,1:]), cls_inds int64 [n]; n_dev (optional device count <= n); image H <= 2048, W <= 2048.
 * counts [n][cap] uint32: the UNCOMPRESSED COCO run lengths of detection d (column-major, starting with the zeros run,
 * exactly maskApi.c rleEncode); run_len [n] their number (0 for d >= *n_dev); *overflow = 1 if some detection needs more
 * than cap counts (run_len then holds the needed size).  The compressed `counts` string of the COCO dict is a pure
 * function of these numbers (rleToString), applied on the host by upsnet_b200.operators.im_post.
 * Workspace: upsnet_im_post_workspace_bytes(n, cap). */
int upsnet_im_post_workspace_bytes(int n, int cap, size_t *bytes);
int upsnet_im_post_rle(const float *mask_probs, int C, int M, const float *boxes, const int64_t *cls_inds, int n,
                       const int *n_dev, int H, int W, uint32_t *counts, int cap, int *run_len, int *overflow,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * Backward kernels of the custom operators (training configuration, BASELINE config #4).  fp32 NCHW, ONE image per call
 * for the deformable kernels (the reference loops over the batch: functions/deform_conv.py:84-104).
 *
 * upsnet_dcn_im2col:       col [Cin*kh*kw, Ho*Wo] = zero-padded bilinear samples (* mask)      -- d(weight) = dY * col^T
 *   replaces: operators/src/deform_conv_kernel.cu:194-242 (K1), mod_deform_conv_kernel.cu:187-249 (K4)
 * upsnet_dcn_col2im:       dx [Cin,H,W] (zeroed by the call) += bilinear weights * dcol (* mask)
 *   replaces: deform_conv_kernel.cu:293-343 (K2), mod_deform_conv_kernel.cu:251-308 (K5)
 * upsnet_dcn_col2im_coord: doffset [2*kh*kw, Ho*Wo] and, with a mask, dmask [kh*kw, Ho*Wo]
 *   replaces: deform_conv_kernel.cu:391-449 (K3), mod_deform_conv_kernel.cu:310-381 (K6)
 * x [Cin,H,W]; offset [2*kh*kw,Ho,Wo]; mask [kh*kw,Ho,Wo] (already activated) or NULL; deformable_groups = 1.
 */
int upsnet_dcn_im2col(const float *x, const float *offset, const float *mask, int Cin, int H, int W, int kh, int kw,
                      int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, float *col, void *stream);
int upsnet_dcn_col2im(const float *dcol, const float *offset, const float *mask, int Cin, int H, int W, int kh, int kw,
                      int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, float *dx, void *stream);
int upsnet_dcn_col2im_coord(const float *dcol, const float *x, const float *offset, const float *mask, int Cin, int H,
                            int W, int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                            float *doffset, float *dmask, void *stream);

/* upsnet_roi_align_backward: dfeat [B,C,H,W] (zeroed by the call) += scatter of dout [R,C,PH,PW] to the bilinear taps.
 * replaces: operators/src/roi_align_kernel.cu:238-348 RoIAlignBackwardFeature (K8). */
int upsnet_roi_align_backward(const float *dout, const float *rois, int R, int B, int C, int H, int W, int PH, int PW,
                              int sampling_ratio, float spatial_scale, float *dfeat, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* UPSNET_B200_H_ */
