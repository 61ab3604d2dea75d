/* loft_b200.h -- C ABI of the B200-native LOFT/FOA training hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t, never
 * allocates, never synchronises the device, and returns 0 or a negative error code
 * (LOFT_ERR_ARG -1 bad argument, LOFT_ERR_SHAPE -2 unsupported shape, LOFT_ERR_CUDA -3 CUDA error;
 * loft_last_error() holds the message).  All tensors are fp32, activations are NHWC
 * (= torch channels_last), conv weights are [Cout][kh][kw][Cin] (= torch OIHW in channels_last).
 *
 * Each group cites the reference call site it replaces (paths under jwwangchn/BONAI @ aeafa46;
 * "mmcv" = the un-vendored mmcv-full==1.0.5 dependency pinned at mmdet/__init__.py:18-26).
 */
#ifndef LOFT_B200_H_
#define LOFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

const char* loft_last_error(void);
int loft_abi_version(void);
/* Leave `n` SMs out of every persistent grid launched from now on (0 = use all): lets a NCCL
 * all-reduce run beside the trunk backward (mmdet/apis/train.py:75-79, DDP overlap).  Returns the
 * previous setting.  Grid sizes are baked into captured CUDA graphs at capture time. */
int loft_reserve_sms(int n);
/* host -> device transfer of the gt-box windows (box grown by `pad` pixels, clipped) of a pinned
 * uint8 bitmap stack [G, H, W] into the same positions of a zeroed dense device stack; boxes_dev is
 * the device copy of the [G, 4] gt boxes.  The input transfer of BitmapMasks
 * (mmdet/core/mask/structures.py:20-60) for masks that vanish outside their boxes */
int loft_h2d_mask_windows(const unsigned char* src_host, const float* boxes_dev,
                          unsigned char* dst_dev, int G, int H, int W, int pad,
                          cudaStream_t stream);

/* Fused epilogue of the dense kernels.  For pixel p, channel c:
 *   acc = sum_k ...;  if (raw_out) raw_out[p,c] = acc;           (pre-BN conv output)
 *   v = acc*scale[c] + shift[c]                                   (BN-eval affine / bias)
 *   v += residual[p,c]  (or residual[up2x(p),c] if res_upsample2x: FPN top-down, fpn.py:185-186)
 *      (res_upsample2x == 2: zero-stuffed -- residual[p/2,c] only where h and w are even: the
 *       backward of a stride-2 1x1 conv on the same input, resnet.py:151-156 downsample path)
 *   if (relu) v = max(v,0);  if (mask) v = mask[p,c] > 0 ? v : 0  (ReLU backward)
 *   out[p,c] = v   (deconv_shuffle: c=(i*2+j)*Co+co scatters to pixel (2h+i,2w+j), channel co)
 *   colsum[g*colsum_gstride + c] += sum_p v  (and colsum2[c]): in a dgrad launch this is the
 *      bias / BN-beta gradient of the layer that produced the dgrad's input, for free
 */
typedef struct loft_epilogue_t {
  float* raw_out;
  const float* scale;
  const float* shift;
  const float* residual;
  const float* mask;
  long long ldr;
  int res_upsample2x;
  int relu;
  int deconv_shuffle;
  int round_out; /* round out[] to TF32 (round-to-nearest) for the consuming tensor-core op */
  float* colsum;
  float* colsum2;
  long long colsum_gstride;
} loft_epilogue_t;

/* ---- dense contractions on tcgen05 (TF32 operands, fp32 accumulate in TMEM) -------------------
 * replace cuDNN/cuBLAS behind torch.nn.Conv2d / nn.Linear / ConvTranspose2d at
 * resnet.py:163-203,260-300; fpn.py:116-132,164-216; rpn_head.py:26-44;
 * convfc_bbox_head.py:118-173; fcn_mask_head.py:64-126; offset_head_expand_feature.py:72-161. */
int loft_gemm_fprop(const float* x, const float* w, float* y, long long P, int K, int Cout,
                    long long ldx, long long ldw, long long ldy, int H, int W,
                    const loft_epilogue_t* epi, cudaStream_t stream);
int loft_gemm_dgrad(const float* dy, const float* w, float* dx, long long P, int Cin, int Cout,
                    long long lddy, long long ldw, long long lddx, const loft_epilogue_t* epi,
                    cudaStream_t stream);
/* gemm_dgrad with the pixel rows laid out as an [N,H,W] grid (needed by the zero-stuffed
 * residual mode of the epilogue) */
int loft_gemm_dgrad_hw(const float* dy, const float* w, float* dx, long long P, int Cin, int Cout,
                       long long lddy, long long ldw, long long lddx, int H, int W,
                       const loft_epilogue_t* epi, cudaStream_t stream);
int loft_gemm_wgrad(const float* dy, const float* x, float* dw, long long P, int Cin, int Cout,
                    long long lddy, long long ldx, long long lddw, cudaStream_t stream);
int loft_conv3x3_fprop(const float* x, const float* w, float* y, int N, int H, int W, int Cin,
                       int Cout, const loft_epilogue_t* epi, cudaStream_t stream);
int loft_conv3x3_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int Cin,
                       int Cout, const loft_epilogue_t* epi, cudaStream_t stream);
int loft_conv3x3_wgrad(const float* dy, const float* x, float* dw, int N, int H, int W, int Cin,
                       int Cout, cudaStream_t stream);
/* grouped forms: the N images are G consecutive groups; group g reads the weights at
 * w + g*w_gstride floats (scale/shift at + g*vec_gstride) -- the four FOA branches
 * (offset_head_expand_feature.py:134-161) in one launch per layer */
int loft_conv3x3_fprop_grouped(const float* x, const float* w, float* y, int N, int H, int W,
                               int Cin, int Cout, int G, long long w_gstride,
                               long long vec_gstride, const loft_epilogue_t* epi,
                               cudaStream_t stream);
int loft_conv3x3_dgrad_grouped(const float* dy, const float* w, float* dx, int N, int H, int W,
                               int Cin, int Cout, int G, long long w_gstride,
                               const loft_epilogue_t* epi, cudaStream_t stream);
int loft_conv3x3_wgrad_grouped(const float* dy, const float* x, float* dw, int N, int H, int W,
                               int Cin, int Cout, int G, long long dw_gstride,
                               cudaStream_t stream);
/* bring-up only: override UMMA descriptor fields (-1 = keep default) */
void loft_debug_set_desc(long long a_desc, long long b_desc, long long a_kstep, long long b_kstep,
                         long long idesc);
/* profiling only: 8 globaltimer stamps (+ k-loop cycle counts in -DLOFT_KTRACE builds) per CTA of
 * every later GEMM launch into `buf` (device memory, 128 B per CTA; NULL = off) --
 * tools/gemm_timeline.py */
void loft_debug_set_trace(unsigned long long* buf);

/* ---- HBM-bound layout / activation / optimizer kernels (elementwise.cu) ------------------------
 * BN-eval fold + backward: resnet.py:260-300,640-649; im2col for the 7x7/2 stem and the three
 * stride-2 3x3 convs: resnet.py:525-571,151-203; maxpool resnet.py:571,631; FPN extra level and
 * top-down backward: fpn.py:185-199; FOA rotation: offset_head_expand_feature.py:163-196
 * (affine_grid+grid_sample == rot90); optimizer: mmcv OptimizerHook(grad_clip) +
 * torch.optim.SGD as configured by configs/_base_/schedules/schedule_2x_bonai.py:2-3. */
int loft_fill(float* p, long long n, float v, cudaStream_t stream);
/* several independent loft_copy2d jobs in one launch (the per-step refresh of fused / padded head
 * weights and the scatter of their gradients) */
typedef struct {
  const float* src;
  long long lds;
  float* dst;
  long long ldd;
  long long rows;
  int cols;
  int accumulate;
  int round_tf32;
} loft_copy2d_t;
int loft_copy2d_multi(const loft_copy2d_t* items, int n, cudaStream_t stream);
int loft_copy2d(const float* src, long long lds, float* dst, long long ldd, long long rows, int cols,
                int accumulate, int round_tf32, cudaStream_t stream);
int loft_permute_acb(const float* src, float* dst, int A, int B, int C, int accumulate,
                     int round_tf32, cudaStream_t stream);
int loft_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                 float eps, float* scale, float* shift, float* rstd, int C, cudaStream_t stream);
/* BN folded into the conv weights the tensor cores read: T[row] = tf32(scale[ch] * P[row]) for the
 * rows (output channels) listed in the table; after backward, bn_finalize turns the accumulated
 * dW' rows into dW and adds dgamma[ch] += rstd*(W.dW' - mean*dbeta[ch]) -- no pass over the
 * activations (replaces BatchNorm2d eval backward, resnet.py:260-300 with norm_eval=True). */
int loft_bn_fold_weights(const float* P, float* T, const long long* row_off, const int* row_k,
                         const int* row_ch, const float* scale, long long rows,
                         cudaStream_t stream);
int loft_bn_finalize(const float* P, float* G, const long long* row_off, const int* row_k,
                     const int* row_ch, const float* scale, const float* rstd, const float* mean,
                     const float* dbeta, float* dgamma, long long rows, cudaStream_t stream);
int loft_act_bwd(const float* dy, const float* y, const float* z, const float* scale,
                 const float* mean, const float* rstd, float* dz, float* dres, float* dgamma,
                 float* dbeta, long long P, int C, int relu, cudaStream_t stream);
int loft_im2col(const float* x, float* col, int N, int H, int W, int C, int kh, int kw, int stride,
                int pad, int Kpad, int nchw_input, cudaStream_t stream);
/* direct 7x7/2 stem conv (replaces nn.Conv2d(3, 64, 7, 2, 3) + eval BN + ReLU of
 * mmdet/models/backbones/resnet.py:525-571,623-630 with no im2col matrix):
 * loft_stem_pack writes the NCHW fp32 image as xp[N][2][(H+7)/2][W+8][4] (zero-padded NHWC4,
 * TF32-rounded, even / odd padded rows in separate planes); loft_stem_conv7x7 reads it through an
 * overlapping-stride tensor map; w is [Cout][7][8][4] with zero pads, y is NHWC [N][Ho][Wo][Cout] */
int loft_stem_pack(const float* x, float* xp, int N, int H, int W, int C, cudaStream_t stream);
int loft_stem_conv7x7(const float* xp, const float* w, float* y, int N, int H, int W, int Cout,
                      const loft_epilogue_t* epi, cudaStream_t stream);
/* rows of a [*, cols] fp32 matrix by index (cols % 4 == 0): dst[r] = src[idx[r]] and
 * dst[idx[r]] += src[r] (idx unique).  Replaces the offset branch's own RoIAlign over the positive
 * RoIs (mmdet/models/roi_heads/loft_roi_head.py:118-120: offset_roi_extractor == bbox_roi_extractor,
 * bonai_loft_foa_r50_fpn_basic.py:37-42,69-74) by the positives' rows of bbox_feats */
int loft_gather_rows(const float* src, const long long* idx, float* dst, long long rows,
                     long long cols, cudaStream_t stream);
int loft_scatter_add_rows(const float* src, const long long* idx, float* dst, long long rows,
                          long long cols, cudaStream_t stream);
/* FOA input assembly fused with the row gather (offset_head_expand_feature.py:163-214: one
 * affine_grid + grid_sample rotation per branch, == rot90 for multiples of 90 degrees):
 * y[b*P + p] = rot90(x[idx[p]], k[b]) for nb <= 4 branches, x / y NHWC [*, S, S, C];
 * backward: gx[idx[p]] += sum_b rot90(gy[b*P + p], -k[b]) (idx unique) */
int loft_gather_rot(const float* x, const long long* idx, float* y, long long P, int S, int C,
                    const int* k, int nb, cudaStream_t stream);
int loft_scatter_rot_add(const float* gy, const long long* idx, float* gx, long long P, int S, int C,
                         const int* k, int nb, cudaStream_t stream);
/* 1x1 conv / linear head with at most 4 output channels over P rows of C channels, e.g. the
 * class-agnostic mask logits nn.Conv2d(256, 1, 1) (fcn_mask_head.py:118-126): w is [4][C] of which
 * the first n_out rows are real (the others zero, never read), y / dy are [P][4].  Backward in one
 * pass over x: dx = (premask ? x > 0 : 1) * (dy . w) rounded to TF32 (dx may be NULL), dw[n_out][C]
 * += dy^T x, db[4] += sum dy, colsum[C] += sum_p dx (bias gradient of the producing layer);
 * C in {128, 256, 512} for the backward.  s2d_h, s2d_w > 0: the rows of x / dx are the UN-shuffled
 * output [(n, h, w), (i, j)] of a 2x2 / stride-2 deconv GEMM over an s2d_h x s2d_w grid (its
 * gradient needs exactly that layout), while y / dy are in image order [n, 2h + i, 2w + j] */
int loft_narrow_head_fwd(const float* x, const float* w, const float* b, float* y, long long P,
                         int C, int n_out, int s2d_h, int s2d_w, cudaStream_t stream);
int loft_narrow_head_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw,
                         float* db, float* colsum, long long P, int C, int n_out, int premask,
                         int s2d_h, int s2d_w, cudaStream_t stream);
int loft_col2im(const float* dcol, float* dx, const float* mask, int N, int H, int W, int C, int kh,
                int kw, int stride, int pad, int Kpad, cudaStream_t stream);
int loft_maxpool3x3s2(const float* x, float* y, int N, int H, int W, int C, cudaStream_t stream);
int loft_subsample2(const float* x, float* y, int N, int H, int W, int C, cudaStream_t stream);
int loft_subsample2_bwd(const float* dy, float* dx, const float* mask, int N, int H, int W, int C,
                        cudaStream_t stream);
int loft_sum2x2_add(const float* fine, const float* base, float* out, int N, int H, int W, int C,
                    cudaStream_t stream);
int loft_rot90(const float* x, float* y, long long K, int S, int C, int k, cudaStream_t stream);
int loft_add(const float* a, const float* b, float* out, long long n, int round_tf32,
             cudaStream_t stream);
/* out = a + b over [rows, C] with colsum[c] += sum_rows out[., c] in the same pass (C/4 | 256) */
int loft_add_colsum(const float* a, const float* b, float* out, long long rows, int C,
                    float* colsum, int round_tf32, cudaStream_t stream);
int loft_grad_sqnorm(const float* g, long long n, double* out, cudaStream_t stream);
int loft_sgd_clip_step(float* p, const float* g, float* m, float* p_tf32, long long n, float lr,
                       float momentum, float weight_decay, float max_norm, float grad_scale,
                       const double* sqnorm, cudaStream_t stream);

/* ---- RoIAlign over the FPN pyramid + mask-target sampling (roi_align.cu) -----------------------
 * replace mmcv.ops.RoIAlign / roi_align at single_level_roi_extractor.py:32-80 and
 * core/mask/structures.py:261-291 (+ mask_target.py:31-62).  rois [K,5] = (batch, x1,y1,x2,y2);
 * out is [K,S,S,C] (NHWC). */
int loft_roi_align_fwd(const float* const* feats, const int* Hs, const int* Ws, const float* scales,
                       int num_levels, const float* rois, long long K, int S, int C,
                       float finest_scale, float* out, int* levels_out, cudaStream_t stream);
int loft_roi_align_bwd(float* const* grads, const int* Hs, const int* Ws, const float* scales,
                       int num_levels, const float* rois, long long K, int S, int C,
                       float finest_scale, const float* dout, cudaStream_t stream);
int loft_mask_target(const uint8_t* masks, const float* boxes, const long long* gt_inds, long long P,
                     int S, int H, int W, float* out, cudaStream_t stream);

/* ---- assignment / proposals / NMS / target encoders (detect.cu) ---------------------------------
 * max_iou_assigner.py:127-212 + iou2d_calculator.py:39-130; rpn_head.py:79-168 +
 * anchor_generator.py:142-271 + delta_xywh_bbox_coder.py:119-197; mmcv.ops.batched_nms;
 * delta_xywh_bbox_coder.py:74-116; offset_head_expand_feature.py:271-344 +
 * delta_xy_offset_coder.py:46-65. */
size_t loft_iou_assign_workspace(long long n, int G);
int loft_iou_assign(const float* boxes, long long n, const float* gts, int G, float pos_thr,
                    float neg_thr, float min_pos, int match_low_quality, long long* gt_inds,
                    float* max_overlaps, void* workspace, size_t ws_bytes, cudaStream_t stream);
int loft_rpn_decode(const float* head_out, int ld, int reg_off, const long long* topk_idx, int k,
                    int fw, int A, const float* base_anchors, float stride, float max_ratio,
                    float img_h, float img_w, float* boxes_out, int batch, long long head_stride,
                    long long idx_stride, long long out_stride, cudaStream_t stream);
/* RCNN RoI sampling: BaseSampler.sample + RandomSampler (samplers/base_sampler.py:34-101,
 * random_sampler.py:31-75) with add_gt_as_proposals, one block per image, after loft_iou_assign.
 * Candidates = [gt boxes, proposals]; up to num_pos_max positives, then negatives up to `num` in
 * total, each drawn uniformly without replacement (counter-based hash of `seed`), emitted in
 * ascending candidate order, positives first.  props [B,K,prop_ld], num_valid [B] or NULL,
 * prop_gt_inds [B,K], gts [sum G,4] + gt_off [B+1] (device), max_gt = largest G (host).  Outputs:
 * sel / out_gt [B,num] i64, out_boxes [B,num,4], out_isgt [B,num] u8, cnt [B,2] i32. */
int loft_rcnn_sample(const float* props, long long prop_stride, int K, int prop_ld,
                     const int* num_valid, const long long* prop_gt_inds, const float* gts,
                     const int* gt_off, int B, int max_gt, int num, int num_pos_max,
                     unsigned long long seed, long long* sel, float* out_boxes, long long* out_gt,
                     unsigned char* out_isgt, int* cnt, cudaStream_t stream);
size_t loft_nms_workspace(int n);
int loft_nms_sorted(const float* boxes, const long long* idxs, int B, int n, float iou_thr,
                    int max_keep, long long* keep, int* num_keep, void* workspace, size_t ws_bytes,
                    cudaStream_t stream);
/* per-level ("segmented") form of the same batched NMS for the RPN: boxes [B,n,4] are
 * SEGMENT-MAJOR (level 0 first, ...; seg_off = L+1 host ints), each segment sorted by score;
 * order [B,n] lists segment-major indices in global score order.  Identical keep / num_keep to
 * loft_nms_sorted on the globally sorted boxes with idxs = level, at ~1/5 of the pair tests. */
size_t loft_nms_segmented_workspace(const int* seg_off, int L);
int loft_nms_segmented(const float* boxes, const int* seg_off, int L, const long long* order, int B,
                       int n, float iou_thr, int max_keep, long long* keep, int* num_keep,
                       void* workspace, size_t ws_bytes, cudaStream_t stream);
/* test-time linear soft-NMS (mmcv.ops.soft_nms via core/post_processing/bbox_nms.py:63 with
 * bonai_loft_foa_r50_fpn_basic.py:138); CPU-only in mmcv 1.0.5, one-block kernel here */
int loft_soft_nms_linear(const float* boxes, const float* scores, const long long* idxs, int n,
                         float iou_thr, float min_score, int max_keep, float* dets, long long* keep,
                         int* num_keep, cudaStream_t stream);
/* test-time mask paste: FCNMaskHead.get_seg_masks + _do_paste_mask (fcn_mask_head.py:151-308).
 * logits: element (n, i) of the M*M mask of detection n at logits[n*ld_n + i*ld_px]; boxes row n at
 * boxes + n*ld_box (x0,y0,x1,y1); out: uint8 [N, img_h, img_w], ZERO-FILLED by the caller -- only
 * each detection's box window is written (thr >= 0: {0,1} = prob >= thr; thr < 0: prob*255). */
int loft_paste_masks(const float* logits, long long ld_n, int ld_px, const float* boxes, int ld_box,
                     int N, int M, int img_h, int img_w, float thr, unsigned char* out,
                     cudaStream_t stream);
/* OffsetHeadExpandFeature.offset_fusion('max') over the 4 rotated branches (pred rows b*n + i) +
 * DeltaXYOffsetCoder.decode (offset_head_expand_feature.py:346-448, delta_xy_offset_coder.py:67-88);
 * max_x <= 0: no clamp.  out [n, 2]. */
int loft_offset_fusion_decode(const float* pred, int ld, long long n, const float* boxes, int ld_box,
                              float std_x, float std_y, float max_x, float max_y, float* out,
                              cudaStream_t stream);
int loft_bbox_encode(const float* props, const float* gts, long long n, float s0, float s1, float s2,
                     float s3, float* out, cudaStream_t stream);
int loft_offset_target(const float* props, const float* gt_offsets, const long long* gt_inds,
                       long long P, float std_x, float std_y, float* out, cudaStream_t stream);

/* ---- device side of the post-load training pipeline (pipeline.cu) -----------------------------
 * configs/_base_/datasets/bonai_instance.py:5-17 on a 1024^2 BONAI tile (Resize is the identity):
 * RandomFlip (transforms.py:484-488, mmcv.imflip) -> Normalize (:655-676, mmcv.imnormalize: BGR->RGB,
 * subtract mean in fp32, multiply by 1/std as a double scalar) -> Pad to a multiple of 32 with 0
 * (:571-600) -> HWC->CHW (formating.py:191-230) in one pass over the uint8 image;
 * BitmapMasks.flip + pad (core/mask/structures.py:218-240) in one pass over the uint8 bitmaps.
 * flip: 0 none, 1 horizontal, 2 vertical.  mean3 / std3 are HOST pointers to 3 floats (RGB order). */
int loft_image_prep(const uint8_t* img_hwc, float* out_chw, int H, int W, int Hp, int Wp,
                    const float* mean3, const float* std3, int to_rgb, int flip,
                    cudaStream_t stream);
int loft_mask_flip_pad(const uint8_t* in, uint8_t* out, long long G, int H, int W, int Hp, int Wp,
                       int flip, cudaStream_t stream);

/* Resize(keep_ratio) for tiles not already at img_scale (transforms.py:186-215,244-253 ->
 * mmcv.imrescale -> cv2.resize): uint8 HWC image with INTER_LINEAR (OpenCV's 11-bit fixed-point
 * arithmetic restated), uint8 [G,H,W] bitmaps with INTER_NEAREST.  The BONAI configuration never
 * takes this path (1024^2 tiles at img_scale 1024). */
int loft_resize_bilinear_u8(const uint8_t* src_hwc, uint8_t* dst_hwc, int sh, int sw, int dh, int dw,
                            cudaStream_t stream);
int loft_resize_nearest_u8(const uint8_t* src, uint8_t* dst, long long G, int sh, int sw, int dh,
                           int dw, cudaStream_t stream);

/* Polygon -> bitmap: LoadAnnotations(poly2mask=True) (pipelines/loading.py:301-326,345-368), i.e.
 * pycocotools 2.0.x `decode(merge(frPyObjects(polygons, h, w)))` (common/maskApi.c rleFrPoly /
 * rleMerge / rleDecode; the dependency is not vendored in the reference tree).  xy: vertices of all
 * parts as (x, y) doubles; part_off[n_parts+1]: first vertex of each part; inst_part_off[n_inst+1]:
 * first part of each instance; out: uint8 [n_inst, H, W] row-major, every byte written.  scratch:
 * loft_poly_scratch_bytes() bytes of device memory; *err (device int, zero it first) is set to
 * 1 + part index if a part has more than 8192 run boundaries (its bitmap is then incomplete). */
long long loft_poly_scratch_bytes(long long total_vertices, int n_parts);
int loft_poly_rasterize(const double* xy, const long long* part_off, const int* inst_part_off,
                        int n_inst, int n_parts, long long total_vertices, int H, int W,
                        uint8_t* out, void* scratch, int* err, cudaStream_t stream);

/* ---- losses (loss.cu) ----------------------------------------------------------------------------
 * mode 0 = BCE-with-logits (cross_entropy_loss.py:58-125), 1 = L1, 2 = SmoothL1
 * (smooth_l1_loss.py:8-42); sums are accumulated into device scalars (weight_reduce_loss,
 * losses/utils.py:26-52); gscale is the device-resident upstream gradient (NULL = 1). */
/* AnchorHead.loss over all pyramid levels in one launch (anchor_head.py:382-497): sums[0..n) =
 * per-level sigmoid-BCE * cls_scale, sums[n..2n) = per-level L1 (mode 1) / SmoothL1 (mode 2) *
 * bbox_scale; if `grad` of a level is set, d(total)/d(fused head output) is written there in full
 * ([rows, ld], padding columns zero).  Fused head output row = (image, y, x), columns [0,A) cls,
 * [A,5A) deltas; labels / label_w are [rows*A], bbox_t / bbox_w [rows*4A]. */
typedef struct {
  const float* out;
  const float* labels;
  const float* label_w;
  const float* bbox_t;
  const float* bbox_w;
  float* grad;
  long long rows;
} loft_rpn_level_t;
/* denom: optional device scalar (the avg_factor loft_rpn_targets leaves on the device); both
 * scales are divided by max(*denom, 1) */
int loft_rpn_loss_fused(const loft_rpn_level_t* levels, int n_levels, int A, int ld, int mode_bbox,
                        float beta, float cls_scale, float bbox_scale, const float* denom,
                        float* sums, cudaStream_t stream);
/* RPN sampling + targets of a batch without a host read-back: replaces
 * AnchorHead._get_targets_single after the assigner + RandomSampler.sample + images_to_levels
 * (mmdet/models/dense_heads/anchor_head.py:206-278,363-380, core/bbox/samplers/random_sampler.py:31-75).
 * gt_inds [B, A]: MaxIoUAssigner output per image; lvl_off[n_levels+1]: anchor offsets of the
 * levels; outputs: flat B*A (labels, label_w) / B*A*4 (bbox_t, bbox_w) floats, level l at
 * [B*lvl_off[l], B*lvl_off[l+1]) in (image, anchor) order; total[0] = sum over images of
 * max(#pos,1) + max(#neg,1) */
size_t loft_rpn_targets_workspace(int B);
int loft_rpn_targets(const float* anchors, const long long* gt_inds, const float* gts,
                     const int* gt_off, const long long* lvl_off, int n_levels, int B, long long A,
                     int num, int num_pos_max, unsigned long long seed, float s0, float s1,
                     float s2, float s3, float pos_weight, float* labels, float* label_w,
                     float* bbox_t, float* bbox_w, float* total, void* workspace, size_t ws_bytes,
                     cudaStream_t stream);
int loft_rpn_targets_overflowed(const void* workspace, cudaStream_t stream);
int loft_elem_loss_fwd(int mode, const float* pred, long long ld, int col_off, int ncols,
                       long long rows, const float* target, const float* weight, float beta,
                       float scale, float* out_sum, cudaStream_t stream);
int loft_elem_loss_bwd(int mode, const float* pred, long long ld, int col_off, int ncols,
                       long long rows, const float* target, const float* weight, float beta,
                       float scale, const float* gscale, float* dpred, cudaStream_t stream);
int loft_softmax_ce_fwd(const float* logits, long long ld, int C, long long n,
                        const long long* labels, const float* weight, float scale, float* out2,
                        cudaStream_t stream);
int loft_softmax_ce_bwd(const float* logits, long long ld, int C, long long n,
                        const long long* labels, const float* weight, float scale,
                        const float* gscale, float* dlogits, cudaStream_t stream);
int loft_sigmoid_focal_loss_fwd(const float* x, const long long* target, const float* weight,
                                long long n, int C, float gamma, float alpha, float* loss,
                                cudaStream_t stream);
int loft_sigmoid_focal_loss_bwd(const float* x, const long long* target, const float* weight,
                                long long n, int C, float gamma, float alpha, const float* dloss,
                                float* dx, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LOFT_B200_H_ */
