/*
 * ryolo.h -- C-ABI of the B200-native rotate-yolov3 hot path (libryolo.so).
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / a CUDA stream handle
 * (passed as void* so that this header needs no CUDA include), allocates nothing behind the
 * caller's back (scratch comes from a caller-owned workspace whose size is queried first),
 * never throws, and returns an int status: 0 = ok, <0 = RYOLO_E_* (see ryolo_last_error()).
 * All data pointers are DEVICE pointers unless the parameter name ends in `_host`.
 * There is NO CPU fallback: without a CUDA device every compute call returns RYOLO_E_CUDA.
 *
 * Reference interfaces replaced (paths relative to the reference repo, ming71/rotate-yolov3):
 *   ryolo_rnms*            <- r_nms(dets, thr)          utils/nms/src/rotate_polygon_nms.cpp:7-16
 *                             nms_cuda()                utils/nms/src/rotate_polygon_nms_kernel.cu:323-384
 *                             rotate_nms_kernel         utils/nms/src/rotate_polygon_nms_kernel.cu:262-308
 *   ryolo_riou_*           <- skew_bbox_iou(box1, box2, GIoU)   utils/utils.py:290-320
 *                             skewiou / get_rotated_coors        utils/utils.py:663-699, 702-725
 *   ryolo_nms_filter       <- non_max_suppression() candidate filter  utils/nms/nms.py:34-40,55
 *   ryolo_conv_*           <- Conv2d -> BatchNorm2d -> PReLU blocks built by create_modules
 *                             model/models.py:49-66; folded BN per utils/torch_utils.py:45-69
 *   ryolo_yolo_decode      <- YOLOLayer.forward eval branch   model/models.py:183-227,
 *                             create_grids                     model/model_utils.py:16-35
 */
#ifndef RYOLO_H_
#define RYOLO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RYOLO_OK 0
#define RYOLO_E_ARG (-1)       /* bad argument (null pointer, negative size, unknown mode) */
#define RYOLO_E_WORKSPACE (-2) /* workspace too small / misaligned                          */
#define RYOLO_E_CUDA (-3)      /* CUDA runtime error (text in ryolo_last_error)             */
#define RYOLO_E_UNSUPPORTED (-4)

/* ABI version of this header; bumped on any signature change. */
int ryolo_abi_version(void);
/* Thread-local text of the last failure on the calling thread ("" if none). */
const char* ryolo_last_error(void);
/* Number of this library's kernel launches issued by the calling process so far
 * (bench.py's `gpu_launches` evidence). */
uint64_t ryolo_launch_count(void);
/* Keep n SMs free of this library's persistent GEMM kernels (conv fwd / dgrad / wgrad grids = SM count - n) from now on:
 * a data-parallel training step reserves a few SMs during backward so that the concurrent NCCL all-reduce kernels are
 * resident next to them instead of forcing a second wave (process-wide setting, 0 <= n <= 64; 0 = use every SM). */
int ryolo_set_reserved_sms(int n);

/* ------------------------------------------------------------------------------------------ *
 * Rotated NMS  (reference: r_nms / nms_cuda, rotate_polygon_nms_kernel.cu:323-384)
 *
 * dets     [n,6] fp32 (cx, cy, w, h, theta_rad, score), row-major, device.
 * thr      suppress j iff IoU(i,j) > thr (strict), i = higher score; NaN IoU never suppresses.
 * keep_out [n] int64, device: receives the ORIGINAL indices of the kept boxes, ASCENDING
 *          (reference :380-383); only the first *num_keep entries are defined.
 * num_keep device int32[1].
 * The IoU arithmetic replicates the reference device code operation-for-operation, including
 * the FMA contractions nvcc's default -fmad=true applies to it (DESIGN.md "pinned arithmetic"),
 * so the kept index list is bit-identical to the reference kernel's on the same inputs when the
 * scores are tie-free (the reference's torch sort is unstable; ours is a stable descending sort).
 * The mask is evaluated lazily in row chunks of 1024 boxes, alternating with the greedy scan:
 * rows and columns already suppressed by earlier chunks are skipped -- their mask bits can never
 * influence the scan, so the kept list is unchanged while dense candidate sets cost far less.
 * Asynchronous on `stream`; no host synchronisation inside.
 * ------------------------------------------------------------------------------------------ */
size_t ryolo_rnms_workspace_bytes(int n);
int ryolo_rnms(const float* dets, int n, float thr, int64_t* keep_out, int32_t* num_keep,
               void* workspace, size_t workspace_bytes, void* stream);

/* Same result as ryolo_rnms, but computes EVERY upper-triangle mask word in one launch like the
 * reference kernel (no skipping of already-suppressed rows/columns).  Slower; exists so that the
 * parity tests can compare the whole mask with the reference's. */
int ryolo_rnms_full_mask(const float* dets, int n, float thr, int64_t* keep_out, int32_t* num_keep,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Batched ("segmented") form: `segments` independent NMS problems -- one per (image, class) -- in ONE set of launches
 * (reference: the per-image / per-class Python loop of non_max_suppression, utils/nms/nms.py:28-69, calling r_nms once
 * per class).  Segment s owns rows [s*cap, s*cap + n_dev[s]) of dets [segments*cap, 6]; its counts come from DEVICE
 * memory (ryolo_detect_select), so nothing synchronises with the host.  At most `limit` boxes per segment -- the
 * best-scored, ties by row index -- enter the NMS.  keep_out [segments, cap] int64: kept row indices WITHIN the segment,
 * ascending; num_keep [segments] int32.  Same kernels and same pinned arithmetic as ryolo_rnms (which is the
 * segments = 1, host-count case): per segment the kept list is bit-identical to r_nms on that segment's boxes. */
size_t ryolo_rnms_batched_workspace_bytes(int segments, int cap);
int ryolo_rnms_batched(const float* dets, int segments, int cap, const int32_t* n_dev, int limit, float thr,
                       int64_t* keep_out, int32_t* num_keep, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Batched candidate selection in front of it (reference: utils/nms/nms.py:34-40 per image; detect.py:204-213).
 * io [batch, p, 6+nc] fp32 decoded predictions (NOT modified).  Per image: rows passing the reference's filter
 * (conf = obj * max class conf > conf_thres, w and h > min_wh, all values finite) whose conf lies in or above the
 * histogram bin (4096 linear bins on [0,1]) of the limit-th largest conf are written IN INPUT ORDER to
 * dets_out [batch, cap, 6] = (x, y, w, h, theta, conf); n_out [batch] int32 = their number (<= cap; cap >= limit,
 * the surplus over limit is at most one bin's population and is cut by ryolo_rnms_batched's `limit`). */
size_t ryolo_detect_select_workspace_bytes(int batch, int p);
int ryolo_detect_select(const float* io, int batch, int p, int nc, float conf_thres, float min_wh, int limit,
                        float* dets_out, int cap, int32_t* n_out, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Introspection of the last ryolo_rnms_full_mask call that used `workspace` (valid until the workspace is
 * reused): device pointers to the score-sorted boxes [n,6], the sort permutation order[n]
 * (int32, sorted position -> original index) and the suppression mask [n, ceil(n/64)] uint64 in the
 * reference's layout (only words with column-block >= row-block are defined, as in the
 * reference host scan :371-374).  Used by the parity tests to compare masks bit-for-bit. */
int ryolo_rnms_debug_views(void* workspace, int n, const float** sorted_boxes,
                           const int32_t** order, const unsigned long long** mask);

/* ------------------------------------------------------------------------------------------ *
 * Rotated IoU  (reference: skew_bbox_iou, utils/utils.py:290-320)
 *
 * Boxes are rows of `stride` floats (stride >= 5) whose first five are (cx, cy, w, h, theta_rad);
 * corners per get_rotated_coors (utils/utils.py:702-725).  mode: 0 = 'iou'
 * (inter / (A1 + A2 - inter)), 1 = 'giou' as the reference defines it (inter / area of the
 * axis-aligned envelope of the 8 corners, utils/utils.py:682-685).  Zero-area boxes and a zero
 * denominator give 0 (utils/utils.py:672-673, 693-694).  fp32 in/out; geometry is an fp32
 * convex clip in box1's local frame (tolerance vs the float64 oracle: 1e-4 rel + 1e-6 abs).
 * ------------------------------------------------------------------------------------------ */
#define RYOLO_IOU_MODE_IOU 0
#define RYOLO_IOU_MODE_GIOU 1
/* out[i] = iou(a_i, b_i), i < n  (the reference's N-vs-N form) */
int ryolo_riou_paired(const float* a, const float* b, int n, int stride_a, int stride_b, int mode,
                      float* out, void* stream);
/* out[i*m + j] = iou(a_i, b_j)  (the reference's 1-vs-N form applied to every row of a) */
int ryolo_riou_pairwise(const float* a, int n, int stride_a, const float* b, int m, int stride_b,
                        int mode, float* out, void* stream);

/* Rotated IoU as a differentiable op (the README's "riou loss", which model/loss.py does not
 * implement): iou_out[i] = IoU(a_i, b_i) (mode 'iou' semantics as above) and, given grad_out[i] = dL/d iou_i (NULL = 1),
 * the analytic gradients grad_a / grad_b [n, 5] w.r.t. (cx, cy, w, h, theta) of each box (boundary-velocity formula:
 * d area(a ∩ b) = integral over the part of the moving box's boundary inside the other box of its normal velocity;
 * riou_grad.cu).  Any of iou_out / grad_a / grad_b may be NULL. */
int ryolo_riou_paired_grad(const float* a, const float* b, int n, int stride_a, int stride_b,
                           const float* grad_out, float* iou_out, float* grad_a, float* grad_b, void* stream);

/* Greedy prediction <-> ground-truth assignment of the mAP evaluation (reference test.py:134-151) on the device.
 * iou [p, t] (row stride iou_stride) = ryolo_riou_pairwise of the image's confidence-sorted predictions against its
 * targets; pcls: class of prediction i at pcls[i * pcls_stride]; tcls [t].  In order, a prediction is correct when the
 * best-IoU target OF ITS CLASS (first maximum) has IoU > iou_thres and was not claimed before; the walk stops once every
 * target is claimed.  correct [p] uint8; claimed_scratch [t] uint8. */
int ryolo_match_detections(const float* iou, int p, int t, int iou_stride, const float* pcls, int pcls_stride,
                           const float* tcls, float iou_thres, unsigned char* claimed_scratch,
                           unsigned char* correct, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * non_max_suppression candidate filter (reference: utils/nms/nms.py:34-40, 55)
 *
 * pred [p, 6+nc] fp32 (x,y,w,h,theta,obj, cls...), one image.  Exactly like the reference:
 * class_conf/class_pred = max/argmax over the class columns (first max on ties), pred[:,5] *=
 * class_conf IN PLACE (nms.py:35), keep rows with conf > conf_thres, w > min_wh, h > min_wh and
 * all 6+nc values finite.  Survivors are written in input order to out [*, 8] =
 * (x,y,w,h,theta,conf,class_conf,class) and counted in num_out (device int32[1]).
 * Stable (order-preserving) compaction; capacity = max rows `out` can hold (extra rows dropped,
 * num_out still reports the true count so that the caller can detect overflow).
 * ------------------------------------------------------------------------------------------ */
size_t ryolo_nms_filter_workspace_bytes(int p);
int ryolo_nms_filter(float* pred, int p, int nc, float conf_thres, float min_wh, float* out,
                     int capacity, int32_t* num_out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ------------------------------------------------------------------------------------------ *
 * YOLO head decode (reference: YOLOLayer.forward eval branch, model/models.py:198-227)
 *
 * p        [bs, na*(nc+6), ny, nx] fp32 NCHW raw head output (the conv's natural layout).
 * anchors  [na, 3] fp32 (w_px, h_px, theta): anchor_vec = (w/stride, h/stride, theta)
 *          (model/model_utils.py:30-32).
 * io_out   [bs, io_rows_total, nc+6] fp32; this layer writes rows [row_offset, row_offset+na*ny*nx)
 *          of every image, in the reference's (a, y, x) order, so the three layers concatenate
 *          by row_offset exactly like torch.cat(io, 1) (model/models.py:298).
 * p_out    optional [bs, na, ny, nx, nc+6] fp32 permuted raw copy (the second tensor the
 *          reference returns); may be NULL.
 * arc_default != 0 -> sigmoid on obj and class columns ('default' arcs); nc == 1 forces the
 * class column to 1 (models.py:220-221).
 * ------------------------------------------------------------------------------------------ */
int ryolo_yolo_decode(const float* p, int bs, int na, int nc, int ny, int nx, const float* anchors,
                      float stride, float context_factor, int arc_default, float* io_out,
                      int io_rows_total, int row_offset, float* p_out, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Conv2d -> folded BatchNorm -> PReLU (+ residual, + x2 upsample) blocks
 * (reference: create_modules, model/models.py:49-66; shortcut add :281-282; nn.Upsample :93-94;
 *  BN folding utils/torch_utils.py:45-69).  sm_100a tcgen05 implicit GEMM, bf16 in / fp32 accumulate.
 *
 * Activations live in HBM as "padded NHWC": [B, H+2, W+2, Cs] bf16 with a one-pixel ZERO halo
 * (zeroed once by the owner of the buffer; these kernels never write it), channels contiguous
 * with channel stride Cs >= C, so that a 3x3/stride-1 conv is nine row-shifted GEMMs over the
 * flat pixel index and needs no im2col buffer, and channel concatenation (route layers,
 * models.py:269-278) is a write at a channel offset into a wider buffer.  Channel counts are
 * padded to multiples of 64 with zeros (the K chunk of the GEMM).
 * ------------------------------------------------------------------------------------------ */
#define RYOLO_DT_BF16 0
#define RYOLO_DT_F32 1

typedef struct ryolo_conv_desc {
  int32_t batch;       /* B                                                                   */
  int32_t in_h, in_w;  /* input spatial size (unpadded)                                       */
  int32_t cin;         /* channels read by this conv                                          */
  int32_t cin_stride;  /* channel stride of the input buffer (>= cin, multiple of 8)           */
  int32_t cout;        /* filters                                                             */
  int32_t cout_stride; /* channel stride of the bf16 output buffer (>= round_up(cout, 32); channels
                          [cout, round_up(cout,32)) are written as zeros, nothing beyond)     */
  int32_t ksize;       /* 1 or 3 (pad = (k-1)/2 as in models.py:53); 2 = 2x2 taps at offsets
                          {-1,0}^2 and -2 = offsets {0,+1}^2 (the space-to-depth form of a
                          3x3/stride-2 conv and its adjoint, see ryolo_space_to_depth)         */
  int32_t stride;      /* 1 or 2                                                              */
  int32_t has_act;     /* 1: PReLU with scalar `slope` (cfg activation=leaky), 0: linear      */
  float slope;         /* PReLU weight (nn.PReLU(num_parameters=1), models.py:65)             */
  int32_t has_residual;/* 1: add `residual` after the activation (shortcut, models.py:281-282) */
  int32_t res_stride;  /* channel stride of the residual buffer (padded NHWC at OUTPUT size)  */
  int32_t upsample2x;  /* 1: replicate each output pixel into the 2x2 block of a 2H x 2W output
                          buffer (nearest x2 folded into the producer)                        */
  int32_t out_dtype;   /* RYOLO_DT_BF16: padded NHWC; RYOLO_DT_F32: plain NCHW [B,cout,H,W]
                          (the three linear heads feeding ryolo_yolo_decode)                  */
} ryolo_conv_desc;

/* Pack reference-layout weights [cout, cin, k, k] fp32 (nn.Conv2d.weight), times an optional
 * per-filter scale (folded BN: gamma / sqrt(var + eps)), into the bf16 GEMM operand
 * [k*k][cout_pad][cin_pad] (zero padded).  Device pointers. */
size_t ryolo_conv_packed_weight_bytes(const ryolo_conv_desc* d);
int ryolo_conv_pack_weights(const ryolo_conv_desc* d, const float* weight, const float* scale,
                            void* packed_out, void* stream);
/* Same with the operand transform done in the packing kernel (no framework-side flips / transposes / gathers):
 * mode 0 = plain; 1 = dgrad operand of a plain conv (`weight` is the FORWARD weight, `d` the dgrad
 * descriptor: taps mirrored, matrix transposed); 2 = space-to-depth form of a 3x3/stride-2 conv
 * (`weight` [cout, C, 3, 3], d->cin = 4C, d->ksize = 2); 3 = dgrad operand of mode 2 (d->cout = 4C,
 * d->ksize = -2).  ryolo_conv_unpack_wgrad is the inverse gather for weight GRADIENTS: wgrad output
 * [taps][cout_pad][cin_pad] -> nn.Conv2d layout [cout][cin][k][k] (mode 0) or, from the
 * space-to-depth form, [cout][C][3][3] (mode 2, cin = C). */
int ryolo_conv_pack_weights_ex(const ryolo_conv_desc* d, const float* weight, const float* scale,
                               void* packed_out, int mode, void* stream);
int ryolo_conv_unpack_wgrad(const float* dw, int cout_pad, int cin_pad, int mode, int cout, int cin,
                            int ksize, float* grad, void* stream);
/* Multi-tensor forms (one launch for all layers of a training step; the per-layer launches are a fixed ~2 ms per step
 * that caps strong scaling).  Job tables and prefix sums (prefix[j] = first global element of job j, prefix[njobs] =
 * total) live in DEVICE memory and are built once per plan.  Pack: `packed` receives taps*cout_pad*cin_pad bf16 for the
 * operand described like ryolo_conv_pack_weights_ex (ks = taps per side of the PACKED operand: 1, 3 or 2; cout/cin of
 * the packed operand; mode 0..3).  Unpack: `grad` receives cout*cin*ks*ks fp32 in nn.Conv2d layout (mode 0 / 2 as
 * ryolo_conv_unpack_wgrad; ks = kernel size of the nn.Conv2d weight). */
typedef struct ryolo_pack_job {
  const float* weight;
  void* packed;
  int32_t cout, cin, ks, cout_pad, cin_pad, mode;
} ryolo_pack_job;
typedef struct ryolo_unpack_job {
  const float* dw;
  float* grad;
  int32_t cout_pad, cin_pad, mode, cout, cin, ks;
} ryolo_unpack_job;
int ryolo_conv_pack_weights_multi(const ryolo_pack_job* jobs_dev, int njobs, const long long* prefix_dev,
                                  long long total, void* stream);
int ryolo_conv_unpack_wgrad_multi(const ryolo_unpack_job* jobs_dev, int njobs, const long long* prefix_dev,
                                  long long total, void* stream);
/* y = act(conv(x, W') + bias) [+ residual].  bias: fp32 [cout_pad] (zero beyond cout). */
size_t ryolo_conv_workspace_bytes(const ryolo_conv_desc* d);
int ryolo_conv_bn_act_fwd(const ryolo_conv_desc* d, const void* x, const void* packed_w,
                          const float* bias, const void* residual, void* y, void* workspace,
                          size_t workspace_bytes, void* stream);
/* First layer (cin = 3, 3x3, stride 1; reference model/models.py:45-74 block 0): fp32 NCHW image
 * [B,3,H,W] -> bf16 padded NHWC with channel stride cout_stride (channels >= cout zero-filled).
 * weight [cout,3,3,3] fp32 with BN folded (rounded to bf16 in the kernel like every other layer's
 * weights), bias [cout], cout = 16 | 32.  The im2col row of each pixel is built in registers
 * (image split into bf16 hi + lo halves: 16 mantissa bits) and contracted with tcgen05.mma.
 * ryolo_conv_first_s2d_fwd writes the same values straight into the space-to-depth buffer of a
 * following 3x3/stride-2 layer (layout of ryolo_space_to_depth; H, W even; xs_cstride >= 4*cout). */
int ryolo_conv_first_fwd(const float* img, int batch, int h, int w, const float* weight,
                         const float* bias, int cout, float slope, void* y, int cout_stride,
                         void* stream);
int ryolo_conv_first_s2d_fwd(const float* img, int batch, int h, int w, const float* weight,
                             const float* bias, int cout, float slope, void* xs, int xs_cstride,
                             void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Training path of the conv blocks (reference: autograd through nn.Conv2d / BatchNorm2d / PReLU,
 * train.py:268-282).  Same padded-NHWC bf16 layout; gradients of activations are bf16, gradients
 * of parameters fp32.
 * ------------------------------------------------------------------------------------------ */
/* dW[tap][co][ci] (fp32, [k*k][cout_pad][cin_pad], ACCUMULATED into a caller-zeroed buffer) =
 * sum over flat padded pixels p of dz[p, co] * x[p + tap_offset, ci].  dz / x: padded NHWC bf16 at
 * the conv's INPUT resolution (a stride-2 conv passes its output gradient zero-inserted onto the
 * input grid).  tcgen05 GEMM with K = pixels, MN-major operands, split-K + fp32 atomics. */
int ryolo_conv_wgrad(const void* dz, int dz_cstride, int cout_pad, const void* x, int x_cstride,
                     int cin_pad, int batch, int in_h, int in_w, int ksize, float* dw, void* stream);

/* Per-channel batch statistics of a raw conv output z (padded NHWC bf16, interior pixels):
 * sums[0..c) = sum z, sums[c..2c) = sum z^2 (fp32; zeroed inside). */
int ryolo_bn_stats(const void* z, int z_cstride, int batch, int h, int w, int c, float* sums,
                   void* stream);
/* sums -> per-channel mean, invstd, scale = gamma*invstd, shift = beta - mean*scale, and (when the
 * pointers are non-null) the nn.BatchNorm2d running statistics update with momentum and the
 * unbiased variance (reference model/models.py:62: momentum 0.1, eps 1e-5). */
/* ryolo_bn_stats + ryolo_bn_finalize in one call: when z's rows are contiguous (z_cstride == c) the last CTA of the
 * statistics kernel to finish finalises (one launch per block instead of two); otherwise the two kernels run back to
 * back.  sums: [2*c + 1] fp32 scratch (zeroed by the call; the extra word is the CTA ticket). */
int ryolo_bn_stats_finalize(const void* z, int z_cstride, int batch, int h, int w, int c, float* sums, float eps,
                            float momentum, const float* gamma, const float* beta, float* mean, float* invstd,
                            float* scale, float* shift, float* running_mean, float* running_var, void* stream);
int ryolo_bn_finalize(const float* sums, int c, float count, float eps, float momentum,
                      const float* gamma, const float* beta, float* mean, float* invstd,
                      float* scale, float* shift, float* running_mean, float* running_var,
                      void* stream);
/* y = prelu(z * scale + shift) [+ residual], scale/shift = folded batch-stat BN (gamma*invstd,
 * beta - mean*gamma*invstd); upsample2x writes every pixel to its 2x2 block of a 2h x 2w y. */
int ryolo_bn_act_fwd(const void* z, int z_cstride, int batch, int h, int w, int c,
                     const float* scale, const float* shift, float slope, int has_act,
                     const void* residual, int res_cstride, void* y, int y_cstride, int upsample2x,
                     const float* slope_dev, void* stream);
/* Same, and additionally writes y in the space-to-depth layout of ryolo_space_to_depth into xs (the operand of a
 * following 3x3/stride-2 block): the separate s2d pass over y disappears.  h, w even. */
int ryolo_bn_act_fwd_s2d(const void* z, int z_cstride, int batch, int h, int w, int c,
                         const float* scale, const float* shift, float slope, int has_act,
                         const void* residual, int res_cstride, void* y, int y_cstride,
                         const float* slope_dev, void* xs, int xs_cstride, void* stream);
/* slope_dev (both directions, optional): device scalar holding the PReLU slope (nn.PReLU.weight);
 * when non-null it replaces `slope`, so a training step never copies the parameter to the host.
 * Backward of the same block.  dy: gradient of y (at 2h x 2w when upsample2x == 1; upsample2x == 2: dy is read in the
 * space-to-depth layout [B, h/2+2, w/2+2, >=4c] a consuming 3x3/stride-2 block's dgrad wrote, no depth_to_space pass).  z_dz: in = z, out =
 * dz (gradient of the raw conv output, written in place).  sums (fp32 [2c+1], zeroed inside):
 * [0..c) = d(beta), [c..2c) = d(gamma), [2c] = d(slope).  gres (optional): gradient buffer of the
 * shortcut source, receives (+)= dy. */
int ryolo_bn_act_bwd(const void* dy, int dy_cstride, int upsample2x, void* z_dz, int z_cstride,
                     int batch, int h, int w, int c, const float* scale, const float* shift,
                     const float* mean, const float* invstd, float slope, int has_act, int has_bn,
                     float* sums, void* gres, int gres_cstride, int gres_accumulate,
                     const float* slope_dev, void* stream);
/* Adjoint of a stride-2 conv's pixel selection: src (h x w) -> even interior pixels of the
 * pre-zeroed dst (dst_h x dst_w). */
int ryolo_zero_insert2x(const void* src, int src_cstride, int batch, int h, int w, int c, void* dst,
                        int dst_cstride, int dst_h, int dst_w, void* stream);
/* Space-to-depth of a padded-NHWC tensor with even H, W:
 * xs[b, Y, X, (py*2+px)*C + c] = x[b, 2Y+py, 2X+px, c]  ([B, H/2+2, W/2+2, >=4C]).  A 3x3/stride-2/pad-1
 * conv on x equals a ksize=2 (offsets -1..0), stride-1 conv on xs with the weight blocks
 * W2[(qy,qx)][co][(py,px,c)] = W[co][c][kh][kw], kh = {(0,1):0,(1,0):1,(1,1):2}[(qy,py)] (zero for
 * (0,0)), same for kw.  ryolo_depth_to_space is the adjoint (optionally accumulating). */
int ryolo_space_to_depth(const void* x, int x_cstride, int batch, int h, int w, int c, void* xs,
                         int xs_cstride, void* stream);
int ryolo_depth_to_space(const void* dxs, int dxs_cstride, int batch, int h, int w, int c, void* gx,
                         int gx_cstride, int accumulate, void* stream);
/* Squeeze-and-excitation block (reference SELayer, model/models.py:16-31; `[se]` blocks of cfg/ICDAR/yolov3_608_se.cfg,
 * cfg/HRSC+/yolov3_512_se.cfg): x[b,:,:,c] *= sigmoid(W2 . relu(W1 . mean_hw x[b]))[c], in place on a padded-NHWC bf16
 * tensor.  w1 [reduced, c], w2 [c, reduced] fp32 (nn.Linear weights, no bias); sums_scratch / scale_scratch [batch*c]
 * fp32.  Eval path. */
int ryolo_se_block(void* x, int x_cstride, int batch, int h, int w, int c, const float* w1, const float* w2,
                   int reduced, float* sums_scratch, float* scale_scratch, void* stream);
/* 2x2 max pooling of a padded-NHWC bf16 tensor (reference: the maxpool blocks of
 * cfg/yolov3-tiny.cfg, model/models.py:79-87).  stride 2: [in_h/2, in_w/2] output; stride 1: same
 * size, the window past the right/bottom edge reads the zero halo exactly like the reference's
 * nn.ZeroPad2d((0,1,0,1)) + MaxPool2d(2, 1). */
int ryolo_maxpool2x2(const void* x, int x_cstride, int batch, int in_h, int in_w, int c, int stride,
                     void* y, int y_cstride, void* stream);
/* fp32 NCHW [B,C,H,W] -> bf16 padded NHWC interior, channels [0,C) (head gradients). */
int ryolo_nchw_to_padded(const float* src, int batch, int c, int h, int w, void* dst,
                         int dst_cstride, void* stream);
/* Gradient of a YOLO head in the layout autograd delivers it, fp32 [B, na, ny, nx, no] (training-mode
 * head output, reference model/models.py:190-192), -> bf16 padded NHWC [B, ny+2, nx+2, dst_cstride]
 * with channel = a*no + k (the head conv's filter index); channels >= na*no are left untouched. */
int ryolo_head_grad_to_padded(const float* g, int batch, int na, int no, int ny, int nx, void* dst,
                              int dst_cstride, void* stream);
/* Same, for the cotangent in the layout the head convolution WROTE its output in: fp32 NCHW [B, C = na*no, ny, nx]
 * (what arrives when the consumer differentiates the permuted VIEW of that buffer, e.g. the fused loss below).
 * bias_grad (nullable): [C] fp32, the per-channel sums of g are ADDED to it (the head convolution's bias gradient);
 * C <= 2048. */
int ryolo_head_grad_nchw_to_padded(const float* g, int batch, int c, int ny, int nx, void* dst,
                                   int dst_cstride, float* bias_grad, void* stream);
/* Objectness term of compute_loss (reference model/loss.py:340-348: BCEWithLogitsLoss(pos_weight)(pi[..., 5], tobj)
 * over every cell of a head map).  x: logical [batch, na, ny, nx, no] fp32 tensor addressed through `strides` (5 element
 * strides, so both the contiguous tensor and the permuted view of the NCHW head buffer are read in place); ch = the
 * objectness channel; tobj: contiguous [batch, na, ny, nx].  fwd ADDS the sum of the element losses to *sum_out
 * (float64, caller zeroes it and divides by the cell count); bwd writes the whole cotangent tensor `grad` (strides of x):
 * channel ch = *scale_dev * d(loss)/dx, every other channel 0. */
int ryolo_obj_bce_fwd(const float* x, const long long* strides, int batch, int na, int ny, int nx, int no, int ch,
                      const float* tobj, float pos_weight, double* sum_out, void* stream);
int ryolo_obj_bce_bwd(const float* x, const long long* strides, int batch, int na, int ny, int nx, int no, int ch,
                      const float* tobj, float pos_weight, const float* scale_dev, float* grad, void* stream);
/* The matched (anchor, target) rows of compute_loss (reference model/loss.py:300-338 indexes `pi[b, a, gj, gi]`): int64
 * device index vectors b / a / gj / gi and a 0/1 byte mask, `rows` entries each.  gather: out[r, k] = x[b, a, gj, gi, k]
 * (0 for masked-out or out-of-range rows); set_tobj: tobj[b, a, gj, gi] = 1 (tobj contiguous [batch, na, ny, nx]);
 * scatter_add: grad[b, a, gj, gi, k] += vals[r, k] * *scale_dev (atomic; duplicated cells accumulate like autograd's
 * index backward).  x / grad are addressed through 5 element strides like ryolo_obj_bce_*. */
int ryolo_loss_rows_gather(const float* x, const long long* strides, int batch, int na, int ny, int nx, int no,
                           const long long* b, const long long* a, const long long* gj, const long long* gi,
                           const unsigned char* mask, int rows, float* out, void* stream);
int ryolo_loss_rows_scatter_add(float* grad, const long long* strides, int batch, int na, int ny, int nx, int no,
                                const long long* b, const long long* a, const long long* gj, const long long* gi,
                                const unsigned char* mask, int rows, const float* vals, const float* scale_dev,
                                void* stream);
int ryolo_loss_rows_set_tobj(float* tobj, int batch, int na, int ny, int nx, const long long* b, const long long* a,
                             const long long* gj, const long long* gi, const unsigned char* mask, int rows, void* stream);
/* im2col of the 3-channel fp32 image for the first 3x3 conv: bf16 padded NHWC with 64 channels
 * (27 real, column = c*9 + kh*3 + kw), so that the first layer runs on the same GEMM kernels. */
int ryolo_im2col_first(const float* img, int batch, int h, int w, void* dst, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Parity-precision path (Darknet(..., precision="parity")): fp32-grade conv blocks on the bf16
 * tensor pipe, for the 1e-4 tolerance north_star states against the reference's fp32 nn.Conv2d
 * (model/models.py:55-65).  GEMM operands are SPLIT bf16 tensors: padded NHWC like above, but
 * every value v is two planes hi = bf16(v), lo = bf16(v - hi), `lo` channels apart in the same
 * row ("split view": pointer to channel 0 of plane 0, channel stride, plane offset `lo`, number of
 * planes: plane k = bf16 of the remainder left by planes < k, k*lo channels further; three planes
 * carry all 24 bits of an fp32).  Conv results, BN / PReLU / shortcut arithmetic and all gradients
 * of activations are fp32 padded NHWC; reductions are fp64.  a*w is evaluated as a sum of TERMS
 * a_i*w_j (a term table, 2 bits per term per operand: i + j <= 2 for three planes), exact products;
 * the TMEM accumulator is drained every 8 k-steps into round-to-nearest fp32 register sums (the tensor
 * pipe's own accumulator truncates).
 * ------------------------------------------------------------------------------------------ */
/* weight [cout,cin,k,k] fp32 -> bf16 [k*k][round_up(cout,64)][nterms*round_up(cin,64)]; term t of the
 * K range holds plane (w_plane_code >> 2t) & 3.  mode as ryolo_conv_pack_weights_ex
 * (ksize 2 / -2 for the space-to-depth forms; cout/cin are those of the PACKED operand). */
size_t ryolo_px_packed_weight_bytes(int cout, int cin, int ksize, int nterms);
int ryolo_px_pack_weights(const float* weight, int cout, int cin, int ksize, int nterms,
                          int w_plane_code, int mode, void* packed_out, void* stream);
/* out[p, co] (+)= sum over terms, taps, channels.  x_base: start of the split buffer row 0, the view
 * starts x_ch_off channels in; (a_plane_code >> 2t) & 3 is the activation plane of term t.  out: fp32
 * padded NHWC on the same grid (stride 1; stride-2 convs run on space-to-depth inputs), columns
 * [0, out_cols) written, accumulate != 0 adds to the existing contents (dgrad). */
int ryolo_px_conv(const void* x_base, int x_cstride, int x_ch_off, int plane_stride, int cin,
                  const void* packed_w, int nterms, int a_plane_code, int batch, int in_h, int in_w,
                  int ksize, int cout, float* out, int out_cstride, int out_cols, int accumulate,
                  void* stream);
/* wgrad of split operands = ryolo_conv_wgrad over all planes of dz and x ([taps][np*plane_out][np*plane_in]);
 * this sums the nplanes^2 plane blocks into nn.Conv2d layout (mode 0 / 2 as ryolo_conv_unpack_wgrad); cin_off:
 * first channel of x's view inside its plane (a conv reading a slice of a concat buffer). */
int ryolo_px_unpack_wgrad(const float* dw, int plane_out, int plane_in, int nplanes, int mode, int cout,
                          int cin, int ksize, int cin_off, float* grad, void* stream);
/* batch statistics of fp32 z: sums[0..c) = sum, [c..2c) = sum of squares (fp64, zeroed inside) */
int ryolo_px_bn_stats(const float* z, int z_cstride, int batch, int h, int w, int c, double* sums,
                      void* stream);
/* eval_mode 0: batch statistics from sums (running statistics updated); 1: running statistics. */
int ryolo_px_bn_finalize(const double* sums, int c, double count, float eps, float momentum,
                         const float* gamma, const float* beta, int eval_mode, float* mean,
                         float* invstd, float* scale, float* shift, float* running_mean,
                         float* running_var, void* stream);
/* y = prelu(z*scale + shift) [+ residual] -> split view (slope_dev NULL: linear) */
int ryolo_px_bn_act_fwd(const float* z, int z_cstride, int batch, int h, int w, int c,
                        const float* scale, const float* shift, const float* slope_dev,
                        const void* residual, int res_cstride, int res_lo, void* y, int y_cstride,
                        int y_lo, int upsample2x, int nplanes, void* stream);
/* backward: dy fp32 (2h x 2w when upsample2x), z fp32 -> sums fp64 [2c+1] (d beta, d gamma, d slope),
 * dz split view, shortcut gradient gres fp32 (+)= dy. */
int ryolo_px_bn_act_bwd(const float* dy, int dy_cstride, int upsample2x, const float* z, int z_cstride,
                        int batch, int h, int w, int c, const float* scale, const float* shift,
                        const float* mean, const float* invstd, const float* slope_dev,
                        int batch_stats, double* sums, void* dz, int dz_cstride, int dz_lo, float* gres,
                        int gres_cstride, int gres_accumulate, int nplanes, void* stream);
int ryolo_px_im2col_first(const float* img, int batch, int h, int w, void* dst, int dst_cstride,
                          int dst_lo, int nplanes, void* stream);
/* fp32 padded NHWC (+ bias[c]) -> fp32 NCHW [B,c,h,w] (the linear heads feeding ryolo_yolo_decode) */
int ryolo_px_to_nchw(const float* z, int z_cstride, int batch, int c, int h, int w, const float* bias,
                     float* out, void* stream);
int ryolo_px_head_grad(const float* g, int batch, int na, int no, int ny, int nx, void* dst,
                       int dst_cstride, int dst_lo, int nplanes, void* stream);
int ryolo_px_depth_to_space(const float* dxs, int dxs_cstride, int batch, int h, int w, int c, float* gx,
                            int gx_cstride, int accumulate, void* stream);
int ryolo_px_split_from_nchw(const float* src, int batch, int c, int h, int w, void* dst,
                             int dst_cstride, int dst_lo, int nplanes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RYOLO_H_ */
