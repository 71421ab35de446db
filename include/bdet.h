/*
 * bdet.h -- C ABI of libbdet.so: B200 (sm_100a) box-op hot path for BaseDet.
 *
 * The reference (megvii-research/basedet) has NO FFI boundary today: its box ops are
 * thin Python over megengine.functional.  Each entry point below names
 * the reference Python interface (file:line) whose arithmetic it replaces; the drop-in
 * Python layer in basedet_b200/{structures,layers} keeps those signatures and calls
 * these functions through ctypes (INTEGRATION.md shows the binding a maintainer adds).
 *
 * Conventions
 *   - every pointer is a RAW DEVICE pointer unless the parameter name ends in `_host`;
 *   - fp32 values, int32 indices / labels, row-major, contiguous unless an `ld` is given;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and never
 *     synchronise, allocate, or keep global state; the caller owns all buffers;
 *   - variable-length results return their count through a device int32*;
 *   - functions return BDET_OK or a negative BDET_E* code; bdet_last_error() gives the
 *     thread-local message;
 *   - there is no CPU fallback and no other backend: without an sm_100 device the
 *     launches fail and the error is reported.
 */
#ifndef BDET_H_
#define BDET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDET_ABI_VERSION 1

#define BDET_OK 0
#define BDET_EINVAL (-1)       /* bad argument (mirrors the reference's Python asserts) */
#define BDET_ECUDA (-2)        /* CUDA runtime / launch error */
#define BDET_EWORKSPACE (-3)   /* workspace too small */
#define BDET_EUNSUPPORTED (-4) /* size outside what the kernels cover */

#define BDET_MAX_LEVELS 8
#define BDET_MAX_MATCH_LABELS 8

typedef void* bdet_stream_t;

int bdet_abi_version(void);
const char* bdet_last_error(void);
/* SM count / compute capability of the current device (host query, no launch). */
int bdet_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);

/* ------------------------------------------------------------------ a1/a2: anchors
 * DefaultAnchorGenerator.generate_anchors_by_features  layers/common/anchor_generator.py:111-122
 * create_anchor_grid                                   layers/common/anchor_generator.py:23-30
 * All levels in one launch.  Level l writes H*W*n_base[l] boxes at out + 4*out_offset[l]
 * (offsets in boxes), order (h, w, base).  x = fp32(shift + w*stride) evaluated in fp64.
 * base_host: concatenated (sum n_base, 4) fp32 base anchors (host memory, <= 64 rows). */
int bdet_anchors_grid(float* out, int n_levels, const int* hw_host /*2*n_levels: H,W*/,
                      const double* stride_host, const double* shift_host, const int* n_base_host,
                      const float* base_host, const int64_t* out_offset_host, bdet_stream_t stream);

/* AnchorPointGenerator (mode 0)  layers/common/anchor_generator.py:152-165  (x,y) repeated num_anchors
 * FastPointGenerator   (mode 1)  layers/common/anchor_generator.py:175-182  (i*stride, j*stride), (w,h) mesh quirk kept */
int bdet_points_grid(float* out, int n_levels, const int* hw_host, const double* stride_host,
                     const double* shift_host, int num_anchors, int mode,
                     const int64_t* out_offset_host /*in points*/, bdet_stream_t stream);

/* ------------------------------------------------------------------ a3/a4: pairwise box ops
 * box_iou  structures/op_patch.py:33-97      (mode BDET_PAIR_IOU)
 * box_ioa  structures/op_patch.py:169-227    (mode BDET_PAIR_IOA)
 * Boxes.intersection structures/boxes.py:114-130 (BDET_PAIR_INTER); Boxes.giou :74-95 (BDET_PAIR_GIOU)
 * boxes1: (N, >=4) with row stride ld1 floats; boxes2: (M, >=4) with row stride ld2 floats.
 * out: (N, M) with row stride ldo >= M floats.  When ldo is a multiple of 4 (rows padded to 16 bytes) and out is
 * 16-byte aligned the kernel writes 128-bit stores (up to 3 padding floats per row may be written).  Batched form: batch b reads boxes1 + b*bs1, boxes2 + b*bs2 (bs2 = 0 shares
 * boxes2), writes out + b*bs_out, and uses n1_dev[b] rows when n1_dev != NULL (rows >= n1_dev[b] untouched). */
#define BDET_PAIR_IOU 0
#define BDET_PAIR_IOA 1
#define BDET_PAIR_INTER 2
#define BDET_PAIR_GIOU 3
int bdet_pairwise(const float* boxes1, int ld1, int N, const float* boxes2, int ld2, int M, float* out,
                  int64_t ldo, int mode, bdet_stream_t stream);
int bdet_pairwise_batched(const float* boxes1, int ld1, int64_t bs1, const int* n1_dev, int N,
                          const float* boxes2, int ld2, int64_t bs2, int M, float* out, int64_t ldo,
                          int64_t bs_out, int B, int mode, bdet_stream_t stream);
/* box_center structures/op_patch.py:100-130: (N,4) -> (N,2);  point_distance :133-166: (N,2)x(M,2) -> (N,M) */
int bdet_box_center(const float* boxes, int ld, int N, float* out, bdet_stream_t stream);
int bdet_point_distance(const float* p1, int N, const float* p2, int M, float* out, bdet_stream_t stream);

/* ------------------------------------------------------------------ a5: Matcher
 * Matcher.__call__  layers/common/matcher.py:31-51
 * matrix (G, A) fp32 with row stride ld >= A [batched: (B, Gmax, A) with g_dev[b] valid rows, NULL -> Gmax].
 * thresholds_host: n_labels-1 user thresholds (the +-inf the constructor adds are implied);
 * labels_host: n_labels ints.  Outputs match_idx (B,A) int32 = first argmax over G, labels (B,A) int32.
 * Single pass over the matrix + a fix-up that re-reads only the row segments holding a row maximum. */
size_t bdet_match_workspace(int Gmax, int A, int B);
int bdet_match(const float* matrix, int64_t ld, int64_t batch_stride, const int* g_dev, int Gmax, int A, int B,
               const float* thresholds_host, const int* labels_host, int n_labels, int allow_low_quality,
               int* match_idx, int* labels, void* workspace, size_t workspace_bytes, bdet_stream_t stream);
/* (R, G) layout used by RCNN.get_ground_truth  layers/head/rcnn.py:113-116: max / first argmax over axis 1.
 * One warp per row, warp-shuffle argmax. */
int bdet_match_rows(const float* matrix, int R, int G, float* max_out, int* argmax_out, bdet_stream_t stream);

/* ------------------------------------------------------------------ a7-a9: coders
 * BoxCoder.encode structures/boxcoder.py:61-73; BoxCoder.decode :75-98
 * encode: t = ((delta(bbox, gt)) - mean) / std.  If gather_idx != NULL, gt row = gt[gather_idx[i]] with
 * row stride gt_ld floats (fuses `gt_boxes[match_indices]`, models/det/retinanet.py:220-224). */
int bdet_box_encode(const float* bbox, const float* gt, int gt_ld, const int* gather_idx, int N,
                    const float* mean_host, const float* std_host, float* out, bdet_stream_t stream);
/* decode: anchors (N,4), deltas (N,4k) -> out (N,4k).  writeback != 0 stores deltas*std+mean back into
 * `deltas` like the reference's in-place update (boxcoder.py:76-77).
 * If sel_idx != NULL (n_sel entries): out (n_sel,4) = decode(anchors[sel_idx[i]/sel_div], deltas[...]) --
 * the decode-then-gather of models/det/retinanet.py:194-196 without decoding the other 99 %. */
int bdet_box_decode(const float* anchors, float* deltas, int N, int k, const float* mean_host,
                    const float* std_host, float* out, int writeback, const int* sel_idx, int n_sel,
                    int sel_div, bdet_stream_t stream);
/* SumBoxCoder structures/boxcoder.py:115-127 */
int bdet_sum_encode(const float* anchors, const float* gt, int N, const float* mean_host,
                    const float* std_host, float* out, bdet_stream_t stream);
int bdet_sum_decode(const float* anchors, float* deltas, int N, const float* mean_host,
                    const float* std_host, float* out, int writeback, bdet_stream_t stream);
/* PointCoder structures/boxcoder.py:132-141.  encode: points (A,2), gt (G,>=4, ld) -> (G,A,4);
 * decode: points (N,2), deltas (N,4k) -> (N,4k); sel_idx as in bdet_box_decode. */
int bdet_point_encode(const float* points, int A, const float* gt, int gt_ld, int G, float* out,
                      bdet_stream_t stream);
/* row-wise form used after matching (models/det/fcos.py:268): points (N,2), gt (N,>=4, ld) -> (N,4) */
int bdet_point_encode_rows(const float* points, const float* gt, int gt_ld, int N, float* out, bdet_stream_t stream);
int bdet_point_decode(const float* points, const float* deltas, int N, int k, float* out,
                      const int* sel_idx, int n_sel, int sel_div, bdet_stream_t stream);

/* ------------------------------------------------------------------ fused target assignment
 * RetinaNet.get_ground_truth  models/det/retinanet.py:211-232 (and RPN.get_ground_truth rpn.py:215-226
 * up to, not including, sample_labels): IoU (G,A) -> Matcher -> labels[fg]=class -> BoxCoder.encode,
 * for all B images, WITHOUT materialising the (G,A) matrix.  Results are bit-identical to the
 * bdet_pairwise + bdet_match + bdet_box_encode sequence.
 * anchors (A,4) shared; gt (B,Gmax,5) rows [x1,y1,x2,y2,class]; num_gt_dev (B) int32.
 * apply_class != 0: labels==1 become int32(gt class) (retinanet.py:222-223); 0: raw matcher labels (RPN).
 * Outputs: labels (B,A) int32, match_idx (B,A) int32, offsets (B,A,4) fp32.
 * Images with num_gt == 0 (the reference raises there): labels = label of IoU 0, idx 0, offsets 0. */
size_t bdet_assign_targets_workspace(int Gmax, int A, int B);
int bdet_assign_targets(const float* anchors, int A, const float* gt, int Gmax, const int* num_gt_dev, int B,
                        const float* thresholds_host, const int* labels_host, int n_labels,
                        int allow_low_quality, int apply_class, const float* mean_host,
                        const float* std_host, int* labels, int* match_idx, float* offsets,
                        void* workspace, size_t workspace_bytes, bdet_stream_t stream);
/* The same with the anchors generated in registers from the grid description of bdet_anchors_grid (bit-identical
 * anchors; level l holds H*W*n_base[l] anchors in (h, w, base) order, levels concatenated): no anchor tensor, no separate
 * launch -- models/det/retinanet.py:116 regenerates the anchors every forward.  counts (B,3) int32, optional: the census
 * of the final labels (< 0, == 0, > 0) that retinanet.py:142-146 / rpn.py:231 needs, accumulated by the same kernels. */
int bdet_assign_targets_grid(int n_levels, const int* hw_host, const double* stride_host, const double* shift_host,
                             const int* n_base_host, const float* base_host, const float* gt, int Gmax,
                             const int* num_gt_dev, int B, const float* thresholds_host, const int* labels_host,
                             int n_labels, int allow_low_quality, int apply_class, const float* mean_host,
                             const float* std_host, int* labels, int* match_idx, float* offsets, int* counts,
                             void* workspace, size_t workspace_bytes, bdet_stream_t stream);

/* ------------------------------------------------------------------ 8(f)-1: anchor-free dense-head target assignment
 * points (A,2) = all levels concatenated (level l owns [level_start[l], level_start[l+1])), gt (B,Gmax,5) rows
 * [x1,y1,x2,y2,class], num_gt (B).  Outputs for all B images: labels (B,A) int32 (0 = background), offsets (B,A,4) =
 * PointCoder.encode(point, matched GT) (GT 0 for background points, as the reference's argmin/argmax of an empty
 * selection gives index 0), ctrness (B,A), match_idx (B,A) optional.  No (G,A) tensor is materialised.
 * FCOS.get_ground_truth  models/det/fcos.py:222-293: a GT owns a point when max(l,t,r,b) lies in
 *   [size_lo[l], size_hi[l]] of the point's level (:245-249) and the point is inside the GT's centre box
 *   [max(c - radius[l], tl), min(c + radius[l], br)] (:251-264; all radius 0 / NULL: inside the GT itself, :266);
 *   among owners the smallest area wins, first index on ties (:268-272).  radius[l] = stride_l * CENTER_SAMPLING_RADIUS.
 * ATSS.get_ground_truth  models/det/atss.py:17-86: per GT and level the topk points nearest to the GT centre (:39-44),
 *   IoU of the GT with their square anchors point +- half_size[l] (= stride_l * SCALE / 2, :31-37), threshold =
 *   mean + std of those IoUs (:50-51, sequential fp32 sums), positives = candidates with IoU >= threshold that lie inside
 *   the GT (:52-61); a point takes the GT of highest IoU, first index on ties (:63).
 * Images with num_gt == 0 (the reference raises there): labels 0, offsets 0, ctrness 0. */
int bdet_fcos_targets(const float* points, int A, const int* level_start_host, int L, const float* radius_host,
                      const float* size_lo_host, const float* size_hi_host, const float* gt, int Gmax,
                      const int* num_gt_dev, int B, int* labels, float* offsets, float* ctrness, int* match_idx,
                      bdet_stream_t stream);
size_t bdet_atss_targets_workspace(int A, int B);
int bdet_atss_targets(const float* points, int A, const int* level_start_host, int L, const float* half_size_host,
                      int topk, const float* gt, int Gmax, const int* num_gt_dev, int B, int* labels, float* offsets,
                      float* ctrness, int* match_idx, void* workspace, size_t workspace_bytes, bdet_stream_t stream);

/* ------------------------------------------------------------------ 8(f)-2: fg / bg subsampling
 * sample_labels  layers/common/sampling.py:7-30 (used by models/det/rpn.py:228-232 and layers/head/rcnn.py:124-127).
 * labels (B,A) int32 IN PLACE, per image: if more than num_samples elements equal label_value, the surplus -- those
 * with the LARGEST variates -- is set to ignore_label.  RNG contract: noise (B,A) holds one uniform [0,1) variate per
 * element, drawn by the caller's framework; the reference's uniform(size=num_valid) is the variates of the selected
 * positions in index order, so feeding it those reproduces the reference decision for decision (ties: lower index
 * goes first).  num_samples_dev (B) optional device array overrides num_samples per image (rpn.py:231: the number of
 * negatives depends on the positives that survived). */
int bdet_sample_labels(int* labels, const float* noise, int A, int B, int label_value, int ignore_label,
                       int num_samples, const int* num_samples_dev, bdet_stream_t stream);
/* RCNN.get_ground_truth (training)  layers/head/rcnn.py:95-147, in three steps that never build the (R,G) matrix:
 * bdet_rcnn_match: rois (B,Rmax,5) rows [batch,x1,y1,x2,y2] (image b's first n_rois[b] rows; what
 *   rpn_rois[rpn_rois[:,0]==b] selects) + gt (B,Gmax,5) -> all_rois (B,N,5) = proposals then [b, gt box] rows (:108-112),
 *   N = Rmax+Gmax, n_all (B), assign (B,N) = argmax IoU over the image's GT (first index), matched_class (B,N) fp32,
 *   fg_mask / bg_mask (B,N) int32 0/1: (max >= fg_thresh) & (class >= 0), bg_low <= max < bg_high (:114-123).
 * then bdet_sample_labels(fg_mask, noise_fg, N, B, 1, 0, int(num_rois*fg_ratio)) and, with the per-image budget
 *   num_rois - sum(fg_mask) in a device array, bdet_sample_labels(bg_mask, ...) (:125-128);
 * bdet_rcnn_collect: kept rows (fg|bg) in order -> out_rois (B,num_out,5), out_labels (B,num_out) (class, 0 for bg),
 *   out_targets (B,num_out,4) = BoxCoder.encode(roi, matched GT) with the RCNN mean/std, out_count (B); rows past the
 *   count are zero (:130-137).  The reference concatenates the per-image results; slice with out_count to do the same. */
int bdet_rcnn_match(const float* rois, const int* n_rois_dev, int Rmax, const float* gt, const int* num_gt_dev,
                    int Gmax, int B, float fg_thresh, float bg_thresh_low, float bg_thresh_high, float* all_rois,
                    int* n_all, int* assign, float* matched_class, int* fg_mask, int* bg_mask, bdet_stream_t stream);
int bdet_rcnn_collect(const float* all_rois, const int* n_all, const int* assign, const float* matched_class,
                      const int* fg_mask, const int* bg_mask, int N, const float* gt, int Gmax, int B,
                      const float* mean_host, const float* std_host, int num_out, float* out_rois, int* out_labels,
                      float* out_targets, int* out_count, bdet_stream_t stream);

/* ------------------------------------------------------------------ 8(f)-3: dynamic-k matching (OTA / YOLOX)
 * OTATopkMatcher.__call__  layers/common/matcher.py:134-161.  cost, ious: (G,A) fp32 with row strides ldc / ldi (the
 * matrices models/det/ota.py:76-180 builds from its losses).  Per GT: k = clip(int32(sum of its candidate_k largest
 * IoUs), 1) anchors of smallest cost are matched; an anchor matched by several GTs goes to the GT of smallest cost
 * over all GTs; matched_gt (A) int32 = GT index, G for background.  candidate_k <= 16. */
size_t bdet_ota_topk_match_workspace(int A);
int bdet_ota_topk_match(const float* cost, int ldc, const float* ious, int ldi, int G, int A, int candidate_k,
                        int* matched_gt, void* workspace, size_t workspace_bytes, bdet_stream_t stream);

/* OTA cost construction, models/det/ota.py:91-152 with layers/losses/sigmoid_focal_loss.py:30-36 and iou_loss.py:9-43,96:
 * per (GT g, point a)  cost = (sum_c focal(logit[a,c], onehot_g) + reg_weight * -log(max(iou_ltrb(pred[a], enc(a, g)), eps)))
 * + 1e6 * !(a inside box g and inside its centre box of radius[a]),  ious = that ltrb IoU.  points (A,2); radius (A) =
 * stride * 2.5 of the point's level; gt (G,5) classes 1..C; cls_logits (A,C); pred_deltas (A,4) ltrb; cost, ious (G,A).
 * Floating point through logf / expf: tolerance-gated (1e-5), feed bdet_ota_topk_match; bdet_ota_collect then writes the
 * targets of ota.py:160-175: gt_classes (A) fp32 (0 = background), box_targets (A,4), iou_targets (A). */
size_t bdet_ota_cost_workspace(int A);
int bdet_ota_cost(const float* points, const float* radius, int A, const float* gt, int G, const float* cls_logits,
                  int num_classes, const float* pred_deltas, double alpha, double gamma, double reg_weight, float* cost,
                  float* ious, void* workspace, size_t workspace_bytes, bdet_stream_t stream);
int bdet_ota_collect(const int* matched_gt, const float* points, int A, const float* gt, int G, const float* ious,
                     float* gt_classes, float* box_targets, float* iou_targets, bdet_stream_t stream);

/* FreeAnchor box ops, models/det/free_anchor.py:48-113 (one image).
 * bdet_free_anchor_box_prob (:54-86): box_prob (A,C) = scatter of clip((IoU(gt g, pred a) - t1) / (t2[g] - t1), 0, 1) to
 * [a, class(g) - 1] over the non-zero entries in ascending (g, a) order (the last GT of a class wins), t2[g] =
 * clip(max_a IoU, thresh2_lower, 1), thresh2_lower = fp32(t1 + clamp_eps) summed in double by the caller; includes the
 * reference's empty-set workaround (:71-73, :85-86).  pred_boxes (A,4) = BoxCoder.decode(anchors, pred_offsets).
 * bdet_free_anchor_bags (:95-113): for matched_idx (G,K) (the per-GT top-K anchors by IoU: bdet_pairwise + bdet_topk)
 * matched_score (G,K) = pred_scores[idx, class(g) - 1] and matched_offsets (G*K,4) = BoxCoder.encode(anchors[idx], gt g). */
size_t bdet_free_anchor_box_prob_workspace(int G);
int bdet_free_anchor_box_prob(const float* pred_boxes, int A, const float* gt, int G, int num_classes,
                              float box_iou_thresh, float thresh2_lower, float clamp_eps, float* box_prob,
                              void* workspace, size_t workspace_bytes, bdet_stream_t stream);
int bdet_free_anchor_bags(const int* matched_idx, int G, int K, const float* anchors, const float* gt,
                          const float* pred_scores, int num_classes, const float* mean_host, const float* std_host,
                          float* matched_score, float* matched_offsets, bdet_stream_t stream);

/* Dense tail: bdet_select_decode -> bdet_nms_runs -> bdet_finalize_detections of a dense head in ONE kernel (one CTA per
 * image; RetinaNet / FCOS inference, retinanet.py:193-209, fcos.py:204-221, post_processing.py:17-47,78-103).  Arguments
 * as bdet_select_decode_ws (label = idx % div, no size filter) plus the NMS threshold, max_out and im_info rows
 * [h, w, orig_h, orig_w] for the final scale / clip; dets (B, max_out, 6) rows [x1,y1,x2,y2,score,label] zero padded,
 * det_count (B).  Needs bdet_dense_tail_smem(L, k, max_out) <= 200 KB (5 levels x 1000 candidates: 142 KB); identical
 * detections to the separate kernels. */
size_t bdet_dense_tail_smem(int L, int k, int max_out);
int bdet_dense_tail(const float* const* anchors_host, const float* const* deltas_host, const int* n_l_host, const int* hw_host,
                    int L, int B, int k, int div, int coder, const int* topk_idx, const float* topk_val, const int* topk_cnt,
                    const float* mean_host, const float* std_host, const float* im_info, int info_ld, float iou_thresh,
                    int max_out, float* dets, int* det_count, bdet_stream_t stream);

/* ------------------------------------------------------------------ 8(f)-4: COCO result records on the device
 * COCOEvaluator.format  evaluators/coco_eval.py:111-138.  dets (B,K,6) rows [x1,y1,x2,y2,score,label] with counts (B)
 * valid rows (bdet_finalize_detections' layout) -> compact records in image order: image id, bbox [x, y, w, h] and score
 * as float64 (the reference converts to float64 before the subtraction), category id = category_ids[label]
 * (classes_originID; -1 if the label is out of range) or label + 1 when category_ids is NULL; *total = record count. */
int bdet_coco_format(const float* dets, const int* counts, int B, int K, const int* image_ids, const int* category_ids,
                     int num_classes, int* rec_image_id, double* rec_bbox_xywh, double* rec_score, int* rec_category_id,
                     int* total, bdet_stream_t stream);

/* ------------------------------------------------------------------ a10: score filter + top-k
 * F.topk(scores, k, descending=True) as used in models/det/rpn.py:155 and retinanet.py:189-190.
 * Segmented: segment s covers scores[seg_start[s] .. seg_start[s] + seg_len[s]) (element offsets from `scores`;
 * segments may lie in different allocations).  For each segment writes min(k, n_s) (value, index-within-segment)
 * pairs sorted by (value desc, index asc) at out_* + s*k, and the count at out_count[s].
 * Radix select on unique (value, index) keys + shared-memory bitonic sort; k <= 16384. */
size_t bdet_topk_workspace(int64_t total, int n_seg, int k);
int bdet_topk(const float* scores, const int64_t* seg_start_host, const int64_t* seg_len_host, int n_seg, int k,
              float* out_vals, int* out_idx, int* out_count, void* workspace, size_t workspace_bytes,
              bdet_stream_t stream);
/* sigmoid -> score > thr -> top-k of models/det/retinanet.py:181-191 (mode BDET_SCORE_SIGMOID) and
 * sqrt(sigmoid(cls)*sigmoid(ctr)) of models/det/fcos.py:194-202 (mode BDET_SCORE_FCOS; ctrness has one value per C
 * logits: element e of segment s uses ctrness[ctr_start[s] + e / C]; ctr_start_host NULL -> seg_start / C).
 * index = flat index within the segment (label = idx % C, box = idx / C).  BDET_SCORE_RAW treats logits as scores. */
#define BDET_SCORE_RAW 0
#define BDET_SCORE_SIGMOID 1
#define BDET_SCORE_FCOS 2
size_t bdet_score_filter_topk_workspace(int64_t total, int n_seg, int k);
int bdet_score_filter_topk(const float* logits, const float* ctrness, int C, const int64_t* seg_start_host,
                           const int64_t* seg_len_host, const int64_t* ctr_start_host, int n_seg, float threshold,
                           int k, int mode, float* out_scores, int* out_idx, int* out_count, void* workspace,
                           size_t workspace_bytes, bdet_stream_t stream);
/* Same, reading the head outputs as the network writes them: segment s is a (num_anchors*C, H, W)
 * block starting at logits + seg_start[s] with seg_hw[s] = H*W; centerness (FCOS) a (num_anchors, H, W) block at
 * ctrness + ctr_start[s].  The reported indices are those of the reference's permuted layout
 * (permute_to_N_Any_K, layers/common/function.py:26-32: idx = ((h*W + w)*A + a)*C + c), so results are identical to
 * permuting first -- without the transpose pass. */
int bdet_score_filter_topk_nchw(const float* logits, const float* ctrness, int C, int num_anchors,
                                const int64_t* seg_start_host, const int* seg_hw_host, const int64_t* ctr_start_host,
                                int n_seg, float threshold, int k, int mode, float* out_scores, int* out_idx,
                                int* out_count, void* workspace, size_t workspace_bytes, bdet_stream_t stream);
/* F.sigmoid / fcos score as a plain elementwise op (so tests can feed bit-identical scores to the oracle). */
int bdet_scores(const float* logits, const float* ctrness, int C, int64_t n, int mode, float* out,
                bdet_stream_t stream);

/* ------------------------------------------------------------------ batched glue: top-k -> boxes -> NMS -> detections
 * bdet_select_decode: for image b, level l (in order) and each of its topk_cnt[b,l] selected flat indices idx:
 *   box   = decode(anchors_l[idx / div], deltas_l[b, idx / div])     (coder 0 BoxCoder boxcoder.py:75-98, 1 PointCoder :135-141)
 *   label = idx % div (label_mode 0, retinanet.py:194) or the level id as fp32 (label_mode 1, rpn.py:160)
 *   im_info != NULL: drop candidates whose box, clipped to im_info[b, :2] = (h, w), has h <= 0 or w <= 0 (rpn.py:168-171)
 * and writes them compacted, levels concatenated in order (post_processing.py:63-67 / rpn.py:163-165):
 *   boxes (B, L*k, 4), scores (B, L*k), labels (B, L*k), count (B).  anchors_host / deltas_host: L device pointers to
 *   (n_l, 4|2) and (B, n_l, 4).  topk_* as written by bdet_topk / bdet_score_filter_topk with segment s = b*L + l.
 *   run_end (B, L) optional: exclusive end of level l's candidates in image b's list (input of bdet_nms_runs). */
int bdet_select_decode(const float* const* anchors_host, const float* const* deltas_host, const int* n_l_host, int L,
                       int B, int k, int div, int coder, int label_mode, const int* topk_idx, const float* topk_val,
                       const int* topk_cnt, const float* mean_host, const float* std_host, const float* im_info,
                       int info_ld, float* boxes, float* scores, void* labels, int* count, int* run_end,
                       bdet_stream_t stream);
/* hw_host[l] > 0: deltas of level l are the head output (B, A*4, H, W) with H*W = hw_host[l] (n_l = H*W*A). */
int bdet_select_decode_nchw(const float* const* anchors_host, const float* const* deltas_host, const int* n_l_host,
                            const int* hw_host, int L, int B, int k, int div, int coder, int label_mode,
                            const int* topk_idx, const float* topk_val, const int* topk_cnt, const float* mean_host,
                            const float* std_host, const float* im_info, int info_ld, float* boxes, float* scores,
                            void* labels, int* count, int* run_end, bdet_stream_t stream);
/* Same with a scratch buffer: with the size filter (im_info != NULL) the ordered compaction then runs as two fully
 * parallel launches instead of one CTA per image.  workspace NULL = bdet_select_decode_nchw. */
size_t bdet_select_decode_workspace(int L, int B, int k);
int bdet_select_decode_ws(const float* const* anchors_host, const float* const* deltas_host, const int* n_l_host,
                          const int* hw_host, int L, int B, int k, int div, int coder, int label_mode,
                          const int* topk_idx, const float* topk_val, const int* topk_cnt, const float* mean_host,
                          const float* std_host, const float* im_info, int info_ld, float* boxes, float* scores,
                          void* labels, int* count, int* run_end, void* workspace, size_t workspace_bytes,
                          bdet_stream_t stream);
/* mode 0: post_processing.py:96-101 -- out (B, max_out, 6) = [box scaled by (orig/resized) and clipped to the original
 *         image, score, label] for the kept indices (im_info (B, >=4) [h, w, orig_h, orig_w]; NULL = no scale/clip);
 * mode 1: rpn.py:179-183 -- out (B, max_out, 5) = [batch index, x1, y1, x2, y2].  Rows >= keep_count[b] are zero. */
int bdet_finalize_detections(const float* boxes, const float* scores, const void* labels, int labels_is_float, int N,
                             const int* keep, int keep_ld, const int* keep_count, const float* im_info, int info_ld,
                             int B, int max_out, int mode, float* out, bdet_stream_t stream);

/* ------------------------------------------------------------------ a11: class-aware batched NMS
 * batched_nms  layers/common/post_processing.py:17-47  ->  F.vision.nms
 * boxes (B,Nmax,4), scores (B,Nmax), idxs (B,Nmax) int32 or fp32 (idxs_is_float; NULL = single class),
 * n_dev (B) valid counts (NULL -> Nmax).  The class offset `idxs*(max(boxes)+1)` is applied in fp32 exactly
 * as the reference does.  keep (B, max_out_cap) int32 = ORIGINAL indices of kept boxes in score-descending
 * order; keep_count (B).  max_output <= 0 means unlimited (cap = Nmax). */
size_t bdet_nms_workspace(int Nmax, int B);
int bdet_nms(const float* boxes, const float* scores, const void* idxs, int idxs_is_float, const int* n_dev,
             int Nmax, int B, float iou_thresh, int max_output, int* keep, int keep_ld, int* keep_count,
             void* workspace, size_t workspace_bytes, bdet_stream_t stream);
/* Same, for inputs that are already n_runs back-to-back runs per image, each in (score desc, index asc) order -- the
 * level concat of per-level top-k results (post_processing.py:63-67, rpn.py:163-165).  run_end (B, n_runs) int32:
 * exclusive end of every run, run_end[b][n_runs-1] == n_dev[b].  The argsort becomes a merge by ranking; the promise
 * is checked on the device and an unsorted run silently takes the general sort, so results never differ from bdet_nms. */
int bdet_nms_runs(const float* boxes, const float* scores, const void* idxs, int idxs_is_float, const int* n_dev,
                  const int* run_end, int n_runs, int Nmax, int B, float iou_thresh, int max_output, int* keep,
                  int keep_ld, int* keep_count, void* workspace, size_t workspace_bytes, bdet_stream_t stream);

/* ------------------------------------------------------------------ a12: Boxes.scale / clip / filter_by_size
 * structures/boxes.py:193-212, :152-177, :132-150.  boxes (N,4) in place: x*=sw, y*=sh then clip to
 * [0,clip_w]x[0,clip_h] (clip skipped when clip_w < 0).  keep_mask (N) uint8 optional: (w>0)&(h>0) style. */
int bdet_boxes_scale_clip(float* boxes, int N, float scale_w, float scale_h, float clip_w, float clip_h,
                          bdet_stream_t stream);
int bdet_boxes_filter_by_size(const float* boxes, int N, float size0, float size1, uint8_t* keep_mask,
                              bdet_stream_t stream);

/* ------------------------------------------------------------------ a13/a14: ROI pooling
 * assign_rois  layers/common/roi_pool.py:12-25: level = clamp(floor(4 + log(sqrt(area)/224)/ln2), lo, hi) - lo */
int bdet_roi_assign_levels(const float* rois, int K, int min_level, int max_level, int* levels,
                           bdet_stream_t stream);
/* roi_pool -> F.nn.roi_align(mode="average", aligned=True)  layers/common/roi_pool.py:35-78
 * All FPN levels in one launch, output in ORIGINAL roi order (the reference's dummy-roi / argsort glue
 * :28-31,:74-76 is not needed).  feats_host: n_levels device pointers to (B,C,H_l,W_l) fp32 NCHW;
 * hw_host: H_l,W_l; scale_host: spatial scale (1/stride) per level; rois (K,5) [batch,x1,y1,x2,y2];
 * levels (K) int32 (NULL -> all level 0).  out (K,C,PH,PW). */
int bdet_roi_align_fwd(const float* const* feats_host, int n_levels, const int* hw_host,
                       const float* scale_host, int B, int C, const float* rois, const int* levels, int K,
                       int PH, int PW, int sample_h, int sample_w, int aligned, float* out,
                       bdet_stream_t stream);
/* Backward w.r.t. the features (rois carry no gradient, roi_pool.py:56).
 * accumulate == 0: dfeats are OVERWRITTEN with the gradient (zeros where no ROI reaches); != 0: added to.
 * With a workspace (bdet_roi_align_bwd_workspace bytes) the gather form runs: ROIs are binned per 16x32-pixel tile,
 * one CTA accumulates its tile x 32-channel chunk in shared memory and writes every element exactly once -- no global
 * atomics, no memset, deterministic.  workspace == NULL selects the scatter form (shared-memory footprint
 * accumulation + red.global.add), which needs no workspace. */
size_t bdet_roi_align_bwd_workspace(int n_levels, const int* hw_host, int B, int K);
int bdet_roi_align_bwd(float* const* dfeats_host, int n_levels, const int* hw_host, const float* scale_host,
                       int B, int C, const float* rois, const int* levels, int K, int PH, int PW,
                       int sample_h, int sample_w, int aligned, const float* dout, int accumulate,
                       void* workspace, size_t workspace_bytes, bdet_stream_t stream);

/* Processing order for the ROI kernels: perm (K) int32 = the rois sorted by (image, level, 32 x 32-pixel tile of the roi
 * centre).  A roi touches its footprint in all C planes of its level, so rois that run at the same time should be
 * neighbours: with the permutation the planes stay in L2 instead of being re-read from HBM (K <= 16 384; one CTA per
 * image).
 * bdet_roi_align_fwd_perm / _bwd_perm = the calls above with CTA i working on roi perm[i] (perm NULL = identity); results
 * are in the original roi order either way (the backward's fp32 reductions in a different order: same 1e-5 gate). */
int bdet_roi_order(int n_levels, const int* hw_host, const float* scale_host, int B, const float* rois, const int* levels,
                   int K, int PH, int PW, int aligned, int* perm, bdet_stream_t stream);
int bdet_roi_align_fwd_perm(const float* const* feats_host, int n_levels, const int* hw_host, const float* scale_host,
                            int B, int C, const float* rois, const int* levels, int K, int PH, int PW, int sample_h,
                            int sample_w, int aligned, float* out, const int* perm, bdet_stream_t stream);
int bdet_roi_align_bwd_perm(float* const* dfeats_host, int n_levels, const int* hw_host, const float* scale_host, int B,
                            int C, const float* rois, const int* levels, int K, int PH, int PW, int sample_h, int sample_w,
                            int aligned, const float* dout, int accumulate, const int* perm, void* workspace,
                            size_t workspace_bytes, bdet_stream_t stream);

/* Max ROI pooling: roi_pool(..., pooler_type="roi_pool"), layers/common/roi_pool.py:62-63 -> F.nn.roi_pooling(mode="max").
 * Caffe / MegDNN rule (pinned by the reference's test tests/layers/test_roi_pool.py:48-61): corners rounded to pixels,
 * size = end - start + 1 (>= 1), bin [floor(p*size/P), ceil((p+1)*size/P)) clipped to the map, 0 for an empty bin.
 * argmax (K,C,PH,PW) int32 (optional in the forward) = flat y*W+x index of the maximum, -1 for an empty bin; the backward
 * adds dout to dfeat at it (fp32 atomics; dfeats zeroed first unless accumulate). */
int bdet_roi_maxpool_fwd(const float* const* feats_host, int n_levels, const int* hw_host, const float* scale_host, int B,
                         int C, const float* rois, const int* levels, int K, int PH, int PW, float* out, int* argmax,
                         bdet_stream_t stream);
int bdet_roi_maxpool_bwd(float* const* dfeats_host, int n_levels, const int* hw_host, int B, int C, const float* rois,
                         const int* levels, int K, int PH, int PW, const float* dout, const int* argmax, int accumulate,
                         bdet_stream_t stream);

/* ------------------------------------------------------------------ small Boxes / glue ops
 * Boxes.width / height / area  structures/boxes.py:36-52 (mode 0 / 1 / 2) -> out (N) */
int bdet_box_props(const float* boxes, int ld, int N, int mode, float* out, bdet_stream_t stream);
/* BoxConverter.convert  structures/box_convert.py:51-82; modes BoxMode XYXY=0, XYWH=1, XcYcWH=2; (N,4) -> (N,4) */
int bdet_box_convert(const float* boxes, int N, int from_mode, int to_mode, float* out, bdet_stream_t stream);
/* non_zeros -> F.cond_take(mask != 0, x)  layers/common/function.py:19-23: ordered stream compaction.
 * x (n) fp32; mask (n) uint8 or NULL (then the mask is x != 0).  out_vals / out_idx (n) receive the selected
 * values and their ascending flat indices; count_dev (1) the number selected. */
size_t bdet_cond_take_workspace(int64_t n);
int bdet_cond_take(const float* x, const uint8_t* mask, int64_t n, float* out_vals, int* out_idx, int* count_dev,
                   void* workspace, size_t workspace_bytes, bdet_stream_t stream);
/* Per-image label census: counts (B,3) = #(label < 0), #(label == 0), #(label > 0); the `num_fg` normaliser of
 * models/det/retinanet.py:142-146 and the sampling counts of models/det/rpn.py:232-236. */
int bdet_count_labels(const int* labels, int A, int B, int* counts, bdet_stream_t stream);

/* ------------------------------------------------------------------ measurement hooks (bench.py only)
 * While profiling is on, every named kernel launch inside the library is bracketed by a CUDA event pair on the
 * launching stream.  bdet_profile_collect synchronises those events and returns the summed duration and the
 * number of launches whose kernel name matches `name` (NULL = all).  bdet_profile_select(name) restricts the event
 * pairs to one kernel (the others are still counted) so that the measurement does not stretch the step it measures;
 * NULL brackets every kernel again.  Off by default; thread-local. */
/* Streaming probe: mode 0 write-only (st.v4), 1 read-only (ld.v4), 2 copy, 3 cudaMemsetAsync -- the device's
 * write / read / copy HBM ceilings that the kernels' achieved GB/s are put next to (profiles/). */
int bdet_bw_probe(void* dst, const void* src, size_t bytes, int mode, int ctas_per_sm, bdet_stream_t stream);
int bdet_profile_begin(void);
int bdet_profile_select(const char* name);
int bdet_profile_collect(const char* name, float* total_ms_host, int* launches_host);
/* "name total_ms launches\n" for every bracketed kernel since bdet_profile_begin(), into a host buffer */
int bdet_profile_report(char* buf_host, size_t buf_bytes);
int bdet_profile_end(void);

#ifdef __cplusplus
}
#endif
#endif /* BDET_H_ */
