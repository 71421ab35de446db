"""Batched pipelines: the per-image / per-level Python loops of the reference's callers, as a handful of launches.

Each function states the reference call stack it replaces.  Results are per-image identical to
running the reference's single-image code B times (the reference itself refuses multi-batch inference,
models/det/retinanet.py:174).  All arithmetic happens in libbdet.so."""
import math

import torch

from . import _lib, ops


def retinanet_targets(anchors, gt_boxes, num_gt, thresholds=(0.4, 0.5), labels=(0, -1, 1), allow_low_quality=True,
                      reg_mean=(0, 0, 0, 0), reg_std=(1, 1, 1, 1), apply_class=True, plan=None):
    """RetinaNet.get_ground_truth, models/det/retinanet.py:211-232 (apply_class=False: RPN.get_ground_truth up to
    sample_labels, models/det/rpn.py:215-226).  -> labels (B,A) int32, offsets (B,A,4), match_indices (B,A)."""
    lab, idx, off = ops.assign_targets(anchors, gt_boxes, num_gt, list(thresholds), list(labels), allow_low_quality,
                                       apply_class, reg_mean, reg_std, plan=plan)
    return lab, off, idx


def _flat_levels(tensors):
    return [t.reshape(t.shape[0], -1) for t in tensors]


def dense_postprocess(logits_list, offsets_list, anchors_list, img_info, cls_threshold=0.05, iou_threshold=0.5,
                      max_detections=100, topk=1000, ctrness_list=None, reg_mean=(0, 0, 0, 0), reg_std=(1, 1, 1, 1),
                      fused_tail=False):
    """RetinaNet.inference / FCOS.inference minus the network, for a whole batch.

    models/det/retinanet.py:181-209 (ctrness_list None: sigmoid scores, BoxCoder) or models/det/fcos.py:191-221
    (ctrness_list given: sqrt(sigmoid(cls)*sigmoid(ctr)), PointCoder) + layers/common/post_processing.py:50-103.
    logits_list[l] (B, n_l, C); offsets_list[l] (B, n_l, 4); anchors_list[l] (n_l, 4) boxes or (n_l, 2) points;
    img_info (B, >=4) rows [h, w, orig_h, orig_w].
    Returns dets (B, max_detections, 6) rows [x1, y1, x2, y2, score, label] (zero padded) and counts (B,).
    ``fused_tail``: decode + sort + NMS + finalize as one kernel per image (bdet_dense_tail) when the candidates fit shared
    memory; identical detections.  Off by default: measured on B200 the single kernel takes 43 us against 47 us of kernel time
    for the four launches it replaces, which overlap their launch latencies inside a CUDA graph (config 1: 0.101 vs 0.094 ms
    per image) -- every phase is a per-image serial chain either way."""
    L = len(logits_list)
    C = logits_list[0].shape[-1]
    lg = [t.float().contiguous() for t in logits_list]
    base, starts, lens = ops._segments(_flat_levels(lg))
    if ctrness_list is not None:
        ct = [t.float().contiguous() for t in ctrness_list]
        cbase, cstarts, _ = ops._segments(_flat_levels(ct))
        top = ops.score_filter_topk_raw(base, starts, lens, cls_threshold, topk, _lib.SCORE_FCOS, cbase, cstarts, C)
        coder = 1
    else:
        top = ops.score_filter_topk_raw(base, starts, lens, cls_threshold, topk, _lib.SCORE_SIGMOID, None, None, C)
        coder = 0
    if fused_tail:
        out = ops.dense_tail(anchors_list, offsets_list, top, topk, C, coder, iou_threshold, max_detections, img_info, reg_mean,
                             reg_std)
        if out is not None:
            return out
    boxes, scores, labels, count, runs = ops.select_decode(anchors_list, offsets_list, top, topk, C, coder, 0, reg_mean,
                                                           reg_std, with_runs=True)
    # every level's candidates arrive score-sorted from the top-k: NMS merges the L runs instead of re-sorting
    keep, keep_cnt = ops.nms_batched(boxes, scores, labels, iou_threshold, max_detections, num=count, runs=runs)
    dets = ops.finalize_detections(boxes, scores, labels, keep, keep_cnt, max_detections, img_info, mode=0)
    return dets, keep_cnt


def dense_postprocess_nchw(head_logits, head_offsets, anchors_list, img_info, num_classes, cls_threshold=0.05,
                           iou_threshold=0.5, max_detections=100, topk=1000, head_ctrness=None, reg_mean=(0, 0, 0, 0),
                           reg_std=(1, 1, 1, 1), fused_tail=False):
    """``dense_postprocess`` fed with the head outputs as the network writes them -- head_logits[l] (B, A*C, H, W),
    head_offsets[l] (B, A*4, H, W), head_ctrness[l] (B, A, H, W) -- i.e. without the reference's
    ``permute_to_N_Any_K`` passes (layers/common/function.py:26-32, models/det/retinanet.py:119-124).  The filter sweeps
    the NCHW memory linearly and re-indexes only the survivors; results are identical to permuting first."""
    mode = _lib.SCORE_FCOS if head_ctrness is not None else _lib.SCORE_SIGMOID
    top = ops.score_filter_topk_nchw(head_logits, cls_threshold, topk, num_classes, mode, head_ctrness)
    if fused_tail:
        out = ops.dense_tail(anchors_list, head_offsets, top, topk, num_classes, 1 if head_ctrness is not None else 0,
                             iou_threshold, max_detections, img_info, reg_mean, reg_std, nchw=True)
        if out is not None:
            return out
    boxes, scores, labels, count, runs = ops.select_decode(anchors_list, head_offsets, top, topk, num_classes,
                                                           1 if head_ctrness is not None else 0, 0, reg_mean, reg_std,
                                                           with_runs=True, nchw=True)
    keep, keep_cnt = ops.nms_batched(boxes, scores, labels, iou_threshold, max_detections, num=count, runs=runs)
    dets = ops.finalize_detections(boxes, scores, labels, keep, keep_cnt, max_detections, img_info, mode=0)
    return dets, keep_cnt


def rpn_proposals(scores_list, offsets_list, anchors_list, im_info, prev_nms_topk=2000, post_nms_topk=1000,
                  nms_threshold=0.7, reg_mean=(0, 0, 0, 0), reg_std=(1, 1, 1, 1)):
    """RPN.find_top_rpn_proposals, models/det/rpn.py:134-186, for a whole batch.

    scores_list[l] (B, n_l) objectness logits in (h, w, anchor) order; offsets_list[l] (B, n_l, 4); anchors_list[l]
    (n_l, 4); im_info (B, >=2) rows [h, w, ...].  k is clamped to the level size.
    Returns rois (B, post_nms_topk, 5) rows [batch, x1, y1, x2, y2] (zero padded) and counts (B,)."""
    sc = [t.float().contiguous() for t in scores_list]
    base, starts, lens = ops._segments(_flat_levels(sc))
    top = ops.topk_raw(base, starts, lens, prev_nms_topk)
    boxes, scores, levels, count, runs = ops.select_decode(anchors_list, offsets_list, top, prev_nms_topk, 1, 0, 1, reg_mean,
                                                           reg_std, im_info=im_info, with_runs=True)
    keep, keep_cnt = ops.nms_batched(boxes, scores, levels, nms_threshold, post_nms_topk, num=count, runs=runs)
    rois = ops.finalize_detections(boxes, scores, levels, keep, keep_cnt, post_nms_topk, None, mode=1)
    return rois, keep_cnt


def rpn_targets(anchors, gt_boxes, num_gt, noise_pos, noise_neg, thresholds=(0.3, 0.7), labels=(0, -1, 1),
                allow_low_quality=True, num_sample_anchors=256, positive_anchor_ratio=0.5, reg_mean=(0, 0, 0, 0),
                reg_std=(1, 1, 1, 1), plan=None):
    """RPN.get_ground_truth, models/det/rpn.py:215-240 (defaults configs/det_model/faster_rcnn_cfg.py:23-24,60-64):
    fused IoU -> Matcher -> encode, then sample_labels for positives and negatives with caller-supplied uniform
    variates noise_pos / noise_neg (B, A) (RNG contract in include/bdet.h).  -> labels (B, A) in {-1, 0, 1}, offsets."""
    lab, idx, off = ops.assign_targets(anchors, gt_boxes, num_gt, list(thresholds), list(labels), allow_low_quality,
                                       False, reg_mean, reg_std, plan=plan)
    num_pos = int(positive_anchor_ratio * num_sample_anchors)                  # rpn.py:29
    ops.sample_labels(lab, noise_pos, num_pos, 1, -1)                          # :229
    num_neg = num_sample_anchors - ops.count_labels(lab)[:, 2]                 # :231 (labels == 1).sum(), stays on device
    ops.sample_labels(lab, noise_neg, num_neg.to(torch.int32), 0, -1)          # :232
    return lab, off


def rcnn_targets(rois, n_rois, gt_boxes, num_gt, noise_fg, noise_bg, num_rois=512, fg_ratio=0.5, fg_thresh=0.5,
                 bg_thresh_high=0.5, bg_thresh_low=0.0, reg_mean=(0, 0, 0, 0), reg_std=(0.1, 0.1, 0.2, 0.2)):
    """RCNN.get_ground_truth (training), layers/head/rcnn.py:95-147 (defaults configs/det_model/faster_rcnn_cfg.py:37-41,
    56-59), for the whole batch.  rois (B, Rmax, 5) padded as ``rpn_proposals`` returns them with counts n_rois (B,);
    noise_fg / noise_bg (B, Rmax + Gmax): uniform variates of the two sample_labels calls (RNG contract in
    include/bdet.h).  -> rois (B, num_rois, 5), labels (B, num_rois) int32, bbox_targets (B, num_rois, 4), count (B,);
    the reference returns the per-image results concatenated: ``x[b, :count[b]]``."""
    m = ops.rcnn_match(rois, n_rois, gt_boxes, num_gt, fg_thresh, bg_thresh_low, bg_thresh_high)
    ops.sample_labels(m["fg"], noise_fg, int(num_rois * fg_ratio), 1, 0)        # rcnn.py:125-126
    num_bg = num_rois - ops.count_labels(m["fg"])[:, 2]                         # :127 fg_inds_mask.sum(), on the device
    ops.sample_labels(m["bg"], noise_bg, num_bg.to(torch.int32), 1, 0)          # :128
    return ops.rcnn_collect(m, gt_boxes, num_rois, reg_mean, reg_std)


def fcos_targets(points_list, gt_boxes, num_gt, strides=(8, 16, 32, 64, 128),
                 sizes_of_interest=((-1, 64), (64, 128), (128, 256), (256, 512), (512, float("inf"))),
                 center_sampling_radius=1.5, plan=None):
    """FCOS.get_ground_truth, models/det/fcos.py:222-293 (defaults: configs/det_model/fcos_cfg.py:41-47).
    -> labels (B, A) int32, offsets (B, A, 4), ctrness (B, A)."""
    lab, off, ctr, _ = ops.fcos_targets(points_list, gt_boxes, num_gt, list(strides), [tuple(s) for s in sizes_of_interest],
                                        center_sampling_radius, plan=plan)
    return lab, off, ctr


def atss_targets(points_list, gt_boxes, num_gt, strides=(8, 16, 32, 64, 128), anchor_scale=8, topk=9, plan=None):
    """ATSS.get_ground_truth, models/det/atss.py:17-86 (defaults: configs/det_model/atss_cfg.py:8-11).
    -> labels (B, A) int32, offsets (B, A, 4), ctrness (B, A)."""
    lab, off, ctr, _ = ops.atss_targets(points_list, gt_boxes, num_gt, list(strides), anchor_scale, topk, plan=plan)
    return lab, off, ctr


def free_anchor_targets(anchors, pred_offsets, pred_scores, gt_boxes, num_classes, box_iou_thresh=0.6, bucket_size=50,
                        reg_mean=(0, 0, 0, 0), reg_std=(0.1, 0.1, 0.2, 0.2)):
    """The box ops of one image of FreeAnchor.get_losses, models/det/free_anchor.py:48-113 (defaults:
    configs/det_model/freeanchor_cfg.py): anchors (A,4), pred_offsets (A,4), pred_scores (A,C) = sigmoid(logits),
    gt_boxes (G,5).  -> box_prob (A,C), matched_idx (G,bucket) [sorted by (IoU desc, index asc); the reference's
    F.topk(no_sort=True) leaves the order inside a bag unspecified], matched_score (G,bucket), matched_offsets (G*bucket,4)."""
    pred_box = ops.box_decode(anchors, pred_offsets, reg_mean, reg_std)                   # :55 (no in-place rescale: detached)
    box_prob = ops.free_anchor_box_prob(pred_box, gt_boxes, num_classes, box_iou_thresh)  # :57-86
    quality = ops.pairwise(gt_boxes[:, :4], anchors)                                      # :91
    G, A = quality.shape
    k = min(int(bucket_size), A)
    q = quality.contiguous() if quality.stride(0) != A else quality
    _, idx, _ = ops.topk_segments(q.reshape(-1), [A] * G, k)                              # :93-95
    ms, mo = ops.free_anchor_bags(idx, anchors, gt_boxes, pred_scores, reg_mean, reg_std)  # :99-114
    return box_prob, idx, ms, mo


def ota_targets(points_list, strides, gt_boxes, cls_logits, pred_deltas, candidate_k=10, alpha=0.25, gamma=2.0,
                reg_weight=1.5, center_sampling_radius=2.5):
    """One image of OTA.get_ground_truth with the top-k matcher, models/det/ota.py:91-175: points_list[l] (n_l,2),
    gt_boxes (G,5), cls_logits (A,C), pred_deltas (A,4) ltrb.  -> gt_classes (A,) fp32 (0 = background), box_targets
    (A,4), iou_targets (A,), matched (A,) int32 (G = background), cost (G,A), ious (G,A)."""
    pts = torch.cat([p.reshape(-1, 2) for p in points_list]) if len(points_list) > 1 else points_list[0]
    radius = torch.cat([torch.full((p.shape[0],), float(s * center_sampling_radius), dtype=torch.float32, device=pts.device)
                        for p, s in zip(points_list, strides)])
    cost, ious = ops.ota_cost(pts, radius, gt_boxes, cls_logits, pred_deltas, alpha, gamma, reg_weight)   # :91-152
    matched = ops.ota_topk_match(cost, ious, candidate_k)                                                 # :158
    cls_t, box_t, iou_t = ops.ota_collect(matched, pts, gt_boxes, ious)                                   # :160-175
    return cls_t, box_t, iou_t, matched, cost, ious


# Measured on B200 (profiles/r02_strong_scaling.md): generating the anchors in registers and folding the census into the
# assignment kernels wins while the step is launch bound (2 images: 61 -> 54 us) and loses once it is issue bound
# (16 images: 112 -> 131 us: the integer divisions and fp64 grid arithmetic are paid per image instead of once).
GRID_ASSIGN_MAX_BATCH = 4


class _AssignSlot:
    """One set of static input / output buffers of TargetAssigner with its captured graph."""

    def __init__(self, owner, batch, max_gt):
        dev = owner.device
        self.gt = torch.zeros((batch, max_gt, 5), dtype=torch.float32, device=dev)
        self.num_gt = torch.zeros((batch,), dtype=torch.int32, device=dev)
        self.counts = torch.zeros((batch, 3), dtype=torch.int32, device=dev)
        self.plan = ops.AssignPlan(owner.num_anchors, max_gt, batch, dev)
        self.loaded = torch.cuda.Event()   # inputs of this slot are on the device
        self.done = torch.cuda.Event()     # the step that used this slot has finished
        self.graph = None

    def step(self, owner):
        thr, lab, lq, cls, mean, std = owner.cfg
        if self.gt.shape[0] <= GRID_ASSIGN_MAX_BATCH:
            ops.assign_targets_grid(owner.gen._plan(owner.sizes), self.gt, self.num_gt, thr, lab, lq, cls, mean, std,
                                    plan=self.plan, counts=self.counts)
        else:
            anchors = owner.gen.generate_all_level_anchors(owner.sizes, owner.device)
            ops.assign_targets(anchors, self.gt, self.num_gt, thr, lab, lq, cls, mean, std, plan=self.plan)
            ops.count_labels(self.plan.labels, out=self.counts)


class TargetAssigner:
    """The whole `get_ground_truth` step of a dense head (anchors -> IoU -> Matcher -> labels -> BoxCoder.encode ->
    label census) for a FIXED batch shape, captured once into a CUDA graph and replayed every iteration.

    models/det/retinanet.py:116 (anchors regenerated per forward), :211-232 (targets), :142-146 (num_fg).  The step
    is three short kernels plus the census; replaying them as one graph removes the per-launch host work that bounds
    the eager path when the inputs arrive from the host every iteration.  There are ``depth`` (default 2) sets of
    static buffers used in turn: the host-to-device copies of step i+1 run on a copy stream while step i computes,
    so a loop fed from (pinned) host memory runs at the speed of the kernels.  Returns
      labels (B, A) int32, match_idx (B, A) int32, offsets (B, A, 4) fp32, counts (B, 3) int32 (<0, ==0, >0),
    valid until the ``depth``-th following call reuses their slot."""

    def __init__(self, anchor_generator, feature_sizes, batch, max_gt, thresholds=(0.4, 0.5), labels=(0, -1, 1),
                 allow_low_quality=True, apply_class=True, reg_mean=(0, 0, 0, 0), reg_std=(1, 1, 1, 1), device=None,
                 depth=2):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.gen, self.sizes = anchor_generator, [tuple(s) for s in feature_sizes]
        self.cfg = (list(thresholds), list(labels), bool(allow_low_quality), bool(apply_class), tuple(reg_mean), tuple(reg_std))
        with torch.cuda.device(self.device):
            self.num_anchors = self.gen.generate_all_level_anchors(self.sizes, self.device).shape[0]
            self.slots = [_AssignSlot(self, batch, max_gt) for _ in range(max(1, int(depth)))]
            self.copy_stream = torch.cuda.Stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up outside the capture (lazy module loading, allocator)
                    self.slots[0].step(self)
            torch.cuda.current_stream(self.device).wait_stream(side)
            for slot in self.slots:
                slot.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(slot.graph):
                    slot.step(self)
                slot.done.record()
        self.turn = 0
        self.kernels_per_replay = 2 if batch <= GRID_ASSIGN_MAX_BATCH else 4

    def __call__(self, gt_boxes, num_gt):
        """gt_boxes (B, max_gt, 5) fp32 and num_gt (B,) int32, host (pinned) or device."""
        slot = self.slots[self.turn]
        self.turn = (self.turn + 1) % len(self.slots)
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot.done)  # the step that last read these input buffers is over
            slot.gt.copy_(gt_boxes, non_blocking=True)
            slot.num_gt.copy_(num_gt, non_blocking=True)
            slot.loaded.record()
        main.wait_event(slot.loaded)
        slot.graph.replay()
        slot.done.record(main)
        return slot.plan.labels, slot.plan.idx, slot.plan.offsets, slot.counts


class GraphedPipeline:
    """Capture one of the pipelines above on FIXED input tensors into a CUDA graph and replay it every iteration.

    The post-processing / proposal pipelines are chains of ~6-30 short launches (filter, cluster select, decode, NMS
    chunks, finalize); replayed as one graph they cost the kernels only -- no per-launch host work, no allocator
    traffic.  ``fn(*args, **kwargs)`` is run twice for warm-up and once under capture; the tensors among the arguments
    must keep their addresses (write the next batch INTO them: ``t.copy_(new)``), the returned tensors are the same
    objects after every ``replay()``.  Example::

        post = GraphedPipeline(dense_postprocess_nchw, head_logits, head_offsets, anchors, img_info, 80)
        for batch in loader:
            run_network_into(head_logits, head_offsets)     # or copy_ into them
            dets, counts = post.replay()
    """

    def __init__(self, fn, *args, **kwargs):
        tensors = [a for a in args if torch.is_tensor(a)] + [t for a in args if isinstance(a, (list, tuple)) for t in a
                                                             if torch.is_tensor(t)]
        assert tensors and tensors[0].is_cuda, "GraphedPipeline needs CUDA tensor arguments"
        self.device = tensors[0].device
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn(*args, **kwargs)
            torch.cuda.current_stream(self.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = fn(*args, **kwargs)

    def replay(self):
        self.graph.replay()
        return self.outputs


def roi_pool_forward_backward(features, rois, strides, pool_shape=(7, 7), dout=None):
    """RCNN.forward's roi_pool (layers/head/rcnn.py:56 -> layers/common/roi_pool.py:35-78) and, if ``dout`` is
    given, the gradient w.r.t. the features that MegEngine's autodiff would produce."""
    levels = None
    if len(strides) > 1:
        levels = ops.roi_assign_levels(rois, int(math.log2(strides[0])), int(math.log2(strides[-1])))
    scales = [1.0 / s for s in strides]
    out = ops.roi_align_fwd(features, rois, levels, scales, pool_shape)
    if dout is None:
        return out
    grads = ops.roi_align_bwd(dout, [tuple(f.shape) for f in features], rois, levels, scales, pool_shape)
    return out, grads


_SIDE_STREAMS = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(torch.device("cuda", key))
    return _SIDE_STREAMS[key]


def frcnn_train_box_ops(anchor_generator, feature_sizes, rpn_scores, rpn_offsets, features, gt_boxes, num_gt, im_info,
                        noise_rpn, noise_rcnn, dout=None, rpn_strides=(4, 8, 16, 32, 64), rcnn_strides=(4, 8, 16, 32), prev_nms_topk=2000,
                        post_nms_topk=1000, nms_threshold=0.7, num_rois=512, pool_shape=(7, 7), plan=None, dfeats=None,
                        roi_order=True):
    """Every box op of one Faster R-CNN FPN training step, for the whole batch (BASELINE configs[2]):

      FasterRCNN.get_losses, models/det/faster_rcnn.py:73-94
        RPN.forward, models/det/rpn.py:91-132: anchors (per forward, :109) -> find_top_rpn_proposals (:134-186:
          top-k per level, BoxCoder.decode, clip / filter_by_size, batched_nms 0.7, <= post_nms_topk rois)
          + get_ground_truth (:215-240: IoU, Matcher(.3/.7, low-quality), BoxCoder.encode, sample_labels x 2)
        RCNN.get_ground_truth, layers/head/rcnn.py:95-147 (rois + gt -> IoU, argmax, fg / bg sampling, targets)
        roi_pool, layers/common/roi_pool.py:35-78 (assign_rois + ROIAlign 7x7 on the 512 sampled rois per image)
        ... and, with ``dout`` (the gradient of the box head w.r.t. the pooled features), the ROIAlign backward
        that MegEngine's autodiff runs for the same step.

    feature_sizes[l] = (H_l, W_l) of the RPN levels (``features[l].shape[-2:]`` in rpn.py:109), rpn_scores[l] (B, n_l)
    objectness logits in (h, w, anchor) order, rpn_offsets[l] (B, n_l, 4), features[l]
    (B, C, H_l, W_l) for the RCNN levels, gt_boxes (B, Gmax, 5), num_gt (B,), im_info (B, >= 2), noise_rpn (B, 2, A)
    and noise_rcnn (B, 2, post_nms_topk + Gmax): the uniform variates of the four sample_labels calls.
    Returns a dict of device tensors; nothing is copied to the host."""
    dev = gt_boxes.device
    B = gt_boxes.shape[0]
    # anchors: one launch for all levels; the per-level tensors are views of the flat buffer
    n_l = [int(s.shape[1]) for s in rpn_scores]
    anchors_all = anchor_generator.generate_all_level_anchors(feature_sizes, dev)
    # Two branches of the step are independent of the proposal -> RCNN -> ROIAlign chain: the RPN targets (anchors and GT
    # only) and the zero fill of dfeat.  Both go to a side stream (a second branch of the graph when captured): the chain
    # is a sequence of short latency-bound launches that leaves most of the SMs and of the HBM bandwidth idle.
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        rpn_labels, rpn_targets_ = rpn_targets(anchors_all, gt_boxes, num_gt, noise_rpn[:, 0], noise_rpn[:, 1], plan=plan)
        if dout is not None:
            if dfeats is None:
                dfeats = [torch.empty_like(f) for f in features]
            for d in dfeats:
                d.zero_()
    offs = [0]
    for n in n_l:
        offs.append(offs[-1] + n)
    anchors_list = [anchors_all[offs[i]:offs[i + 1]] for i in range(len(n_l))]
    # RPN proposals (no gradient flows through them, rpn.py:186)
    rois, n_rois = rpn_proposals(rpn_scores, rpn_offsets, anchors_list, im_info, prev_nms_topk, post_nms_topk, nms_threshold)
    # RCNN targets on the proposals
    s_rois, s_labels, s_targets, s_count = rcnn_targets(rois, n_rois, gt_boxes, num_gt, noise_rcnn[:, 0], noise_rcnn[:, 1],
                                                        num_rois=num_rois)
    # ROIAlign on the sampled rois.  The reference concatenates the per-image lists (rcnn.py:139-146); rows beyond
    # count[b] are zero rois of image b here (fixed shapes), which pool to whatever a zero box pools to and carry no
    # gradient because their dout rows are zero in a real step.
    flat_rois = s_rois.reshape(B * num_rois, 5)
    levels = ops.roi_assign_levels(flat_rois, int(math.log2(rcnn_strides[0])), int(math.log2(rcnn_strides[-1])))
    scales = [1.0 / s for s in rcnn_strides]
    # one processing order for the forward and the backward: rois sorted by (image, level, tile) keep the planes in L2
    perm = ops.roi_order([tuple(f.shape) for f in features], flat_rois, levels, scales, pool_shape) if roi_order else None
    pooled = ops.roi_align_fwd(features, flat_rois, levels, scales, pool_shape, perm=perm)
    out = dict(rois=rois, n_rois=n_rois, rpn_labels=rpn_labels, rpn_targets=rpn_targets_, rcnn_rois=s_rois,
               rcnn_labels=s_labels, rcnn_targets=s_targets, rcnn_count=s_count, pooled=pooled, levels=levels)
    main.wait_stream(side)  # join: RPN targets done, dfeat zeroed
    if dout is not None:
        out["dfeats"] = ops.roi_align_bwd(dout, None, flat_rois, levels, scales, pool_shape, dfeats=dfeats, accumulate=True,
                                          perm=perm)
    return out

