"""TEST INFRASTRUCTURE ONLY -- one-image CPU restatements of the BASELINE configs' box-op pipelines.

Composed from oracle/ref_ops.py (numpy, pinned to the reference through tests/golden) and oracle/c/oracle.c (the
pthread-parallel C restatement, checked against ref_ops) for the passes that are too slow in numpy.  Used (a) by
tests/ as the checker of the batched CUDA pipelines at full size and (b) by bench.py's cpu_baseline / --impl reference
legs as the stand-in for the reference's MegEngine CPU path (MegEngine cannot be installed offline).  The product
(basedet_b200/) never imports this module.
"""
import numpy as np

from basedet_b200 import workloads as W

from . import c_oracle as C
from . import ref_ops as R

f32 = np.float32


# ------------------------------------------------------------------------------------------------ config 3
def rpn_proposals_image(scores_l, deltas_l, anchors_l, im_hw, pre_k, post_k, thr):
    """RPN.find_top_rpn_proposals for one image, models/det/rpn.py:141-186 (proposals keep their un-clipped
    coordinates; the clipped boxes only feed the size filter).  -> (n, 4) proposals in NMS order."""
    props, scs, lvls = [], [], []
    for level, (s, d, a) in enumerate(zip(scores_l, deltas_l, anchors_l)):
        boxes, _ = R.boxcoder_decode(a, d)                                      # :149-150
        v, order = R.topk_desc(s, pre_k)                                        # :154-155
        props.append(boxes[order])
        scs.append(v)
        lvls.append(np.full(len(v), level, f32))                                # :160
    props, scs, lvls = np.concatenate(props), np.concatenate(scs), np.concatenate(lvls)
    keep_mask = R.boxes_filter_by_size(R.boxes_clip(props, im_hw))              # :168-171
    props, scs, lvls = props[keep_mask], scs[keep_mask], lvls[keep_mask]
    if len(props) == 0:
        return props
    shifted = R.nms_offset_boxes(props, lvls)                                   # post_processing.py:44-46
    keep = C.nms(shifted, scs, thr, post_k)                                     # F.vision.nms (ASSUMED-5), C sweep
    return props[keep]


def rpn_targets_image(anchors, gt5, noise_pos, noise_neg, thresholds=(0.3, 0.7), labels=(0, -1, 1), allow_lq=True,
                      num_sample=256, pos_ratio=0.5):
    """RPN.get_ground_truth for one image, models/det/rpn.py:215-240, IoU / Matcher passes in C."""
    overlaps = C.box_iou(gt5[:, :4], anchors)                                   # :223
    idx, lab = C.matcher(overlaps, list(thresholds), list(labels), allow_lq)    # :224
    off = R.boxcoder_encode(anchors, gt5[idx][:, :4])                           # :226
    lab = R.sample_labels(lab, int(pos_ratio * num_sample), 1, -1, noise_pos)   # :229
    lab = R.sample_labels(lab, num_sample - int((lab == 1).sum()), 0, -1, noise_neg)   # :231-232
    return lab, off


def frcnn_image_chain(inp, feats=None, dout=None, bid=0, hw=W.FRCNN_HW, with_roi=True, given_rois=None, given_sampled=None):
    """Every box op of one Faster R-CNN training step for ONE image (the per-image body of the reference's loops):
    anchors -> RPN proposals -> RPN targets -> RCNN targets -> roi_pool (+ its backward).  ``inp`` = one image of
    ``workloads.frcnn_image``; feats[l] (1, C, H_l, W_l); dout (num_rois, C, 7, 7).  ``bid``: the batch index written
    into the roi rows (rpn.py:176-177).  ``given_rois`` (n, 5) / ``given_sampled`` (m, 5) replace this function's own
    proposals / sampled rois as the INPUT of the following stage (discrete decisions are compared stage by
    stage on bit-identical inputs; decode goes through expf, which differs by an ulp between libms)."""
    sizes = W.frcnn_level_sizes(*hw)
    anchors_l = R.default_anchors(sizes, W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, W.FRCNN_OFFSET)
    anchors = np.concatenate(anchors_l)
    g5 = inp["gt"][: int(inp["num_gt"])]
    props = rpn_proposals_image(inp["scores"], inp["deltas"], anchors_l, hw, W.FRCNN_PRE_NMS, W.FRCNN_POST_NMS,
                                W.FRCNN_NMS_THR)
    rois = np.concatenate([np.full((len(props), 1), f32(bid), f32), props], axis=1)
    out_rois = rois
    if given_rois is not None:
        rois = np.asarray(given_rois, f32)
    rpn_lab, rpn_off = rpn_targets_image(anchors, g5, inp["noise_rpn"][0], inp["noise_rpn"][1])
    n_all = len(rois) + len(g5)
    (s_rois, s_lab, s_tgt), = R.rcnn_targets([rois], [g5], [len(g5)], [inp["noise_rcnn"][0][:n_all]],
                                             [inp["noise_rcnn"][1][:n_all]], num_rois=W.FRCNN_NUM_ROIS)
    s_rois[:, 0] = f32(bid)          # R.rcnn_targets numbers the gt rows by the position in ITS list (0 here)
    out = dict(rois=out_rois, rpn_labels=rpn_lab, rpn_targets=rpn_off, rcnn_rois=s_rois, rcnn_labels=s_lab, rcnn_targets=s_tgt)
    if not with_roi or feats is None:
        return out
    strides = W.FRCNN_RCNN_STRIDES
    local = (s_rois if given_sampled is None else np.asarray(given_sampled, f32)).copy()
    local[:, 0] = 0                  # feats holds this image only
    levels = R.assign_levels(local, strides)
    C_ = feats[0].shape[1]
    pooled = np.zeros((len(local), C_, 7, 7), f32)
    grads = [np.zeros(f.shape, np.float64) for f in feats] if dout is not None else None
    for l, f in enumerate(feats):
        sel = np.flatnonzero(levels == l)
        if len(sel) == 0:
            continue
        pooled[sel] = C.roi_align_fwd(f, local[sel], (7, 7), 1.0 / strides[l])
        if dout is not None:
            grads[l] = C.roi_align_bwd(dout[sel], f.shape, local[sel], (7, 7), 1.0 / strides[l])
    out.update(levels=levels, pooled=pooled)
    if grads is not None:
        out["dfeats"] = grads
    return out


# ------------------------------------------------------------------------------------------------ configs 1 / 4
def dense_image_postprocess(logits_l, offsets_l, anchors_l, im_info, cls_thr=0.05, iou_thr=0.5, max_dets=100, topk=1000,
                            ctrness_l=None):
    """RetinaNet.inference / FCOS.inference minus the network for one image: models/det/retinanet.py:181-209 or
    models/det/fcos.py:191-221 + layers/common/post_processing.py:50-103.  -> (boxes (n,4), scores (n,), labels (n,))."""
    tb, ts, tl = [], [], []
    for l, (lg, off, anc) in enumerate(zip(logits_l, offsets_l, anchors_l)):
        Cn = lg.shape[-1]
        if ctrness_l is None:
            sc = R.sigmoid_f32(lg.reshape(-1))
        else:
            sc = R.fcos_scores(lg, ctrness_l[l]).reshape(-1)
        keep, vals = R.filter_topk_scores(sc, cls_thr, topk)
        if len(keep) == 0:
            continue
        if ctrness_l is None:
            boxes, _ = R.boxcoder_decode(anc, off)
        else:
            boxes = R.pointcoder_decode(anc, off)
        tb.append(boxes[keep // Cn])
        ts.append(vals)
        tl.append((keep % Cn).astype(np.int32))
    if not tb:
        return np.zeros((0, 4), f32), np.zeros((0,), f32), np.zeros((0,), np.int32)
    kb, ks, kl, _ = R.post_processing(np.concatenate(tb), np.concatenate(ts), np.concatenate(tl),
                                      np.asarray(im_info, f32).reshape(1, -1), iou_thr, max_dets)
    return kb, ks, kl


# ------------------------------------------------------------------------------------------------ configs 2 / 5
def retinanet_targets_image(anchors, gt5, scratch=None):
    """RetinaNet.get_ground_truth for one image (retinanet.py:211-232), all passes in C with a materialised matrix."""
    return C.retinanet_targets_one(anchors, gt5, W.RETINANET_MATCHER["thresholds"], W.RETINANET_MATCHER["labels"],
                                   W.RETINANET_MATCHER["allow_low_quality"], scratch=scratch)


def stress_image_ops(inp):
    """Config 5 for one image: the (500, 200k) IoU matrix + Matcher, and single-class NMS 0.5 over 100k boxes."""
    iou = C.box_iou(inp["gt"][:, :4], inp["anchors"])
    idx, lab = C.matcher(iou, [0.4, 0.5], [0, -1, 1], True)
    keep = C.nms(inp["boxes"], inp["scores"], 0.5, None)
    return iou, idx, lab, keep
