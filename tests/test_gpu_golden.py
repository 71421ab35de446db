"""GPU: the CUDA library, called through the drop-in layer, against golden vectors produced by executing the
reference's own source files (tests/golden/gen_golden.py).  Integer / index outputs and add-sub-mul-div float
outputs are bit-exact; logf / expf / sigmoid based outputs within 1e-6 (scale-relative)."""
import os

import numpy as np
import pytest
import torch

from basedet_b200 import ops
from basedet_b200 import workloads as W
from basedet_b200.layers import (AnchorPointGenerator, DefaultAnchorGenerator, FastPointGenerator, Matcher, batched_nms,
                                 non_zeros, post_processing, roi_pool)
from basedet_b200.structures import BoxCoder, BoxConverter, Boxes, Container, PointCoder, SumBoxCoder, point_distance

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def N(t):
    return t.detach().cpu().numpy()


def same(got, ref):
    got, ref = np.ascontiguousarray(N(got) if isinstance(got, torch.Tensor) else got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.dtype.kind == "f":
        nan = np.isnan(ref)
        assert np.array_equal(np.isnan(got), nan)
        assert np.array_equal(got[~nan].view(np.uint32), np.ascontiguousarray(ref[~nan]).view(np.uint32))
    else:
        assert np.array_equal(got.astype(ref.dtype), ref)


def close(got, ref, tol=1e-6, scale=None):
    got, ref = N(got).astype(np.float64), np.asarray(ref, np.float64)
    if scale is None:
        scale = np.maximum(np.abs(ref), 1.0)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.max((np.abs(got - ref) / scale)[fin], initial=0.0) <= tol


def box_scale(ref):
    ref = np.asarray(ref, np.float64).reshape(-1, 4)
    return np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)


def test_pairwise_family(cuda):
    b1, b2 = Boxes(T(GOLD["pair_b1"])), Boxes(T(GOLD["pair_b2"]))
    same(b1.iou(b2), GOLD["pair_iou"])
    same(b1.ioa(b2), GOLD["pair_ioa"])
    same(b1.intersection(b2), GOLD["pair_inter"])
    same(b1.giou(b2), GOLD["pair_giou"])
    same(b2.centers, GOLD["pair_centers"])
    same(b2.area, GOLD["pair_area"])
    same(b2.width, GOLD["pair_width"])
    same(b2.height, GOLD["pair_height"])
    close(point_distance(T(GOLD["pd_p1"]), T(GOLD["pd_p2"])), GOLD["pd_out"])


def test_boxes_misc_and_convert(cuda):
    raw = GOLD["misc_boxes"]
    same(Boxes(T(raw)).clip((250.0, 333.0)), GOLD["misc_clip"])
    same(Boxes(T(raw)).scale((1.25, 0.75)), GOLD["misc_scale"])
    same(Boxes(T(raw)).clip((250.0, 333.0)).filter_by_size(), GOLD["misc_filter"])
    for mode in ("xyxy2xywh", "xywh2xyxy", "xyxy2xcycwh", "xcycwh2xyxy", "xywh2xcycwh", "xcycwh2xywh"):
        same(BoxConverter.convert(T(GOLD["pair_b2"]), mode), GOLD["conv_" + mode])


def test_anchor_generators(cuda):
    feats = [torch.empty((1, 1, int(h), int(w)), device=cuda) for h, w in GOLD["anc_sizes"]]
    for i, a in enumerate(DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)(feats)):
        same(a, GOLD["anc_retina_%d" % i])
    ffeats = [torch.empty((1, 1, int(h), int(w)), device=cuda) for h, w in GOLD["anc_fsizes"]]
    for i, a in enumerate(DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)(ffeats)):
        same(a, GOLD["anc_rpn_%d" % i])
    for i, a in enumerate(AnchorPointGenerator(1, tuple(W.RETINANET_STRIDES), 0.5)(feats)):
        same(a, GOLD["anc_points_%d" % i])
    for i, a in enumerate(AnchorPointGenerator(3, (8, 16), 0.0)(feats[:2])):
        same(a, GOLD["anc_points3_%d" % i])
    for i, a in enumerate(FastPointGenerator((8, 16, 32))(feats[:3])):
        same(a, GOLD["anc_fast_%d" % i])


@pytest.mark.parametrize("tag,thr,labs,lq", [("retina", [0.4, 0.5], [0, -1, 1], True), ("rpn", [0.3, 0.7], [0, -1, 1], True),
                                             ("nolq", [0.4, 0.5], [0, -1, 1], False), ("two", [0.5], [0, 1], True)])
def test_matcher(cuda, tag, thr, labs, lq):
    idx, lab = Matcher(list(thr), list(labs), lq)(T(GOLD["match_m"]))
    same(idx, GOLD["match_%s_idx" % tag])
    same(lab, GOLD["match_%s_lab" % tag])


@pytest.mark.parametrize("tag,mean,std", [("unit", (0., 0., 0., 0.), (1., 1., 1., 1.)), ("rcnn", (0., 0., 0., 0.), (.1, .1, .2, .2)),
                                          ("odd", (0.1, -0.1, 0.05, 0.0), (0.5, 0.25, 2.0, 1.0))])
def test_coders(cuda, tag, mean, std):
    an, gt, d = T(GOLD["coder_anchors"]), T(GOLD["coder_gt"]), T(GOLD["coder_deltas"])
    bc = BoxCoder(mean, std)
    close(bc.encode(an, gt), GOLD["coder_enc_" + tag])
    dec = bc.decode(an, d)
    close(dec, GOLD["coder_dec_" + tag], scale=np.repeat(box_scale(GOLD["coder_dec_" + tag]), 4, 1).reshape(dec.shape))
    same(d, GOLD["coder_dec_inplace_" + tag])          # the caller's deltas are rescaled in place (boxcoder.py:76-77)
    sc = SumBoxCoder(mean, std)
    same(sc.encode(an, gt), GOLD["coder_sumenc_" + tag])
    same(sc.decode(an, T(GOLD["coder_deltas"])), GOLD["coder_sumdec_" + tag])


def test_point_coder(cuda):
    pc = PointCoder()
    pts = T(GOLD["pc_pts"])
    same(pc.encode(pts, T(GOLD["pc_gt"]).unsqueeze(1)), GOLD["pc_enc"])
    same(pc.encode(pts, T(GOLD["pc_gt"][GOLD["pc_ridx"]])), GOLD["pc_enc_rows"])
    same(pc.decode(pts, T(GOLD["pc_deltas"])), GOLD["pc_dec"])


def test_retinanet_get_ground_truth_dropin_and_fused(cuda):
    anchors = T(GOLD["coder_anchors"])
    gt5, ng = GOLD["gt_boxes"], GOLD["gt_num"]
    matcher = Matcher([0.4, 0.5], [0, -1, 1], True)
    box_coder = BoxCoder((0., 0., 0., 0.), (1., 1., 1., 1.))
    for b in range(gt5.shape[0]):
        gt_boxes = T(gt5[b])[: int(ng[b])]
        overlaps = Boxes(gt_boxes[:, :4]).iou(Boxes(anchors))
        match_indices, labels = matcher(overlaps)
        gt_boxes_matched = gt_boxes[match_indices.long()]
        fg_mask = labels == 1
        labels[fg_mask] = gt_boxes_matched[fg_mask, 4].to(torch.int32)
        offsets = box_coder.encode(anchors, gt_boxes_matched[:, :4])
        same(match_indices, GOLD["gt_match_idx"][b])
        same(labels, GOLD["gt_labels"][b])
        close(offsets, GOLD["gt_offsets"][b])
    lab, idx, off = ops.assign_targets(anchors, T(gt5), T(ng), [0.4, 0.5], [0, -1, 1], True, True)
    same(idx, GOLD["gt_match_idx"])
    same(lab, GOLD["gt_labels"])
    close(off, GOLD["gt_offsets"])


def test_nms_and_post_processing(cuda):
    b, s, l = T(GOLD["nms_boxes"]), T(GOLD["nms_scores"]), T(GOLD["nms_labels"])
    same(batched_nms(b, s, l, 0.5), GOLD["nms_keep_05"])
    same(batched_nms(b, s, l, 0.6, 50), GOLD["nms_keep_06_max50"])
    same(batched_nms(b, s, l.float(), 0.7, 100), GOLD["nms_keep_float_levels"])
    same(batched_nms(b, T(GOLD["nms_scores_tied"]), l, 0.5), GOLD["nms_keep_tied"])
    cont = Container(boxes=Boxes(T(GOLD["nms_boxes"])), box_scores=s, box_labels=l)
    res = post_processing(cont, T(GOLD["pp_img_info"]), 0.5, max_detections_per_image=30)
    same(res.boxes, GOLD["pp_boxes"])
    same(res.box_scores, GOLD["pp_scores"])
    same(res.box_labels, GOLD["pp_labels"])


def test_level_select_glue(cuda):
    """retinanet.py:181-196 for one level; top-k / candidate indices are asserted on the GOLDEN score tensor
    (bit-identical input), the CUDA sigmoid separately within 1e-6."""
    close(ops.scores(T(GOLD["lvl_logits"]).reshape(-1)), GOLD["lvl_scores"], tol=1e-6, scale=np.maximum(GOLD["lvl_scores"], 1e-30))
    scores = T(GOLD["lvl_scores"])
    _, keep_idx = non_zeros(scores > 0.05)
    topk_num = min(keep_idx.shape[0], 100)
    _, topk_idx, _ = ops.topk_segments(scores[keep_idx.long()], [keep_idx.shape[0]], topk_num)
    keep_idx = keep_idx[topk_idx[0].long()]
    same(keep_idx, GOLD["lvl_keep_idx"])
    same(scores[keep_idx.long()], GOLD["lvl_keep_scores"])
    same(keep_idx % 20, GOLD["lvl_labels"])
    boxes = BoxCoder().decode(T(GOLD["lvl_anchors"]), T(GOLD["lvl_offsets"]).reshape(-1, 4))
    close(boxes[(keep_idx // 20).long()], GOLD["lvl_boxes"], scale=box_scale(GOLD["lvl_boxes"]))
    # fused path on raw scores gives the same candidates
    vals, idx, cnt = ops.score_filter_topk(scores, [scores.numel()], 0.05, 100, mode=0)
    same(idx[0, : int(cnt[0])], GOLD["lvl_keep_idx"])
    same(vals[0, : int(cnt[0])], GOLD["lvl_keep_scores"])


def test_roi_pool(cuda):
    feats = [T(GOLD["roi_feat_%d" % i]) for i in range(4)]
    rois = T(GOLD["roi_rois"])
    same(ops.roi_assign_levels(rois, 2, 5), GOLD["roi_levels"])
    out = roi_pool(feats, rois, [4, 8, 16, 32], (7, 7), "roi_align")
    close(out, GOLD["roi_out"], tol=1e-5)
    same(out, GOLD["roi_out"])  # same op order, no FMA contraction: bit-exact in practice


def test_rpn_get_ground_truth_with_sampling_golden(cuda):
    """RPN.get_ground_truth + sample_labels (rpn.py:215-240, sampling.py:7-30) vs the vectors the reference source
    produced with the same explicit variates."""
    from basedet_b200 import pipelines
    sizes = [tuple(int(v) for v in x) for x in GOLD["samp_sizes"]]
    gen = DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
    anchors = gen.generate_all_level_anchors(sizes, cuda)
    Tc = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda)  # noqa: E731
    lab, off = pipelines.rpn_targets(anchors, Tc(GOLD["samp_gt"]), Tc(GOLD["samp_num"]), Tc(GOLD["samp_noise_pos"]),
                                     Tc(GOLD["samp_noise_neg"]), (0.3, 0.7), (0, -1, 1), True, 48, 6 / 48)
    assert np.array_equal(lab.cpu().numpy(), GOLD["samp_labels"])
    assert np.max(np.abs(off.cpu().numpy() - GOLD["samp_offsets"])) <= 1e-6


def test_rcnn_get_ground_truth_golden(cuda):
    """RCNN.get_ground_truth (layers/head/rcnn.py:95-147) vs the vectors produced by the reference method itself."""
    from basedet_b200 import pipelines
    Tc = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda)  # noqa: E731
    rois, labels, targets, count = pipelines.rcnn_targets(Tc(GOLD["rcnn_rois"]), Tc(GOLD["rcnn_nrois"]), Tc(GOLD["rcnn_gt"]),
                                                          Tc(GOLD["rcnn_num"]), Tc(GOLD["rcnn_noise_fg"]), Tc(GOLD["rcnn_noise_bg"]),
                                                          32, 0.25, 0.5, 0.5, 0.0)
    count = count.cpu().numpy()
    cat = lambda x: np.concatenate([x[b, : count[b]].cpu().numpy() for b in range(len(count))])  # noqa: E731
    assert np.array_equal(cat(rois), GOLD["rcnn_out_rois"])
    assert np.array_equal(cat(labels), GOLD["rcnn_out_labels"])
    assert np.max(np.abs(cat(targets) - GOLD["rcnn_out_targets"])) <= 1e-5   # logf / divide by std 0.1, 0.2
    for b in range(len(count)):
        assert float(rois[b, count[b]:].abs().sum()) == 0.0


@pytest.mark.parametrize("tag", ["ota_a", "ota_b"])
def test_ota_topk_matcher_golden(cuda, tag):
    """layers.OTATopkMatcher (matcher.py:129-161) vs the vectors produced by the reference class."""
    from basedet_b200.layers import OTATopkMatcher
    Tc = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda)  # noqa: E731
    out = OTATopkMatcher(10)(Tc(GOLD[tag + "_cost"]), Tc(GOLD[tag + "_ious"]))
    assert out.dtype == torch.int32 and np.array_equal(out.cpu().numpy(), GOLD[tag + "_match"])
