"""CPU: the oracle (oracle/ref_ops.py) against golden vectors produced by EXECUTING the reference's own source
files (tests/golden/gen_golden.py, numpy megengine shim).  Pins the oracle's op order / indexing / broadcasting
to the reference composition; leaf-op assumptions (ASSUMED-1..7) are shared by construction."""
import os

import numpy as np
import pytest

from basedet_b200 import workloads as W
from oracle import ref_ops as R

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.dtype.kind == "f":
        nan = np.isnan(a) & np.isnan(b)
        assert np.array_equal(bits(a)[~nan], bits(b)[~nan])
    else:
        assert np.array_equal(a, b)


def test_pairwise_family():
    b1, b2 = GOLD["pair_b1"], GOLD["pair_b2"]
    same(R.box_iou(b1, b2), GOLD["pair_iou"])
    same(R.box_ioa(b1, b2), GOLD["pair_ioa"])
    same(R.box_intersection(b1, b2), GOLD["pair_inter"])
    same(R.box_giou(b1, b2), GOLD["pair_giou"])
    same(R.box_center(b2), GOLD["pair_centers"])
    same((b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1]), GOLD["pair_area"])
    same(R.point_distance(GOLD["pd_p1"], GOLD["pd_p2"]), GOLD["pd_out"])


def test_boxes_misc_and_convert():
    raw = GOLD["misc_boxes"]
    same(R.boxes_clip(raw, (250.0, 333.0)), GOLD["misc_clip"])
    same(R.boxes_scale(raw, (1.25, 0.75)), GOLD["misc_scale"])
    same(R.boxes_filter_by_size(R.boxes_clip(raw, (250.0, 333.0))), GOLD["misc_filter"])
    for mode in ("xyxy2xywh", "xywh2xyxy", "xyxy2xcycwh", "xcycwh2xyxy", "xywh2xcycwh", "xcycwh2xywh"):
        same(R.box_convert(GOLD["pair_b2"], mode), GOLD["conv_" + mode])


def test_anchor_generators():
    sizes = [tuple(s) for s in GOLD["anc_sizes"]]
    for i, a in enumerate(R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)):
        same(a, GOLD["anc_retina_%d" % i])
    fs = [tuple(s) for s in GOLD["anc_fsizes"]]
    for i, a in enumerate(R.default_anchors(fs, W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)):
        same(a, GOLD["anc_rpn_%d" % i])
    for i, a in enumerate(R.anchor_points(sizes, 1, W.RETINANET_STRIDES, 0.5)):
        same(a, GOLD["anc_points_%d" % i])
    for i, a in enumerate(R.anchor_points(sizes[:2], 3, [8, 16], 0.0)):
        same(a, GOLD["anc_points3_%d" % i])
    for i, a in enumerate(R.fast_points(sizes[:3], [8, 16, 32])):
        same(a, GOLD["anc_fast_%d" % i])


@pytest.mark.parametrize("tag,thr,labs,lq", [("retina", [0.4, 0.5], [0, -1, 1], True), ("rpn", [0.3, 0.7], [0, -1, 1], True),
                                             ("nolq", [0.4, 0.5], [0, -1, 1], False), ("two", [0.5], [0, 1], True)])
def test_matcher(tag, thr, labs, lq):
    idx, lab = R.matcher(GOLD["match_m"], thr, labs, lq)
    same(idx, GOLD["match_%s_idx" % tag])
    same(lab, GOLD["match_%s_lab" % tag])


@pytest.mark.parametrize("tag,mean,std", [("unit", (0., 0., 0., 0.), (1., 1., 1., 1.)), ("rcnn", (0., 0., 0., 0.), (.1, .1, .2, .2)),
                                          ("odd", (0.1, -0.1, 0.05, 0.0), (0.5, 0.25, 2.0, 1.0))])
def test_coders(tag, mean, std):
    an, gt, d = GOLD["coder_anchors"], GOLD["coder_gt"], GOLD["coder_deltas"]
    same(R.boxcoder_encode(an, gt, mean, std), GOLD["coder_enc_" + tag])
    dec, dd = R.boxcoder_decode(an, d, mean, std)
    same(dec, GOLD["coder_dec_" + tag])
    same(dd, GOLD["coder_dec_inplace_" + tag])
    same(R.sumcoder_encode(an, gt, mean, std), GOLD["coder_sumenc_" + tag])
    same(R.sumcoder_decode(an, d, mean, std)[0], GOLD["coder_sumdec_" + tag])


def test_point_coder():
    pts, gt = GOLD["pc_pts"], GOLD["pc_gt"]
    same(R.pointcoder_encode(pts, gt[:, None, :]), GOLD["pc_enc"])
    same(R.pointcoder_encode(pts, gt[GOLD["pc_ridx"]]), GOLD["pc_enc_rows"])
    same(R.pointcoder_decode(pts, GOLD["pc_deltas"]), GOLD["pc_dec"])


def test_retinanet_get_ground_truth():
    lab, off, idx = R.retinanet_targets(GOLD["coder_anchors"], GOLD["gt_boxes"], GOLD["gt_num"], [0.4, 0.5], [0, -1, 1], True)
    same(idx, GOLD["gt_match_idx"])
    same(lab, GOLD["gt_labels"])
    same(off, GOLD["gt_offsets"])


def _dense_points():
    return [GOLD["dense_points_%d" % i] for i in range(5)]


SOI = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, float("inf")]]


@pytest.mark.parametrize("tag,radius", [("fcos_r15", 1.5), ("fcos_r0", 0)])
def test_fcos_get_ground_truth(tag, radius):
    """FCOS.get_ground_truth itself (fcos.py:222-293, run from the reference file) == oracle restatement."""
    lab, off, ctr, _ = R.fcos_targets(_dense_points(), GOLD["dense_gt"], GOLD["dense_num"], W.RETINANET_STRIDES, SOI, radius)
    same(lab, GOLD[tag + "_labels"])
    same(off, GOLD[tag + "_offsets"])
    same(ctr, GOLD[tag + "_ctrness"])
    assert (lab > 0).sum() > 20


def test_atss_get_ground_truth():
    """ATSS.get_ground_truth itself (atss.py:17-86) == oracle restatement."""
    lab, off, ctr, _ = R.atss_targets(_dense_points(), GOLD["dense_gt"], GOLD["dense_num"], W.RETINANET_STRIDES, 8, 9)
    same(lab, GOLD["atss_labels"])
    same(off, GOLD["atss_offsets"])
    same(ctr, GOLD["atss_ctrness"])
    assert (lab > 0).sum() > 20


def test_rpn_get_ground_truth_with_sample_labels():
    """RPN.get_ground_truth + sample_labels run from the reference files (variates fed explicitly) == oracle."""
    sizes = [tuple(x) for x in GOLD["samp_sizes"]]
    anchors = np.concatenate(R.default_anchors(sizes, W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5))
    lab, off = R.rpn_targets(anchors, GOLD["samp_gt"], GOLD["samp_num"], [0.3, 0.7], [0, -1, 1], True, 48, 6,
                             GOLD["samp_noise_pos"], GOLD["samp_noise_neg"])
    same(lab, GOLD["samp_labels"])
    same(off, GOLD["samp_offsets"])
    assert np.all((lab == 1).sum(1) <= 6) and np.all((lab >= 0).sum(1) == 48)


def test_rcnn_get_ground_truth():
    """RCNN.get_ground_truth (rcnn.py:95-147) run from the reference file == oracle restatement."""
    rois, cnt, N = GOLD["rcnn_rois"], GOLD["rcnn_nrois"], GOLD["rcnn_noise_fg"].shape[1]
    rois_list = [rois[b, : cnt[b]] for b in range(len(cnt))]
    n_all = [int(cnt[b] + GOLD["rcnn_num"][b]) for b in range(len(cnt))]
    out = R.rcnn_targets(rois_list, GOLD["rcnn_gt"], GOLD["rcnn_num"], [GOLD["rcnn_noise_fg"][b, :n] for b, n in enumerate(n_all)],
                         [GOLD["rcnn_noise_bg"][b, :n] for b, n in enumerate(n_all)], 32, 0.25, 0.5, 0.5, 0.0)
    same(np.concatenate([o[0] for o in out]), GOLD["rcnn_out_rois"])
    same(np.concatenate([o[1] for o in out]), GOLD["rcnn_out_labels"])
    same(np.concatenate([o[2] for o in out]), GOLD["rcnn_out_targets"])
    assert (GOLD["rcnn_out_labels"] > 0).sum() >= 8 and N == rois.shape[1] + GOLD["rcnn_gt"].shape[1]


@pytest.mark.parametrize("tag", ["ota_a", "ota_b"])
def test_ota_topk_matcher(tag):
    """OTATopkMatcher (matcher.py:129-161) run from the reference file == oracle restatement."""
    same(R.ota_topk_match(GOLD[tag + "_cost"], GOLD[tag + "_ious"], 10), GOLD[tag + "_match"])
    assert (GOLD[tag + "_match"] < GOLD[tag + "_cost"].shape[0]).sum() > 20


def test_nms_and_post_processing():
    b, s, l = GOLD["nms_boxes"], GOLD["nms_scores"], GOLD["nms_labels"]
    same(R.batched_nms(b, s, l, 0.5), GOLD["nms_keep_05"])
    same(R.batched_nms(b, s, l, 0.6, 50), GOLD["nms_keep_06_max50"])
    same(R.batched_nms(b, s, l.astype(np.float32), 0.7, 100), GOLD["nms_keep_float_levels"])
    same(R.batched_nms(b, GOLD["nms_scores_tied"], l, 0.5), GOLD["nms_keep_tied"])
    kb, ks, kl, _ = R.post_processing(b, s, l, GOLD["pp_img_info"], 0.5, 30)
    same(kb, GOLD["pp_boxes"])
    same(ks, GOLD["pp_scores"])
    same(kl, GOLD["pp_labels"])


def test_level_select_glue():
    r = R.retinanet_level_select(GOLD["lvl_logits"], GOLD["lvl_offsets"], GOLD["lvl_anchors"], 0.05, topk=100)
    boxes, sc, labels, keep = r
    same(R.sigmoid_f32(GOLD["lvl_logits"]).reshape(-1), GOLD["lvl_scores"])
    same(keep, GOLD["lvl_keep_idx"])
    same(sc, GOLD["lvl_keep_scores"])
    same(labels, GOLD["lvl_labels"])
    same(boxes, GOLD["lvl_boxes"])


def test_roi_pool():
    feats = [GOLD["roi_feat_%d" % i] for i in range(4)]
    same(R.assign_levels(GOLD["roi_rois"], [4, 8, 16, 32]), GOLD["roi_levels"])
    same(R.roi_pool(feats, GOLD["roi_rois"], [4, 8, 16, 32], (7, 7)), GOLD["roi_out"])
