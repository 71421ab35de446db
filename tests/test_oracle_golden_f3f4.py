"""CPU: the oracle restatements of the widened rows 8(f)-3 / 8(f)-4 (FreeAnchor box probabilities and bags, OTA cost
construction + dynamic-k targets, COCO result records) against golden vectors produced by the reference's own statements
(tests/golden/gen_golden_f3f4.py)."""
import os

import numpy as np

from basedet_b200 import workloads as W
from oracle import ref_ops as R

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_f3f4.npz"))


def test_free_anchor_box_prob_and_bags_bit_identical():
    for case in ("normal", "fill"):
        p = "fa_%s_" % case
        scores = R.sigmoid_f32(GOLD[p + "logits"])
        bp, idx, sc, off, fill = R.free_anchor_targets(GOLD[p + "anchors"], GOLD[p + "offsets"], scores, GOLD[p + "gt"], 8)
        assert fill == bool(GOLD[p + "fill"])
        assert np.array_equal(bp, GOLD[p + "box_prob"]), case
        assert np.array_equal(idx, GOLD[p + "matched_idx"]), case           # the shim's no_sort top-k is (desc, index asc) too
        assert np.array_equal(sc, GOLD[p + "matched_score"]), case
        assert np.array_equal(off, GOLD[p + "matched_offsets"]), case


def test_ota_ground_truth_bit_identical():
    hw = tuple(int(v) for v in GOLD["ota_hw"])
    strides = [8, 16, 32, 64, 128]
    pts = R.anchor_points(W.retinanet_level_sizes(*hw), 1, strides, 0.5)
    for b in range(2):
        n = int(GOLD["ota_num_gt"][b])
        cls = np.concatenate([GOLD["ota_cls_%d" % l][b] for l in range(5)])
        dl = np.concatenate([GOLD["ota_delta_%d" % l][b] for l in range(5)])
        ct, bt, it, _ = R.ota_targets(pts, strides, GOLD["ota_gt"][b, :n], cls, dl, 6)
        assert np.array_equal(ct, GOLD["ota_gt_classes"][b])
        assert np.array_equal(bt, GOLD["ota_gt_deltas"][b])
        assert np.array_equal(it, GOLD["ota_gt_ious"][b])


def test_coco_format_records():
    img, box, sc, cat = R.coco_format(GOLD["coco_dets"], GOLD["coco_cnt"], GOLD["coco_image_ids"], GOLD["coco_origin"])
    assert np.array_equal(img, GOLD["coco_rec_image"]) and np.array_equal(cat, GOLD["coco_rec_cat"])
    assert np.array_equal(box, GOLD["coco_rec_bbox"]) and np.array_equal(sc, GOLD["coco_rec_score"])
    assert np.array_equal(R.coco_format(GOLD["coco_dets"], GOLD["coco_cnt"], GOLD["coco_image_ids"])[3], GOLD["coco_rec_cat_plus1"])
