"""GPU parity of the widened rows 8(f)-3 / 8(f)-4: FreeAnchor box probabilities + bags, OTA cost construction +
dynamic-k targets, COCO result records -- against the golden vectors the reference's own statements produced
(tests/golden/gen_golden_f3f4.py) and against the oracle on larger seeded inputs."""
import os

import numpy as np
import pytest
import torch

from basedet_b200 import ops, pipelines
from basedet_b200 import workloads as W
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_f3f4.npz"))


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0), initial=0.0))


@pytest.mark.parametrize("case", ["normal", "fill"])
def test_free_anchor_golden(case):
    p = "fa_%s_" % case
    anchors, gt, offs = GOLD[p + "anchors"], GOLD[p + "gt"], GOLD[p + "offsets"]
    scores = R.sigmoid_f32(GOLD[p + "logits"])
    bp, idx, ms, mo = pipelines.free_anchor_targets(T(anchors), T(offs), T(scores), T(gt), 8)
    # decode goes through expf: the box probabilities are compared at 1e-5 where both are non-zero, and the support may
    # differ only where the reference value sits at the clip edge (IoU within an ulp of the threshold)
    ref = GOLD[p + "box_prob"]
    got = bp.cpu().numpy()
    both = (ref != 0) & (got != 0)
    assert rel(got[both], ref[both]) <= 1e-5
    edge = (ref != 0) ^ (got != 0)
    assert np.all(np.maximum(ref, got)[edge] <= 1e-4), case
    assert np.array_equal(idx.cpu().numpy(), GOLD[p + "matched_idx"])           # IoU is exact: same bags, same order
    assert np.array_equal(ms.cpu().numpy(), GOLD[p + "matched_score"])
    assert rel(mo.cpu().numpy(), GOLD[p + "matched_offsets"]) <= 1e-6


def test_free_anchor_box_prob_exact_on_given_boxes():
    """Op-level gate (H9): on GIVEN decoded boxes the scatter is bit-exact, including class collisions (last GT wins)
    and the empty-set workaround."""
    rng = np.random.default_rng(5)
    A, G, C = 5000, 40, 12
    pred = W.make_gt(rng, A, 400, 600, 8, 200)[:, :4].copy()
    gt = W.make_gt(rng, G, 400, 600, 16, 200)
    gt[:, 4] = rng.integers(1, 4, G)                                            # 3 classes over 40 GT: many collisions
    for fill in (False, True):
        g5 = gt.copy()
        if fill:
            g5[:, :4] += 5000.0                                                 # nothing overlaps
        got = ops.free_anchor_box_prob(T(pred), T(g5), C).cpu().numpy()
        overlaps = R.box_iou(g5[:, :4], pred)
        t1 = np.float32(0.6)
        t2 = np.minimum(np.maximum(overlaps.max(axis=1, keepdims=True), np.float32(0.6 + 1e-7)), np.float32(1.0))
        prob = np.minimum(np.maximum(((overlaps - t1) / (t2 - t1)).astype(np.float32), np.float32(0)), np.float32(1))
        f = bool(prob.max() <= np.float32(1e-7))
        assert f == fill
        if f:
            prob[0, 0] = np.float32(0.001)
        ref = np.zeros((A, C), np.float32)
        labels = g5[:, 4].astype(np.int32) - 1
        for g, a in zip(*np.nonzero(prob)):
            ref[a, labels[g]] = prob[g, a]
        if f:
            ref[0, 0] = 0
        assert np.array_equal(got, ref), fill


def test_ota_ground_truth_golden():
    hw = tuple(int(v) for v in GOLD["ota_hw"])
    strides = [8, 16, 32, 64, 128]
    pts = R.anchor_points(W.retinanet_level_sizes(*hw), 1, strides, 0.5)
    for b in range(2):
        n = int(GOLD["ota_num_gt"][b])
        gt = GOLD["ota_gt"][b, :n]
        cls = np.concatenate([GOLD["ota_cls_%d" % l][b] for l in range(5)])
        dl = np.concatenate([GOLD["ota_delta_%d" % l][b] for l in range(5)])
        ct, bt, it, matched, cost, ious = pipelines.ota_targets([T(p) for p in pts], strides, T(gt), T(cls), T(dl))
        rc, ri, _, _, _ = R.ota_cost(pts, strides, gt, cls, dl, 6)
        assert rel(cost.cpu().numpy(), rc) <= 1e-5 and rel(ious.cpu().numpy(), ri) <= 1e-6
        # discrete decisions on bit-identical inputs (H9): the oracle matcher fed with the GPU's matrices
        assert np.array_equal(matched.cpu().numpy(), R.ota_topk_match(cost.cpu().numpy(), ious.cpu().numpy(), 10))
        # and end to end against the reference's own output
        assert np.array_equal(ct.cpu().numpy(), GOLD["ota_gt_classes"][b])
        assert np.array_equal(bt.cpu().numpy(), GOLD["ota_gt_deltas"][b])
        assert rel(it.cpu().numpy(), GOLD["ota_gt_ious"][b]) <= 1e-6


def test_ota_cost_config4_shape_vs_oracle():
    """22 400 points x 100 GT x 80 classes (the config-4 pyramid): cost / IoU matrices against the oracle."""
    rng = np.random.default_rng(8)
    strides = [8, 16, 32, 64, 128]
    pts = R.anchor_points(W.retinanet_level_sizes(800, 1344), 1, strides, 0.5)
    A = sum(p.shape[0] for p in pts)
    gt = W.make_gt(rng, 100, 800, 1344)
    cls = rng.normal(-3, 2, (A, 80)).astype(np.float32)
    dl = np.concatenate([np.abs(rng.normal(0, 1, (p.shape[0], 4)) * s * 2).astype(np.float32) for p, s in zip(pts, strides)])
    ct, bt, it, matched, cost, ious = pipelines.ota_targets([T(p) for p in pts], strides, T(gt), T(cls), T(dl))
    rc, ri, _, gd, _ = R.ota_cost(pts, strides, gt, cls, dl, 80)
    assert rel(cost.cpu().numpy(), rc) <= 1e-5 and rel(ious.cpu().numpy(), ri) <= 1e-6
    m = R.ota_topk_match(cost.cpu().numpy(), ious.cpu().numpy(), 10)
    assert np.array_equal(matched.cpu().numpy(), m)
    fg = m != 100
    assert fg.sum() > 0
    assert np.array_equal(ct.cpu().numpy()[fg], gt[m[fg], 4]) and not ct.cpu().numpy()[~fg].any()
    assert np.array_equal(bt.cpu().numpy()[fg], gd[m[fg], np.arange(A)[fg]])


def test_coco_format_golden():
    img, box, sc, cat = ops.coco_format(T(GOLD["coco_dets"]), T(GOLD["coco_cnt"]), T(GOLD["coco_image_ids"]), T(GOLD["coco_origin"]))
    assert np.array_equal(img.cpu().numpy(), GOLD["coco_rec_image"]) and np.array_equal(cat.cpu().numpy(), GOLD["coco_rec_cat"])
    assert np.array_equal(box.cpu().numpy(), GOLD["coco_rec_bbox"]) and np.array_equal(sc.cpu().numpy(), GOLD["coco_rec_score"])
    cat1 = ops.coco_format(T(GOLD["coco_dets"]), T(GOLD["coco_cnt"]), T(GOLD["coco_image_ids"]))[3]
    assert np.array_equal(cat1.cpu().numpy(), GOLD["coco_rec_cat_plus1"])
    # batch of 64 x 100 straight from the dense post-processing layout
    rng = np.random.default_rng(3)
    dets = rng.uniform(0, 500, (64, 100, 6)).astype(np.float32)
    dets[:, :, 5] = rng.integers(0, 80, (64, 100))
    cnt = rng.integers(0, 101, 64).astype(np.int32)
    ids = rng.integers(1, 10 ** 6, 64).astype(np.int32)
    got = ops.coco_format(T(dets), T(cnt), T(ids))
    ref = R.coco_format(dets, cnt, ids)
    for g_, r_ in zip(got, ref):
        assert np.array_equal(g_.cpu().numpy(), r_)
