"""Oracle vs the reference's own known-answer tests (verbatim data) -- CPU only.

tests/structures/test_boxes.py:38-94, tests/layers/test_postprocess.py:13-28,
tests/layers/test_roi_pool.py:32-45,64-75 of megvii-research/basedet.
"""
import numpy as np

from oracle import ref_ops as R

BOXES1 = np.array([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 1.0, 1.0]], dtype=np.float32)
BOXES2 = np.array(
    [
        [0.0, 0.0, 1.0, 1.0],
        [0.0, 0.0, 0.5, 1.0],
        [0.0, 0.0, 1.0, 0.5],
        [0.0, 0.0, 0.5, 0.5],
        [0.5, 0.5, 1.0, 1.0],
        [0.5, 0.5, 1.5, 1.5],
    ],
    dtype=np.float32,
)
NMS_BOXES = np.array(
    [
        [0.0, 0.0, 100.0, 100.0],
        [0.0, 0.0, 100.5, 100.0],
        [0.0, 0.0, 201.0, 200.5],
        [0.0, 0.0, 200.5, 200.5],
        [0.5, 0.5, 100.0, 101.0],
        [0.5, 0.5, 120.5, 120.5],
    ],
    dtype=np.float32,
)
NMS_SCORES = np.array([0.9, 0.8, 0.3, 0.7, 0.6, 0.4], dtype=np.float32)
NMS_LABELS = np.array([1, 1, 1, 2, 2, 2], dtype=np.int32)
ROI_FEAT = np.arange(25, dtype=np.float32).reshape(1, 1, 5, 5)
ROI = np.array([[0, 1, 1, 3, 3]], dtype=np.float32)
ROI_ALIGN_EXPECTED = np.array(
    [[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]]
)


def test_iou_kat():
    expected = np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / (2 - 0.25)]] * 2)
    assert np.allclose(R.box_iou(BOXES1, BOXES2), expected)


def test_ioa_kat():
    expected = np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25]] * 2).T
    assert np.allclose(R.box_ioa(BOXES2, BOXES1), expected)


def test_intersection_kat():
    expected = np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25]] * 2)
    assert np.allclose(R.box_intersection(BOXES1, BOXES2), expected)


def test_scale_kat():
    assert np.allclose(R.boxes_scale(BOXES1, (2, 2)), BOXES1 * 2)


def test_center_kat():
    assert np.allclose(R.box_center(BOXES1), np.array([[0.5, 0.5], [0.5, 0.5]]))


def test_batched_nms_kat():
    keep = R.batched_nms(NMS_BOXES, NMS_SCORES, NMS_LABELS, iou_thresh=0.4)
    assert list(keep) == [0, 3, 4, 2]


def test_roi_align_kat():
    out = R.roi_pool([ROI_FEAT], ROI, strides=[1], pool_shape=4, pooler_type="roi_align")
    assert out.shape == (1, 1, 4, 4)
    assert np.allclose(out[0, 0], ROI_ALIGN_EXPECTED)


def test_roi_max_pool_kat():
    # test_roi_pool.py:48-61 (pooler_type="roi_pool" -> F.nn.roi_pooling(mode="max")): pins ASSUMED-13
    expected = np.array([[6.0, 7.0, 8.0, 8.0], [11.0, 12.0, 13.0, 13.0], [16.0, 17.0, 18.0, 18.0], [16.0, 17.0, 18.0, 18.0]])
    out = R.roi_pool([ROI_FEAT], ROI, strides=[1], pool_shape=4, pooler_type="roi_pool")
    assert np.array_equal(out[0, 0], expected)


def test_roi_align_resize_kat():
    # test_roi_pool.py:64-75: 2x nearest-upsampled... the reference uses F.vision.interpolate (bilinear,
    # align_corners=False).  Equivariance holds for a linear ramp, which arange is.
    out = R.roi_pool([ROI_FEAT], ROI, strides=[1], pool_shape=4)
    import torch
    import torch.nn.functional as F

    feat2x = F.interpolate(torch.from_numpy(ROI_FEAT), scale_factor=2, mode="bilinear", align_corners=False).numpy()
    out2x = R.roi_pool([feat2x], ROI, strides=[1 / 2], pool_shape=4)
    assert np.allclose(out2x, out)


def test_anchor_counts():
    from basedet_b200 import workloads as W

    for hw, total in (((800, 800), 120087), ((800, 1344), 201600)):
        sizes = W.retinanet_level_sizes(*hw)
        anchors = R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
        assert sum(a.shape[0] for a in anchors) == total
    sizes = W.frcnn_level_sizes(800, 1344)
    anchors = R.default_anchors(sizes, W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
    assert [a.shape[0] for a in anchors] == [201600, 50400, 12600, 3150, 819]
