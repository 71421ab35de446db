"""CPU, build container only: the committed golden vectors are what the reference's own source produces NOW, and the
oracle agrees with the live reference code on additional seeds.  Skipped where /root/reference is absent (GPU box)."""

import numpy as np
import pytest

from basedet_b200 import workloads as W
from oracle import ref_ops as R
from oracle import ref_runner

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="reference tree not present")


def test_committed_golden_vectors_are_current():
    from tests.golden import gen_golden

    fresh = gen_golden.build()
    gold = np.load(gen_golden.OUT)
    assert set(fresh) == set(gold.files)
    for k, v in fresh.items():
        a, b = np.asarray(v), gold[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), k


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_live_reference_on_random_inputs(seed):
    ref = ref_runner.load()
    T = ref.Tensor
    rng = np.random.default_rng(seed)
    sizes = W.retinanet_level_sizes(64, 96)
    anchors = np.concatenate(R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5))
    gt = W.make_gt(rng, 11, 64, 96, 4, 80)
    iou = ref.boxes.Boxes(T(gt[:, :4].copy())).iou(ref.boxes.Boxes(T(anchors.copy()))).numpy()
    assert np.array_equal(iou, R.box_iou(gt[:, :4], anchors))
    idx, lab = ref.matcher.Matcher([0.3, 0.7], [0, -1, 1], True)(T(iou.copy()))
    ridx, rlab = R.matcher(iou, [0.3, 0.7], [0, -1, 1], True)
    assert np.array_equal(idx.numpy(), ridx) and np.array_equal(lab.numpy(), rlab)
    n = 500
    boxes = (W.make_gt(rng, 30, 200, 200, 8, 100)[rng.integers(0, 30, n), :4] + rng.normal(0, 3, (n, 4))).astype(np.float32)
    scores = W.distinct_scores(rng, n)
    labels = rng.integers(0, 4, n).astype(np.int32)
    keep = ref.post_processing.batched_nms(T(boxes.copy()), T(scores.copy()), T(labels.copy()), 0.5, 100).numpy()
    assert np.array_equal(keep, R.batched_nms(boxes, scores, labels, 0.5, 100))


def test_reference_py_cpu_nms_agrees_with_oracle_nms():
    """The one hot-path function of the reference that is pure numpy (post_processing.py:106-132)."""
    ref = ref_runner.load()
    rng = np.random.default_rng(4)
    n = 400
    boxes = (W.make_gt(rng, 25, 300, 300, 8, 120)[rng.integers(0, 25, n), :4] + rng.normal(0, 4, (n, 4))).astype(np.float32)
    boxes[:, 2:] = np.maximum(boxes[:, 2:], boxes[:, :2] + 1)
    scores = W.distinct_scores(rng, n)
    keep = ref.post_processing.py_cpu_nms(np.concatenate([boxes, scores[:, None]], 1), 0.5)
    assert [int(k) for k in keep] == R.nms(boxes, scores, 0.5).tolist()
