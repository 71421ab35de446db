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


class _Cfg(dict):
    __getattr__ = dict.__getitem__


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_dense_head_targets_oracle_matches_live_reference_methods(seed):
    """FCOS / ATSS get_ground_truth and OTATopkMatcher, run from the reference files on fresh seeds (beyond the committed
    golden vectors), against the oracle restatements."""
    import types

    ref = ref_runner.load()
    T = ref.Tensor
    strides = [8, 16, 32, 64, 128]
    soi = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, float("inf")]]
    hw = (128 + 32 * (seed % 3), 192)
    sizes = W.retinanet_level_sizes(*hw)
    pts = R.anchor_points(sizes, 1, strides, 0.5)
    gt, ng = W.target_assign_batch(2, 10, hw[0], hw[1], seed0=seed * 10, ragged=True)
    fcos = ref_runner.load_method("models/det/fcos.py", "FCOS", "get_ground_truth")
    atss = ref_runner.load_method("models/det/atss.py", "ATSS", "get_ground_truth")
    head = types.SimpleNamespace(strides=strides)
    for radius in (1.5, 0):
        me = types.SimpleNamespace(cfg=_Cfg(MODEL=_Cfg(HEAD=_Cfg(CENTER_SAMPLING_RADIUS=radius, OBJECT_SIZES_OF_INTEREST=soi))),
                                   head=head, box_coder=ref.boxcoder.PointCoder())
        out = [o.numpy() for o in fcos(me, [T(p) for p in pts], T(gt), [int(n) for n in ng])]
        mine = R.fcos_targets(pts, gt, ng, strides, soi, radius)
        for a, b in zip(out, mine[:3]):
            assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
    me = types.SimpleNamespace(cfg=_Cfg(MODEL=_Cfg(ANCHOR=_Cfg(SCALE=8, TOPK=9))), head=head, box_coder=ref.boxcoder.PointCoder())
    out = [o.numpy() for o in atss(me, [T(p) for p in pts], T(gt), [int(n) for n in ng])]
    mine = R.atss_targets(pts, gt, ng, strides, 8, 9)
    for a, b in zip(out, mine[:3]):
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
    rng = np.random.default_rng(seed)
    G, A = 6 + seed % 5, 900
    ious = (rng.uniform(0, 1, (G, A)) ** 3).astype(np.float32)
    cost = (np.floor(rng.uniform(0, 5, (G, A)) * 64) / 64).astype(np.float32)
    got = ref.matcher.OTATopkMatcher(10)(T(cost.copy()), T(ious.copy())).numpy()
    assert np.array_equal(got, R.ota_topk_match(cost, ious, 10))
