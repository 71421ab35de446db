"""GPU parity of the anchor-free target assignment: FCOS / ATSS get_ground_truth, fused kernels vs
the oracle restatement (itself pinned against the reference methods, tests/test_oracle_golden.py) and vs the golden
vectors produced by the reference's own source.  Labels / match indices bit-exact; offsets and centerness are
add / sub / div / sqrt of fp32 values in the reference's op order, so they are compared bit-exactly too."""
import os

import numpy as np
import pytest
import torch

from basedet_b200 import ops, pipelines
from basedet_b200 import workloads as W
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
STRIDES = [8, 16, 32, 64, 128]
SOI = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, float("inf")]]


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def same(got, ref):
    got = got.cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(got, ref, equal_nan=ref.dtype.kind == "f")


def gold_points():
    return [GOLD["dense_points_%d" % i] for i in range(5)]


@pytest.mark.parametrize("tag,radius", [("fcos_r15", 1.5), ("fcos_r0", 0)])
def test_fcos_targets_golden(tag, radius):
    lab, off, ctr = pipelines.fcos_targets([T(p) for p in gold_points()], T(GOLD["dense_gt"]), T(GOLD["dense_num"]),
                                           STRIDES, SOI, radius)
    same(lab, GOLD[tag + "_labels"])
    same(off, GOLD[tag + "_offsets"])
    same(ctr, GOLD[tag + "_ctrness"])


def test_atss_targets_golden():
    lab, off, ctr = pipelines.atss_targets([T(p) for p in gold_points()], T(GOLD["dense_gt"]), T(GOLD["dense_num"]), STRIDES, 8, 9)
    same(lab, GOLD["atss_labels"])
    same(off, GOLD["atss_offsets"])
    same(ctr, GOLD["atss_ctrness"])


def _case(hw, B, G, seed, ragged=True, tiny=False):
    sizes = W.retinanet_level_sizes(*hw)
    pts = R.anchor_points(sizes, 1, STRIDES, 0.5)
    gt, ng = W.target_assign_batch(B, G, hw[0], hw[1], seed0=seed, ragged=ragged)
    if tiny:  # crowded small objects: many equal-area / equal-IoU ties, points on box borders
        gt[:, :, :4] = np.round(gt[:, :, :4] / 8) * 8
        gt[:, 1::2, :4] = gt[:, 0::2, :4][:, : gt[:, 1::2].shape[1]]  # duplicated boxes -> first index must win
    return pts, gt, ng


@pytest.mark.parametrize("hw,B,G,seed,tiny", [((96, 128), 2, 5, 11, False), ((320, 480), 3, 40, 12, False),
                                               ((320, 480), 2, 40, 13, True), ((800, 1344), 1, 100, 14, False)])
@pytest.mark.parametrize("radius", [1.5, 0])
def test_fcos_targets_vs_oracle(hw, B, G, seed, tiny, radius):
    pts, gt, ng = _case(hw, B, G, seed, tiny=tiny)
    lab, off, ctr, idx = ops.fcos_targets([T(p) for p in pts], T(gt), T(ng), STRIDES, SOI, radius)
    rl, ro, rc, ri = R.fcos_targets(pts, gt, ng, STRIDES, SOI, radius)
    same(idx, ri)
    same(lab, rl)
    same(off, ro)
    same(ctr, rc)
    assert (rl > 0).sum() > 0


@pytest.mark.parametrize("hw,B,G,seed,tiny", [((96, 128), 2, 5, 21, False), ((320, 480), 3, 40, 22, False),
                                               ((320, 480), 2, 40, 23, True), ((800, 1344), 1, 100, 24, False)])
@pytest.mark.parametrize("topk", [9, 1, 16])
def test_atss_targets_vs_oracle(hw, B, G, seed, tiny, topk):
    pts, gt, ng = _case(hw, B, G, seed, tiny=tiny)
    lab, off, ctr, idx = ops.atss_targets([T(p) for p in pts], T(gt), T(ng), STRIDES, 8, topk)
    rl, ro, rc, ri = R.atss_targets(pts, gt, ng, STRIDES, 8, topk)
    same(idx, ri)
    same(lab, rl)
    same(off, ro)
    same(ctr, rc)
    assert (rl > 0).sum() > 0


def test_dense_targets_levels_smaller_than_topk_and_empty_images():
    """P7 of a 96x128 image has 1 point (< topk): k is clamped per level; an image without GT gets all-background."""
    sizes = W.retinanet_level_sizes(96, 128)
    pts = R.anchor_points(sizes, 1, STRIDES, 0.5)
    assert min(len(p) for p in pts) < 9
    gt, ng = W.target_assign_batch(3, 6, 96, 128, seed0=31)
    ng[1] = 0
    for fn, args in ((ops.fcos_targets, (STRIDES, SOI, 1.5)), (ops.atss_targets, (STRIDES, 8, 9))):
        lab, off, ctr, idx = fn([T(p) for p in pts], T(gt), T(ng), *args)
        assert int(lab[1].abs().sum()) == 0 and float(off[1].abs().sum()) == 0.0 and float(ctr[1].abs().sum()) == 0.0
        ref = (R.fcos_targets if fn is ops.fcos_targets else R.atss_targets)(pts, gt[[0, 2]], ng[[0, 2]], *args)
        same(lab[[0, 2]], ref[0])
        same(off[[0, 2]], ref[1])
        same(ctr[[0, 2]], ref[2])


def test_dense_targets_config4_full_size_properties():
    """Config-4 shape (64 images x 22 400 points, 100 GT): size-independent properties + oracle on two images."""
    sizes = W.retinanet_level_sizes(800, 1344)
    pts = R.anchor_points(sizes, 1, STRIDES, 0.5)
    B = 64
    gt, ng = W.target_assign_batch(B, 100, 800, 1344, seed0=4000, ragged=True)
    P = [T(p) for p in pts]
    allp = np.concatenate(pts)
    for name, fn, args, ofn in (("fcos", ops.fcos_targets, (STRIDES, SOI, 1.5), R.fcos_targets),
                                ("atss", ops.atss_targets, (STRIDES, 8, 9), R.atss_targets)):
        lab, off, ctr, idx = [x.cpu().numpy() for x in fn(P, T(gt), T(ng), *args)]
        for b in (0, 37):
            rl, ro, rc, ri = ofn(pts, gt[b:b + 1], ng[b:b + 1], *args)
            assert np.array_equal(lab[b], rl[0]) and np.array_equal(idx[b], ri[0]), name
            assert np.array_equal(off[b], ro[0]) and np.array_equal(ctr[b], rc[0], equal_nan=True), name
        fg = lab > 0
        assert fg.sum() > 1000
        assert np.all(idx < ng[:, None]) and np.all(idx >= 0)
        # foreground points lie strictly inside their matched GT, carry its class, and decode back to it
        m = gt[np.arange(B)[:, None], idx]
        assert np.all(off[fg] > 0)
        assert np.array_equal(lab[fg], m[fg][:, 4].astype(np.int32))
        px = np.broadcast_to(allp[None], (B,) + allp.shape)
        dec = np.stack([px[..., 0] - off[..., 0], px[..., 1] - off[..., 1], px[..., 0] + off[..., 2], px[..., 1] + off[..., 3]], -1)
        assert np.max(np.abs(dec[fg] - m[fg][:, :4])) <= 1e-3
        assert np.all((ctr[fg] > 0) & (ctr[fg] <= 1))


def test_atss_equidistant_points_take_the_index_order_fallback():
    """Hundreds of points at the same distance from a GT centre (duplicates, rings): F.topk ascending keeps the lowest
    indices (ASSUMED-9); exercises the path for > 256 points inside the distance bound."""
    rng = np.random.default_rng(77)
    ring = np.stack([200 + 64 * np.cos(np.linspace(0, 2 * np.pi, 600, endpoint=False)),
                     160 + 64 * np.sin(np.linspace(0, 2 * np.pi, 600, endpoint=False))], 1)
    pts = [np.concatenate([np.full((700, 2), 100.0), rng.uniform(0, 400, (300, 2))]).astype(np.float32),
           np.round(ring).astype(np.float32),
           rng.uniform(0, 400, (50, 2)).astype(np.float32)]
    gt = np.zeros((2, 3, 5), np.float32)
    gt[0, 0] = [60, 60, 140, 140, 3]      # centre exactly on the 700 duplicates
    gt[0, 1] = [136, 96, 264, 224, 5]     # centre (200, 160): the rounded ring is full of equal distances
    gt[0, 2] = [10, 10, 390, 390, 7]
    gt[1, 0] = [90, 90, 110, 110, 2]
    ng = np.array([3, 1], np.int32)
    strides = [8, 16, 32]
    for topk in (9, 16):
        lab, off, ctr, idx = ops.atss_targets([T(p) for p in pts], T(gt), T(ng), strides, 8, topk)
        rl, ro, rc, ri = R.atss_targets(pts, gt, ng, strides, 8, topk)
        same(idx, ri)
        same(lab, rl)
        same(off, ro)
        same(ctr, rc)
