"""GPU parity: anchors, pairwise IoU family, Matcher, coders, fused target assignment vs the oracle.

Bit-exact gates: IoU (expected bit-exact with -fmad=false), matcher indices/labels,
anchors, point coder.  encode/decode: <= 1e-6 relative to max(|ref|, 1) (logf/expf ulp).
"""
import numpy as np
import pytest
import torch

from basedet_b200 import _lib, ops
from basedet_b200 import workloads as W
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def T(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))) if ref.size else 0.0


def box_err(got, ref):
    """Decoded-box error relative to the box's own scale: corners cancel to ~0 (pcx - 0.5*pw), so the
    1e-6 gate of SURVEY 8d is taken relative to max(|coords of that box|, 1), not to the single corner."""
    got = np.asarray(got, dtype=np.float64).reshape(-1, 4)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, 4)
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
    return float(np.max(np.abs(got - ref) / scale)) if ref.size else 0.0


def retina_anchors_np(hw=(800, 800)):
    sizes = W.retinanet_level_sizes(*hw)
    return sizes, R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)


# ------------------------------------------------------------------ anchors
@pytest.mark.parametrize("hw", [(800, 800), (800, 1344), (64, 96)])
def test_anchors_retinanet_bit_exact(cuda, hw):
    sizes, ref = retina_anchors_np(hw)
    base = [R.generate_base_anchors(s, W.RETINANET_RATIOS[0]) for s in W.RETINANET_SCALES]
    got = ops.anchors_grid(sizes, W.RETINANET_STRIDES, [0.5 * s for s in W.RETINANET_STRIDES], base, cuda)
    for g, r in zip(got, ref):
        assert np.array_equal(g.cpu().numpy(), r)


def test_points_grid_bit_exact(cuda):
    sizes = W.retinanet_level_sizes(800, 1344)
    ref = R.anchor_points(sizes, 1, W.RETINANET_STRIDES, 0.5)
    got = ops.points_grid(sizes, W.RETINANET_STRIDES, [0.5 * s for s in W.RETINANET_STRIDES], 1, 0, cuda)
    for g, r in zip(got, ref):
        assert np.array_equal(g.cpu().numpy(), r)
    ref3 = R.anchor_points(sizes[:2], 3, [8, 16], 0.0)
    got3 = ops.points_grid(sizes[:2], [8, 16], [0.0, 0.0], 3, 0, cuda)
    for g, r in zip(got3, ref3):
        assert np.array_equal(g.cpu().numpy(), r)
    sz = [(5, 7), (3, 4)]
    reff = R.fast_points(sz, [8, 16])
    gotf = ops.points_grid(sz, [8, 16], [0.0, 0.0], 1, 1, cuda)
    for g, r in zip(gotf, reff):
        assert np.array_equal(g.cpu().numpy(), r)


# ------------------------------------------------------------------ pairwise
def test_pairwise_reference_kat(cuda):
    from tests.test_oracle_kat import BOXES1, BOXES2

    iou = ops.pairwise(T(BOXES1, cuda), T(BOXES2, cuda)).cpu().numpy()
    assert np.allclose(iou, np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / 1.75]] * 2))
    ioa = ops.pairwise(T(BOXES2, cuda), T(BOXES1, cuda), _lib.PAIR_IOA).cpu().numpy()
    assert np.allclose(ioa, np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25]] * 2).T)
    inter = ops.pairwise(T(BOXES1, cuda), T(BOXES2, cuda), _lib.PAIR_INTER).cpu().numpy()
    assert np.allclose(inter, np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25]] * 2))
    assert np.allclose(ops.box_center(T(BOXES1, cuda)).cpu().numpy(), 0.5)


@pytest.mark.parametrize("n,m", [(1, 1), (3, 5), (100, 4099), (37, 1024), (257, 33)])
def test_pairwise_bit_exact_random(cuda, n, m):
    rng = np.random.default_rng(n * 1000 + m)
    b1 = W.make_gt(rng, n, 800, 800)[:, :4]
    b2 = W.make_gt(rng, m, 800, 800, 8, 300)[:, :4]
    # exact duplicates, touching boxes and degenerate (zero-area) boxes
    b2[0] = b1[0]
    if m > 2:
        b2[1] = [b1[0][2], b1[0][1], b1[0][2] + 10, b1[0][3]]
        b2[2] = [5, 5, 5, 5]
    for mode, fn in ((_lib.PAIR_IOU, R.box_iou), (_lib.PAIR_IOA, R.box_ioa), (_lib.PAIR_INTER, R.box_intersection)):
        got = ops.pairwise(T(b1, cuda), T(b2, cuda), mode).cpu().numpy()
        ref = fn(b1, b2)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), mode
    got = ops.pairwise(T(b1, cuda), T(b2, cuda), _lib.PAIR_GIOU).cpu().numpy()
    ref = R.box_giou(b1, b2)
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), ok)
    assert np.array_equal(got[ok].view(np.uint32), ref[ok].view(np.uint32))


def test_pairwise_strided_gt_view(cuda):
    rng = np.random.default_rng(5)
    gt = W.make_gt(rng, 17, 800, 800)
    anchors = np.concatenate(retina_anchors_np((128, 160))[1])
    g = T(gt, cuda)
    got = ops.pairwise(g[:, :4], T(anchors, cuda)).cpu().numpy()
    assert np.array_equal(got, R.box_iou(gt[:, :4], anchors))


def test_pairwise_config2_shape_bit_exact(cuda):
    """BASELINE config 2 shape for one image: (100, 120087)."""
    anchors = np.concatenate(retina_anchors_np((800, 800))[1])
    gt, _ = W.target_assign_batch(1)
    got = ops.pairwise(T(gt[0, :, :4], cuda), T(anchors, cuda)).cpu().numpy()
    ref = R.box_iou(gt[0, :, :4], anchors)
    assert got.shape == (100, 120087)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_point_distance(cuda):
    rng = np.random.default_rng(3)
    p1 = rng.uniform(0, 800, (33, 2)).astype(np.float32)
    p2 = rng.uniform(0, 800, (517, 2)).astype(np.float32)
    got = ops.point_distance(T(p1, cuda), T(p2, cuda)).cpu().numpy()
    assert rel_err(got, R.point_distance(p1, p2)) <= 1e-6


# ------------------------------------------------------------------ matcher
@pytest.mark.parametrize("g,a,lq", [(1, 1, True), (7, 333, True), (100, 5000, True), (100, 5000, False), (33, 70001, True)])
def test_matcher_bit_exact(cuda, g, a, lq):
    rng = np.random.default_rng(g * 7 + a)
    m = rng.uniform(0, 1, (g, a)).astype(np.float32)
    m[:, rng.integers(0, a, a // 2)] = 0.0          # many all-zero columns -> argmax must be 0
    if g > 2 and a > 10:
        m[1, 3] = m[1].max()                         # tie at a row maximum: both columns are low-quality matches
        m[2, :] = 0.0                                # degenerate GT: every anchor ties at 0
        m[0, 5] = m[min(3, g - 1), 5] = 0.45         # tie inside a column: first index wins
    idx, lab = ops.match(T(m, cuda), [0.4, 0.5], [0, -1, 1], lq)
    ridx, rlab = R.matcher(m, [0.4, 0.5], [0, -1, 1], lq)
    assert np.array_equal(idx.cpu().numpy(), ridx)
    assert np.array_equal(lab.cpu().numpy(), rlab)


def test_matcher_on_real_iou_and_batched(cuda):
    anchors = np.concatenate(retina_anchors_np((256, 320))[1])
    gt, ng = W.target_assign_batch(3, num_gt=20, img_h=256, img_w=320, ragged=True)
    iou = ops.pairwise_batched(T(gt, cuda), T(ng, cuda), T(anchors, cuda))
    idx, lab = ops.match(iou, [0.4, 0.5], [0, -1, 1], True, num_g=T(ng, cuda))
    for b in range(3):
        ref = R.box_iou(gt[b, : ng[b], :4], anchors)
        assert np.array_equal(iou[b, : ng[b]].cpu().numpy(), ref)
        ridx, rlab = R.matcher(ref, [0.4, 0.5], [0, -1, 1], True)
        assert np.array_equal(idx[b].cpu().numpy(), ridx)
        assert np.array_equal(lab[b].cpu().numpy(), rlab)


def test_matcher_threshold_edges(cuda):
    m = np.array([[0.4, 0.5, 0.39999998, 0.49999997, 0.0, 1.0, np.inf]], dtype=np.float32)
    for thr, labs in (([0.4, 0.5], [0, -1, 1]), ([0.3, 0.7], [0, -1, 1]), ([0.5], [0, 1]), ([0.5, 0.5], [0, -1, 1])):
        idx, lab = ops.match(T(m, cuda), thr, labs, False)
        ridx, rlab = R.matcher(m, thr, labs, False)
        assert np.array_equal(lab.cpu().numpy(), rlab), (thr, lab, rlab)


def test_match_rows(cuda):
    rng = np.random.default_rng(11)
    for r, g in ((5, 1), (100, 7), (513, 100), (9, 1000)):
        m = rng.uniform(0, 1, (r, g)).astype(np.float32)
        m[0, :] = 0.0
        if g > 3:
            m[1, 2] = m[1, 3] = 2.0
        mx, am = ops.match_rows(T(m, cuda))
        rmx, ram = R.matcher_rows(m)
        assert np.array_equal(mx.cpu().numpy(), rmx)
        assert np.array_equal(am.cpu().numpy(), ram)


# ------------------------------------------------------------------ coders
def test_box_encode_decode(cuda):
    rng = np.random.default_rng(21)
    anchors = np.concatenate(retina_anchors_np((256, 256))[1])
    n = anchors.shape[0]
    gt = W.make_gt(rng, n, 256, 256)[:, :4]
    for mean, std in (((0, 0, 0, 0), (1, 1, 1, 1)), ((0.0, 0.1, -0.1, 0.0), (0.1, 0.1, 0.2, 0.2))):
        enc = ops.box_encode(T(anchors, cuda), T(gt, cuda), mean, std).cpu().numpy()
        assert rel_err(enc, R.boxcoder_encode(anchors, gt, mean, std)) <= 1e-6
        deltas = W.deltas_level(rng, n)
        d = T(deltas, cuda)
        dec = ops.box_decode(T(anchors, cuda), d, mean, std, writeback=True).cpu().numpy()
        rdec, rd = R.boxcoder_decode(anchors, deltas, mean, std)
        assert box_err(dec, rdec) <= 1e-6
        assert np.array_equal(d.cpu().numpy(), rd)  # in-place rescale, boxcoder.py:76-77
    # gather form == encode(anchors, gt[idx])
    gsmall = W.make_gt(rng, 13, 256, 256)
    idx = rng.integers(0, 13, n).astype(np.int32)
    enc = ops.box_encode(T(anchors, cuda), T(gsmall, cuda)[:, :4], (0, 0, 0, 0), (1, 1, 1, 1), gather_idx=T(idx, cuda))
    assert rel_err(enc.cpu().numpy(), R.boxcoder_encode(anchors, gsmall[idx, :4])) <= 1e-6
    # (N, 4k) deltas and selection decode
    d8 = W.deltas_level(rng, n * 2).reshape(n, 8)
    dec8 = ops.box_decode(T(anchors, cuda), T(d8, cuda), (0, 0, 0, 0), (1, 1, 1, 1)).cpu().numpy()
    assert box_err(dec8, R.boxcoder_decode(anchors, d8)[0]) <= 1e-6
    sel = rng.integers(0, n * 80, 777).astype(np.int32)
    decs = ops.box_decode(T(anchors, cuda), T(deltas, cuda), (0, 0, 0, 0), (1, 1, 1, 1), sel_idx=T(sel, cuda), sel_div=80)
    assert box_err(decs.cpu().numpy(), R.boxcoder_decode(anchors, deltas)[0][sel // 80]) <= 1e-6


def test_point_and_sum_coders_bit_exact(cuda):
    rng = np.random.default_rng(22)
    pts = np.concatenate(R.anchor_points(W.retinanet_level_sizes(256, 320), 1, W.RETINANET_STRIDES, 0.5))
    gt = W.make_gt(rng, 9, 256, 320)
    enc = ops.point_encode(T(pts, cuda), T(gt, cuda)[:, :4]).cpu().numpy()
    assert np.array_equal(enc, R.pointcoder_encode(pts, gt[:, None, :4]))
    d = np.abs(rng.normal(0, 30, (pts.shape[0], 4))).astype(np.float32)
    dec = ops.point_decode(T(pts, cuda), T(d, cuda)).cpu().numpy()
    assert np.array_equal(dec, R.pointcoder_decode(pts, d))
    a = W.make_gt(rng, 100, 256, 320)[:, :4]
    g = W.make_gt(rng, 100, 256, 320)[:, :4]
    mean, std = (0.0, 0.5, 0.0, -0.5), (0.1, 0.2, 0.3, 0.4)
    assert np.array_equal(ops.sum_encode(T(a, cuda), T(g, cuda), mean, std).cpu().numpy(), R.sumcoder_encode(a, g, mean, std))
    dd = T(d[:100], cuda)
    out = ops.sum_decode(T(a, cuda), dd, mean, std, writeback=True).cpu().numpy()
    rout, rd = R.sumcoder_decode(a, d[:100], mean, std)
    assert np.array_equal(out, rout) and np.array_equal(dd.cpu().numpy(), rd)


def test_scale_clip_filter(cuda):
    rng = np.random.default_rng(23)
    b = (W.make_gt(rng, 300, 800, 1344)[:, :4] + rng.normal(0, 40, (300, 4))).astype(np.float32)
    t = T(b, cuda)
    ops.boxes_scale_clip(t, 0.7, 1.3, 500.0, 375.0)
    ref = R.boxes_clip(R.boxes_scale(b, (1.3, 0.7)), (375.0, 500.0))
    assert np.array_equal(t.cpu().numpy(), ref)
    keep = ops.boxes_filter_by_size(T(ref, cuda)).cpu().numpy()
    assert np.array_equal(keep, R.boxes_filter_by_size(ref))


# ------------------------------------------------------------------ fused target assignment
def _check_assign(cuda, anchors, gt, ng, matcher, apply_class=True, mean=(0, 0, 0, 0), std=(1, 1, 1, 1)):
    lab, idx, off = ops.assign_targets(T(anchors, cuda), T(gt, cuda), T(ng, cuda), matcher["thresholds"], matcher["labels"],
                                       matcher["allow_low_quality"], apply_class, mean, std)
    lab, idx, off = lab.cpu().numpy(), idx.cpu().numpy(), off.cpu().numpy()
    for b in range(gt.shape[0]):
        g = gt[b, : ng[b]]
        iou = R.box_iou(g[:, :4], anchors)
        ridx, rlab = R.matcher(iou, matcher["thresholds"], matcher["labels"], matcher["allow_low_quality"])
        if apply_class:
            fg = rlab == 1
            rlab[fg] = g[ridx][fg, 4].astype(np.int32)
        roff = R.boxcoder_encode(anchors, g[ridx, :4], mean, std)
        assert np.array_equal(idx[b], ridx), b
        assert np.array_equal(lab[b], rlab), b
        ok = np.isfinite(roff)
        assert np.array_equal(np.isfinite(off[b]), ok)
        assert rel_err(off[b][ok], roff[ok]) <= 1e-6


def test_assign_targets_small_ragged(cuda):
    anchors = np.concatenate(retina_anchors_np((256, 320))[1])
    gt, ng = W.target_assign_batch(4, num_gt=30, img_h=256, img_w=320, ragged=True)
    _check_assign(cuda, anchors, gt, ng, W.RETINANET_MATCHER)
    _check_assign(cuda, anchors, gt, ng, W.RPN_MATCHER, apply_class=False)
    _check_assign(cuda, anchors, gt, ng, dict(thresholds=[0.5], labels=[0, 1], allow_low_quality=False),
                  mean=(0.0, 0.0, 0.1, 0.1), std=(0.1, 0.1, 0.2, 0.2))


def test_assign_targets_degenerate_gt(cuda):
    """Zero-area / far-away GT rows have row maximum 0: every zero-IoU anchor becomes positive;
    duplicate GTs tie everywhere: first index wins."""
    anchors = np.concatenate(retina_anchors_np((128, 128))[1])
    gt = np.zeros((2, 6, 5), dtype=np.float32)
    gt[0, 0] = [10, 10, 60, 70, 3]
    gt[0, 1] = [10, 10, 60, 70, 9]           # duplicate of row 0
    gt[0, 2] = [5000, 5000, 5100, 5100, 4]   # overlaps nothing
    gt[0, 3] = [30, 30, 30, 30, 5]           # zero area
    gt[1, 0] = [0, 0, 128, 128, 7]
    gt[1, 1] = [20, 30, 90, 100, 2]
    ng = np.array([4, 2], dtype=np.int32)
    _check_assign(cuda, anchors, gt, ng, W.RETINANET_MATCHER)


def test_assign_targets_config2_one_image(cuda):
    anchors = np.concatenate(retina_anchors_np((800, 800))[1])
    gt, ng = W.target_assign_batch(2)
    _check_assign(cuda, anchors, gt, ng, W.RETINANET_MATCHER)


@pytest.mark.parametrize("B,A", [(1, 1), (3, 7), (2, 4099), (16, 120087), (5, 8192)])
def test_count_labels_census(cuda, B, A):
    """bdet_count_labels: per-image (#<0, #==0, #>0), rows at any alignment (retinanet.py:142-146 num_fg)."""
    rng = np.random.default_rng(B * A)
    lab = rng.integers(-1, 81, (B, A)).astype(np.int32)
    lab[rng.random((B, A)) < 0.7] = 0
    got = ops.count_labels(torch.from_numpy(lab).to(cuda)).cpu().numpy()
    ref = np.stack([(lab < 0).sum(1), (lab == 0).sum(1), (lab > 0).sum(1)], 1)
    assert np.array_equal(got, ref)


# ------------------------------------------------------------------ sample_labels / RPN.get_ground_truth
@pytest.mark.parametrize("A,B,ns,quant", [(50, 2, 8, 0), (5000, 3, 128, 0), (5000, 3, 128, 16), (268569, 2, 256, 0),
                                           (268569, 2, 256, 64), (4096, 4, 0, 0), (300, 2, 1000, 0)])
def test_sample_labels_bit_exact(cuda, A, B, ns, quant):
    """layers/common/sampling.py:7-30 with explicit variates: the kept / ignored sets are exactly the oracle's,
    including ties between equal variates (lower index is ignored first) and budgets above the population."""
    from basedet_b200 import workloads as W  # noqa: F401
    rng = np.random.default_rng(A + ns + quant)
    lab = rng.choice(np.array([-1, 0, 1], np.int32), size=(B, A), p=[0.1, 0.8, 0.1]).astype(np.int32)
    noise = rng.uniform(0, 1, (B, A)).astype(np.float32)
    if quant:
        noise = (np.floor(noise * quant) / quant).astype(np.float32)   # many equal variates, some exactly 0
    for value, ignore in ((0, -1), (1, -1), (1, 0)):
        ref = np.stack([R.sample_labels(lab[b], ns, value, ignore, noise[b]) for b in range(B)])
        got = ops.sample_labels(torch.from_numpy(lab.copy()).to(cuda), torch.from_numpy(noise).to(cuda), ns, value, ignore)
        assert np.array_equal(got.cpu().numpy(), ref)
    budgets = rng.integers(0, 400, B).astype(np.int32)               # per-image budgets from a device tensor
    ref = np.stack([R.sample_labels(lab[b], int(budgets[b]), 0, -1, noise[b]) for b in range(B)])
    got = ops.sample_labels(torch.from_numpy(lab.copy()).to(cuda), torch.from_numpy(noise).to(cuda),
                            torch.from_numpy(budgets).to(cuda), 0, -1)
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("hw,B,G", [((96, 128), 2, 6), ((320, 480), 3, 25), ((800, 1344), 2, 60)])
def test_rpn_get_ground_truth_with_sampling(cuda, hw, B, G):
    """RPN.get_ground_truth (rpn.py:215-240): fused match/encode + two sample_labels calls == oracle."""
    from basedet_b200 import pipelines
    from basedet_b200 import workloads as W
    sizes = W.frcnn_level_sizes(*hw)
    anchors = np.concatenate(R.default_anchors(sizes, W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5))
    gt, ng = W.target_assign_batch(B, G, hw[0], hw[1], seed0=hw[0], ragged=True)
    rng = np.random.default_rng(hw[1])
    A = anchors.shape[0]
    npz, nnz = rng.uniform(0, 1, (B, A)).astype(np.float32), rng.uniform(0, 1, (B, A)).astype(np.float32)
    Tc = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda)  # noqa: E731
    for total, ratio in ((256, 0.5), (64, 0.25)):
        lab, off = pipelines.rpn_targets(Tc(anchors), Tc(gt), Tc(ng), Tc(npz), Tc(nnz), (0.3, 0.7), (0, -1, 1), True, total, ratio)
        rl, ro = R.rpn_targets(anchors, gt, ng, [0.3, 0.7], [0, -1, 1], True, total, int(ratio * total), npz, nnz)
        assert np.array_equal(lab.cpu().numpy(), rl)
        assert np.max(np.abs(off.cpu().numpy() - ro)) <= 1e-6
        assert np.all((rl == 1).sum(1) <= int(ratio * total)) and np.all((rl >= 0).sum(1) <= total)


@pytest.mark.parametrize("B,Rmax,G,num,ratio", [(2, 100, 5, 32, 0.25), (4, 1000, 40, 512, 0.5), (16, 2000, 100, 512, 0.5)])
def test_rcnn_get_ground_truth_vs_oracle(cuda, B, Rmax, G, num, ratio):
    """RCNN.get_ground_truth (rcnn.py:95-147) at proposal counts up to config 3 (16 x 2000 proposals, 100 GT)."""
    from basedet_b200 import pipelines
    from basedet_b200 import workloads as W
    rng = np.random.default_rng(B * Rmax + G)
    gt, ng = W.target_assign_batch(B, G, 800, 1344, seed0=Rmax, ragged=True)
    rois = np.zeros((B, Rmax, 5), np.float32)
    cnt = rng.integers(Rmax // 2, Rmax + 1, B).astype(np.int32)
    cnt[0] = Rmax
    for b in range(B):
        r = W.make_rois(rng, int(cnt[b]), 1, 800, 1344, 8, 600)
        r[:, 0] = b
        near = min(int(cnt[b]) // 3, 400)
        j = rng.integers(0, ng[b], near)
        r[:near, 1:] = gt[b, j, :4] + rng.normal(0, 6, (near, 4)).astype(np.float32)
        rois[b, : cnt[b]] = r
    N = Rmax + G
    nfg, nbg = rng.uniform(0, 1, (B, N)).astype(np.float32), rng.uniform(0, 1, (B, N)).astype(np.float32)
    Tc = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda)  # noqa: E731
    orois, olab, otgt, ocnt = pipelines.rcnn_targets(Tc(rois), Tc(cnt), Tc(gt), Tc(ng), Tc(nfg), Tc(nbg), num, ratio)
    ocnt = ocnt.cpu().numpy()
    ref = R.rcnn_targets([rois[b, : cnt[b]] for b in range(B)], gt, ng, [nfg[b, : cnt[b] + ng[b]] for b in range(B)],
                         [nbg[b, : cnt[b] + ng[b]] for b in range(B)], num, ratio)
    for b in range(B):
        rr, rl, rt = ref[b]
        assert ocnt[b] == len(rl)
        assert np.array_equal(orois[b, : ocnt[b]].cpu().numpy(), rr)
        assert np.array_equal(olab[b, : ocnt[b]].cpu().numpy(), rl)
        # logf differs by <= 1 ulp between CUDA and libm and the result is divided by std = 0.1 / 0.2; degenerate
        # proposals (non-positive width after the jitter) give inf / NaN targets in both
        assert np.allclose(otgt[b, : ocnt[b]].cpu().numpy(), rt, rtol=2e-6, atol=1e-5, equal_nan=True)
        assert (rl > 0).sum() <= int(num * ratio) and len(rl) <= num


def test_new_target_entry_points_degenerate_inputs(cuda):
    """Empty / degenerate shapes of the 8(f) entry points: no GT, no proposals, zero budgets, one element."""
    from basedet_b200 import pipelines
    Tc = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda)  # noqa: E731
    # sample_labels: nothing to do / everything ignored
    lab = np.array([[1, 0, 1, -1, 0]], np.int32)
    nz = np.array([[0.5, 0.1, 0.9, 0.3, 0.2]], np.float32)
    assert np.array_equal(ops.sample_labels(Tc(lab.copy()), Tc(nz), 5, 1, -1).cpu().numpy(), lab)
    assert np.array_equal(ops.sample_labels(Tc(lab.copy()), Tc(nz), 0, 1, -1).cpu().numpy(), R.sample_labels(lab[0], 0, 1, -1, nz[0])[None])
    assert np.array_equal(ops.sample_labels(Tc(lab.copy()), Tc(nz), 1, 0, 7).cpu().numpy(), R.sample_labels(lab[0], 1, 0, 7, nz[0])[None])
    # RCNN glue: an image without GT contributes nothing, an image without proposals still keeps its GT rows
    gt = np.zeros((2, 3, 5), np.float32)
    gt[1, 0] = [10, 10, 60, 70, 4]
    gt[1, 1] = [100, 20, 180, 90, 9]
    ng = np.array([0, 2], np.int32)
    rois = np.zeros((2, 6, 5), np.float32)
    rois[0, :, 1:] = [5, 5, 50, 50]
    nr = np.array([6, 0], np.int32)
    nz2 = np.random.default_rng(0).uniform(0, 1, (2, 9)).astype(np.float32)
    orois, olab, otgt, cnt = pipelines.rcnn_targets(Tc(rois), Tc(nr), Tc(gt), Tc(ng), Tc(nz2), Tc(nz2), 8, 0.5)
    cnt = cnt.cpu().numpy()
    assert cnt[0] == 0
    ref = R.rcnn_targets([rois[1, :0]], gt[1:], ng[1:], [nz2[1, :2]], [nz2[1, :2]], 8, 0.5)[0]
    assert cnt[1] == len(ref[1]) == 2
    assert np.array_equal(orois[1, :2].cpu().numpy()[:, 1:], ref[0][:, 1:]) and np.array_equal(olab[1, :2].cpu().numpy(), ref[1])
    assert np.allclose(otgt[1, :2].cpu().numpy(), ref[2], atol=1e-6)


@pytest.mark.parametrize("G,A,q,kc", [(1, 40, 0, 10), (12, 9, 0, 10), (50, 22400, 0, 10), (50, 22400, 16, 10), (100, 8400, 0, 16),
                                      (30, 5000, 0, 1), (8, 3000, -1, 10)])
def test_ota_topk_match_vs_oracle(cuda, G, A, q, kc):
    """Dynamic-k matching (matcher.py:134-161) at OTA / YOLOX sizes: ties from quantised inputs, anchors claimed by
    several GTs, rows shorter than candidate_k, and (q = -1) a cost row whose small values all sit in one thread's
    strided slice (the > 256-keys-inside-the-bound path)."""
    rng = np.random.default_rng(G * A + kc)
    ious = (rng.uniform(0, 1, (G, A)) ** 3).astype(np.float32)
    cost = rng.uniform(0, 5, (G, A)).astype(np.float32)
    if q > 0:
        ious, cost = (np.floor(ious * q) / q).astype(np.float32), (np.floor(cost * q) / q).astype(np.float32)
    if q < 0:
        cost[:, 5::256] = rng.uniform(-9, -8, cost[:, 5::256].shape).astype(np.float32)   # slice of thread 5 only... 
        cost[:, 261::256] -= 3.0
        ious[:, 7::256] = 0.99
    cost[:, ::7] += 1e6
    got = ops.ota_topk_match(torch.from_numpy(cost).to(cuda), torch.from_numpy(ious).to(cuda), kc).cpu().numpy()
    ref = R.ota_topk_match(cost, ious, kc)
    assert np.array_equal(got, ref)
    assert (ref < G).sum() >= 1


def test_assign_targets_grid_equals_tensor_path(cuda):
    """bdet_assign_targets_grid (anchors generated in registers + label census folded into the kernels) is bit-identical
    to bdet_assign_targets on the materialised anchors, for RetinaNet (class labels) and RPN (no class) settings, ragged
    GT counts and an image without GT; the census equals count_labels."""
    from basedet_b200.layers import DefaultAnchorGenerator

    for scales, ratios, strides, sizes, thr, cls in (
            (W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, W.retinanet_level_sizes(256, 320), [0.4, 0.5], True),
            (W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, W.frcnn_level_sizes(192, 256), [0.3, 0.7], False)):
        gen = DefaultAnchorGenerator(scales, ratios, strides, 0.5)
        anchors = gen.generate_all_level_anchors(sizes, cuda)
        gt, ng = W.target_assign_batch(5, num_gt=20, img_h=sizes[0][0] * strides[0], img_w=sizes[0][1] * strides[0], ragged=True)
        ng[3] = 0
        gt_d, ng_d = torch.from_numpy(gt).to(cuda), torch.from_numpy(ng).to(cuda)
        ref = [t.clone() for t in ops.assign_targets(anchors, gt_d, ng_d, thr, [0, -1, 1], True, cls)]
        counts = torch.full((5, 3), -7, dtype=torch.int32, device=cuda)
        got = ops.assign_targets_grid(gen._plan(sizes), gt_d, ng_d, thr, [0, -1, 1], True, cls, counts=counts)
        for g_, r_ in zip(got, ref):
            assert torch.equal(g_, r_)
        assert torch.equal(counts, ops.count_labels(ref[0]))
        assert int(counts.sum()) == 5 * anchors.shape[0]
