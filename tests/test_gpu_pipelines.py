"""GPU parity of the batched pipelines (configs 1, 3, 4, 5 of BASELINE.json) against the oracle composed exactly
like the reference's callers, at reduced AND full sizes; full-size checks that the oracle cannot finish in seconds
use size-independent properties."""
import numpy as np
import pytest
import torch

from basedet_b200 import _lib, ops, pipelines
from basedet_b200 import workloads as W
from oracle import c_oracle as C
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def box_close(got, ref, tol=1e-6):
    got, ref = np.asarray(got, np.float64).reshape(-1, 4), np.asarray(ref, np.float64).reshape(-1, 4)
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
    return np.max(np.abs(got - ref) / scale, initial=0.0) <= tol


def retina_inputs(rng, B, hw, C=80, mean=-6.0):
    sizes = W.retinanet_level_sizes(*hw)
    anchors = R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
    logits = [np.stack([W.logits_level(rng, h * w * 9, C, mean) for _ in range(B)]) for h, w in sizes]
    deltas = [np.stack([W.deltas_level(rng, h * w * 9) for _ in range(B)]) for h, w in sizes]
    return sizes, anchors, logits, deltas


def check_dense(B, hw, C=80, mean=-6.0, seed=0):
    rng = np.random.default_rng(seed)
    sizes, anchors, logits, deltas = retina_inputs(rng, B, hw, C, mean)
    info = np.array([[hw[0], hw[1], 612.0 + 7 * b, 612.0 + 11 * b, 0.0] for b in range(B)], np.float32)
    dets, cnt = pipelines.dense_postprocess([T(x) for x in logits], [T(x) for x in deltas], [T(a) for a in anchors], T(info),
                                            0.05, 0.5, 100, 1000)
    dets, cnt = dets.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        sc = [ops.scores(T(l[b]).reshape(-1)).cpu().numpy() for l in logits]  # bit-identical score tensors (H9)
        rb, rs, rl, _ = R.retinanet_postprocess([l[b] for l in logits], [d[b] for d in deltas], anchors, info[b:b + 1], 0.05, 0.5,
                                                100, 1000, scores_list=sc)
        n = len(rs)
        assert cnt[b] == n, (b, cnt[b], n)
        assert np.array_equal(dets[b, :n, 4], rs)
        assert np.array_equal(dets[b, :n, 5].astype(np.int32), rl)
        assert box_close(dets[b, :n, :4], rb)
        assert not dets[b, n:].any()
    return cnt


def test_retinanet_postprocess_small_batch():
    cnt = check_dense(3, (256, 320), mean=-5.0)
    assert cnt.min() > 0


def test_retinanet_postprocess_config1_full_size():
    """BASELINE config 1: one 800x800 image, 120 087 anchors x 80 classes, top-1000/level, NMS 0.5, 100 dets."""
    check_dense(1, (800, 800), seed=1)


def test_retinanet_postprocess_empty_levels():
    check_dense(2, (128, 128), mean=-12.0)  # nothing above the threshold anywhere -> empty detections


def test_fcos_postprocess_batch():
    """models/det/fcos.py:191-221 + post_processing with NMS 0.6 (fcos_cfg.py:56)."""
    rng = np.random.default_rng(2)
    B, C, hw = 3, 80, (256, 320)
    sizes = W.retinanet_level_sizes(*hw)
    points = R.anchor_points(sizes, 1, W.RETINANET_STRIDES, 0.5)
    logits = [np.stack([W.logits_level(rng, h * w, C, -4.5) for _ in range(B)]) for h, w in sizes]
    ctr = [rng.normal(0, 1, (B, h * w, 1)).astype(np.float32) for h, w in sizes]
    ltrb = [(np.abs(rng.normal(0, 1, (B, h * w, 4))) * s * 4).astype(np.float32) for (h, w), s in zip(sizes, W.RETINANET_STRIDES)]
    info = np.array([[hw[0], hw[1], 480.0, 600.0, 0.0]] * B, np.float32)
    dets, cnt = pipelines.dense_postprocess([T(x) for x in logits], [T(x) for x in ltrb], [T(p) for p in points], T(info), 0.05,
                                            0.6, 100, 1000, ctrness_list=[T(c) for c in ctr])
    dets, cnt = dets.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        tb, ts, tl = [], [], []
        for l in range(len(sizes)):
            s = ops.scores(T(logits[l][b]), _lib.SCORE_FCOS, T(ctr[l][b]), C).cpu().numpy().reshape(-1)
            keep, sc = R.filter_topk_scores(s, 0.05, 1000)
            if keep.size == 0:
                continue
            boxes = R.pointcoder_decode(points[l], ltrb[l][b])
            tb.append(boxes[keep // C]); ts.append(sc); tl.append((keep % C).astype(np.int32))
        rb, rs, rl, _ = R.post_processing(np.concatenate(tb), np.concatenate(ts), np.concatenate(tl), info[b:b + 1], 0.6, 100)
        n = len(rs)
        assert cnt[b] == n
        assert np.array_equal(dets[b, :n, 4], rs) and np.array_equal(dets[b, :n, 5].astype(np.int32), rl)
        assert np.array_equal(dets[b, :n, :4], rb)  # PointCoder: add/sub only -> bit-exact


def oracle_rpn(scores_l, deltas_l, anchors_l, im_hw, pre_k, post_k, thr):
    """models/det/rpn.py:141-186 for one image (proposals keep their un-clipped coordinates)."""
    props, scs, lvls = [], [], []
    for level, (s, d, a) in enumerate(zip(scores_l, deltas_l, anchors_l)):
        boxes, _ = R.boxcoder_decode(a, d)
        v, order = R.topk_desc(s, pre_k)
        props.append(boxes[order]); scs.append(v); lvls.append(np.full(len(v), level, np.float32))
    props, scs, lvls = np.concatenate(props), np.concatenate(scs), np.concatenate(lvls)
    keep_mask = R.boxes_filter_by_size(R.boxes_clip(props, im_hw))
    props, scs, lvls = props[keep_mask], scs[keep_mask], lvls[keep_mask]
    keep = R.batched_nms(props, scs, lvls, thr, post_k)
    return props[keep]


@pytest.mark.parametrize("hw,B,pre_k,post_k", [((128, 160), 2, 300, 100), ((800, 1344), 1, 2000, 1000)])
def test_rpn_proposals(hw, B, pre_k, post_k):
    """Config 3 (second case = full size: 268 569 anchors, top-2000/level, NMS 0.7, 1000 proposals)."""
    rng = np.random.default_rng(3)
    sizes = W.frcnn_level_sizes(*hw)
    anchors = R.default_anchors(sizes, W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
    n_l = [a.shape[0] for a in anchors]
    scores = [np.stack([W.distinct_scores(rng, n, -9.0, 3.0) for _ in range(B)]) for n in n_l]
    deltas = [np.stack([(rng.normal(0, 0.2, (n, 4))).astype(np.float32) for _ in range(B)]) for n in n_l]
    info = np.array([[hw[0], hw[1], hw[0], hw[1], 0.0]] * B, np.float32)
    rois, cnt = pipelines.rpn_proposals([T(s) for s in scores], [T(d) for d in deltas], [T(a) for a in anchors], T(info), pre_k,
                                        post_k, 0.7)
    rois, cnt = rois.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        ref = oracle_rpn([s[b] for s in scores], [d[b] for d in deltas], anchors, hw, pre_k, post_k, 0.7)
        n = ref.shape[0]
        assert cnt[b] == n, (cnt[b], n)
        assert (rois[b, :n, 0] == b).all()
        # decode goes through expf: boxes within 1e-6 (box scale); the NMS decisions themselves are asserted by the count
        # and by the position-wise agreement of every kept box
        assert box_close(rois[b, :n, 1:], ref, 2e-6)


def test_roi_pool_config3_full_size_one_image():
    """512 ROIs x 256 channels x 7x7 on the 800x1344 pyramid of one image, forward + backward vs the C oracle."""
    rng = np.random.default_rng(4)
    hw, Cn, K = (800, 1344), 256, 512
    sizes = [(-(-hw[0] // s), -(-hw[1] // s)) for s in W.FRCNN_RCNN_STRIDES]
    feats = [rng.normal(0, 1, (1, Cn, h, w)).astype(np.float32) for h, w in sizes]
    rois = W.make_rois(rng, K, 1, hw[0], hw[1], 8, 600)
    dout = rng.normal(0, 1, (K, Cn, 7, 7)).astype(np.float32)
    out, grads = pipelines.roi_pool_forward_backward([T(f) for f in feats], T(rois), W.FRCNN_RCNN_STRIDES, (7, 7), T(dout))
    out = out.cpu().numpy()
    levels = R.assign_levels(rois, W.FRCNN_RCNN_STRIDES)
    assert np.array_equal(ops.roi_assign_levels(T(rois), 2, 5).cpu().numpy(), levels)
    for l, f in enumerate(feats):
        sel = np.flatnonzero(levels == l)
        ref = C.roi_align_fwd(f, rois[sel], (7, 7), 1.0 / W.FRCNN_RCNN_STRIDES[l])
        assert np.max(np.abs(out[sel] - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-5
        gref = C.roi_align_bwd(dout[sel], f.shape, rois[sel], (7, 7), 1.0 / W.FRCNN_RCNN_STRIDES[l])  # fp64-accumulated
        g = grads[l].cpu().numpy()
        assert np.max(np.abs(g - gref)) / max(np.abs(gref).max(), 1.0) <= 1e-5, l


def test_stress_iou_200k_x_500_properties():
    """Config 5 IoU at full size: spot rows against the oracle + size-independent properties."""
    rng = np.random.default_rng(5)
    a = W.make_gt(rng, 200000, 800, 1333, 8, 128)[:, :4]
    g = W.make_gt(rng, 500, 800, 1333)[:, :4]
    iou = ops.pairwise(T(g), T(a))
    rows = rng.integers(0, 500, 6)
    assert np.array_equal(iou[rows.tolist()].cpu().numpy(), R.box_iou(g[rows], a))
    assert float(iou.min()) >= 0.0 and float(iou.max()) <= 1.0
    sym = ops.pairwise(T(a[:4000]), T(g))          # IoU is symmetric in its arguments (commutative fp32 ops)
    assert torch.equal(sym.t().contiguous(), iou[:, :4000].contiguous())
    assert torch.equal(ops.pairwise(T(g), T(g)).diagonal(), torch.ones(500, device="cuda"))  # iou(b, b) == 1


def test_stress_nms_100k_properties():
    """Config 5 NMS (100 000 boxes, single class, no max_output): exact oracle equality is checked at N = 20 000
    (test_gpu_postprocess); here the full size is checked through properties of greedy NMS."""
    rng = np.random.default_rng(6)
    n = 100000
    boxes = W.make_gt(rng, n, 800, 1333, 8, 128)[:, :4]
    scores = W.distinct_scores(rng, n)
    keep, cnt = ops.nms_batched(T(boxes)[None], T(scores)[None], None, 0.5, None)
    k = keep[0, : int(cnt[0])].cpu().numpy()
    assert len(np.unique(k)) == len(k)
    assert np.all(np.diff(scores[k]) < 0)                      # score-descending
    kb, ks = boxes[k], scores[k]
    keep2, cnt2 = ops.nms_batched(T(kb)[None], T(ks)[None], None, 0.5, None)   # idempotence
    assert int(cnt2[0]) == len(k) and np.array_equal(keep2[0, : len(k)].cpu().numpy(), np.arange(len(k)))
    kept_mask = np.zeros(n, bool)
    kept_mask[k] = True
    removed = rng.choice(np.flatnonzero(~kept_mask), 300, replace=False)
    for r in removed:                                           # every removed box is suppressed by a better kept box
        better = k[scores[k] > scores[r]]
        assert (R.box_iou(boxes[r:r + 1], boxes[better])[0] > np.float32(0.5)).any()
    for r in rng.choice(k, 300, replace=False):                 # no kept box is suppressed by a better kept box
        better = k[scores[k] > scores[r]]
        assert not (R.box_iou(boxes[r:r + 1], boxes[better])[0] > np.float32(0.5)).any()


def test_target_assigner_graph_replay_matches_oracle():
    """pipelines.TargetAssigner (the step captured in a CUDA graph, host-pinned inputs) == eager == oracle, over
    several replays with different ground truth (ragged num_gt included)."""
    from basedet_b200.layers import DefaultAnchorGenerator
    hw, B, G = (256, 320), 3, 24
    sizes = W.retinanet_level_sizes(*hw)
    gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
    anchors = np.concatenate(R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5))
    ta = pipelines.TargetAssigner(gen, sizes, B, G)
    for it in range(3):
        gt, ng = W.target_assign_batch(B, G, hw[0], hw[1], seed0=500 + 10 * it, ragged=(it > 0))
        gt_h, ng_h = torch.from_numpy(gt).pin_memory(), torch.from_numpy(ng).pin_memory()
        lab, idx, off, counts = ta(gt_h, ng_h)
        torch.cuda.synchronize()
        rl, ro, ri = R.retinanet_targets(anchors, gt, ng, [0.4, 0.5], [0, -1, 1], True)
        assert np.array_equal(lab.cpu().numpy(), rl)
        assert np.array_equal(idx.cpu().numpy(), ri)
        assert np.max(np.abs(off.cpu().numpy() - ro)) <= 1e-6
        c = counts.cpu().numpy()
        assert np.array_equal(c, np.stack([(rl < 0).sum(1), (rl == 0).sum(1), (rl > 0).sum(1)], 1))
        el, ei, eo = ops.assign_targets(T(anchors), T(gt), T(ng), [0.4, 0.5], [0, -1, 1], True, True)
        assert torch.equal(el, lab) and torch.equal(ei, idx) and torch.equal(eo, off)


@pytest.mark.parametrize("fcos", [False, True])
def test_dense_postprocess_from_nchw_head_outputs(fcos):
    """reading (B, A*C, H, W) head outputs directly == permute_to_N_Any_K (function.py:26-32) first."""
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    B, C, hw = 3, 80, (320, 416)
    sizes = W.retinanet_level_sizes(*hw)
    A = 1 if fcos else 9
    if fcos:
        anchors = [T(p) for p in R.anchor_points(sizes, 1, W.RETINANET_STRIDES, 0.5)]
    else:
        anchors = [T(a) for a in R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)]
    logits = [torch.randn((B, A * C, h, w), device="cuda", generator=g) * 1.25 - (4.0 if fcos else 6.0) for h, w in sizes]
    offsets = [torch.randn((B, A * 4, h, w), device="cuda", generator=g).abs() * (8 if fcos else 0.15) for h, w in sizes]
    ctr = [torch.randn((B, A, h, w), device="cuda", generator=g) for h, w in sizes] if fcos else None
    info = T(np.array([[hw[0], hw[1], 400.0, 520.0, 0.0]] * B, np.float32))

    def perm(t, k):  # permute_to_N_Any_K
        n, _, h, w = t.shape
        return t.reshape(n, -1, k, h, w).permute(0, 3, 4, 1, 2).reshape(n, -1, k).contiguous()

    d0, c0 = pipelines.dense_postprocess([perm(x, C) for x in logits], [perm(x, 4) for x in offsets], anchors, info, 0.05, 0.6, 100, 1000,
                                         ctrness_list=[perm(x, 1) for x in ctr] if fcos else None)
    d1, c1 = pipelines.dense_postprocess_nchw(logits, offsets, anchors, info, C, 0.05, 0.6, 100, 1000, head_ctrness=ctr)
    assert torch.equal(c0, c1) and int(c0.sum()) > 50
    assert torch.equal(d0, d1)


def test_graphed_pipelines_replay_equals_eager():
    """pipelines.GraphedPipeline: dense post-processing and RPN proposals captured once, replayed on new inputs written
    into the same tensors == eager results (segment tables travel as kernel parameters, so the capture is legal)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    B, C, hw = 2, 80, (256, 320)
    sizes = W.retinanet_level_sizes(*hw)
    anchors = [T(a) for a in R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)]
    logits = [torch.randn((B, h * w * 9, C), device="cuda", generator=g) * 1.25 - 6.0 for h, w in sizes]
    offsets = [torch.randn((B, h * w * 9, 4), device="cuda", generator=g) * 0.15 for h, w in sizes]
    info = T(np.array([[hw[0], hw[1], 300.0, 380.0, 0.0]] * B, np.float32))
    post = pipelines.GraphedPipeline(pipelines.dense_postprocess, logits, offsets, anchors, info, 0.05, 0.5, 100, 1000)
    for it in range(3):
        for t in logits:
            t.copy_(torch.randn(t.shape, device="cuda", generator=g) * 1.25 - 6.0)
        for t in offsets:
            t.copy_(torch.randn(t.shape, device="cuda", generator=g) * 0.15)
        dets, cnt = post.replay()
        ed, ec = pipelines.dense_postprocess(logits, offsets, anchors, info, 0.05, 0.5, 100, 1000)
        assert torch.equal(cnt, ec) and torch.equal(dets, ed) and int(cnt.sum()) > 0
    fs = W.frcnn_level_sizes(256, 320)
    ranchors = [T(a) for a in R.default_anchors(fs, W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)]
    sc = [torch.randn((B, a.shape[0]), device="cuda", generator=g) * 2 - 3 for a in ranchors]
    dl = [torch.randn((B, a.shape[0], 4), device="cuda", generator=g) * 0.2 for a in ranchors]
    rinfo = T(np.array([[256, 320, 256, 320, 0.0]] * B, np.float32))
    rpn = pipelines.GraphedPipeline(pipelines.rpn_proposals, sc, dl, ranchors, rinfo, 2000, 1000, 0.7)
    for it in range(2):
        for t in sc:
            t.copy_(torch.randn(t.shape, device="cuda", generator=g) * 2 - 3)
        rois, cnt = rpn.replay()
        er, ec = pipelines.rpn_proposals(sc, dl, ranchors, rinfo, 2000, 1000, 0.7)
        assert torch.equal(cnt, ec) and torch.equal(rois, er)


@pytest.mark.parametrize("fcos", [False, True])
def test_dense_tail_kernel_equals_separate_kernels(fcos):
    """bdet_dense_tail (decode + sort + NMS + finalize in one kernel per image) returns bit-identical detections to the
    four separate launches, RetinaNet and FCOS flavours, ragged candidate counts and an image with an empty level."""
    rng = np.random.default_rng(21)
    B, hw, C = 4, (256, 320), 80
    if fcos:
        sizes = W.retinanet_level_sizes(*hw)
        anchors = [T(p) for p in R.anchor_points(sizes, 1, W.RETINANET_STRIDES, 0.5)]
        logits = [T(rng.normal(-4.0, 1.5, (B, h * w, C)).astype(np.float32)) for h, w in sizes]
        ctr = [T(rng.normal(0, 1, (B, h * w, 1)).astype(np.float32)) for h, w in sizes]
        offs = [T((np.abs(rng.normal(0, 1, (B, h * w, 4))) * s * 3).astype(np.float32)) for (h, w), s in zip(sizes, W.RETINANET_STRIDES)]
    else:
        sizes, anc, lg, dl = retina_inputs(rng, B, hw, C, -5.0)
        anchors, logits, offs, ctr = [T(a) for a in anc], [T(x) for x in lg], [T(x) for x in dl], None
    logits[4][1] -= 50.0                                         # image 1: nothing passes on the last level
    info = T(np.array([[hw[0], hw[1], 480.0 + 9 * b, 600.0 + 5 * b, 0.0] for b in range(B)], np.float32))
    thr = 0.6 if fcos else 0.5
    d1, c1 = pipelines.dense_postprocess(logits, offs, anchors, info, 0.05, thr, 100, 1000, ctrness_list=ctr, fused_tail=True)
    d0, c0 = pipelines.dense_postprocess(logits, offs, anchors, info, 0.05, thr, 100, 1000, ctrness_list=ctr, fused_tail=False)
    assert torch.equal(c1, c0) and torch.equal(d1, d0)
    assert int(c1.min()) > 0
