"""The reference's own unit tests, re-stated against the drop-in layer (torch tensors stand in for mge tensors).

tests/structures/test_boxes.py, tests/layers/test_postprocess.py, tests/layers/test_roi_pool.py of
megvii-research/basedet, plus the call sequences of RetinaNet.get_ground_truth / inference written with the
reference's public names."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from basedet_b200 import workloads as W
from basedet_b200.layers import (AnchorPointGenerator, DefaultAnchorGenerator, FastPointGenerator, Matcher, batched_nms,
                                 non_zeros, post_process_with_empty_input, roi_pool)
from basedet_b200.structures import BoxCoder, BoxConverter, Boxes, Container, PointCoder
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


@pytest.fixture()
def boxes(cuda):
    b1 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 1.0, 1.0]], device=cuda)
    b2 = torch.tensor(
        [[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 1.0, 0.5], [0.0, 0.0, 0.5, 0.5],
         [0.5, 0.5, 1.0, 1.0], [0.5, 0.5, 1.5, 1.5]], device=cuda)
    return Boxes(b1), Boxes(b2)


class TestBoxes:  # tests/structures/test_boxes.py
    def test_iou(self, boxes):
        b1, b2 = boxes
        expected = np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / (2 - 0.25)]] * 2)
        assert np.allclose(b1.iou(b2).cpu().numpy(), expected)

    def test_ioa(self, boxes):
        b1, b2 = boxes
        expected = np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25]] * 2).T
        assert np.allclose(b2.ioa(b1).cpu().numpy(), expected)

    def test_interseaction(self, boxes):
        b1, b2 = boxes
        expected = np.array([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25]] * 2)
        assert np.allclose(b1.intersection(b2).cpu().numpy(), expected)

    def test_scale(self, boxes):
        b1, _ = boxes
        new_boxes = b1.scale(2, inplace=False)
        assert np.allclose(new_boxes.cpu().numpy(), b1.cpu().numpy() * 2)

    def test_center(self, boxes):
        b1, _ = boxes
        assert np.allclose(b1.centers.cpu().numpy(), np.array([[0.5, 0.5], [0.5, 0.5]]))

    def test_get_item(self, boxes):
        b1, _ = boxes
        sub_box = b1[:1]
        assert np.allclose(sub_box.cpu().numpy(), np.array([[0.0, 0.0, 1.0, 1.0]]))
        assert isinstance(sub_box, Boxes)
        value = b1[0, 0]
        assert int(value.cpu().numpy()) == 0
        assert not isinstance(b1[:, 0], Boxes)

    def test_props_clip_inplace(self, boxes, cuda):
        _, b2 = boxes
        assert np.allclose(b2.area.cpu().numpy(), [1, 0.5, 0.5, 0.25, 0.25, 1.0])
        assert np.allclose(b2.width.cpu().numpy(), [1, 0.5, 1, 0.5, 0.5, 1.0])
        raw = torch.tensor([[-5.0, 2.0, 50.0, 80.0]], device=cuda)
        bx = Boxes(raw)
        assert bx.data_ptr() == raw.data_ptr()                          # construction is a zero-copy alias ...
        out = bx.clip((60, 40))
        assert out is bx and bx.cpu().numpy().tolist() == [[0.0, 2.0, 40.0, 60.0]]
        # ... and the in-place op detaches: MegEngine tensors are values, the wrapped tensor keeps its coordinates
        # (SURVEY N3 / rpn.py:168: Boxes(proposals).clip(...) leaves `proposals` un-clipped)
        assert raw.cpu().numpy().tolist() == [[-5.0, 2.0, 50.0, 80.0]]
        bx.scale((2.0, 0.5))
        assert bx.cpu().numpy().tolist() == [[0.0, 4.0, 20.0, 120.0]] and raw[0, 0].item() == -5.0
        assert bx.filter_by_size().cpu().numpy().tolist() == [True]


def test_batched_nms(cuda):  # tests/layers/test_postprocess.py
    boxes = torch.tensor(
        [[0.0, 0.0, 100.0, 100.0], [0.0, 0.0, 100.5, 100.0], [0.0, 0.0, 201.0, 200.5], [0.0, 0.0, 200.5, 200.5],
         [0.5, 0.5, 100.0, 101.0], [0.5, 0.5, 120.5, 120.5]], device=cuda)
    scores = torch.tensor([0.9, 0.8, 0.3, 0.7, 0.6, 0.4], device=cuda)
    labels = torch.tensor([1, 1, 1, 2, 2, 2], device=cuda)
    keep_idx = batched_nms(boxes, scores, labels, iou_thresh=0.4).cpu().numpy()
    assert list(keep_idx) == [0, 3, 4, 2]
    with pytest.raises(AssertionError):
        batched_nms(boxes[:, :3], scores, labels, 0.4)
    with pytest.raises(AssertionError):
        batched_nms(boxes, scores[:3], labels, 0.4)


class TestRoIPool:  # tests/layers/test_roi_pool.py
    def setup_method(self):
        self.feat = torch.arange(25, dtype=torch.float32, device="cuda").reshape(1, 1, 5, 5)

    def test_roi_align(self, cuda):
        rois = torch.tensor([[0, 1, 1, 3, 3]], dtype=torch.float32, device=cuda)
        align_results = np.array([[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]])
        out = roi_pool([self.feat], rois, strides=[1], pool_shape=4, pooler_type="roi_align")
        assert np.allclose(out.cpu().numpy(), align_results)

    def test_resize(self, cuda):
        rois = torch.tensor([[0, 1, 1, 3, 3]], dtype=torch.float32, device=cuda)
        output = roi_pool([self.feat], rois, strides=[1], pool_shape=4, pooler_type="roi_align")
        feat2x = F.interpolate(self.feat, scale_factor=2, mode="bilinear", align_corners=False)
        output2x = roi_pool([feat2x], rois, strides=[1 / 2], pool_shape=4, pooler_type="roi_align")
        assert np.allclose(output2x.cpu().numpy(), output.cpu().numpy())

    def test_autograd(self, cuda):
        rng = np.random.default_rng(0)
        feats_np = [rng.normal(0, 1, (2, 4, 40 // (2 ** i), 56 // (2 ** i))).astype(np.float32) for i in range(4)]
        feats = [torch.from_numpy(f).to(cuda).requires_grad_(True) for f in feats_np]
        rois_np = W.make_rois(rng, 16, 2, 160, 224, 8, 200)
        out = roi_pool(feats, torch.from_numpy(rois_np).to(cuda), [4, 8, 16, 32], (7, 7))
        dout = rng.normal(0, 1, out.shape).astype(np.float32)
        out.backward(torch.from_numpy(dout).to(cuda))
        levels = R.assign_levels(rois_np, [4, 8, 16, 32])
        for l, f in enumerate(feats):
            sel = levels == l
            ref = R.roi_align_backward(dout[sel], feats_np[l].shape, rois_np[sel], (7, 7), 1.0 / [4, 8, 16, 32][l])
            assert np.max(np.abs(f.grad.cpu().numpy() - ref)) / max(np.abs(ref).max(), 1.0) <= 1e-5


def test_anchor_generators(cuda):
    feats = [torch.empty((1, 8, h, w), device=cuda) for h, w in W.retinanet_level_sizes(256, 320)]
    gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
    ref = R.default_anchors([tuple(f.shape[-2:]) for f in feats], W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
    for g, r in zip(gen(feats), ref):
        assert np.array_equal(g.cpu().numpy(), r)
    assert gen.anchor_dim == 4
    with pytest.raises(AssertionError):
        gen(feats[:3])
    pg = AnchorPointGenerator(1, tuple(W.RETINANET_STRIDES), 0.5)
    for g, r in zip(pg(feats), R.anchor_points([tuple(f.shape[-2:]) for f in feats], 1, W.RETINANET_STRIDES, 0.5)):
        assert np.array_equal(g.cpu().numpy(), r)
    fp = FastPointGenerator((8, 16, 32))
    for g, r in zip(fp(feats[:3]), R.fast_points([tuple(f.shape[-2:]) for f in feats[:3]], [8, 16, 32])):
        assert np.array_equal(g.cpu().numpy(), r)


def test_retinanet_get_ground_truth_call_sequence(cuda):
    """models/det/retinanet.py:211-232 written with the drop-in names, one image at a time."""
    sizes = W.retinanet_level_sizes(256, 320)
    anchors_np = np.concatenate(R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5))
    gt_np, ng = W.target_assign_batch(2, num_gt=12, img_h=256, img_w=320, ragged=True)
    anchors = torch.from_numpy(anchors_np).to(cuda)
    matcher = Matcher([0.4, 0.5], [0, -1, 1], allow_low_quality_matches=True)
    box_coder = BoxCoder((0.0, 0.0, 0.0, 0.0), (1.0, 1.0, 1.0, 1.0))
    for b in range(2):
        gt_boxes = torch.from_numpy(gt_np[b, : ng[b]]).to(cuda)
        overlaps = Boxes(gt_boxes[:, :4]).iou(Boxes(anchors))
        match_indices, labels = matcher(overlaps)
        gt_boxes_matched = gt_boxes[match_indices.long()]
        fg_mask = labels == 1
        labels[fg_mask] = gt_boxes_matched[fg_mask, 4].to(torch.int32)
        offsets = box_coder.encode(anchors, gt_boxes_matched[:, :4])
        rl, ro, _ = R.retinanet_targets(anchors_np, gt_np[b:b + 1], ng[b:b + 1], [0.4, 0.5], [0, -1, 1], True)
        assert np.array_equal(labels.cpu().numpy(), rl[0])
        err = np.abs(offsets.cpu().numpy() - ro[0]) / np.maximum(np.abs(ro[0]), 1.0)
        assert err.max() <= 1e-6


def test_retinanet_inference_call_sequence(cuda):
    """models/det/retinanet.py:172-209 with the drop-in names: non_zeros -> topk -> decode -> post_process."""
    rng = np.random.default_rng(3)
    sizes = W.retinanet_level_sizes(256, 320)
    anchors_np = R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
    box_coder = BoxCoder()
    img_info = torch.tensor([[256.0, 320.0, 500.0, 640.0, 0.0]], device=cuda)
    total_boxes, total_scores, logits_label = [], [], []
    ref_scores, logits_np, offsets_np = [], [], []
    from basedet_b200 import ops
    for (h, w), an in zip(sizes, anchors_np):
        lg = W.logits_level(rng, h * w * 9, 80, mean=-5.0)
        of = W.deltas_level(rng, h * w * 9)
        logits_np.append(lg); offsets_np.append(of)
        logits = torch.from_numpy(lg).to(cuda)
        scores = ops.scores(logits.reshape(-1))                    # F.sigmoid(F.flatten(logits))
        ref_scores.append(scores.cpu().numpy())
        _, keep_idx = non_zeros(scores > 0.05)
        if keep_idx.numel() == 0:
            continue
        topk_num = min(keep_idx.shape[0], 1000)
        _, topk_idx, _ = ops.topk_segments(scores[keep_idx.long()], [keep_idx.shape[0]], topk_num)  # F.topk(descending=True)
        keep_idx = keep_idx[topk_idx[0].long()]
        total_scores.append(scores[keep_idx.long()])
        logits_label.append(keep_idx % 80)
        boxes = box_coder.decode(torch.from_numpy(an).to(cuda), torch.from_numpy(of).to(cuda).reshape(-1, 4))
        total_boxes.append(boxes[(keep_idx // 80).long()])
    dets = post_process_with_empty_input(total_boxes, total_scores, logits_label, img_info, 0.5, 100)
    rb, rs, rl, _ = R.retinanet_postprocess(logits_np, offsets_np, anchors_np, img_info.cpu().numpy(), 0.05, 0.5, 100,
                                            scores_list=ref_scores)
    assert isinstance(dets, Container) and isinstance(dets.boxes, Boxes)
    assert np.array_equal(dets.box_labels.cpu().numpy(), rl)
    assert np.array_equal(dets.box_scores.cpu().numpy(), rs)
    scale = np.maximum(np.abs(rb).max(axis=1, keepdims=True), 1.0)
    assert (np.abs(dets.boxes.cpu().numpy() - rb) / scale).max() <= 1e-6
    empty = post_process_with_empty_input([], [], [], img_info)
    assert empty.boxes.numel() == 0


def test_coders_and_convert(cuda):
    rng = np.random.default_rng(5)
    pts = np.concatenate(R.anchor_points([(16, 20)], 1, [8], 0.5))
    gt = W.make_gt(rng, 7, 128, 160)
    pc = PointCoder()
    enc = pc.encode(torch.from_numpy(pts).to(cuda), torch.from_numpy(gt[:, None, :4]).to(cuda))
    assert np.array_equal(enc.cpu().numpy(), R.pointcoder_encode(pts, gt[:, None, :4]))
    idx = rng.integers(0, 7, pts.shape[0])
    enc_rows = pc.encode(torch.from_numpy(pts).to(cuda), torch.from_numpy(gt[idx, :4]).to(cuda))
    assert np.array_equal(enc_rows.cpu().numpy(), R.pointcoder_encode(pts, gt[idx, :4]))
    b = gt[:, :4]
    for mode in ("xyxy2xywh", "xywh2xyxy", "xyxy2xcycwh", "xcycwh2xyxy", "xywh2xcycwh", "xcycwh2xywh", "xyxy2xyxy"):
        got = BoxConverter.convert(torch.from_numpy(b).to(cuda), mode).cpu().numpy()
        assert np.array_equal(got, R.box_convert(b, mode)), mode


def test_non_zeros(cuda):
    rng = np.random.default_rng(6)
    for n in (1, 100, 2048, 2049, 100000):
        x = rng.normal(0, 1, n).astype(np.float32)
        x[rng.random(n) < 0.7] = 0
        vals, idx = non_zeros(torch.from_numpy(x).to(cuda))
        rv, ri = R.cond_take(x != 0, x)
        assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(vals.cpu().numpy(), rv)
    z = torch.zeros(10, device=cuda)
    assert non_zeros(z)[1].numel() == 0


def test_cpu_tensors_are_rejected():
    """No CPU fallback: host tensors fail loudly."""
    from basedet_b200 import ops

    with pytest.raises(RuntimeError):
        ops.pairwise(torch.zeros(2, 4), torch.zeros(3, 4))


def test_sample_labels_dropin(cuda):
    """layers.sample_labels keeps the reference call (sampling.py:7): in-place int labels, bool masks, budget respected,
    nothing touched when the population fits, explicit variates reproduce the oracle."""
    from basedet_b200.layers import sample_labels
    from oracle import ref_ops as R
    rng = np.random.default_rng(3)
    lab = rng.choice(np.array([-1, 0, 1], np.int32), size=4000, p=[0.2, 0.7, 0.1]).astype(np.int32)
    noise = rng.uniform(0, 1, 4000).astype(np.float32)
    t = torch.from_numpy(lab.copy()).to(cuda)
    out = sample_labels(t, 64, 1, -1, noise=torch.from_numpy(noise).to(cuda))
    assert out is t and np.array_equal(t.cpu().numpy(), R.sample_labels(lab, 64, 1, -1, noise))
    t2 = torch.from_numpy(lab.copy()).to(cuda)
    sample_labels(t2, 100, 0)                                  # variates from torch's generator
    got = t2.cpu().numpy()
    assert (got == 0).sum() == 100 and np.array_equal(got[lab != 0], lab[lab != 0]) and set(got[lab == 0]) <= {0, -1}
    m = torch.from_numpy(lab == 1).to(cuda)
    ms = sample_labels(m, 10, True, False)
    assert ms.dtype == torch.bool and int(ms.sum()) == 10 and bool((ms & ~m).sum() == 0)
    t3 = torch.from_numpy(lab.copy()).to(cuda)
    sample_labels(t3, 10 ** 6, 0)
    assert np.array_equal(t3.cpu().numpy(), lab)


def test_dropin_find_top_rpn_proposals_equals_fused_pipeline(cuda):
    """models/det/rpn.py:141-186 written against the drop-in layer (Boxes.clip / filter_by_size / BoxCoder.decode /
    batched_nms), line for line, equals pipelines.rpn_proposals: same rois in the same order, un-clipped coordinates
    (the two paths must agree on the Boxes in-place semantics)."""
    from basedet_b200 import pipelines

    rng = np.random.default_rng(9)
    hw, B, pre_k, post_k, thr = (128, 160), 2, 300, 100, 0.7
    sizes = W.frcnn_level_sizes(*hw)
    gen = DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
    anchors_list = gen.generate_anchors_by_features(sizes, cuda)
    n_l = [a.shape[0] for a in anchors_list]
    scores = [torch.from_numpy(np.stack([W.distinct_scores(rng, n, -9.0, 3.0) for _ in range(B)])).to(cuda) for n in n_l]
    deltas = [torch.from_numpy(rng.normal(0, 0.5, (B, n, 4)).astype(np.float32)).to(cuda) for n in n_l]   # wide: many clip
    im_info = torch.tensor([[hw[0], hw[1], hw[0], hw[1], 0.0]] * B, device=cuda)
    rois, cnt = pipelines.rpn_proposals(scores, deltas, anchors_list, im_info, pre_k, post_k, thr)
    box_coder = BoxCoder((0.0, 0.0, 0.0, 0.0), (1.0, 1.0, 1.0, 1.0))
    for bid in range(B):
        props, scs, lvls = [], [], []
        for level, (s, d, a) in enumerate(zip(scores, deltas, anchors_list)):
            proposals = box_coder.decode(a, d[bid].clone())
            k = min(pre_k, s.shape[1])
            sc, order = torch.topk(s[bid], k)                      # F.topk(descending=True); scores are distinct
            props.append(proposals[order]); scs.append(sc); lvls.append(torch.full_like(sc, level))
        proposals, sc, levels = torch.cat(props), torch.cat(scs), torch.cat(lvls)
        proposal_boxes = Boxes(proposals).clip(im_info[bid][:2])    # rpn.py:168
        keep_mask = proposal_boxes.filter_by_size()
        proposals, sc, levels = proposals[keep_mask], sc[keep_mask], levels[keep_mask]
        keep = batched_nms(proposals, sc, levels, thr, post_k).long()
        ref = proposals[keep]
        n = int(cnt[bid])
        assert n == ref.shape[0]
        assert (rois[bid, :n, 0] == bid).all()
        assert torch.equal(rois[bid, :n, 1:], ref)                  # same kernels underneath: bit-identical, un-clipped
        assert (ref[:, 0] < 0).any() or (ref[:, 2] > hw[1]).any()   # the case is exercised: some rois stick out
