"""TEST INFRASTRUCTURE ONLY -- numpy fp32 restatement of BaseDet's box-op hot path.

Every function follows the cited reference lines *in operation order*, one fp32
rounding per arithmetic op (numpy float32 arrays never contract to FMA).  All
``path:line`` citations are relative to the reference tree (megvii-research/basedet).

MegEngine itself (``megengine.functional``) is not available offline, so the
semantics of its leaf ops are restated from its published behaviour and are
flagged ``ASSUMED`` below; DESIGN.md lists which of them are pinned by the
reference's own tests.

  ASSUMED-1  Elemwise MAX / MIN are ``x>y?x:y`` / ``x<y?x:y``  (so max(NaN, 0) = 0).
  ASSUMED-2  argmax returns the FIRST (lowest) index among equal maxima.
  ASSUMED-3  argsort / topk(descending) order is (value desc, index asc), i.e. stable.
  ASSUMED-4  cond_take returns flat indices in ascending order (int32).
  ASSUMED-5  F.vision.nms: IoU = inter / (Sa + Sb - inter), suppress iff IoU > thresh,
             greedy in score-descending order, truncated to max_output.
  ASSUMED-6  F.nn.roi_align(aligned=True): offset 0.5, zero padding for taps outside
             the map (no clamping), lerp written as a + (b - a) * t, average = sum / S^2.
  ASSUMED-7  F.arange(start, stop, step) = fp32(start + i * step) evaluated in float64.
  ASSUMED-8  F.mean / F.std over the 45 ATSS candidates: fp32 sum accumulated sequentially in index
             order, one divide by n; std = sqrt(mean((x - mean) ** 2)) (population).  MegDNN's reduce
             order is unknown; any fixed order differs from another by <= a few ulp of the threshold.
  ASSUMED-10 F.topk takes a signed k (MegDNN TopK): a negative k selects the |k| LARGEST elements, ordered
             (value desc, index asc).  ``sample_labels`` (sampling.py:27) relies on it.
  ASSUMED-9  ``x ** 2`` on a tensor is ``x * x`` (one rounding); F.topk(descending=False) orders by
             (value asc, index asc); argmin returns the FIRST index among equal minima.
  ASSUMED-11 Advanced-index assignment ``t[i, j] = v`` with repeated (i, j) pairs keeps the LAST value in index
             order (FreeAnchor's box-probability scatter, free_anchor.py:78).
  ASSUMED-13 F.nn.roi_pooling(mode="max") is MegDNN ROIPooling = the Caffe rule (rounded corners, size = end - start + 1,
             floor / ceil bin edges, 0 for empty bins); pinned by tests/layers/test_roi_pool.py:48-61.
  ASSUMED-12 F.logsigmoid(x) = min(x, 0) - log1p(exp(-|x|)) in fp32; ``x ** gamma`` with gamma = 2 is x * x.
"""
import math

import numpy as np

f32 = np.float32
_ZERO = f32(0.0)


# --------------------------------------------------------------------------- leaf ops
def emax(x, y):
    """MegDNN Elemwise MAX (ASSUMED-1)."""
    return np.where(x > y, x, y).astype(f32)


def emin(x, y):
    """MegDNN Elemwise MIN (ASSUMED-1)."""
    return np.where(x < y, x, y).astype(f32)


def arange_f32(start, stop, step):
    """megengine.functional.arange (ASSUMED-7)."""
    num = int(math.ceil((stop - start) / step))
    num = max(num, 0)
    return (start + np.arange(num, dtype=np.float64) * step).astype(f32)


def meshgrid(x, y):
    """basedet/layers/common/function.py:47-54."""
    assert x.ndim == 1 and y.ndim == 1
    shape = (y.shape[0], x.shape[0])
    return np.broadcast_to(x, shape), np.broadcast_to(y.reshape(-1, 1), shape)


def cond_take(mask, x):
    """F.cond_take (ASSUMED-4): (values, flat ascending int32 indices)."""
    idx = np.flatnonzero(mask.reshape(-1)).astype(np.int32)
    return x.reshape(-1)[idx], idx


def argsort_desc(scores):
    """Stable descending argsort: (score desc, index asc) (ASSUMED-3)."""
    scores = np.asarray(scores, dtype=f32)
    # -0.0 and +0.0 compare equal; NaN-free inputs assumed.
    return np.argsort(-scores, kind="stable").astype(np.int32)


def topk_desc(scores, k):
    """F.topk(scores, k, descending=True) -> (values, int32 indices), sorted (ASSUMED-3).

    ``k`` is clamped to ``len(scores)`` (reference relies on MegEngine doing so at
    RPN P6, basedet/models/det/rpn.py:155 with 819 < 1000).
    """
    order = argsort_desc(scores)[: min(int(k), len(scores))]
    return np.asarray(scores, dtype=f32)[order], order


# --------------------------------------------------------------------------- anchors
def generate_base_anchors(scales, ratios):
    """basedet/layers/common/anchor_generator.py:99-109 (float64 math, fp32 result :95).

    ``scales`` / ``ratios`` are first rounded to fp32 then widened again, exactly as
    ``np.array(..., dtype=np.float32)`` (:73-74) followed by ``.tolist()`` (:83) does.
    """
    scales = np.array(scales, dtype=f32).tolist()
    ratios = np.array(ratios, dtype=f32).tolist()
    out = []
    areas = [s ** 2.0 for s in scales]
    for area in areas:
        for ratio in ratios:
            w = math.sqrt(area / ratio)
            h = ratio * w
            out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return np.array(out, dtype=f32).reshape(-1, 4)


def create_anchor_grid(featmap_size, offsets, stride):
    """basedet/layers/common/anchor_generator.py:23-30.

    ``featmap_size`` is (H, W); the reference's step_x/step_y naming swap (:24) is
    neutralised by ``meshgrid(grid_y, grid_x)`` (:29): x varies along W, y along H.
    """
    step_x, step_y = featmap_size
    shift = offsets * stride
    grid_x = arange_f32(shift, step_x * stride + shift, stride)
    grid_y = arange_f32(shift, step_y * stride + shift, stride)
    grids_x, grids_y = meshgrid(grid_y, grid_x)
    return grids_x.reshape(-1), grids_y.reshape(-1)


def default_anchors(sizes, anchor_scales, anchor_ratios, strides, offset):
    """DefaultAnchorGenerator.generate_anchors_by_features, anchor_generator.py:86-122."""
    n = len(strides)
    scales = list(anchor_scales) * n if len(anchor_scales) == 1 else list(anchor_scales)
    ratios = list(anchor_ratios) * n if len(anchor_ratios) == 1 else list(anchor_ratios)
    assert len(scales) == n and len(ratios) == n and len(sizes) == n
    out = []
    for size, stride, sc, ra in zip(sizes, strides, scales, ratios):
        base = generate_base_anchors(sc, ra)
        gx, gy = create_anchor_grid(size, offset, stride)
        grids = np.stack([gx, gy, gx, gy], axis=1).astype(f32)
        out.append((grids.reshape(-1, 1, 4) + base.reshape(1, -1, 4)).reshape(-1, 4).astype(f32))
    return out


def anchor_points(sizes, num_anchors, strides, offset):
    """AnchorPointGenerator.generate_anchors_by_features, anchor_generator.py:152-165."""
    out = []
    for size, stride in zip(sizes, strides):
        gx, gy = create_anchor_grid(size, offset, stride)
        grids = np.stack([gx, gy], axis=1).astype(f32)
        out.append(np.repeat(grids[:, None, :], num_anchors, axis=1).reshape(-1, 2))
    return out


def fast_points(sizes, strides):
    """FastPointGenerator.__call__, anchor_generator.py:175-182.

    Quirk kept: ``meshgrid(arange(h), arange(w))`` yields a (w, h) mesh, so the
    output row j*h + i holds (i*stride, j*stride) with i < h, j < w.
    """
    out = []
    for (h, w), stride in zip(sizes, strides):
        gx, gy = meshgrid(np.arange(h, dtype=f32), np.arange(w, dtype=f32))
        grids = np.stack((gx, gy), axis=-1).reshape(-1, 2).astype(f32)
        out.append((grids * f32(stride)).astype(f32))
    return out


# --------------------------------------------------------------------------- pairwise box ops
def _pair_inter(b1, b2):
    """Shared front half of IOU / IOA subgraphs, op_patch.py:50-61 / :187-197."""
    b1 = np.asarray(b1, dtype=f32)[:, None, :]
    b2 = np.asarray(b2, dtype=f32)[None, :, :]
    iw = emin(b1[..., 2], b2[..., 2]) - emax(b1[..., 0], b2[..., 0])
    ih = emin(b1[..., 3], b2[..., 3]) - emax(b1[..., 1], b2[..., 1])
    iw = emax(iw, _ZERO)
    ih = emax(ih, _ZERO)
    return (iw * ih).astype(f32), b1, b2


def box_iou(boxes1, boxes2):
    """basedet/structures/op_patch.py:33-97 (IOU subgraph) -> (N, M) fp32."""
    inter, b1, b2 = _pair_inter(boxes1, boxes2)
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    union = (a1 + a2).astype(f32)
    union = (union - inter).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = (inter / union).astype(f32)
    return emax(iou, _ZERO)


def box_ioa(boxes1, boxes2):
    """basedet/structures/op_patch.py:169-227 (IOA subgraph): inter / area(boxes2)."""
    inter, _, b2 = _pair_inter(boxes1, boxes2)
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        ioa = (inter / a2).astype(f32)
    return emax(ioa, _ZERO)


def box_intersection(boxes1, boxes2):
    """Boxes.intersection, basedet/structures/boxes.py:114-130."""
    return _pair_inter(boxes1, boxes2)[0]


def box_giou(boxes1, boxes2):
    """Boxes.giou, basedet/structures/boxes.py:74-95 (iou NOT clamped)."""
    b1 = np.asarray(boxes1, dtype=f32)
    b2 = np.asarray(boxes2, dtype=f32)
    inter = box_intersection(b1, b2)
    a1 = ((b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1]))[:, None]
    a2 = ((b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1]))[None, :]
    union = ((a1 + a2).astype(f32) - inter).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = (inter / union).astype(f32)
        e1, e2 = b1[:, None, :], b2[None, :, :]
        lt = np.minimum(e1[..., :2], e2[..., :2])
        rb = np.maximum(e1[..., 2:], e2[..., 2:])
        wh = np.maximum((rb - lt).astype(f32), _ZERO)  # F.clip(lower=0)
        area = (wh[..., 0] * wh[..., 1]).astype(f32)
        return (iou - ((area - union).astype(f32) / area).astype(f32)).astype(f32)


def box_center(boxes):
    """basedet/structures/op_patch.py:100-113: (tl + br) / 2."""
    b = np.asarray(boxes, dtype=f32)
    return ((b[:, :2] + b[:, -2:]).astype(f32) / f32(2)).astype(f32)


def point_distance(p1, p2):
    """basedet/structures/op_patch.py:133-149: pow(sum(pow(diff, 2)), 0.5)."""
    p1 = np.asarray(p1, dtype=f32)[:, None, :]
    p2 = np.asarray(p2, dtype=f32)
    diff = (p1 - p2).astype(f32)
    sq = (diff * diff).astype(f32)  # pow(x, 2) == x*x exactly in IEEE
    s = (sq[..., 0] + sq[..., 1]).astype(f32)
    return np.sqrt(s).astype(f32)  # pow(x, 0.5): sqrt restated (ulp-level, tolerance-tested)


# --------------------------------------------------------------------------- Boxes helpers
def boxes_clip(boxes, sizes):
    """Boxes.clip, basedet/structures/boxes.py:152-177; sizes = (h, w)."""
    h, w = (f32(sizes[0]), f32(sizes[1]))
    b = np.asarray(boxes, dtype=f32)
    out = np.empty_like(b)
    out[:, 0] = np.minimum(np.maximum(b[:, 0], _ZERO), w)
    out[:, 1] = np.minimum(np.maximum(b[:, 1], _ZERO), h)
    out[:, 2] = np.minimum(np.maximum(b[:, 2], _ZERO), w)
    out[:, 3] = np.minimum(np.maximum(b[:, 3], _ZERO), h)
    return out


def boxes_scale(boxes, scale_ratios):
    """Boxes.scale, basedet/structures/boxes.py:193-212; ratios = (scale_h, scale_w)."""
    sh, sw = f32(scale_ratios[0]), f32(scale_ratios[1])
    return (np.asarray(boxes, dtype=f32) * np.array([sw, sh, sw, sh], dtype=f32)).astype(f32)


def boxes_filter_by_size(boxes, sizes=0):
    """Boxes.filter_by_size, basedet/structures/boxes.py:132-150 (keeps the h/w name swap)."""
    if isinstance(sizes, (int, float)):
        sizes = (sizes, sizes)
    b = np.asarray(boxes, dtype=f32)
    h, w = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    return (w > f32(sizes[0])) & (h > f32(sizes[1]))


# --------------------------------------------------------------------------- coders
def _ltrb_to_cs(b):
    """BoxCoder._box_ltrb_to_cs_opr, basedet/structures/boxcoder.py:44-59."""
    w = (b[:, 2] - b[:, 0]).astype(f32)
    h = (b[:, 3] - b[:, 1]).astype(f32)
    cx = (b[:, 0] + (f32(0.5) * w).astype(f32)).astype(f32)
    cy = (b[:, 1] + (f32(0.5) * h).astype(f32)).astype(f32)
    return w, h, cx, cy


def boxcoder_encode(bbox, gt, mean=(0, 0, 0, 0), std=(1, 1, 1, 1)):
    """BoxCoder.encode, basedet/structures/boxcoder.py:61-73 (true divide by std)."""
    bbox = np.asarray(bbox, dtype=f32)
    gt = np.asarray(gt, dtype=f32)
    bw, bh, bcx, bcy = _ltrb_to_cs(bbox)
    gw, gh, gcx, gcy = _ltrb_to_cs(gt)
    with np.errstate(divide="ignore", invalid="ignore"):
        dx = ((gcx - bcx).astype(f32) / bw).astype(f32)
        dy = ((gcy - bcy).astype(f32) / bh).astype(f32)
        dw = np.log((gw / bw).astype(f32)).astype(f32)
        dh = np.log((gh / bh).astype(f32)).astype(f32)
        t = np.stack([dx, dy, dw, dh], axis=1)
        t = (t - np.asarray(mean, dtype=f32).reshape(1, 4)).astype(f32)
        t = (t / np.asarray(std, dtype=f32).reshape(1, 4)).astype(f32)
    return t


def boxcoder_decode(anchors, deltas, mean=(0, 0, 0, 0), std=(1, 1, 1, 1)):
    """BoxCoder.decode, basedet/structures/boxcoder.py:75-98 -> (boxes (N,4k), rescaled deltas).

    The reference rescales ``deltas`` IN PLACE (:76-77); the rescaled array is
    returned second so tests can check the write-back.  No clamp on dw/dh.
    """
    anchors = np.asarray(anchors, dtype=f32)
    d = np.asarray(deltas, dtype=f32)
    k = d.shape[1] // 4
    d = (d * np.tile(np.asarray(std, dtype=f32), k).reshape(1, -1)).astype(f32)
    d = (d + np.tile(np.asarray(mean, dtype=f32), k).reshape(1, -1)).astype(f32)
    aw, ah, acx, acy = [v[:, None] for v in _ltrb_to_cs(anchors)]
    with np.errstate(over="ignore", invalid="ignore"):
        pcx = (acx + (d[:, 0::4] * aw).astype(f32)).astype(f32)
        pcy = (acy + (d[:, 1::4] * ah).astype(f32)).astype(f32)
        pw = (aw * np.exp(d[:, 2::4]).astype(f32)).astype(f32)
        ph = (ah * np.exp(d[:, 3::4]).astype(f32)).astype(f32)
        half_w = (f32(0.5) * pw).astype(f32)
        half_h = (f32(0.5) * ph).astype(f32)
        x1 = (pcx - half_w).astype(f32)
        y1 = (pcy - half_h).astype(f32)
        x2 = (pcx + half_w).astype(f32)
        y2 = (pcy + half_h).astype(f32)
    box = np.stack([x1, y1, x2, y2], axis=2).reshape(d.shape[0], -1).astype(f32)
    return box, d


def sumcoder_encode(anchors, gt, mean=(0, 0, 0, 0), std=(1, 1, 1, 1)):
    """SumBoxCoder.encode, basedet/structures/boxcoder.py:115-120."""
    t = (np.asarray(gt, dtype=f32) - np.asarray(anchors, dtype=f32)).astype(f32)
    t = (t - np.asarray(mean, dtype=f32).reshape(1, 4)).astype(f32)
    return (t / np.asarray(std, dtype=f32).reshape(1, 4)).astype(f32)


def sumcoder_decode(anchors, deltas, mean=(0, 0, 0, 0), std=(1, 1, 1, 1)):
    """SumBoxCoder.decode, basedet/structures/boxcoder.py:122-127."""
    d = (np.asarray(deltas, dtype=f32) * np.asarray(std, dtype=f32).reshape(1, 4)).astype(f32)
    d = (d + np.asarray(mean, dtype=f32).reshape(1, 4)).astype(f32)
    return (np.asarray(anchors, dtype=f32) + d).astype(f32), d


def pointcoder_encode(point, gt):
    """PointCoder.encode, basedet/structures/boxcoder.py:132-133 (broadcasting concat)."""
    point = np.asarray(point, dtype=f32)
    gt = np.asarray(gt, dtype=f32)
    lt = (point - gt[..., :2]).astype(f32)
    rb = (gt[..., 2:] - point).astype(f32)
    return np.concatenate([lt, rb], axis=-1).astype(f32)


def pointcoder_decode(anchors, deltas):
    """PointCoder.decode, basedet/structures/boxcoder.py:135-141."""
    a = np.asarray(anchors, dtype=f32)
    d = np.asarray(deltas, dtype=f32)
    out = np.stack(
        [
            a[:, 0:1] - d[:, 0::4],
            a[:, 1:2] - d[:, 1::4],
            a[:, 0:1] + d[:, 2::4],
            a[:, 1:2] + d[:, 3::4],
        ],
        axis=2,
    )
    return out.reshape(d.shape).astype(f32)


# --------------------------------------------------------------------------- box_convert
def box_convert(boxes, mode="xywh2xyxy"):
    """BoxConverter.convert, basedet/structures/box_convert.py:51-82 ((N,4) boxes)."""
    src, dst = mode.lower().split("2")
    b = np.asarray(boxes, dtype=f32)
    if src == dst:
        return b
    if src == "xyxy":
        b = np.concatenate([b[:, :2], b[:, 2:3] - b[:, 0:1], b[:, 3:4] - b[:, 1:2]], axis=1)
    elif src == "xcycwh":
        x = b[:, 0:1] - (b[:, 2:3] / f32(2)).astype(f32)
        y = b[:, 1:2] - (b[:, 3:4] / f32(2)).astype(f32)
        b = np.concatenate([x, y, b[:, 2:]], axis=1)
    elif src != "xywh":
        raise NotImplementedError(src)
    b = b.astype(f32)
    if dst == "xyxy":
        b = np.concatenate([b[:, :2], b[:, 0:1] + b[:, 2:3], b[:, 1:2] + b[:, 3:4]], axis=1)
    elif dst == "xcycwh":
        xc = b[:, 0:1] + (b[:, 2:3] / f32(2)).astype(f32)
        yc = b[:, 1:2] + (b[:, 3:4] / f32(2)).astype(f32)
        b = np.concatenate([xc, yc, b[:, 2:]], axis=1)
    return b.astype(f32)


# --------------------------------------------------------------------------- Matcher
def matcher(matrix, thresholds, labels, allow_low_quality_matches=False):
    """Matcher.__call__, basedet/layers/common/matcher.py:31-51.

    ``thresholds`` are the *user* thresholds (without the +-inf the constructor adds, :24-25).
    Returns (match_indices int32 (A,), labels int32 (A,)).  NaN-free matrix assumed.
    """
    matrix = np.asarray(matrix, dtype=f32)
    assert matrix.ndim == 2
    assert len(thresholds) + 1 == len(labels)
    thr = [-float("inf")] + [float(t) for t in thresholds] + [float("inf")]
    max_scores = matrix.max(axis=0)
    match_indices = np.argmax(matrix, axis=0).astype(np.int32)  # ASSUMED-2
    out = np.full(match_indices.shape, -1, dtype=np.int32)
    for label, low, high in zip(labels, thr[:-1], thr[1:]):
        # python-float thresholds become fp32 scalars when compared with an fp32 tensor
        mask = (max_scores >= f32(low)) & (max_scores < f32(high))
        out[mask] = label
    if allow_low_quality_matches:
        mask = (matrix == matrix.max(axis=1, keepdims=True)).sum(axis=0) > 0
        out[mask] = 1
    return match_indices, out


def matcher_rows(matrix):
    """RCNN layout (R, G): max / argmax over axis 1, basedet/layers/head/rcnn.py:113-116."""
    matrix = np.asarray(matrix, dtype=f32)
    return matrix.max(axis=1), np.argmax(matrix, axis=1).astype(np.int32)


def retinanet_targets(anchors, gt_boxes, num_gt, thresholds, labels, allow_lq,
                      mean=(0, 0, 0, 0), std=(1, 1, 1, 1)):
    """RetinaNet.get_ground_truth, basedet/models/det/retinanet.py:211-232.

    anchors (A,4); gt_boxes (B,Gmax,5) rows [x1,y1,x2,y2,class]; num_gt (B,).
    Returns labels (B,A) int32 (class ids for fg, 0 bg, -1 ignore), offsets (B,A,4),
    plus match_indices (B,A) for diagnostics.
    """
    lab_l, off_l, idx_l = [], [], []
    for g, n in zip(gt_boxes, num_gt):
        g = np.asarray(g, dtype=f32)[: int(n)]
        overlaps = box_iou(g[:, :4], anchors)
        idx, lab = matcher(overlaps, thresholds, labels, allow_lq)
        matched = g[idx]
        fg = lab == 1
        lab[fg] = matched[fg, 4].astype(np.int32)
        off_l.append(boxcoder_encode(anchors, matched[:, :4], mean, std))
        lab_l.append(lab)
        idx_l.append(idx)
    return np.stack(lab_l), np.stack(off_l), np.stack(idx_l)


def sample_labels(labels, num_samples, label_value, ignore_label, noise):
    """sample_labels, basedet/layers/common/sampling.py:7-30, with the random draw made explicit.

    RNG contract: ``noise`` has one uniform variate PER ELEMENT of ``labels``; the reference's
    ``uniform(size=num_valid)`` (:24) is ``noise[mask]`` (the variates of the selected positions, in index order).
    Keeps ``num_samples`` elements equal to ``label_value`` and sets the others to ``ignore_label``: the ones with
    the LARGEST variates go (topk with negative k, ASSUMED-10; ties (value desc, index asc), ASSUMED-3).  Returns a
    new array (the reference mutates in place and also returns it)."""
    labels = np.array(labels, copy=True)
    mask = labels == label_value                                                # :19
    num_valid = int(mask.sum())                                                 # :20
    if num_valid <= num_samples:                                                # :21-22
        return labels
    random_tensor = np.zeros(labels.shape, f32)                                 # :24
    random_tensor[mask] = np.asarray(noise, f32)[mask]                          # :25
    k = num_valid - num_samples                                                 # :27 (passed as a negative k)
    invalid = np.argsort(-random_tensor, kind="stable")[:k]
    labels[invalid] = ignore_label                                              # :29
    return labels


def rpn_targets(anchors, gt_boxes, num_gt, thresholds, labels, allow_lq, num_sample_anchors, num_pos_anchor,
                noise_pos, noise_neg, mean=(0, 0, 0, 0), std=(1, 1, 1, 1)):
    """RPN.get_ground_truth, basedet/models/det/rpn.py:215-240: IoU -> Matcher -> encode -> sample positives ->
    sample negatives.  noise_pos / noise_neg (B, A): the variates of the two sample_labels calls (see sample_labels).
    Returns labels (B, A) int32 in {-1, 0, 1} and offsets (B, A, 4)."""
    lab_l, off_l = [], []
    for b, (g, n) in enumerate(zip(gt_boxes, num_gt)):
        g = np.asarray(g, dtype=f32)[: int(n)]
        overlaps = box_iou(g[:, :4], anchors)                                   # :223
        idx, lab = matcher(overlaps, thresholds, labels, allow_lq)              # :224
        off_l.append(boxcoder_encode(anchors, g[idx][:, :4], mean, std))        # :226
        lab = sample_labels(lab, num_pos_anchor, 1, -1, noise_pos[b])           # :229
        num_negative = num_sample_anchors - int((lab == 1).sum())               # :231
        lab = sample_labels(lab, num_negative, 0, -1, noise_neg[b])             # :232
        lab_l.append(lab)
    return np.stack(lab_l), np.stack(off_l)


def rcnn_targets(rois_list, gt_boxes, num_gt, noise_fg, noise_bg, num_rois=512, fg_ratio=0.5, fg_thresh=0.5,
                 bg_thresh_high=0.5, bg_thresh_low=0.0, mean=(0, 0, 0, 0), std=(0.1, 0.1, 0.2, 0.2)):
    """RCNN.get_ground_truth (training branch), basedet/layers/head/rcnn.py:95-147.

    rois_list[b]: (R_b, 5) rows [batch, x1, y1, x2, y2] of image b (what ``rpn_rois[rpn_rois[:, 0] == bid]`` selects);
    noise_fg / noise_bg [b]: one uniform variate per row of all_rois (R_b + G_b), see sample_labels.
    Returns per image (rois (n, 5), labels (n,) int32, bbox_targets (n, 4)); the reference concatenates them."""
    out = []
    for bid, (rois, g5, n) in enumerate(zip(rois_list, gt_boxes, num_gt)):
        g5 = np.asarray(g5, f32)[: int(n)]                                      # :106-107
        gt_rois = np.concatenate([np.full((len(g5), 1), f32(bid), f32), g5[:, :4]], axis=1)   # :108-109
        all_rois = np.concatenate([np.asarray(rois, f32), gt_rois], axis=0)     # :112
        overlaps = box_iou(all_rois[:, 1:], g5[:, :4])                          # :114 (R, G)
        max_ov = overlaps.max(axis=1)                                           # :115
        assign = np.argmax(overlaps, axis=1).astype(np.int32)                   # :116 (ASSUMED-2)
        labels = g5[assign, 4].copy()                                           # :117 (fp32)
        fg = (max_ov >= f32(fg_thresh)) & (labels >= 0)                         # :119
        bg = (max_ov >= f32(bg_thresh_low)) & (max_ov < f32(bg_thresh_high))    # :120-123
        num_fg = int(num_rois * fg_ratio)                                       # :125
        fg_s = sample_labels(fg, num_fg, True, False, noise_fg[bid])            # :126
        num_bg = int(num_rois - fg_s.sum())                                     # :127
        bg_s = sample_labels(bg, num_bg, True, False, noise_bg[bid])            # :128
        labels[bg_s] = 0                                                        # :130
        keep = fg_s | bg_s                                                      # :132
        lab = labels[keep].astype(np.int32)                                     # :133
        kept = all_rois[keep]                                                   # :134
        targets = boxcoder_encode(kept[:, 1:], g5[assign[keep], :4], mean, std).reshape(-1, 4)   # :135-137
        out.append((kept, lab, targets))
    return out


def ota_topk_match(cost, ious, candidate_k=10):
    """OTATopkMatcher.__call__, basedet/layers/common/matcher.py:134-161 (dynamic-k matching of OTA / YOLOX).

    cost, ious: (G, A) fp32.  Returns matched GT index per anchor (A,) int32, G = background.
    The sum of the top-k IoUs (:146) is accumulated sequentially in descending order (ASSUMED-8); its int32 cast
    truncates."""
    cost = np.asarray(cost, f32)
    ious = np.asarray(ious, f32)
    G, A = cost.shape
    matching = np.zeros((G, A), f32)                                            # :143
    k = min(int(candidate_k), A)
    topk_ious = -np.sort(-ious, axis=1, kind="stable")[:, :k]                   # :145 values, descending
    dynamic_ks = np.maximum(seq_sum_f32(topk_ious, 1).astype(np.int32), 1)      # :146
    for g in range(G):                                                          # :147-149
        idx = np.argsort(cost[g], kind="stable")[: dynamic_ks[g]]              # topk ascending (ASSUMED-9)
        matching[g, idx] = 1.0
    multi = matching.sum(0) > 1                                                 # :154
    if multi.sum() > 0:                                                         # :155-158
        cost_argmin = np.argmin(cost[:, multi], axis=0)
        matching[:, multi] = 0.0
        matching[cost_argmin, multi] = 1.0
    full = np.concatenate([matching * 2, np.ones((1, A), f32)], axis=0)         # :160-161
    return np.argmax(full, axis=0).astype(np.int32)                             # first index (ASSUMED-2)


# --------------------------------------------------------------------------- FreeAnchor (8(f)-3)
def free_anchor_targets(anchors, pred_offsets, pred_scores, gt5, num_classes, box_iou_thresh=0.6, bucket_size=50,
                        reg_mean=(0, 0, 0, 0), reg_std=(0.1, 0.1, 0.2, 0.2), clamp_eps=1e-7):
    """The box ops of one image of FreeAnchor.get_losses, basedet/models/det/free_anchor.py:48-113 (defaults:
    configs/det_model/freeanchor_cfg.py:15-24).  anchors (A,4); pred_offsets (A,4); pred_scores (A,C) = sigmoid(logits);
    gt5 (G,5) with classes 1..C.  Returns
      box_prob (A,C)           the box-probability scatter (:54-84), including the reference's empty-set workaround,
      matched_idx (G,bucket)   the bag of every GT: its `bucket_size` best anchors by IoU (:87-92; F.topk(no_sort=True):
                               the order inside a bag is unspecified -- this restatement returns (IoU desc, index asc)),
      matched_score (G,bucket) pred_scores[bag anchor, class of the GT] (:95-101),
      matched_offsets (G*bucket,4) BoxCoder.encode(bag anchors, GT) (:103-110)."""
    anchors = np.asarray(anchors, f32)
    gt5 = np.asarray(gt5, f32)
    A, G = anchors.shape[0], gt5.shape[0]
    labels = gt5[:, 4].astype(np.int32) - 1                                     # :52
    pred_box, _ = boxcoder_decode(anchors, np.array(pred_offsets, f32, copy=True), reg_mean, reg_std)   # :55
    overlaps = box_iou(gt5[:, :4], pred_box)                                    # :57
    thresh1 = f32(box_iou_thresh)
    thresh2 = np.minimum(np.maximum(overlaps.max(axis=1, keepdims=True), f32(box_iou_thresh + clamp_eps)), f32(1.0))   # :60-64
    with np.errstate(all="ignore"):
        prob = ((overlaps - thresh1).astype(f32) / (thresh2 - thresh1).astype(f32)).astype(f32)
    prob = np.minimum(np.maximum(prob, f32(0)), f32(1.0))                       # :65-66
    fill = bool(prob.max() <= f32(clamp_eps))                                   # :71
    if fill:
        prob[0, 0] = f32(0.001)                                                 # :73
    nz = np.flatnonzero(prob.reshape(-1) != 0)                                  # :75 ascending flat indices (ASSUMED-4)
    a_idx, g_idx = nz % A, nz // A                                              # :79-80
    box_prob = np.zeros((A, num_classes), f32)
    for a, g in zip(a_idx.tolist(), g_idx.tolist()):                            # :83 last write wins (ASSUMED-11)
        box_prob[a, labels[g]] = prob[g, a]
    if fill:
        box_prob[0, 0] = 0.0                                                    # :85-86
    quality = box_iou(gt5[:, :4], anchors)                                      # :91
    k = min(int(bucket_size), A)
    matched_idx = np.argsort(-quality, axis=1, kind="stable")[:, :k].astype(np.int32)   # :93-95
    matched_score = np.asarray(pred_scores, f32)[matched_idx, labels[:, None]]  # :99-106
    flat = matched_idx.reshape(-1)
    gt_b = np.broadcast_to(gt5[:, None, :4], (G, k, 4)).reshape(-1, 4)          # :108-110
    matched_offsets = boxcoder_encode(anchors[flat], gt_b, reg_mean, reg_std)   # :111-114
    return box_prob, matched_idx, matched_score, matched_offsets, fill


# --------------------------------------------------------------------------- OTA (8(f)-3)
def logsigmoid_f32(x):
    """F.logsigmoid (ASSUMED-12)."""
    x = np.asarray(x, f32)
    with np.errstate(all="ignore"):
        return (np.minimum(x, f32(0)) - np.log1p(np.exp(-np.abs(x)).astype(f32)).astype(f32)).astype(f32)


def sigmoid_focal_loss(logits, targets, alpha=-1.0, gamma=0.0):
    """basedet/layers/losses/sigmoid_focal_loss.py:30-36 over cross_entropy.py:24-27, fp32."""
    logits, targets = np.asarray(logits, f32), np.asarray(targets, f32)
    scores = sigmoid_f32(logits)
    loss = (-((targets * logsigmoid_f32(logits)).astype(f32)
              + ((f32(1) - targets) * logsigmoid_f32(-logits)).astype(f32)).astype(f32)).astype(f32)
    if gamma != 0:
        base = ((targets * (f32(1) - scores)).astype(f32) + ((f32(1) - targets) * scores).astype(f32)).astype(f32)
        loss = (loss * np.power(base, f32(gamma)).astype(f32)).astype(f32)
    if alpha >= 0:
        loss = (loss * ((targets * f32(alpha)).astype(f32) + ((f32(1) - targets) * f32(1 - alpha)).astype(f32)).astype(f32)).astype(f32)
    return loss


def ltrb_iou(b1, b2, eps):
    """get_ltrb_boxes_iou(iou_type="iou"), basedet/layers/losses/iou_loss.py:9-43."""
    b1, b2 = np.asarray(b1, f32), np.asarray(b2, f32)
    b1 = np.concatenate([-b1[..., :2], b1[..., 2:]], axis=-1)
    b2 = np.concatenate([-b2[..., :2], b2[..., 2:]], axis=-1)
    a1 = (np.maximum(b1[..., 2] - b1[..., 0], f32(0)) * np.maximum(b1[..., 3] - b1[..., 1], f32(0))).astype(f32)
    a2 = (np.maximum(b2[..., 2] - b2[..., 0], f32(0)) * np.maximum(b2[..., 3] - b2[..., 1], f32(0))).astype(f32)
    w = np.maximum(np.minimum(b1[..., 2], b2[..., 2]) - np.maximum(b1[..., 0], b2[..., 0]), f32(0))
    h = np.maximum(np.minimum(b1[..., 3], b2[..., 3]) - np.maximum(b1[..., 1], b2[..., 1]), f32(0))
    inter = (w * h).astype(f32)
    union = ((a1 + a2).astype(f32) - inter).astype(f32)
    with np.errstate(all="ignore"):
        return (inter / np.maximum(union, f32(eps))).astype(f32)


def ota_cost(points_list, strides, gt5, cls_logits, pred_deltas, num_classes, alpha=0.25, gamma=2.0, reg_weight=1.5,
             center_sampling_radius=2.5):
    """The cost / IoU matrices of one image of OTA.get_ground_truth, basedet/models/det/ota.py:91-145.
    points_list[l] (n_l, 2); gt5 (G, 5) classes 1..C; cls_logits (A, C); pred_deltas (A, 4) ltrb.
    -> cost (G, A), ious (G, A), is_in_boxes (G, A) bool, gt_deltas (G, A, 4), loss_cls_bg (A,)."""
    gt5 = np.asarray(gt5, f32)
    pts = np.concatenate([np.asarray(p, f32) for p in points_list], axis=0)
    G, A = gt5.shape[0], pts.shape[0]
    deltas = pointcoder_encode(pts, gt5[:, None, :4])                           # :92  (G, A, 4)
    is_in_boxes = deltas.min(axis=-1) > f32(0.01)                               # :93
    gt_centers = ((gt5[:, :2] + gt5[:, 2:4]).astype(f32) / f32(2)).astype(f32)  # :97
    in_centers = []
    for stride, lp in zip(strides, points_list):                                # :99-109
        radius = f32(stride * center_sampling_radius)
        cb = np.concatenate([np.maximum((gt_centers - radius).astype(f32), gt5[:, :2]),
                             np.minimum((gt_centers + radius).astype(f32), gt5[:, 2:4])], axis=-1)
        cd = pointcoder_encode(np.asarray(lp, f32), cb[:, None, :])
        in_centers.append(cd.min(axis=-1) > 0)
    is_in_boxes = is_in_boxes & np.concatenate(in_centers, axis=1)              # :110-112
    onehot = (np.arange(num_classes)[None, :] == (gt5[:, 4].astype(np.int32) - 1)[:, None]).astype(f32)   # :115-117
    logits = np.asarray(cls_logits, f32)
    loss_cls = seq_sum_f32(sigmoid_focal_loss(np.broadcast_to(logits[None], (G, A, num_classes)),
                                              np.broadcast_to(onehot[:, None, :], (G, A, num_classes)), alpha, gamma), 2)   # :122-127
    loss_cls_bg = seq_sum_f32(sigmoid_focal_loss(logits, np.zeros_like(logits), alpha, gamma), 1)   # :129-134
    gt_delta = pointcoder_encode(pts, gt5[:, None, :4])                         # :136-138
    ious = ltrb_iou(np.broadcast_to(np.asarray(pred_deltas, f32)[None], (G, A, 4)), gt_delta, np.finfo(np.float32).eps)
    with np.errstate(all="ignore"):
        loss_delta = (-np.log(np.maximum(ious, f32(np.finfo(np.float32).eps)))).astype(f32)   # iou_loss.py:96
    cost = ((loss_cls + (f32(reg_weight) * loss_delta).astype(f32)).astype(f32)
            + (f32(1e6) * (~is_in_boxes).astype(f32)).astype(f32)).astype(f32)  # :152
    return cost, ious, is_in_boxes, gt_delta, loss_cls_bg


def ota_targets(points_list, strides, gt5, cls_logits, pred_deltas, num_classes, candidate_k=10, **kw):
    """One image of OTA.get_ground_truth with matching == "topk", ota.py:91-172.
    -> gt_classes (A,) fp32 (0 = background), box targets (A, 4), ious (A,), matched (A,) int32 (G = background)."""
    cost, ious, _, gt_delta, _ = ota_cost(points_list, strides, gt5, cls_logits, pred_deltas, num_classes, **kw)
    G, A = cost.shape
    matched = ota_topk_match(cost, ious, candidate_k)                           # :158
    fg = matched != G                                                           # :161
    cls_t = np.zeros(A, f32)
    cls_t[fg] = np.asarray(gt5, f32)[:, 4][matched[fg]]                         # :162
    box_t = np.zeros((A, 4), f32)
    box_t[fg] = gt_delta[matched[fg], np.arange(A)[fg]]                         # :165-168
    iou_t = np.zeros(A, f32)
    iou_t[fg] = ious[matched[fg], np.arange(A)[fg]]                             # :171-175
    return cls_t, box_t, iou_t, matched


# --------------------------------------------------------------------------- COCO result records (8(f)-4)
def coco_format(dets, counts, image_ids, category_ids=None):
    """COCOEvaluator.format, basedet/evaluators/coco_eval.py:111-138, for padded detections (B, K, 6) rows
    [x1, y1, x2, y2, score, label] with `counts[b]` valid rows.  -> image_id (N,), bbox xywh (N, 4), score (N,),
    category_id (N,): `category_ids[label]` (classes_originID) when given, else label + 1; images without detections
    contribute nothing (:123-124)."""
    dets = np.asarray(dets, f32)
    img, box, sc, cat = [], [], [], []
    for b, n in enumerate(np.asarray(counts).tolist()):
        if n <= 0:
            continue
        d = dets[b, :n].astype(np.float64)                                      # np.array(..., dtype=np.float) (:107)
        d[:, 2:4] = d[:, 2:4] - d[:, 0:2]                                       # :125
        for row in d:
            img.append(int(image_ids[b]))
            box.append(row[:4])
            sc.append(row[4])
            cat.append(int(category_ids[int(row[5])]) if category_ids is not None else int(row[5]) + 1)
    return (np.array(img, np.int32), np.array(box, np.float64).reshape(-1, 4), np.array(sc, np.float64), np.array(cat, np.int32))


def _ctrness(offsets):
    """fcos.py:276-281 / atss.py:72-77: sqrt(max(min(l,r)/max(l,r), 0) * max(min(t,b)/max(t,b), 0)).
    F.maximum(x, 0) and F.clip(x, lower=0) are both the elementwise MAX of ASSUMED-1 (NaN -> 0)."""
    lr, tb = offsets[:, [0, 2]], offsets[:, [1, 3]]
    with np.errstate(all="ignore"):
        a = (lr.min(axis=1) / lr.max(axis=1)).astype(f32)
        b = (tb.min(axis=1) / tb.max(axis=1)).astype(f32)
        return np.sqrt((emax(a, _ZERO) * emax(b, _ZERO)).astype(f32)).astype(f32)


def fcos_targets(points_list, gt_boxes, num_gt, strides, sizes_of_interest, center_sampling_radius):
    """FCOS.get_ground_truth, basedet/models/det/fcos.py:222-293.

    points_list: L arrays (n_l, 2); gt_boxes (B, Gmax, 5); num_gt (B,).  Returns labels (B, A) int32 (class of the
    smallest-area GT that contains the point [inside its centre box when radius > 0] and cares about the level,
    0 = background), offsets (B, A, 4) = PointCoder.encode(point, matched GT) (GT 0 for background, as the
    reference's argmin over an all-inf column gives index 0), ctrness (B, A), match_indices (B, A)."""
    points = np.concatenate([np.asarray(p, f32) for p in points_list], axis=0)
    lo = np.concatenate([np.full(len(p), f32(s[0]), f32) for p, s in zip(points_list, sizes_of_interest)])  # :236-243
    hi = np.concatenate([np.full(len(p), f32(s[1]), f32) for p, s in zip(points_list, sizes_of_interest)])
    lab_l, off_l, ctr_l, idx_l = [], [], [], []
    for g5, n in zip(gt_boxes, num_gt):
        g5 = np.asarray(g5, f32)[: int(n)]
        gt = g5[:, :4]
        offsets = pointcoder_encode(points, gt[:, None, :])                      # :231 (G, A, 4)
        max_off = offsets.max(axis=2)                                           # :245
        cared = (max_off >= lo[None, :]) & (max_off <= hi[None, :])             # :246-249
        if center_sampling_radius > 0:                                          # :251-264
            ctr = box_center(gt)
            parts = []
            for stride, pts in zip(strides, points_list):
                radius = stride * center_sampling_radius                        # python float, cast to fp32 per op
                cb = np.concatenate([emax((ctr - f32(radius)).astype(f32), gt[:, :2]),
                                     emin((ctr + f32(radius)).astype(f32), gt[:, 2:4])], axis=1)
                co = pointcoder_encode(np.asarray(pts, f32), cb[:, None, :])
                parts.append(co.min(axis=2) > 0)
            in_boxes = np.concatenate(parts, axis=1)
        else:
            in_boxes = offsets.min(axis=2) > 0                                  # :266
        area = ((gt[:, 2] - gt[:, 0]).astype(f32) * (gt[:, 3] - gt[:, 1]).astype(f32)).astype(f32)  # boxes.py:36-42
        areas = np.broadcast_to(area[:, None], max_off.shape).copy()            # :268
        areas[~cared] = np.inf                                                  # :269-270
        areas[~in_boxes] = np.inf
        idx = np.argmin(areas, axis=0).astype(np.int32)                         # :272 (first index, ASSUMED-9)
        matched = g5[idx]
        min_area = areas[idx, np.arange(areas.shape[1])]                        # :274
        labels = matched[:, 4].astype(np.int32)                                 # :276
        labels[min_area == np.inf] = 0                                          # :277
        off = pointcoder_encode(points, matched[:, :4])                         # :278
        lab_l.append(labels)
        off_l.append(off)
        ctr_l.append(_ctrness(off))
        idx_l.append(idx)
    return np.stack(lab_l), np.stack(off_l), np.stack(ctr_l), np.stack(idx_l)


def seq_sum_f32(x, axis):
    """ASSUMED-8: fp32 sum accumulated sequentially in index order."""
    return np.cumsum(np.asarray(x, f32), axis=axis, dtype=f32).take(-1, axis=axis)


def atss_targets(points_list, gt_boxes, num_gt, strides, anchor_scale, topk):
    """ATSS.get_ground_truth, basedet/models/det/atss.py:17-86.  Same returns as fcos_targets."""
    points = np.concatenate([np.asarray(p, f32) for p in points_list], axis=0)
    lab_l, off_l, ctr_l, idx_l = [], [], [], []
    for g5, n in zip(gt_boxes, num_gt):
        g5 = np.asarray(g5, f32)[: int(n)]
        gt = g5[:, :4]
        ious, cands, base = [], [], 0
        ctr = box_center(gt)                                                    # :39
        for stride, pts in zip(strides, points_list):
            pts = np.asarray(pts, f32)
            half = f32(stride * anchor_scale / 2)                               # :33-34 python float -> fp32 scalar
            boxes = np.concatenate([(pts - half).astype(f32), (pts + half).astype(f32)], axis=1)
            ious.append(box_iou(gt, boxes))                                     # :31-37 gt_boxes.iou(anchor boxes)
            d = (ctr[:, None, :] - pts[None, :, :]).astype(f32)
            d2 = (d * d).astype(f32)                                            # ** 2 (ASSUMED-9)
            dist = np.sqrt((d2[..., 0] + d2[..., 1]).astype(f32)).astype(f32)   # :40-42
            k = min(int(topk), dist.shape[1])
            order = np.argsort(dist, axis=1, kind="stable")[:, :k]              # :43 topk ascending (ASSUMED-9)
            cands.append(base + order)
            base += len(pts)
        ious = np.concatenate(ious, axis=1)                                     # :46
        cands = np.concatenate(cands, axis=1)                                   # :47
        cand_iou = np.take_along_axis(ious, cands, axis=1)                      # :49
        ncand = f32(cand_iou.shape[1])
        mean = (seq_sum_f32(cand_iou, 1) / ncand).astype(f32)                   # :50 (ASSUMED-8)
        dev = (cand_iou - mean[:, None]).astype(f32)
        std = np.sqrt((seq_sum_f32((dev * dev).astype(f32), 1) / ncand).astype(f32)).astype(f32)  # :51
        thr = (mean + std).astype(f32)
        is_cand = np.zeros(ious.shape, bool)
        np.put_along_axis(is_cand, cands, True, axis=1)                         # :52-54
        fg = is_cand & (ious >= thr[:, None])
        in_boxes = pointcoder_encode(points, gt[:, None, :]).min(axis=2) > 0    # :56-58
        ious = ious.copy()
        ious[~fg] = -1                                                          # :60-61
        ious[~in_boxes] = -1
        idx = np.argmax(ious, axis=0).astype(np.int32)                          # :63 (ASSUMED-2)
        matched = g5[idx]
        max_iou = ious[idx, np.arange(ious.shape[1])]                           # :65
        labels = matched[:, 4].astype(np.int32)                                 # :67
        labels[max_iou == -1] = 0                                               # :68
        off = pointcoder_encode(points, matched[:, :4])                         # :69
        lab_l.append(labels)
        off_l.append(off)
        ctr_l.append(_ctrness(off))
        idx_l.append(idx)
    return np.stack(lab_l), np.stack(off_l), np.stack(ctr_l), np.stack(idx_l)


# --------------------------------------------------------------------------- score filter + top-k
def sigmoid_f32(x):
    """F.sigmoid in fp32: 1 / (1 + exp(-x)) (ulp-level differences vs CUDA expected)."""
    x = np.asarray(x, dtype=f32)
    with np.errstate(over="ignore"):
        return (f32(1) / (f32(1) + np.exp(-x).astype(f32)).astype(f32)).astype(f32)


def filter_topk_scores(scores_flat, cls_threshold, topk=1000):
    """Score filter + top-k on a GIVEN fp32 score vector.

    basedet/models/det/retinanet.py:185-191 / fcos.py:196-202:
    cand = ascending flat idx with score > thr; keep = cand[topk_desc(scores[cand])].
    Returns (keep_idx int32 sorted by score desc, scores[keep_idx]); empty if no candidate.
    """
    scores_flat = np.asarray(scores_flat, dtype=f32).reshape(-1)
    _, keep = cond_take(scores_flat > f32(cls_threshold), scores_flat)
    if keep.size == 0:
        return keep, scores_flat[keep]
    k = min(keep.shape[0], topk)
    _, top = topk_desc(scores_flat[keep], k)
    keep = keep[top]
    return keep, scores_flat[keep]


def retinanet_level_select(logits, offsets, anchors, cls_threshold, topk=1000,
                           mean=(0, 0, 0, 0), std=(1, 1, 1, 1), scores=None):
    """One iteration of the level loop of RetinaNet.inference, retinanet.py:181-196.

    logits (HWA, C); offsets (HWA, 4); anchors (HWA, 4).  ``scores`` may be supplied
    (bit-identical score tensor); otherwise sigmoid_f32(logits).
    Returns (boxes (k,4), scores (k,), labels (k,) int32, keep_idx (k,) int32) or None.
    """
    num_classes = logits.shape[-1]
    s = sigmoid_f32(logits).reshape(-1) if scores is None else np.asarray(scores, f32).reshape(-1)
    keep, sc = filter_topk_scores(s, cls_threshold, topk)
    if keep.size == 0:
        return None
    boxes, _ = boxcoder_decode(anchors, np.asarray(offsets, f32).reshape(-1, 4), mean, std)
    return boxes[keep // num_classes], sc, (keep % num_classes).astype(np.int32), keep


def fcos_scores(logits, ctrness):
    """basedet/models/det/fcos.py:194: sqrt(sigmoid(cls) * sigmoid(ctr)), (HW,C)*(HW,1)."""
    return np.sqrt((sigmoid_f32(logits) * sigmoid_f32(ctrness)).astype(f32)).astype(f32)


# --------------------------------------------------------------------------- NMS
def nms(boxes, scores, iou_thresh, max_output=None):
    """megengine.functional.vision.nms (ASSUMED-3, ASSUMED-5).

    Returns original indices of kept boxes in score-descending order (int32).
    """
    boxes = np.asarray(boxes, dtype=f32)
    scores = np.asarray(scores, dtype=f32)
    n = boxes.shape[0]
    order = argsort_desc(scores)
    b = boxes[order]
    area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(f32)
    removed = np.zeros(n, dtype=bool)
    thr = f32(iou_thresh)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        if max_output is not None and len(keep) >= max_output:
            break
        r = b[i + 1:]
        left = np.maximum(b[i, 0], r[:, 0])
        right = np.minimum(b[i, 2], r[:, 2])
        top = np.maximum(b[i, 1], r[:, 1])
        bottom = np.minimum(b[i, 3], r[:, 3])
        w = np.maximum((right - left).astype(f32), _ZERO)
        h = np.maximum((bottom - top).astype(f32), _ZERO)
        inter = (w * h).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = (inter / ((area[i] + area[i + 1:]).astype(f32) - inter).astype(f32)).astype(f32)
        removed[i + 1:] |= iou > thr
    return order[np.asarray(keep, dtype=np.int64)].astype(np.int32)


def nms_offset_boxes(boxes, idxs):
    """The class-offset trick, basedet/layers/common/post_processing.py:44-46 (fp32)."""
    boxes = np.asarray(boxes, dtype=f32)
    idxs = np.asarray(idxs)
    max_coordinate = boxes.max()
    offsets = (idxs.astype(f32) * (max_coordinate + f32(1)).astype(f32)).astype(f32)
    return (boxes + offsets.reshape(-1, 1)).astype(f32)


def batched_nms(boxes, scores, idxs, iou_thresh, max_output=None):
    """basedet/layers/common/post_processing.py:17-47."""
    boxes = np.asarray(boxes, dtype=f32)
    scores = np.asarray(scores, dtype=f32)
    idxs = np.asarray(idxs)
    assert boxes.ndim == 2 and boxes.shape[1] == 4, "the expected shape of boxes is (N, 4)"
    assert scores.ndim == 1, "the expected shape of scores is (N,)"
    assert idxs.ndim == 1, "the expected shape of idxs is (N,)"
    assert boxes.shape[0] == scores.shape[0] == idxs.shape[0], \
        "number of boxes, scores and idxs are not matched"
    if boxes.shape[0] == 0:
        return np.zeros((0,), dtype=np.int32)
    return nms(nms_offset_boxes(boxes, idxs), scores, iou_thresh, max_output)


def py_cpu_nms(dets, thresh):
    """Restatement of the reference's own numpy helper, post_processing.py:106-132.

    (tests additionally ast-extract the ORIGINAL function when /root/reference exists.)
    """
    dets = np.asarray(dets)
    x1, y1, x2, y2 = [np.ascontiguousarray(dets[:, i]) for i in range(4)]
    areas = (x2 - x1) * (y2 - y1)
    order = dets[:, 4].argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(i)
        order = order[1:]
        xx1, yy1 = np.maximum(x1[i], x1[order]), np.maximum(y1[i], y1[order])
        xx2, yy2 = np.minimum(x2[i], x2[order]), np.minimum(y2[i], y2[order])
        inter = np.maximum(xx2 - xx1, 0) * np.maximum(yy2 - yy1, 0)
        iou = inter / np.maximum(areas[i] + areas[order] - inter, 1e-5)
        order = order[iou <= thresh]
    return keep


def post_processing(boxes, box_scores, box_labels, img_info, iou_threshold,
                    max_detections_per_image=None):
    """basedet/layers/common/post_processing.py:78-103 -> (boxes, scores, labels, keep)."""
    boxes = np.asarray(boxes, dtype=f32)
    img_info = np.asarray(img_info, dtype=f32)
    keep = batched_nms(boxes, box_scores, box_labels, iou_threshold, max_detections_per_image)
    kb = boxes[keep]
    scale_ratios = (img_info[0, 2] / img_info[0, 0], img_info[0, 3] / img_info[0, 1])
    kb = boxes_scale(kb, scale_ratios)
    kb = boxes_clip(kb, img_info[0, 2:4])
    return kb, np.asarray(box_scores, f32)[keep], np.asarray(box_labels)[keep], keep


def retinanet_postprocess(logits_list, offsets_list, anchors_list, img_info, cls_threshold=0.05,
                          iou_threshold=0.5, max_dets=100, topk=1000, scores_list=None):
    """RetinaNet.inference minus the network, retinanet.py:172-209 (single image)."""
    tb, ts, tl = [], [], []
    for lvl, (lg, of, an) in enumerate(zip(logits_list, offsets_list, anchors_list)):
        sc = None if scores_list is None else scores_list[lvl]
        r = retinanet_level_select(lg, of, an, cls_threshold, topk, scores=sc)
        if r is None:
            continue
        tb.append(r[0]); ts.append(r[1]); tl.append(r[2])
    if not tb:
        e = np.zeros((0,), dtype=f32)
        return e.reshape(0, 4), e, e.astype(np.int32), e.astype(np.int32)
    return post_processing(np.concatenate(tb), np.concatenate(ts), np.concatenate(tl),
                           img_info, iou_threshold, max_dets)


# --------------------------------------------------------------------------- ROI pooling
def assign_levels(rois, strides):
    """assign_rois without the dummy rows, basedet/layers/common/roi_pool.py:12-25 -> int32 (K,)."""
    rois = np.asarray(rois, dtype=f32)
    min_level, max_level = int(math.log2(strides[0])), int(math.log2(strides[-1]))
    box_area = ((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2])).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        lvl = np.floor(
            (f32(4) + (np.log((np.sqrt(box_area).astype(f32) / f32(224)).astype(f32)).astype(f32)
                       / f32(math.log(2))).astype(f32)).astype(f32)
        )
        # float -> int32 of NaN / -inf: x86 cvttss2si yields INT_MIN; the clamp makes it min_level.
        lvl = np.where(np.isfinite(lvl), lvl, -1e9)
    lvl = np.clip(lvl, -2 ** 31, 2 ** 31 - 1).astype(np.int64)
    lvl = np.minimum(lvl, max_level)
    lvl = np.maximum(lvl, min_level)
    return (lvl - min_level).astype(np.int32)


def _bilinear_setup(h, w, height, width):
    h0 = np.floor(h).astype(np.int64)
    w0 = np.floor(w).astype(np.int64)
    return h0, w0, h0 + 1, w0 + 1, (h - h0.astype(f32)).astype(f32), (w - w0.astype(f32)).astype(f32)


def roi_align(feat, rois, pool_shape, spatial_scale, sample_points=2, aligned=True):
    """megengine.functional.nn.roi_align(mode="average") (ASSUMED-6), MegDNN roi_align order.

    feat (B,C,H,W) fp32; rois (K,5) [batch, x1, y1, x2, y2] -> (K,C,PH,PW).
    """
    feat = np.asarray(feat, dtype=f32)
    rois = np.asarray(rois, dtype=f32)
    if isinstance(pool_shape, int):
        pool_shape = (pool_shape, pool_shape)
    ph_n, pw_n = pool_shape
    if isinstance(sample_points, int):
        sample_points = (sample_points, sample_points)
    sh, sw = sample_points
    _, C, H, W = feat.shape
    K = rois.shape[0]
    scale = f32(spatial_scale)
    offset = f32(0.5) if aligned else f32(0.0)
    out = np.zeros((K, C, ph_n, pw_n), dtype=f32)
    ph = np.arange(ph_n, dtype=f32).reshape(-1, 1)
    pw = np.arange(pw_n, dtype=f32).reshape(1, -1)
    for k in range(K):
        n = int(rois[k, 0])
        fm = feat[n]
        start_w = (rois[k, 1] * scale - offset).astype(f32)
        start_h = (rois[k, 2] * scale - offset).astype(f32)
        end_w = (rois[k, 3] * scale - offset).astype(f32)
        end_h = (rois[k, 4] * scale - offset).astype(f32)
        roi_w = np.maximum((end_w - start_w).astype(f32), _ZERO)
        roi_h = np.maximum((end_h - start_h).astype(f32), _ZERO)
        bin_h = (roi_h / f32(ph_n)).astype(f32)
        bin_w = (roi_w / f32(pw_n)).astype(f32)
        acc = np.zeros((C, ph_n, pw_n), dtype=f32)
        for iy in range(sh):
            for ix in range(sw):
                fy = (f32(iy + 0.5) / f32(sh)).astype(f32)
                fx = (f32(ix + 0.5) / f32(sw)).astype(f32)
                hc = (start_h + (bin_h * (ph + fy).astype(f32)).astype(f32)).astype(f32)
                wc = (start_w + (bin_w * (pw + fx).astype(f32)).astype(f32)).astype(f32)
                hc = np.broadcast_to(hc, (ph_n, pw_n))
                wc = np.broadcast_to(wc, (ph_n, pw_n))
                h0, w0, h1, w1, lh, lw = _bilinear_setup(hc, wc, H, W)

                def tap(hh, ww):
                    ok = (hh >= 0) & (hh < H) & (ww >= 0) & (ww < W)
                    v = fm[:, np.clip(hh, 0, H - 1), np.clip(ww, 0, W - 1)]
                    return np.where(ok[None], v, _ZERO).astype(f32)

                tl, tr, bl, br = tap(h0, w0), tap(h0, w1), tap(h1, w0), tap(h1, w1)
                top = (tl + ((tr - tl).astype(f32) * lw[None]).astype(f32)).astype(f32)
                bot = (bl + ((br - bl).astype(f32) * lw[None]).astype(f32)).astype(f32)
                val = (top + ((bot - top).astype(f32) * lh[None]).astype(f32)).astype(f32)
                acc = (acc + val).astype(f32)
        out[k] = (acc / f32(sh * sw)).astype(f32)
    return out


def roi_align_backward(dout, feat_shape, rois, pool_shape, spatial_scale, sample_points=2,
                       aligned=True, accumulate=np.float64):
    """Gradient of roi_align w.r.t. feat (ASSUMED-6; reference test-suite never pins it).

    Each of the S^2 samples of a bin passes dout/S^2 to its 4 taps with the bilinear weights
    (1-lh)(1-lw), (1-lh)lw, lh(1-lw), lh*lw; out-of-range taps are dropped.
    ``accumulate``: dtype of the accumulation buffer (float64 = order-independent reference).
    """
    dout = np.asarray(dout, dtype=f32)
    rois = np.asarray(rois, dtype=f32)
    if isinstance(pool_shape, int):
        pool_shape = (pool_shape, pool_shape)
    ph_n, pw_n = pool_shape
    if isinstance(sample_points, int):
        sample_points = (sample_points, sample_points)
    sh, sw = sample_points
    B, C, H, W = feat_shape
    grad = np.zeros((B, C, H * W), dtype=accumulate)
    scale = f32(spatial_scale)
    offset = f32(0.5) if aligned else f32(0.0)
    ph = np.arange(ph_n, dtype=f32).reshape(-1, 1)
    pw = np.arange(pw_n, dtype=f32).reshape(1, -1)
    for k in range(rois.shape[0]):
        n = int(rois[k, 0])
        start_w = (rois[k, 1] * scale - offset).astype(f32)
        start_h = (rois[k, 2] * scale - offset).astype(f32)
        end_w = (rois[k, 3] * scale - offset).astype(f32)
        end_h = (rois[k, 4] * scale - offset).astype(f32)
        roi_w = np.maximum((end_w - start_w).astype(f32), _ZERO)
        roi_h = np.maximum((end_h - start_h).astype(f32), _ZERO)
        bin_h = (roi_h / f32(ph_n)).astype(f32)
        bin_w = (roi_w / f32(pw_n)).astype(f32)
        g = (dout[k] / f32(sh * sw)).astype(f32)  # (C, PH, PW)
        for iy in range(sh):
            for ix in range(sw):
                fy = (f32(iy + 0.5) / f32(sh)).astype(f32)
                fx = (f32(ix + 0.5) / f32(sw)).astype(f32)
                hc = np.broadcast_to((start_h + (bin_h * (ph + fy).astype(f32)).astype(f32)).astype(f32),
                                     (ph_n, pw_n))
                wc = np.broadcast_to((start_w + (bin_w * (pw + fx).astype(f32)).astype(f32)).astype(f32),
                                     (ph_n, pw_n))
                h0, w0, h1, w1, lh, lw = _bilinear_setup(hc, wc, H, W)
                one = f32(1)
                for hh, ww, wt in (
                    (h0, w0, ((one - lh) * (one - lw)).astype(f32)),
                    (h0, w1, ((one - lh) * lw).astype(f32)),
                    (h1, w0, (lh * (one - lw)).astype(f32)),
                    (h1, w1, (lh * lw).astype(f32)),
                ):
                    ok = (hh >= 0) & (hh < H) & (ww >= 0) & (ww < W)
                    if not ok.any():
                        continue
                    flat = (hh * W + ww)[ok]
                    contrib = (g[:, ok] * wt[ok][None]).astype(f32)
                    for c in range(C):
                        np.add.at(grad[n, c], flat, contrib[c].astype(accumulate))
    return grad.reshape(B, C, H, W)


def roi_max_pooling(feat, rois, pool_shape, spatial_scale):
    """megengine.functional.nn.roi_pooling(mode="max") (ASSUMED-13: MegDNN ROIPooling = the Caffe rule; pinned by the
    reference's own known-answer test, tests/layers/test_roi_pool.py:48-61).  feat (B,C,H,W); rois (K,5)."""
    feat = np.asarray(feat, f32)
    rois = np.asarray(rois, f32)
    B, C, H, W = feat.shape
    PH, PW = pool_shape
    out = np.zeros((rois.shape[0], C, PH, PW), f32)
    rnd = lambda v: int(math.floor(abs(float(v)) + 0.5) * (1 if v >= 0 else -1))   # C roundf: half away from zero
    for k, r in enumerate(rois):
        n = int(r[0])
        x0, y0, x1, y1 = (rnd(f32(r[i]) * f32(spatial_scale)) for i in (1, 2, 3, 4))
        rw, rh = max(x1 - x0 + 1, 1), max(y1 - y0 + 1, 1)
        bh, bw = f32(rh) / f32(PH), f32(rw) / f32(PW)
        for ph in range(PH):
            hs = min(max(int(math.floor(f32(ph) * bh)) + y0, 0), H)
            he = min(max(int(math.ceil(f32(ph + 1) * bh)) + y0, 0), H)
            for pw in range(PW):
                ws = min(max(int(math.floor(f32(pw) * bw)) + x0, 0), W)
                we = min(max(int(math.ceil(f32(pw + 1) * bw)) + x0, 0), W)
                if he > hs and we > ws:
                    out[k, :, ph, pw] = feat[n, :, hs:he, ws:we].reshape(C, -1).max(axis=1)
    return out


def roi_pool(features, rois, strides, pool_shape, pooler_type="roi_align"):
    """basedet/layers/common/roi_pool.py:35-78.

    Follows the reference literally: dummy ROI per level (:28-31), per-level pooling,
    argsort re-ordering, dummies dropped (:74-76).
    """
    assert pooler_type in ("roi_align", "roi_pool")
    assert len(strides) == len(features)
    if isinstance(pool_shape, int):
        pool_shape = (pool_shape, pool_shape)
    rois = np.asarray(rois, dtype=f32)
    num_fms = len(strides)
    lvl = np.concatenate([assign_levels(rois, strides), np.arange(num_fms, dtype=np.int32)])
    rois_d = np.concatenate([rois, np.zeros((num_fms, rois.shape[-1]), dtype=f32)])
    pool_list, inds_list = [], []
    for i, (feat, stride) in enumerate(zip(features, strides)):
        inds = np.flatnonzero(lvl == i)
        if pooler_type == "roi_pool":
            pool_list.append(roi_max_pooling(feat, rois_d[inds], pool_shape, 1.0 / stride))
        else:
            pool_list.append(roi_align(feat, rois_d[inds], pool_shape, 1.0 / stride, 2, True))
        inds_list.append(inds)
    fm_order = np.argsort(np.concatenate(inds_list), kind="stable")
    pooled = np.concatenate(pool_list, axis=0)
    return pooled[fm_order][:-num_fms]
