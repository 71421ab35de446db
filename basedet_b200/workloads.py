"""Seeded synthetic inputs for the BASELINE.json configs.  Host-side numpy only.

Used by tests/ (parity inputs) and bench.py (timed inputs).  Nothing here is on the hot path.
"""
import math

import numpy as np

f32 = np.float32

RETINANET_STRIDES = [8, 16, 32, 64, 128]
# models/det/retinanet_cfg.py:21-28 (configs/det_model/retinanet_cfg.py ANCHOR)
RETINANET_SCALES = [[x, x * 2 ** (1.0 / 3), x * 2 ** (2.0 / 3)] for x in [32, 64, 128, 256, 512]]
RETINANET_RATIOS = [[0.5, 1, 2]]
RETINANET_OFFSET = 0.5
RETINANET_MATCHER = dict(thresholds=[0.4, 0.5], labels=[0, -1, 1], allow_low_quality=True)

# configs/det_model/faster_rcnn_cfg.py
FRCNN_RPN_STRIDES = [4, 8, 16, 32, 64]
FRCNN_SCALES = [[x] for x in [32, 64, 128, 256, 512]]
FRCNN_RATIOS = [[0.5, 1, 2]]
FRCNN_OFFSET = 0.5
FRCNN_RCNN_STRIDES = [4, 8, 16, 32]
RPN_MATCHER = dict(thresholds=[0.3, 0.7], labels=[0, -1, 1], allow_low_quality=True)


def retinanet_level_sizes(h, w):
    """FPN P3..P7 sizes: P3-P5 = ceil(dim/stride) on the /32-padded image, P6/P7 from stride-2 pad-1 3x3 convs
    (layers/backbone/fpn_backbone.py:198-199)."""
    sizes = []
    ph, pw = h, w
    for s in (8, 16, 32):
        sizes.append((math.ceil(ph / s), math.ceil(pw / s)))
    for _ in range(2):
        lh, lw = sizes[-1]
        sizes.append(((lh - 1) // 2 + 1, (lw - 1) // 2 + 1))
    return sizes


def frcnn_level_sizes(h, w):
    """P2..P5 = dim/stride, P6 = max_pool(k=1, s=2) of P5 (layers/backbone/fpn_backbone.py:183)."""
    sizes = [(math.ceil(h / s), math.ceil(w / s)) for s in (4, 8, 16, 32)]
    lh, lw = sizes[-1]
    sizes.append(((lh - 1) // 2 + 1, (lw - 1) // 2 + 1))
    return sizes


def make_gt(rng, n, img_h, img_w, size_lo=16.0, size_hi=512.0, num_classes=80):
    """GT boxes (n, 5): centres U(0,W)xU(0,H), sqrt-area log-U[lo,hi], aspect log-U[1/3,3], clipped, class U{1..C}."""
    cx = rng.uniform(0, img_w, n)
    cy = rng.uniform(0, img_h, n)
    size = np.exp(rng.uniform(math.log(size_lo), math.log(size_hi), n))
    aspect = np.exp(rng.uniform(math.log(1 / 3), math.log(3), n))
    w = size * np.sqrt(aspect)
    h = size / np.sqrt(aspect)
    x1 = np.clip(cx - w / 2, 0, img_w)
    y1 = np.clip(cy - h / 2, 0, img_h)
    x2 = np.clip(cx + w / 2, 0, img_w)
    y2 = np.clip(cy + h / 2, 0, img_h)
    # keep boxes non-degenerate after clipping
    x2 = np.maximum(x2, x1 + 1.0)
    y2 = np.maximum(y2, y1 + 1.0)
    cls = rng.integers(1, num_classes + 1, n)
    return np.stack([x1, y1, x2, y2, cls], axis=1).astype(f32)


def target_assign_batch(batch, num_gt=100, img_h=800, img_w=800, seed0=100, ragged=False):
    """Config 2: gt_boxes (B, Gmax, 5) fp32 + num_gt (B,) int32.  seed = seed0 + image index."""
    gt = np.zeros((batch, num_gt, 5), dtype=f32)
    ng = np.zeros((batch,), dtype=np.int32)
    for b in range(batch):
        rng = np.random.default_rng(seed0 + b)
        n = num_gt if not ragged else int(rng.integers(1, num_gt + 1))
        gt[b, :n] = make_gt(rng, n, img_h, img_w)
        ng[b] = n
    return gt, ng


def distinct_scores(rng, n, lo=0.0, hi=1.0):
    """n pairwise-distinct fp32 scores in (lo, hi): random permutation of a strictly monotone grid."""
    grid = np.linspace(lo, hi, n + 2, dtype=np.float64)[1:-1].astype(f32)
    assert np.unique(grid).size == n, "grid too dense for fp32"
    return grid[rng.permutation(n)]


def logits_level(rng, n_anchor, num_classes, mean=-6.0, std=1.25):
    """Head logits ~ N(-6, 1.25^2): ~0.5 % above sigmoid^-1(0.05) (prior_prob 0.01 after training)."""
    return rng.normal(mean, std, size=(n_anchor, num_classes)).astype(f32)


def deltas_level(rng, n_anchor, std=(0.1, 0.1, 0.2, 0.2)):
    return (rng.normal(0, 1, size=(n_anchor, 4)) * np.asarray(std)).astype(f32)


def make_rois(rng, k, batch, img_h, img_w, size_lo=8.0, size_hi=600.0):
    """(k*batch, 5) rois [batch_idx, x1, y1, x2, y2], k per image, sizes log-U[lo,hi], clipped."""
    out = []
    for b in range(batch):
        g = make_gt(rng, k, img_h, img_w, size_lo, size_hi)
        out.append(np.concatenate([np.full((k, 1), b, dtype=f32), g[:, :4]], axis=1))
    return np.concatenate(out, axis=0).astype(f32)


# ---------------------------------------------------------------------------------------------------------------
# Per-image inputs of the BASELINE configs as a function of the GLOBAL image index, so that a batch sharded over N
# ranks (distributed.shard_range) sees exactly the tensors the 1-rank run sees.
FRCNN_HW = (800, 1344)
FRCNN_CHANNELS = 256
FRCNN_NUM_ROIS = 512
FRCNN_PRE_NMS, FRCNN_POST_NMS, FRCNN_NMS_THR = 2000, 1000, 0.7
FCOS_HW = (800, 1344)
FCOS_NMS_THR = 0.6
STRESS_HW = (800, 1333)


def frcnn_image(img, hw=FRCNN_HW, num_gt=100):
    """Config 3, one image: RPN head outputs in (h, w, anchor) order, GT, im_info and the uniform variates of the four
    sample_labels calls (RNG contract of include/bdet.h).  SURVEY 8(d): logits ~ N(-3, 2^2), deltas ~ N(0, 0.2^2)."""
    rng = np.random.default_rng(3000 + img)
    sizes = frcnn_level_sizes(*hw)
    n_l = [h * w * 3 for h, w in sizes]
    A = sum(n_l)
    out = {
        "scores": [rng.normal(-3.0, 2.0, n).astype(f32) for n in n_l],
        "deltas": [rng.normal(0.0, 0.2, (n, 4)).astype(f32) for n in n_l],
        "gt": make_gt(rng, num_gt, hw[0], hw[1]),
        "num_gt": np.int32(num_gt),
        "im_info": np.array([hw[0], hw[1], hw[0], hw[1] - 11, num_gt], dtype=f32),
        "noise_rpn": rng.random((2, A), dtype=f32),
        "noise_rcnn": rng.random((2, FRCNN_POST_NMS + num_gt), dtype=f32),
    }
    return out


def frcnn_batch(images, hw=FRCNN_HW, num_gt=100):
    """Stack ``frcnn_image`` over a list of global image indices -> dict of (B, ...) arrays (levels stay lists)."""
    per = [frcnn_image(i, hw, num_gt) for i in images]
    L = len(per[0]["scores"])
    return {
        "scores": [np.stack([p["scores"][l] for p in per]) for l in range(L)],
        "deltas": [np.stack([p["deltas"][l] for p in per]) for l in range(L)],
        "gt": np.stack([p["gt"] for p in per]),
        "num_gt": np.array([p["num_gt"] for p in per], np.int32),
        "im_info": np.stack([p["im_info"] for p in per]),
        "noise_rpn": np.stack([p["noise_rpn"] for p in per]),
        "noise_rcnn": np.stack([p["noise_rcnn"] for p in per]),
    }


def fcos_image(img, hw=FCOS_HW, num_classes=80):
    """Config 4, one image: FCOS head outputs per level -- logits (n_l, C) ~ N(-6, 1.25^2), centerness (n_l, 1) ~ N(0, 1),
    ltrb offsets (n_l, 4) = |N(0, 1)| * stride * 4."""
    rng = np.random.default_rng(4000 + img)
    sizes = retinanet_level_sizes(*hw)
    out = {"logits": [], "ctrness": [], "offsets": []}
    for (h, w), s in zip(sizes, RETINANET_STRIDES):
        n = h * w
        out["logits"].append(rng.normal(-6.0, 1.25, (n, num_classes)).astype(f32))
        out["ctrness"].append(rng.normal(0.0, 1.0, (n, 1)).astype(f32))
        out["offsets"].append((np.abs(rng.normal(0.0, 1.0, (n, 4))) * (s * 4.0)).astype(f32))
    out["im_info"] = np.array([hw[0], hw[1], hw[0], hw[1] - 11, 0], dtype=f32)
    return out


def fcos_batch(images, hw=FCOS_HW, num_classes=80):
    per = [fcos_image(i, hw, num_classes) for i in images]
    L = len(per[0]["logits"])
    return {k: [np.stack([p[k][l] for p in per]) for l in range(L)] for k in ("logits", "ctrness", "offsets")} | {
        "im_info": np.stack([p["im_info"] for p in per])}


def retinanet_image(img, hw=(800, 800), num_classes=80):
    """Config 1, one image: RetinaNet head outputs per level (n_l, C) / (n_l, 4) (seed 1 / seed 2 rule kept
    per image)."""
    rng = np.random.default_rng(1000 + img)
    sizes = retinanet_level_sizes(*hw)
    out = {"logits": [], "offsets": []}
    for h, w in sizes:
        n = h * w * 9
        out["logits"].append(logits_level(rng, n, num_classes))
        out["offsets"].append(deltas_level(rng, n))
    out["im_info"] = np.array([hw[0], hw[1], 612.0, 612.0, 0], dtype=f32)
    return out


def stress_image(img, n_boxes=100000, n_anchor=200000, n_gt=500, hw=STRESS_HW):
    """Config 5, one image: (200k x 500) IoU operands and 100k single-class NMS candidates with distinct scores;
    sizes log-U[8, 128] for the crowded boxes."""
    rng = np.random.default_rng(5000 + img)
    return {
        "anchors": make_gt(rng, n_anchor, hw[0], hw[1], 8, 128)[:, :4].copy(),
        "gt": make_gt(rng, n_gt, hw[0], hw[1]),
        "boxes": make_gt(rng, n_boxes, hw[0], hw[1], 8, 128)[:, :4].copy(),
        "scores": distinct_scores(rng, n_boxes),
    }
