"""Generate tests/golden/reference_vectors_f3f4.npz for the widened rows 8(f)-3 / 8(f)-4 by EXECUTING THE REFERENCE'S OWN
STATEMENTS under the numpy megengine shim (oracle/mge_shim):

  * FreeAnchor: the per-image body of FreeAnchor.get_losses (models/det/free_anchor.py:48-113: box-probability scatter,
    bag construction, bag targets) is AST-extracted statement by statement -- everything up to the loss terms -- and run
    with a stand-in `self` carrying configs/det_model/freeanchor_cfg.py's values;
  * OTA: OTA.get_ground_truth (models/det/ota.py:76-180) is AST-extracted as a method, with the reference's own
    layers/losses/{sigmoid_focal_loss,iou_loss,cross_entropy}.py and layers/common/matcher.py (OTATopkMatcher);
  * COCO formatting: COCOEvaluator.format (evaluators/coco_eval.py:111-138), AST-extracted, with a stand-in dataset class.

    python tests/golden/gen_golden_f3f4.py          # only where /root/reference exists (build container)
"""
import ast
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from basedet_b200 import workloads as W  # noqa: E402
from oracle import ref_ops as R  # noqa: E402
from oracle import ref_runner  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors_f3f4.npz")
REF = ref_runner.REFERENCE


def free_anchor_fragment():
    """The statements of the per-image loop body of FreeAnchor.get_losses up to (and including) `matched_offsets = ...`."""
    path = os.path.join(REF, "basedet", "models", "det", "free_anchor.py")
    tree = ast.parse(open(path).read(), path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "FreeAnchor")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "get_losses")
    loop = next(n for n in fn.body if isinstance(n, ast.For))
    body = []
    for st in loop.body:
        body.append(st)
        if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", "") == "matched_offsets":
            break
    else:
        raise RuntimeError("free_anchor.py changed: matched_offsets assignment not found")
    return compile(ast.Module(body=body, type_ignores=[]), path, "exec")


def build():
    ref = ref_runner.load()
    T, F = ref.Tensor, ref.F
    g = {}
    rng = np.random.default_rng(777)

    # ------------------------------------------------------------------ FreeAnchor (free_anchor.py:48-113)
    code = free_anchor_fragment()
    hw, C, G = (128, 160), 8, 6
    sizes = W.retinanet_level_sizes(*hw)
    anchors = np.concatenate(R.default_anchors(sizes, W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5))
    A = anchors.shape[0]
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(BUCKET=types.SimpleNamespace(BOX_IOU_THRESH=0.6, BUCKET_SIZE=50)))
    coder = ref.boxcoder.BoxCoder((0.0, 0.0, 0.0, 0.0), (0.1, 0.1, 0.2, 0.2))
    self_ = types.SimpleNamespace(cfg=cfg, box_coder=coder)
    for case in ("normal", "fill"):
        gt = W.make_gt(rng, G, hw[0], hw[1], 16, 100)
        gt[:, 4] = rng.integers(1, C + 1, G)
        gt[1, 4] = gt[0, 4]                                   # two GT of one class: scatter collisions
        gt[1, :4] = gt[0, :4] + np.float32(3.0)
        idx, _ = R.matcher(R.box_iou(gt[:, :4], anchors), [0.4, 0.5], [0, -1, 1], True)
        # predictions: the encoded match plus noise, so that a few decoded boxes overlap their GT by > 0.6
        enc = R.boxcoder_encode(anchors, gt[idx][:, :4], (0, 0, 0, 0), (0.1, 0.1, 0.2, 0.2))
        noise = 0.8 if case == "normal" else 40.0             # "fill": no decoded box reaches IoU 0.6
        offs = (enc + rng.normal(0, noise, enc.shape)).astype(np.float32)
        logits = rng.normal(-2, 2, (1, A, C)).astype(np.float32)
        info = np.array([hw[0], hw[1], hw[0], hw[1], G], np.float32)
        ns = dict(self=self_, F=F, Boxes=ref.boxes.Boxes, anchors=T(anchors.copy()), pred_offsets=T(offs[None].copy()),
                  pred_logits=T(logits.copy()), pred_scores=F.sigmoid(T(logits.copy())), clamp_eps=1e-7, bucket_size=50,
                  box_prob_list=[], batch_id=0, gt_boxes_per_image=T(gt.copy()), info_per_image=T(info))
        exec(code, ns)
        p = "fa_%s_" % case
        g[p + "anchors"], g[p + "gt"], g[p + "offsets"], g[p + "logits"] = anchors, gt, offs, logits[0]
        g[p + "box_prob"] = ns["box_prob_list"][0].numpy()
        g[p + "fill"] = np.array(bool(ns["fill_prob"]))
        g[p + "matched_idx"] = ns["matched_idx"].numpy()
        g[p + "matched_score"] = ns["matched_score"].numpy()
        g[p + "matched_offsets"] = ns["matched_offsets"].numpy()
        print(case, "box_prob nonzeros", int((g[p + "box_prob"] != 0).sum()), "fill", bool(ns["fill_prob"]))

    # ------------------------------------------------------------------ OTA.get_ground_truth (ota.py:76-180)
    import megengine  # the shim (ref_runner.load() has put it into sys.modules)

    losses_dir = os.path.join(REF, "basedet", "layers", "losses")
    sys.modules["basedet.layers.losses"].__path__ = [losses_dir]
    focal = importlib.import_module("basedet.layers.losses.sigmoid_focal_loss").sigmoid_focal_loss
    iou_loss = importlib.import_module("basedet.layers.losses.iou_loss").iou_loss
    layers_ns = types.SimpleNamespace(sigmoid_focal_loss=focal, iou_loss=iou_loss)
    get_gt = ref_runner.load_method("models/det/ota.py", "OTA", "get_ground_truth", {"layers": layers_ns})
    hw, C, G, B = (96, 128), 6, 5, 2
    strides = [8, 16, 32, 64, 128]
    sizes = W.retinanet_level_sizes(*hw)
    shifts = R.anchor_points(sizes, 1, strides, 0.5)
    A = sum(p.shape[0] for p in shifts)
    gt = np.zeros((B, G, 5), np.float32)
    ng = np.array([G, G - 2], np.int32)
    for b in range(B):
        gt[b, : ng[b]] = W.make_gt(rng, int(ng[b]), hw[0], hw[1], 16, 90)
        gt[b, : ng[b], 4] = rng.integers(1, C + 1, int(ng[b]))
    box_cls = [rng.normal(-2, 1.5, (B, p.shape[0], C)).astype(np.float32) for p in shifts]
    box_delta = [np.abs(rng.normal(0, 1, (B, p.shape[0], 4)) * s * 1.5).astype(np.float32) for p, s in zip(shifts, strides)]
    box_iou = [rng.normal(0, 1, (B, p.shape[0], 1)).astype(np.float32) for p in shifts]
    info = np.stack([np.array([hw[0], hw[1], hw[0], hw[1], n], np.float32) for n in ng])
    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(LOSSES=types.SimpleNamespace(FOCAL_LOSS_ALPHA=0.25, FOCAL_LOSS_GAMMA=2)))
    self_ = types.SimpleNamespace(cfg=cfg, box_coder=ref.boxcoder.PointCoder(), reg_weight=1.5, matching="topk",
                                  head=types.SimpleNamespace(strides=strides, num_classes=C),
                                  matcher=ref.matcher.OTATopkMatcher(candidate_k=10))
    cls_t, delta_t, iou_t = get_gt(self_, [T(p.copy()) for p in shifts], [T(x.copy()) for x in box_cls],
                                   [T(x.copy()) for x in box_delta], [T(x.copy()) for x in box_iou],
                                   {"img_info": T(info), "gt_boxes": T(gt.copy())})
    g["ota_gt"], g["ota_num_gt"], g["ota_hw"] = gt, ng, np.array(hw)
    for l in range(len(shifts)):
        g["ota_cls_%d" % l], g["ota_delta_%d" % l] = box_cls[l], box_delta[l]
    g["ota_gt_classes"] = cls_t.numpy().reshape(B, A)
    g["ota_gt_deltas"] = delta_t.numpy().reshape(B, A, 4)
    g["ota_gt_ious"] = iou_t.numpy().reshape(B, A)
    print("ota foreground per image", (g["ota_gt_classes"] > 0).sum(axis=1))

    # ------------------------------------------------------------------ COCOEvaluator.format (coco_eval.py:111-138)
    path = os.path.join(REF, "basedet", "evaluators", "coco_eval.py")
    tree = ast.parse(open(path).read(), path)
    cls_node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "COCOEvaluator")
    fn = next(n for n in cls_node.body if isinstance(n, ast.FunctionDef) and n.name == "format")
    fn.decorator_list = []
    origin = {"class_%d" % i: 3 * i + 1 for i in range(C)}            # a COCO-like non-contiguous category id table
    dataset_class = types.SimpleNamespace(class_names=["class_%d" % i for i in range(C)], classes_originID=origin)
    registers = types.SimpleNamespace(datasets_info={"COCO": {"dataset_type": "X"}}, datasets={"X": dataset_class})
    glb = {"registers": registers, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), glb)
    dets = np.zeros((3, 7, 6), np.float32)
    cnt = np.array([7, 0, 4], np.int32)
    for b in range(3):
        bx = W.make_gt(rng, 7, 300, 400)
        dets[b, :, :4], dets[b, :, 4], dets[b, :, 5] = bx[:, :4], rng.uniform(0.05, 1, 7), rng.integers(0, C, 7)
        dets[b, cnt[b]:] = 0
    image_ids = np.array([139, 285, 632], np.int32)
    results = [{"image_id": int(i), "det_res": np.array(d[:n], dtype=np.float64)} for i, d, n in zip(image_ids, dets, cnt)]
    cfg2 = types.SimpleNamespace(DATA=types.SimpleNamespace(TEST=types.SimpleNamespace(name="coco_2017_val")))
    recs = glb["format"](results, cfg2)
    g["coco_dets"], g["coco_cnt"], g["coco_image_ids"] = dets, cnt, image_ids
    g["coco_origin"] = np.array([origin["class_%d" % i] for i in range(C)], np.int32)
    g["coco_rec_image"] = np.array([r["image_id"] for r in recs], np.int32)
    g["coco_rec_bbox"] = np.array([r["bbox"] for r in recs], np.float64)
    g["coco_rec_score"] = np.array([r["score"] for r in recs], np.float64)
    g["coco_rec_cat"] = np.array([r["category_id"] for r in recs], np.int32)
    for r in results:
        r["det_res"] = np.array(dets[list(image_ids).index(r["image_id"])][: cnt[list(image_ids).index(r["image_id"])]], dtype=np.float64)
    registers.datasets["X"] = types.SimpleNamespace(class_names=dataset_class.class_names)    # no classes_originID: label + 1
    g["coco_rec_cat_plus1"] = np.array([r["category_id"] for r in glb["format"](results, cfg2)], np.int32)
    return g


if __name__ == "__main__":
    out = build()
    np.savez_compressed(OUT, **out)
    print("wrote %s: %d arrays" % (OUT, len(out)))
