"""Generate tests/golden/reference_vectors.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES
(/root/reference/basedet/{structures,layers/common}/*.py) under the numpy megengine shim (oracle/mge_shim).

    python tests/golden/gen_golden.py            # only works where /root/reference exists (build container)

Inputs are seeded; outputs are what the reference code returns.  Model classes cannot be imported (they need
basecore), so the few glue lines of RetinaNet.get_ground_truth / inference are transcribed below with the
reference's own Boxes / Matcher / BoxCoder / batched_nms objects (cited line by line).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from basedet_b200 import workloads as W  # noqa: E402
from oracle import ref_runner  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")


def build():
    ref = ref_runner.load()
    T, F = ref.Tensor, ref.F
    Boxes = ref.boxes.Boxes
    g = {}
    rng = np.random.default_rng(20240)

    # ---- pairwise family (structures/boxes.py, op_patch.py)
    b1 = W.make_gt(rng, 13, 300, 400)[:, :4]
    b2 = W.make_gt(rng, 57, 300, 400, 8, 200)[:, :4]
    b2[0] = b1[0]                      # identical boxes
    b2[1] = [b1[1][2], b1[1][1], b1[1][2] + 9, b1[1][3]]   # touching
    b2[2] = [7, 7, 7, 7]               # zero area
    g["pair_b1"], g["pair_b2"] = b1, b2
    B1, B2 = Boxes(T(b1.copy())), Boxes(T(b2.copy()))
    g["pair_iou"] = B1.iou(B2).numpy()
    g["pair_ioa"] = B1.ioa(B2).numpy()
    g["pair_inter"] = B1.intersection(B2).numpy()
    g["pair_giou"] = B1.giou(B2).numpy()
    g["pair_centers"] = B2.centers.numpy()
    g["pair_area"], g["pair_width"], g["pair_height"] = B2.area.numpy(), B2.width.numpy(), B2.height.numpy()
    p1, p2 = rng.uniform(0, 300, (9, 2)).astype(np.float32), rng.uniform(0, 300, (21, 2)).astype(np.float32)
    g["pd_p1"], g["pd_p2"] = p1, p2
    g["pd_out"] = ref.op_patch.point_distance(T(p1), T(p2)).numpy()

    # ---- Boxes misc (boxes.py:132-212)
    raw = (W.make_gt(rng, 40, 300, 400)[:, :4] + rng.normal(0, 30, (40, 4))).astype(np.float32)
    g["misc_boxes"] = raw
    g["misc_clip"] = Boxes(T(raw.copy())).clip((250.0, 333.0)).numpy()
    g["misc_scale"] = Boxes(T(raw.copy())).scale((1.25, 0.75)).numpy()
    g["misc_filter"] = Boxes(T(raw.copy())).clip((250.0, 333.0)).filter_by_size().numpy()
    for mode in ("xyxy2xywh", "xywh2xyxy", "xyxy2xcycwh", "xcycwh2xyxy", "xywh2xcycwh", "xcycwh2xywh"):
        g["conv_" + mode] = ref.box_convert.BoxConverter.convert(T(b2.copy()), mode).numpy()

    # ---- anchors (layers/common/anchor_generator.py)
    sizes = W.retinanet_level_sizes(96, 128)
    g["anc_sizes"] = np.array(sizes)
    feats = [T(np.zeros((1, 1, h, w), np.float32)) for h, w in sizes]
    gen = ref.anchor_generator.DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
    for i, a in enumerate(gen(feats)):
        g["anc_retina_%d" % i] = a.numpy()
    fsizes = W.frcnn_level_sizes(96, 128)
    g["anc_fsizes"] = np.array(fsizes)
    gen2 = ref.anchor_generator.DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
    for i, a in enumerate(gen2([T(np.zeros((1, 1, h, w), np.float32)) for h, w in fsizes])):
        g["anc_rpn_%d" % i] = a.numpy()
    pg = ref.anchor_generator.AnchorPointGenerator(1, tuple(W.RETINANET_STRIDES), 0.5)
    for i, a in enumerate(pg(feats)):
        g["anc_points_%d" % i] = a.numpy()
    pg3 = ref.anchor_generator.AnchorPointGenerator(3, (8, 16), 0.0)
    for i, a in enumerate(pg3(feats[:2])):
        g["anc_points3_%d" % i] = a.numpy()
    fp = ref.anchor_generator.FastPointGenerator((8, 16, 32))
    for i, a in enumerate(fp(feats[:3])):
        g["anc_fast_%d" % i] = a.numpy()

    # ---- Matcher (layers/common/matcher.py) on random matrices with ties / degenerate rows
    m = rng.uniform(0, 1, (9, 401)).astype(np.float32)
    m[:, rng.integers(0, 401, 200)] = 0.0
    m[1, 3] = m[1].max()
    m[2, :] = 0.0
    m[0, 5] = m[4, 5] = 0.45
    g["match_m"] = m
    for tag, thr, labs, lq in (("retina", [0.4, 0.5], [0, -1, 1], True), ("rpn", [0.3, 0.7], [0, -1, 1], True),
                               ("nolq", [0.4, 0.5], [0, -1, 1], False), ("two", [0.5], [0, 1], True)):
        idx, lab = ref.matcher.Matcher(list(thr), list(labs), lq)(T(m.copy()))
        g["match_%s_idx" % tag], g["match_%s_lab" % tag] = idx.numpy(), lab.numpy()

    # ---- coders (structures/boxcoder.py)
    anchors = np.concatenate([g["anc_retina_%d" % i] for i in range(5)])
    n = anchors.shape[0]
    gtb = W.make_gt(rng, n, 96, 128, 4, 120)[:, :4]
    deltas = W.deltas_level(rng, n)
    g["coder_anchors"], g["coder_gt"], g["coder_deltas"] = anchors, gtb, deltas
    for tag, mean, std in (("unit", (0., 0., 0., 0.), (1., 1., 1., 1.)), ("rcnn", (0., 0., 0., 0.), (.1, .1, .2, .2)),
                           ("odd", (0.1, -0.1, 0.05, 0.0), (0.5, 0.25, 2.0, 1.0))):
        bc = ref.boxcoder.BoxCoder(mean, std)
        g["coder_enc_" + tag] = bc.encode(T(anchors.copy()), T(gtb.copy())).numpy()
        d = T(deltas.copy())
        g["coder_dec_" + tag] = bc.decode(T(anchors.copy()), d).numpy()
        g["coder_dec_inplace_" + tag] = d.numpy()  # boxcoder.py:76-77 mutates the caller's tensor
        sc = ref.boxcoder.SumBoxCoder(mean, std)
        g["coder_sumenc_" + tag] = sc.encode(T(anchors.copy()), T(gtb.copy())).numpy()
        g["coder_sumdec_" + tag] = sc.decode(T(anchors.copy()), T(deltas.copy())).numpy()
    pts = np.concatenate([g["anc_points_%d" % i] for i in range(5)])
    gsmall = W.make_gt(rng, 6, 96, 128)[:, :4]
    g["pc_pts"], g["pc_gt"] = pts, gsmall
    pc = ref.boxcoder.PointCoder()
    g["pc_enc"] = pc.encode(T(pts.copy()), F.expand_dims(T(gsmall.copy()), axis=1)).numpy()   # fcos.py:231
    ridx = rng.integers(0, 6, pts.shape[0])
    g["pc_ridx"] = ridx.astype(np.int32)
    g["pc_enc_rows"] = pc.encode(T(pts.copy()), T(gsmall[ridx].copy())).numpy()              # fcos.py:268
    ltrb = np.abs(rng.normal(0, 20, (pts.shape[0], 4))).astype(np.float32)
    g["pc_deltas"] = ltrb
    g["pc_dec"] = pc.decode(T(pts.copy()), T(ltrb.copy())).numpy()

    # ---- RetinaNet.get_ground_truth (models/det/retinanet.py:211-232), transcribed with reference objects
    gt5, ng = W.target_assign_batch(3, num_gt=7, img_h=96, img_w=128, ragged=True)
    g["gt_boxes"], g["gt_num"] = gt5, ng
    matcher = ref.matcher.Matcher([0.4, 0.5], [0, -1, 1], True)
    box_coder = ref.boxcoder.BoxCoder((0., 0., 0., 0.), (1., 1., 1., 1.))
    labs, offs, idxs = [], [], []
    tanchors = T(anchors.copy())
    for b in range(3):
        gt_boxes = T(gt5[b].copy())[: int(ng[b])]                       # :216
        overlaps = Boxes(gt_boxes[:, :4]).iou(Boxes(tanchors))          # :218
        match_indices, labels = matcher(overlaps)                       # :219
        gt_boxes_matched = gt_boxes[match_indices]                      # :220
        fg_mask = labels == 1                                           # :222
        labels[fg_mask] = gt_boxes_matched[fg_mask, 4].astype("int32")  # :223
        offsets = box_coder.encode(tanchors, gt_boxes_matched[:, :4])   # :224
        labs.append(labels.numpy()); offs.append(offsets.numpy()); idxs.append(match_indices.numpy())
    g["gt_labels"], g["gt_offsets"], g["gt_match_idx"] = np.stack(labs), np.stack(offs), np.stack(idxs)

    # ---- batched_nms / post_processing (layers/common/post_processing.py)
    nb = 300
    centers = W.make_gt(rng, 40, 300, 400, 8, 150)[:, :4]
    boxes = (centers[rng.integers(0, 40, nb)] + rng.normal(0, 3, (nb, 4))).astype(np.float32)
    boxes[:, 2:] = np.maximum(boxes[:, 2:], boxes[:, :2] + 1)
    scores = W.distinct_scores(rng, nb, 0.05, 1.0)
    labels = rng.integers(0, 7, nb).astype(np.int32)
    g["nms_boxes"], g["nms_scores"], g["nms_labels"] = boxes, scores, labels
    g["nms_keep_05"] = ref.post_processing.batched_nms(T(boxes.copy()), T(scores.copy()), T(labels.copy()), 0.5).numpy()
    g["nms_keep_06_max50"] = ref.post_processing.batched_nms(T(boxes.copy()), T(scores.copy()), T(labels.copy()), 0.6, 50).numpy()
    g["nms_keep_float_levels"] = ref.post_processing.batched_nms(T(boxes.copy()), T(scores.copy()),
                                                                 T(labels.astype(np.float32)), 0.7, 100).numpy()
    tied = np.round(scores, 1)
    g["nms_scores_tied"] = tied
    g["nms_keep_tied"] = ref.post_processing.batched_nms(T(boxes.copy()), T(tied.copy()), T(labels.copy()), 0.5).numpy()
    img_info = np.array([[300.0, 400.0, 480.0, 640.0, 0.0]], np.float32)
    g["pp_img_info"] = img_info
    cont = ref.structures.Container(boxes=Boxes(T(boxes.copy())), box_scores=T(scores.copy()), box_labels=T(labels.copy()))
    res = ref.post_processing.post_processing(cont, T(img_info), 0.5, max_detections_per_image=30)
    g["pp_boxes"], g["pp_scores"], g["pp_labels"] = res.boxes.numpy(), res.box_scores.numpy(), res.box_labels.numpy()

    # ---- score filter + top-k glue (models/det/retinanet.py:181-196), one level, given scores
    lg = W.logits_level(rng, 600, 20, mean=-4.0)
    of = W.deltas_level(rng, 600)
    an = anchors[:600]
    g["lvl_logits"], g["lvl_offsets"], g["lvl_anchors"] = lg, of, an
    logits = T(lg.copy())
    num_classes = 20
    scores_t = F.sigmoid(F.flatten(logits))                                       # :183
    g["lvl_scores"] = scores_t.numpy()
    _, keep_idx = ref.function.non_zeros(scores_t > 0.05)                          # :186
    topk_num = min(keep_idx.shape[0], 100)                                        # :189 (1000 in the reference)
    _, topk_idx = F.topk(scores_t[keep_idx], k=topk_num, descending=True)         # :190
    keep_idx = keep_idx[topk_idx]                                                 # :191
    g["lvl_keep_idx"] = keep_idx.numpy()
    g["lvl_keep_scores"] = scores_t[keep_idx].numpy()                             # :193
    g["lvl_labels"] = (keep_idx % num_classes).numpy()                            # :194
    dboxes = ref.boxcoder.BoxCoder().decode(T(an.copy()), T(of.copy()).reshape(-1, 4))   # :195
    g["lvl_boxes"] = dboxes[keep_idx // num_classes].numpy()                      # :196

    # ---- roi_pool (layers/common/roi_pool.py)
    C = 3
    fshapes = [(2, C, 24, 32), (2, C, 12, 16), (2, C, 6, 8), (2, C, 3, 4)]
    rfeats = [rng.normal(0, 1, s).astype(np.float32) for s in fshapes]
    rois = W.make_rois(rng, 9, 2, 96, 128, 6, 140)
    rois[0, 1:] = [-10, -8, 30, 40]
    rois[1, 1:] = [100, 70, 140, 110]
    for i, f in enumerate(rfeats):
        g["roi_feat_%d" % i] = f
    g["roi_rois"] = rois
    g["roi_out"] = ref.roi_pool.roi_pool([T(f.copy()) for f in rfeats], T(rois.copy()), [4, 8, 16, 32], (7, 7), "roi_align").numpy()
    rr, lv = ref.roi_pool.assign_rois(T(rois.copy()), [4, 8, 16, 32])
    g["roi_levels"] = lv.numpy()[: rois.shape[0]]

    # ---- FCOS.get_ground_truth (models/det/fcos.py:222-293) and ATSS.get_ground_truth (models/det/atss.py:17-86):
    # the methods themselves, AST-extracted from the reference files and run with a stand-in `self` that carries only
    # the attributes they read (cfg values from configs/det_model/{fcos,atss}_cfg.py)
    import types

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    dsizes = W.retinanet_level_sizes(160, 224)
    g["dense_sizes"] = np.array(dsizes)
    dfeats = [T(np.zeros((1, 1, h, w), np.float32)) for h, w in dsizes]
    dpts = ref.anchor_generator.AnchorPointGenerator(1, tuple(W.RETINANET_STRIDES), 0.5)(dfeats)
    for i, a in enumerate(dpts):
        g["dense_points_%d" % i] = a.numpy()
    dgt, dng = W.target_assign_batch(3, num_gt=14, img_h=160, img_w=224, seed0=900, ragged=True)
    g["dense_gt"], g["dense_num"] = dgt, dng
    soi = [[-1, 64], [64, 128], [128, 256], [256, 512], [512, float("inf")]]
    fcos_gt = ref_runner.load_method("models/det/fcos.py", "FCOS", "get_ground_truth")
    atss_gt = ref_runner.load_method("models/det/atss.py", "ATSS", "get_ground_truth")
    head = types.SimpleNamespace(strides=list(W.RETINANET_STRIDES))
    for tag, radius in (("fcos_r15", 1.5), ("fcos_r0", 0)):
        me = types.SimpleNamespace(cfg=Cfg(MODEL=Cfg(HEAD=Cfg(CENTER_SAMPLING_RADIUS=radius, OBJECT_SIZES_OF_INTEREST=soi))),
                                   head=head, box_coder=ref.boxcoder.PointCoder())
        lab, off, ctr = fcos_gt(me, [T(p.numpy().copy()) for p in dpts], T(dgt.copy()), [int(n) for n in dng])
        g[tag + "_labels"], g[tag + "_offsets"], g[tag + "_ctrness"] = lab.numpy(), off.numpy(), ctr.numpy()
    me = types.SimpleNamespace(cfg=Cfg(MODEL=Cfg(ANCHOR=Cfg(SCALE=8, TOPK=9))), head=head, box_coder=ref.boxcoder.PointCoder())
    lab, off, ctr = atss_gt(me, [T(p.numpy().copy()) for p in dpts], T(dgt.copy()), [int(n) for n in dng])
    g["atss_labels"], g["atss_offsets"], g["atss_ctrness"] = lab.numpy(), off.numpy(), ctr.numpy()

    # ---- RPN.get_ground_truth (models/det/rpn.py:215-240) with the reference's own sample_labels
    # (layers/common/sampling.py).  The variates are explicit: noise_* (B, A) hold one uniform value per anchor and the
    # shim's megengine.random.uniform is fed noise[mask] in the order the reference draws (oracle RNG contract).
    import importlib

    import megengine.random as mrand

    from oracle import ref_ops as R

    sampling = importlib.import_module("basedet.layers.common.sampling")
    rpn_gt = ref_runner.load_method("models/det/rpn.py", "RPN", "get_ground_truth", {"sample_labels": sampling.sample_labels})
    ssizes = W.frcnn_level_sizes(128, 160)
    g["samp_sizes"] = np.array(ssizes)
    sgen = ref.anchor_generator.DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
    sanchors = [a.numpy() for a in sgen([T(np.zeros((1, 1, h, w), np.float32)) for h, w in ssizes])]
    sall = np.concatenate(sanchors)
    sgt, sng = W.target_assign_batch(3, num_gt=8, img_h=128, img_w=160, seed0=1300, ragged=True)
    srng = np.random.default_rng(1301)
    noise_p = srng.uniform(0, 1, (3, len(sall))).astype(np.float32)
    noise_n = (np.floor(srng.uniform(0, 1, (3, len(sall))) * 512) / 512).astype(np.float32)   # ties among the variates
    g["samp_gt"], g["samp_num"], g["samp_noise_pos"], g["samp_noise_neg"] = sgt, sng, noise_p, noise_n
    npos_cfg, ntot = 6, 48
    for b in range(3):  # which variates the reference will draw, image by image (the oracle is only the bookkeeper here)
        gb = sgt[b, : sng[b]]
        _, lab0 = R.matcher(R.box_iou(gb[:, :4], sall), [0.3, 0.7], [0, -1, 1], True)
        if (lab0 == 1).sum() > npos_cfg:
            mrand.feed(noise_p[b][lab0 == 1])
        lab1 = R.sample_labels(lab0, npos_cfg, 1, -1, noise_p[b])
        if (lab1 == 0).sum() > ntot - (lab1 == 1).sum():
            mrand.feed(noise_n[b][lab1 == 0])
    me = types.SimpleNamespace(matcher=ref.matcher.Matcher([0.3, 0.7], [0, -1, 1], True),
                               box_coder=ref.boxcoder.BoxCoder((0., 0., 0., 0.), (1., 1., 1., 1.)),
                               num_pos_anchor=npos_cfg, num_sample_anchors=ntot)
    rl, ro = rpn_gt(me, [T(a.copy()) for a in sanchors], T(sgt.copy()), [int(n) for n in sng])
    assert not mrand._queue, "the reference drew fewer variates than were fed"
    g["samp_labels"], g["samp_offsets"] = rl.numpy().reshape(3, -1), ro.numpy().reshape(3, -1, 4)

    # ---- RCNN.get_ground_truth, training branch (layers/head/rcnn.py:95-147), same explicit-variate replay
    rcnn_gt = ref_runner.load_method("layers/head/rcnn.py", "RCNN", "get_ground_truth", {"sample_labels": sampling.sample_labels})
    rrng = np.random.default_rng(1400)
    RB, RG, Rmax = 3, 7, 90
    rgt, rng_ = W.target_assign_batch(RB, num_gt=RG, img_h=200, img_w=300, seed0=1401, ragged=True)
    rois_pad = np.zeros((RB, Rmax, 5), np.float32)
    rcount = np.array([60, 90, 75], np.int32)
    for b in range(RB):
        r = W.make_rois(rrng, int(rcount[b]), 1, 200, 300, 8, 150)
        r[:, 0] = b
        j = rrng.integers(0, rng_[b], 25)
        r[:25, 1:] = rgt[b, j, :4] + rrng.normal(0, 4, (25, 4)).astype(np.float32)   # near-GT proposals -> foreground exists
        rois_pad[b, : rcount[b]] = r
    N = Rmax + RG
    nfg = rrng.uniform(0, 1, (RB, N)).astype(np.float32)
    nbg = (np.floor(rrng.uniform(0, 1, (RB, N)) * 64) / 64).astype(np.float32)       # ties among the variates
    g["rcnn_rois"], g["rcnn_nrois"], g["rcnn_gt"], g["rcnn_num"] = rois_pad, rcount, rgt, rng_
    g["rcnn_noise_fg"], g["rcnn_noise_bg"] = nfg, nbg
    NUM, RATIO = 32, 0.25
    for b in range(RB):  # feed the variates in the reference's draw order
        g5 = rgt[b, : rng_[b]]
        allr = np.concatenate([rois_pad[b, : rcount[b]], np.concatenate([np.full((rng_[b], 1), b, np.float32), g5[:, :4]], 1)])
        ov = R.box_iou(allr[:, 1:], g5[:, :4])
        mx, lab = ov.max(1), g5[ov.argmax(1), 4]
        fg, bg = (mx >= 0.5) & (lab >= 0), (mx >= 0.0) & (mx < 0.5)
        n_b = len(allr)
        if fg.sum() > int(NUM * RATIO):
            mrand.feed(nfg[b, :n_b][fg])
        fgs = R.sample_labels(fg, int(NUM * RATIO), True, False, nfg[b, :n_b])
        if bg.sum() > NUM - fgs.sum():
            mrand.feed(nbg[b, :n_b][bg])
    im_info = np.zeros((RB, 5), np.float32)
    im_info[:, 4] = rng_
    me = types.SimpleNamespace(training=True, num_rois=NUM, fg_ratio=RATIO, fg_thresh=0.5, bg_thresh_high=0.5, bg_thresh_low=0.0,
                               box_coder=ref.boxcoder.BoxCoder([0., 0., 0., 0.], [0.1, 0.1, 0.2, 0.2]))
    flat = np.concatenate([rois_pad[b, : rcount[b]] for b in range(RB)])
    rr, rl, rt = rcnn_gt(me, T(flat.copy()), T(im_info), T(rgt.copy()))
    assert not mrand._queue, "the reference drew fewer variates than were fed"
    g["rcnn_out_rois"], g["rcnn_out_labels"], g["rcnn_out_targets"] = rr.numpy(), rl.numpy(), rt.numpy()

    # ---- OTATopkMatcher (layers/common/matcher.py:129-161): the reference class itself
    orng = np.random.default_rng(1500)
    for tag, (G, A, q) in (("ota_a", (9, 700, 0)), ("ota_b", (17, 3000, 32))):
        ious = (orng.uniform(0, 1, (G, A)) ** 3).astype(np.float32)
        cost = orng.uniform(0, 5, (G, A)).astype(np.float32)
        if q:  # quantised: equal costs / IoUs everywhere, anchors claimed by several GTs
            ious, cost = (np.floor(ious * q) / q).astype(np.float32), (np.floor(cost * q) / q).astype(np.float32)
        cost[:, ::7] += 1e6
        g[tag + "_cost"], g[tag + "_ious"] = cost, ious
        g[tag + "_match"] = ref.matcher.OTATopkMatcher(10)(T(cost.copy()), T(ious.copy())).numpy()
    return g


if __name__ == "__main__":
    data = build()
    np.savez_compressed(OUT, **data)
    print("wrote %s: %d arrays, %.1f KB" % (OUT, len(data), os.path.getsize(OUT) / 1024))
