"""Per-kernel timing at the BASELINE config sizes (library event hooks, L2 flushed between iterations).
Writes gpurun_out/perf_all.json; `--ncu` runs each pipeline once (for ncu captures)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import ops, pipelines, _lib, workloads as W
from basedet_b200.layers import DefaultAnchorGenerator

NCU = "--ncu" in sys.argv
ONLY = [a for a in sys.argv[1:] if not a.startswith("--")]
dev = torch.device("cuda:0")
PEAK = 6549.8
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def T(x): return torch.from_numpy(np.ascontiguousarray(x)).to(dev)
results = {}

def run(name, fn, bytes_by_kernel=None, iters=10):
    if ONLY and name not in ONLY: return
    n_it = 1 if NCU else iters
    for _ in range(1 if NCU else 3): fn()
    torch.cuda.synchronize()
    ops.profile_begin()
    tot = []
    for _ in range(n_it):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); tot.append(s.elapsed_time(e))
    out = {"total_ms_median": float(np.median(tot)), "kernels": {}}
    names = set()

    for k in KERNELS:
        ms, n = ops.profile_collect(k)
        if n:
            d = {"avg_us": ms / n * 1e3, "launches_per_iter": n / n_it}
            if bytes_by_kernel and k in bytes_by_kernel:
                d["GBps"] = bytes_by_kernel[k] / (ms / n * 1e-3) / 1e9
                d["frac_of_measured_peak"] = d["GBps"] / PEAK
            out["kernels"][k] = d
    ops.profile_end()
    results[name] = out
    print(name, json.dumps(out))

KERNELS = ["anchors_grid_kernel", "pairwise_kernel", "match_colmax_kernel", "match_lq_kernel", "assign_main_kernel", "assign_lq_kernel",
           "box_encode_kernel", "box_decode_kernel", "score_filter_kernel", "select_sort_kernel", "select_decode_kernel",
           "nms_sort_small_kernel", "nms_tile_sort_kernel", "nms_global_step_kernel", "nms_tile_tail_kernel", "nms_gather_kernel",
           "nms_maxcoord_kernel", "nms_fused_kernel", "nms_chunk_kernel", "nms_sweep_kernel", "finalize_kernel", "roi_assign_levels_kernel",
           "roi_align_fwd_kernel", "roi_align_bwd_kernel", "roi_align_bwd_gather_kernel", "roi_bin_kernel", "roi_bin_scan_kernel",
           "fcos_targets_kernel", "atss_candidates_kernel", "atss_finish_kernel", "count_labels_kernel", "sample_labels_kernel",
           "rcnn_match_kernel", "rcnn_collect_kernel", "ota_rows_kernel", "ota_resolve_kernel"]
rng = np.random.default_rng(0)

# ---- config 2: target assignment, batch 16
sizes = W.retinanet_level_sizes(800, 800)
gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
anchors = gen.generate_all_level_anchors(sizes, dev)
A = anchors.shape[0]; B = 16; G = 100
gt, ng = W.target_assign_batch(B); gt_d, ng_d = T(gt), T(ng)
plan = ops.AssignPlan(A, G, B, dev)
iou = ops._padded_rows((B, G), A, dev)[0]
run("c2_fused", lambda: ops.assign_targets(anchors, gt_d, ng_d, [0.4, 0.5], [0, -1, 1], True, True, plan=plan),
    {"assign_main_kernel": A * 16 + B * (A * 24 + G * 20)})
run("c2_iou", lambda: ops.pairwise_batched(gt_d, ng_d, anchors, out=iou), {"pairwise_kernel": B * (4 * G * A + 16 * G) + 16 * A})
run("c2_match", lambda: ops.match(iou, [0.4, 0.5], [0, -1, 1], True, num_g=ng_d), {"match_colmax_kernel": B * (4 * G * A + 8 * A)})
idx_m, _ = ops.match(iou, [0.4, 0.5], [0, -1, 1], True, num_g=ng_d)
run("c2_encode", lambda: ops.box_encode(anchors, gt_d[0, :, :4], (0, 0, 0, 0), (1, 1, 1, 1), gather_idx=idx_m[0]), {"box_encode_kernel": A * 36})
run("anchors", lambda: gen.generate_all_level_anchors(sizes, dev), {"anchors_grid_kernel": A * 16})

# ---- config 1 / batched retinanet post-processing (B images 800x800)
def dense_inputs(Bd, hw, C=80):
    sz = W.retinanet_level_sizes(*hw)
    anc = gen.generate_anchors_by_features(sz, dev)
    g = torch.Generator(device=dev); g.manual_seed(1)
    lg = [torch.randn((Bd, h * w * 9, C), device=dev, generator=g) * 1.25 - 6.0 for h, w in sz]
    dl = [torch.randn((Bd, h * w * 9, 4), device=dev, generator=g) * 0.15 for h, w in sz]
    info = T(np.array([[hw[0], hw[1], 612.0, 612.0, 0.0]] * Bd, np.float32))
    nlog = sum(h * w * 9 * C for h, w in sz) * Bd
    return anc, lg, dl, info, nlog
for Bd in (1, 8):
    anc, lg, dl, info, nlog = dense_inputs(Bd, (800, 800))
    run("c1_retina_post_b%d" % Bd, lambda: pipelines.dense_postprocess(lg, dl, anc, info, 0.05, 0.5, 100, 1000), {"score_filter_kernel": nlog * 4})
    del lg, dl

# ---- config 4: FCOS batch 64 @ 800x1344 (22400 points)
sz4 = W.retinanet_level_sizes(800, 1344)
pts = ops.points_grid(sz4, W.RETINANET_STRIDES, [0.5 * s for s in W.RETINANET_STRIDES], 1, 0, dev)
g = torch.Generator(device=dev); g.manual_seed(2)
B4 = 64
lg4 = [torch.randn((B4, h * w, 80), device=dev, generator=g) * 1.25 - 6.0 for h, w in sz4]
ct4 = [torch.randn((B4, h * w, 1), device=dev, generator=g) for h, w in sz4]
lt4 = [torch.randn((B4, h * w, 4), device=dev, generator=g).abs() * s * 4 for (h, w), s in zip(sz4, W.RETINANET_STRIDES)]
info4 = T(np.array([[800, 1344, 800, 1333, 0.0]] * B4, np.float32))
run("c4_fcos_b64", lambda: pipelines.dense_postprocess(lg4, lt4, pts, info4, 0.05, 0.6, 100, 1000, ctrness_list=ct4),
    {"score_filter_kernel": B4 * 22400 * 80 * 4 + B4 * 22400 * 4})
del lg4, ct4, lt4
# config-4 shape, training side: FCOS / ATSS target assignment, 64 images x 22 400 points x 100 GT
gt4, ng4 = W.target_assign_batch(B4, 100, 800, 1344, seed0=4000)
gt4_d, ng4_d = T(gt4), T(ng4)
A4 = sum(h * w for h, w in sz4)
SOI = [(-1, 64), (64, 128), (128, 256), (256, 512), (512, float("inf"))]
dplan = ops.DensePlan(A4, B4, dev, atss=True)
dense_bytes = B4 * (A4 * (4 + 16 + 4 + 4) + 100 * 20) + A4 * 8   # labels + offsets + ctrness + match_idx written, gt + points read
run("c4_fcos_targets_b64", lambda: ops.fcos_targets(pts, gt4_d, ng4_d, W.RETINANET_STRIDES, SOI, 1.5, plan=dplan),
    {"fcos_targets_kernel": dense_bytes})
run("c4_atss_targets_b64", lambda: ops.atss_targets(pts, gt4_d, ng4_d, W.RETINANET_STRIDES, 8, 9, plan=dplan),
    {"atss_finish_kernel": dense_bytes + B4 * A4 * 8})

# OTA dynamic-k matching: one image, 100 GT x 22 400 points, cost / IoU matrices given
cost4 = torch.rand((100, A4), device=dev, generator=g) * 5
iou4 = torch.rand((100, A4), device=dev, generator=g) ** 3
run("c4_ota_match_1img", lambda: ops.ota_topk_match(cost4, iou4, 10), {"ota_rows_kernel": 3 * 100 * A4 * 4 + 100 * A4 * 4})

# ---- config 3: RPN proposals + ROIAlign fwd/bwd, batch 16 @ 800x1344
B3 = 16
sz3 = W.frcnn_level_sizes(800, 1344)
gen3 = DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
anc3 = gen3.generate_anchors_by_features(sz3, dev)
g = torch.Generator(device=dev); g.manual_seed(3)
sc3 = [torch.randn((B3, a.shape[0]), device=dev, generator=g) * 2 - 3 for a in anc3]
dl3 = [torch.randn((B3, a.shape[0], 4), device=dev, generator=g) * 0.2 for a in anc3]
info3 = T(np.array([[800, 1344, 800, 1333, 0.0]] * B3, np.float32))
# training-side glue of config 3: RPN targets with sampling, RCNN label assignment
gt3, ng3 = W.target_assign_batch(B3, 100, 800, 1344, seed0=3000)
gt3_d, ng3_d = T(gt3), T(ng3)
anc3_all = torch.cat(anc3)
A3 = anc3_all.shape[0]
nz1, nz2 = torch.rand((B3, A3), device=dev, generator=g), torch.rand((B3, A3), device=dev, generator=g)
plan3 = ops.AssignPlan(A3, 100, B3, dev)
run("c3_rpn_targets_b16", lambda: pipelines.rpn_targets(anc3_all, gt3_d, ng3_d, nz1, nz2, plan=plan3))
rois3 = T(np.stack([W.make_rois(np.random.default_rng(80 + b), 2000, 1, 800, 1344, 8, 600) for b in range(B3)]))
nr3 = torch.full((B3,), 2000, dtype=torch.int32, device=dev)
nz3, nz4 = torch.rand((B3, 2100), device=dev, generator=g), torch.rand((B3, 2100), device=dev, generator=g)
run("c3_rcnn_targets_b16", lambda: pipelines.rcnn_targets(rois3, nr3, gt3_d, ng3_d, nz3, nz4))
run("c3_rpn_train_b16", lambda: pipelines.rpn_proposals(sc3, dl3, anc3, info3, 2000, 1000, 0.7))
run("c3_rpn_test_b16", lambda: pipelines.rpn_proposals(sc3, dl3, anc3, info3, 1000, 1000, 0.7))
Cn, K = 256, 512 * B3
fs = [(-(-800 // s), -(-1344 // s)) for s in W.FRCNN_RCNN_STRIDES]
feats = [torch.randn((B3, Cn, h, w), device=dev, generator=g) for h, w in fs]
rois = T(W.make_rois(rng, 512, B3, 800, 1344, 8, 600))
dout = torch.randn((K, Cn, 7, 7), device=dev, generator=g)
pyr = sum(B3 * Cn * h * w * 4 for h, w in fs)
run("c3_roi_fwd", lambda: pipelines.roi_pool_forward_backward(feats, rois, W.FRCNN_RCNN_STRIDES, (7, 7)),
    {"roi_align_fwd_kernel": K * Cn * 49 * 4 + pyr}, iters=5)
dfe = [torch.empty_like(f) for f in feats]
lv = ops.roi_assign_levels(rois, 2, 5)
wsb = ops._workspace(_lib.load().bdet_roi_align_bwd_workspace(4, _lib.iarr([v for f in feats for v in f.shape[-2:]]), B3, K), dev)
run("c3_roi_bwd_gather", lambda: ops.roi_align_bwd(dout, None, rois, lv, [1 / s for s in W.FRCNN_RCNN_STRIDES], (7, 7), dfeats=dfe, workspace=wsb, gather=True),
    {"roi_align_bwd_gather_kernel": K * Cn * 49 * 4 + pyr}, iters=5)
run("c3_roi_bwd_scatter", lambda: ops.roi_align_bwd(dout, None, rois, lv, [1 / s for s in W.FRCNN_RCNN_STRIDES], (7, 7), dfeats=dfe, gather=False),
    {"roi_align_bwd_kernel": K * Cn * 49 * 4 + 2 * pyr}, iters=3)
del feats, dfe, dout

# ---- config 5: stress
a5 = T(W.make_gt(rng, 200000, 800, 1333, 8, 128)[:, :4]); g5 = T(W.make_gt(rng, 500, 800, 1333)[:, :4])
run("c5_iou_500x200k", lambda: ops.pairwise(g5, a5), {"pairwise_kernel": 500 * 200000 * 4 + 16 * 200500})
b5 = T(np.stack([W.make_gt(np.random.default_rng(50 + i), 100000, 800, 1333, 8, 128)[:, :4] for i in range(2)]))
s5 = T(np.stack([W.distinct_scores(np.random.default_rng(60 + i), 100000) for i in range(2)]))
ws5 = ops._workspace(_lib.load().bdet_nms_workspace(100000, 2), dev)
run("c5_nms_100k_b2", lambda: ops.nms_batched(b5, s5, None, 0.5, None, workspace=ws5), iters=3)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/perf_all.json", "w"), indent=1)
