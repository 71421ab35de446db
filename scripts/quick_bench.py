"""Scratch timing of the first kernels on the GPU box (not the contract bench)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import ops, _lib, workloads as W
from oracle import ref_ops as R

dev = torch.device("cuda:0")
def T(x): return torch.from_numpy(np.ascontiguousarray(x)).to(dev)
def timeit(fn, n=20, warm=5, flush=None):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if flush is not None: flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sizes = W.retinanet_level_sizes(800, 800)
base = [R.generate_base_anchors(s, W.RETINANET_RATIOS[0]) for s in W.RETINANET_SCALES]
anchors = torch.cat(ops.anchors_grid(sizes, W.RETINANET_STRIDES, [4, 8, 16, 32, 64], base, dev))
A = anchors.shape[0]
B = 16
gt, ng = W.target_assign_batch(B)
gt_d, ng_d = T(gt), T(ng)
print("A", A)
res = {}
res["anchors_grid_ms"] = timeit(lambda: ops.anchors_grid(sizes, W.RETINANET_STRIDES, [4, 8, 16, 32, 64], base, dev), flush=flush)
iou = ops._padded_rows((B, 100), A, dev)[0]
res["iou_b16_ms"] = timeit(lambda: ops.pairwise_batched(gt_d, ng_d, anchors, out=iou), flush=flush)
by = B * 100 * A * 4
print("iou GB/s (median)", by / res["iou_b16_ms"][0] / 1e6)
res["match_b16_ms"] = timeit(lambda: ops.match(iou, [0.4, 0.5], [0, -1, 1], True, num_g=ng_d), flush=flush)
print("match GB/s (median)", by / res["match_b16_ms"][0] / 1e6)
res["match_nolq_b16_ms"] = timeit(lambda: ops.match(iou, [0.4, 0.5], [0, -1, 1], False, num_g=ng_d), flush=flush)
idx, lab = ops.match(iou, [0.4, 0.5], [0, -1, 1], True, num_g=ng_d)
res["encode_gather_1img_ms"] = timeit(lambda: ops.box_encode(anchors, gt_d[0, :, :4], (0,0,0,0), (1,1,1,1), gather_idx=idx[0]), flush=flush)
plan = ops.AssignPlan(A, 100, B, dev)
res["assign_fused_b16_ms"] = timeit(lambda: ops.assign_targets(anchors, gt_d, ng_d, [0.4, 0.5], [0, -1, 1], True, True, plan=plan), flush=flush)
res["assign_fused_nolq_b16_ms"] = timeit(lambda: ops.assign_targets(anchors, gt_d, ng_d, [0.4, 0.5], [0, -1, 1], False, True, plan=plan), flush=flush)
print("fused img/s", B / res["assign_fused_b16_ms"][0] * 1e3, "bytes-based GB/s", B * A * 40 / res["assign_fused_b16_ms"][0] / 1e6)
# single image variants
iou1 = ops._padded_rows((1, 100), A, dev)[0]
res["iou_b1_ms"] = timeit(lambda: ops.pairwise_batched(gt_d[:1], ng_d[:1], anchors, out=iou1), flush=flush)
res["match_b1_ms"] = timeit(lambda: ops.match(iou1, [0.4, 0.5], [0, -1, 1], True, num_g=ng_d[:1]), flush=flush)
plan1 = ops.AssignPlan(A, 100, 1, dev)
res["assign_fused_b1_ms"] = timeit(lambda: ops.assign_targets(anchors, gt_d[:1], ng_d[:1], [0.4, 0.5], [0, -1, 1], True, True, plan=plan1), flush=flush)
# stress config 5: 500 x 200000
rng = np.random.default_rng(0)
a5 = T(W.make_gt(rng, 200000, 800, 1344, 8, 128)[:, :4]); g5 = T(W.make_gt(rng, 500, 800, 1344)[:, :4])
out5 = None
res["iou_500x200k_ms"] = timeit(lambda: ops.pairwise(g5, a5), flush=flush)
print("iou stress GB/s", 500 * 200000 * 4 / res["iou_500x200k_ms"][0] / 1e6)
# cpu oracle timing, 1 image
t0 = time.perf_counter(); io = R.box_iou(gt[0, :, :4], anchors.cpu().numpy()); t1 = time.perf_counter()
R.matcher(io, [0.4, 0.5], [0, -1, 1], True); t2 = time.perf_counter()
res["cpu_oracle_iou_s"], res["cpu_oracle_match_s"] = t1 - t0, t2 - t1
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/quick_bench.json", "w"), indent=1)
# per-kernel breakdown via the library's event hooks
ops.profile_begin()
for _ in range(20):
    flush.zero_()
    ops.assign_targets(anchors, gt_d, ng_d, [0.4, 0.5], [0, -1, 1], True, True, plan=plan)
    flush.zero_()
    ops.pairwise_batched(gt_d, ng_d, anchors, out=iou)
    flush.zero_()
    ops.match(iou, [0.4, 0.5], [0, -1, 1], True, num_g=ng_d)
torch.cuda.synchronize()
for k in ("assign_main_kernel", "assign_lq_kernel", "pairwise_kernel", "match_colmax_kernel", "match_lq_kernel"):
    ms, n = ops.profile_collect(k)
    print("kernel %-22s avg %.2f us over %d launches" % (k, ms / max(n, 1) * 1e3, n))
ops.profile_end()
