"""ROIAlign forward / backward timing at the config-3 size (8 192 ROIs x 256 ch) on SURVEY 8(d)'s synthetic ROIs
(sizes log-U[8, 600]) and on small ROIs; run once with BDET_ROI_TMA=0 and once with 1 to compare the two paths."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import ops, workloads as W
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B, Cn = int(os.environ.get("PERF_B", "16")), 256
NOFLUSH = os.environ.get("PERF_NOFLUSH") == "1"  # with PERF_B=1 the pyramid (91 MB) stays in L2: separates DRAM from SM limits
fs = [(-(-800 // s), -(-1344 // s)) for s in W.FRCNN_RCNN_STRIDES]
g = torch.Generator(device=dev); g.manual_seed(3)
feats = [torch.randn((B, Cn, h, w), device=dev, generator=g) for h, w in fs]
dfe = [torch.empty_like(f) for f in feats]
res = {"tma": os.environ.get("BDET_ROI_TMA", "1"), "B": B, "noflush": NOFLUSH}
for name, lo, hi in (("logU_8_600", 8, 600), ("small_8_64", 8, 64)):
    rois = torch.from_numpy(W.make_rois(np.random.default_rng(0), 8192 // B, B, 800, 1344, lo, hi)).to(dev)
    K = rois.shape[0]
    dout = torch.randn((K, Cn, 7, 7), device=dev, generator=g)
    lv = ops.roi_assign_levels(rois, 2, 5)
    sc = [1 / s for s in W.FRCNN_RCNN_STRIDES]
    shapes = [tuple(f.shape) for f in feats]
    perm = ops.roi_order(shapes, rois, lv, sc, (7, 7))
    for what, fn in (("fwd", lambda: ops.roi_align_fwd(feats, rois, lv, sc, (7, 7))),
                     ("bwd", lambda: ops.roi_align_bwd(dout, None, rois, lv, sc, (7, 7), dfeats=dfe)),
                     ("bwd_gather", lambda: ops.roi_align_bwd(dout, None, rois, lv, sc, (7, 7), dfeats=dfe, gather=True)),
                     ("order", lambda: ops.roi_order(shapes, rois, lv, sc, (7, 7))),
                     ("fwd_ordered", lambda: ops.roi_align_fwd(feats, rois, lv, sc, (7, 7), perm=perm)),
                     ("bwd_ordered", lambda: ops.roi_align_bwd(dout, None, rois, lv, sc, (7, 7), dfeats=dfe, perm=perm))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ops.profile_begin()
        ts = []
        for _ in range(8):
            if not NOFLUSH: flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        rep = ops.profile_report(); ops.profile_end()
        res["%s_%s" % (name, what)] = {"total_ms_median": float(np.median(ts)),
                                       "kernels": {k: v[0] / v[1] * 1e3 for k, v in rep.items()}}
        print(name, what, res["%s_%s" % (name, what)], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_roi_tma%s.json" % res["tma"], "w"), indent=1)
