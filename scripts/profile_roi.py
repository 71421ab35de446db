"""ncu driver: ROIAlign fwd, bwd (scatter), bwd (gather) at B=4 of config 3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import ops, workloads as W
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
B = 4
fs = [(-(-800 // s), -(-1344 // s)) for s in W.FRCNN_RCNN_STRIDES]
feats = [torch.randn((B, 256, h, w), device=dev, generator=g) for h, w in fs]
rois = torch.from_numpy(W.make_rois(np.random.default_rng(0), 512, B, 800, 1344, 8, 600)).to(dev)
dout = torch.randn((512 * B, 256, 7, 7), device=dev, generator=g)
lv = ops.roi_assign_levels(rois, 2, 5)
sc = [1 / s for s in W.FRCNN_RCNN_STRIDES]
dfe = [torch.empty_like(f) for f in feats]
for _ in range(2):
    ops.roi_align_fwd(feats, rois, lv, sc, (7, 7))
    ops.roi_align_bwd(dout, None, rois, lv, sc, (7, 7), dfeats=dfe, gather=False)
    ops.roi_align_bwd(dout, None, rois, lv, sc, (7, 7), dfeats=dfe, gather=True)
torch.cuda.synchronize()
