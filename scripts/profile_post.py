"""Short driver for ncu: one retinanet post-processing batch (B=2, 800x800), one RPN batch (B=2), ROIAlign fwd/bwd (B=2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import ops, pipelines, workloads as W
from basedet_b200.layers import DefaultAnchorGenerator
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
sz = W.retinanet_level_sizes(800, 800)
anc = gen.generate_anchors_by_features(sz, dev)
B = 2
lg = [torch.randn((B, h * w * 9, 80), device=dev, generator=g) * 1.25 - 6.0 for h, w in sz]
dl = [torch.randn((B, h * w * 9, 4), device=dev, generator=g) * 0.15 for h, w in sz]
info = torch.tensor([[800, 800, 612.0, 612.0, 0.0]] * B, device=dev)
sz3 = W.frcnn_level_sizes(800, 1344)
gen3 = DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
anc3 = gen3.generate_anchors_by_features(sz3, dev)
sc3 = [torch.randn((B, a.shape[0]), device=dev, generator=g) * 2 - 3 for a in anc3]
dl3 = [torch.randn((B, a.shape[0], 4), device=dev, generator=g) * 0.2 for a in anc3]
info3 = torch.tensor([[800, 1344, 800, 1333, 0.0]] * B, device=dev)
fs = [(-(-800 // s), -(-1344 // s)) for s in W.FRCNN_RCNN_STRIDES]
feats = [torch.randn((B, 256, h, w), device=dev, generator=g) for h, w in fs]
rois = torch.from_numpy(W.make_rois(np.random.default_rng(0), 512, B, 800, 1344, 8, 600)).to(dev)
dout = torch.randn((512 * B, 256, 7, 7), device=dev, generator=g)
for _ in range(2):
    pipelines.dense_postprocess(lg, dl, anc, info, 0.05, 0.5, 100, 1000)
    pipelines.rpn_proposals(sc3, dl3, anc3, info3, 2000, 1000, 0.7)
    pipelines.roi_pool_forward_backward(feats, rois, W.FRCNN_RCNN_STRIDES, (7, 7), dout)
torch.cuda.synchronize()
