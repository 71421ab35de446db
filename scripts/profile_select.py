"""Short driver for ncu: the RPN top-k (16 images x 5 levels, k=2000) through bdet_topk only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basedet_b200 import ops, workloads as W
from basedet_b200.layers import DefaultAnchorGenerator
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
B = int(os.environ.get("B", "16"))
sz3 = W.frcnn_level_sizes(800, 1344)
gen3 = DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
anc3 = gen3.generate_anchors_by_features(sz3, dev)
sc3 = [torch.randn((B, a.shape[0]), device=dev, generator=g) * 2 - 3 for a in anc3]
base, starts, lens = ops._segments(sc3)
for _ in range(3):
    out = ops.topk_raw(base, starts, lens, 2000)
torch.cuda.synchronize()
