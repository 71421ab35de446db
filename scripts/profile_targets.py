"""Short driver for ncu: two rounds of every target-assignment kernel at the config-2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import ops, workloads as W
from basedet_b200.layers import DefaultAnchorGenerator
dev = torch.device("cuda:0")
sizes = W.retinanet_level_sizes(800, 800)
gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
anchors = gen.generate_all_level_anchors(sizes, dev)
gt, ng = W.target_assign_batch(16)
gt_d, ng_d = torch.from_numpy(gt).to(dev), torch.from_numpy(ng).to(dev)
A = anchors.shape[0]
plan = ops.AssignPlan(A, 100, 16, dev)
iou = ops._padded_rows((16, 100), A, dev)[0]
for _ in range(3):
    ops.assign_targets(anchors, gt_d, ng_d, [0.4, 0.5], [0, -1, 1], True, True, plan=plan)
    ops.pairwise_batched(gt_d, ng_d, anchors, out=iou)
    ops.match(iou, [0.4, 0.5], [0, -1, 1], True, num_g=ng_d)
torch.cuda.synchronize()
