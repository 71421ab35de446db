"""Eager vs CUDA-graph replay of the pipelines (events around the call, L2 flushed between iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import pipelines, workloads as W
from basedet_b200.layers import DefaultAnchorGenerator
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=30):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return round(float(np.median(ts[3:])), 4)
gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
sz = W.retinanet_level_sizes(800, 800)
anc = gen.generate_anchors_by_features(sz, dev)
for B in (1, 8):
    lg = [torch.randn((B, h * w * 9, 80), device=dev, generator=g) * 1.25 - 6.0 for h, w in sz]
    dl = [torch.randn((B, h * w * 9, 4), device=dev, generator=g) * 0.15 for h, w in sz]
    info = torch.tensor([[800, 800, 612.0, 612.0, 0.0]] * B, device=dev)
    args = (lg, dl, anc, info, 0.05, 0.5, 100, 1000)
    gp = pipelines.GraphedPipeline(pipelines.dense_postprocess, *args)
    print("retinanet post B=%d  eager ms %s  graph ms %s" % (B, timed(lambda: pipelines.dense_postprocess(*args)), timed(gp.replay)))
    del lg, dl, gp
B = 16
sz3 = W.frcnn_level_sizes(800, 1344)
gen3 = DefaultAnchorGenerator(W.FRCNN_SCALES, W.FRCNN_RATIOS, W.FRCNN_RPN_STRIDES, 0.5)
anc3 = gen3.generate_anchors_by_features(sz3, dev)
sc3 = [torch.randn((B, a.shape[0]), device=dev, generator=g) * 2 - 3 for a in anc3]
dl3 = [torch.randn((B, a.shape[0], 4), device=dev, generator=g) * 0.2 for a in anc3]
info3 = torch.tensor([[800, 1344, 800, 1333, 0.0]] * B, device=dev)
args = (sc3, dl3, anc3, info3, 2000, 1000, 0.7)
gp = pipelines.GraphedPipeline(pipelines.rpn_proposals, *args)
print("rpn proposals B=16  eager ms %s  graph ms %s" % (timed(lambda: pipelines.rpn_proposals(*args)), timed(gp.replay)))
