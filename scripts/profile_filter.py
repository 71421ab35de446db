"""ncu driver: score filter + top-k at config 1, 8 images (the B=8 line of perf_all)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basedet_b200 import ops, _lib, workloads as W
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
B = int(os.environ.get("B", "8"))
sz = W.retinanet_level_sizes(800, 800)
lg = [torch.randn((B, h * w * 9, 80), device=dev, generator=g) * 1.25 - 6.0 for h, w in sz]
base, starts, lens = ops._segments([t.reshape(B, -1) for t in lg])
for _ in range(3):
    ops.score_filter_topk_raw(base, starts, lens, 0.05, 1000, _lib.SCORE_SIGMOID, None, None, 80)
torch.cuda.synchronize()
