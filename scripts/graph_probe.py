"""Does the target-assignment step capture into a CUDA graph, and what does replay save? (scratch)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import ops, workloads as W
from basedet_b200.layers import DefaultAnchorGenerator
dev = torch.device("cuda:0")
sizes = W.retinanet_level_sizes(800, 800)
gen = DefaultAnchorGenerator(W.RETINANET_SCALES, W.RETINANET_RATIOS, W.RETINANET_STRIDES, 0.5)
B, G = 16, 100
gt, ng = W.target_assign_batch(B)
gt_d, ng_d = torch.from_numpy(gt).to(dev), torch.from_numpy(ng).to(dev)
A = sum(h * w * 9 for h, w in sizes)
plan = ops.AssignPlan(A, G, B, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    anchors = gen.generate_all_level_anchors(sizes, dev)
    return ops.assign_targets(anchors, gt_d, ng_d, [0.4, 0.5], [0, -1, 1], True, True, plan=plan)
for _ in range(5): out = step()
torch.cuda.synchronize()
ref = [o.clone() for o in out if torch.is_tensor(o)]
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    gout = step()
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
for a, b in zip(ref, [o for o in gout if torch.is_tensor(o)]):
    assert torch.equal(a, b), "graph replay differs"
def timed(fn, n=300):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ts.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ts]))
def wall(fn, n=1000):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("eager  events ms", timed(step), "wall ms/step", wall(step))
print("graph  events ms", timed(g.replay), "wall ms/step", wall(g.replay))
