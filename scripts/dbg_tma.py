import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import ops, workloads as W
from oracle import c_oracle as C
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
Cn = int(os.environ.get("DBG_C", "8"))
feat = rng.normal(0, 1, (1, Cn, 40, 56)).astype(np.float32)
rois = W.make_rois(rng, 4, 1, 160, 224, 8, 100)
what = os.environ.get("DBG_WHAT", "fwd")
if what == "fwd":
    out = ops.roi_align_fwd([torch.from_numpy(feat).to(dev)], torch.from_numpy(rois).to(dev), None, [0.25], (7, 7))
    torch.cuda.synchronize()
    ref = C.roi_align_fwd(feat, rois, (7, 7), 0.25)
    print("fwd equal:", np.array_equal(out.cpu().numpy(), ref), np.abs(out.cpu().numpy() - ref).max())
else:
    dout = rng.normal(0, 1, (4, Cn, 7, 7)).astype(np.float32)
    g = ops.roi_align_bwd(torch.from_numpy(dout).to(dev), [feat.shape], torch.from_numpy(rois).to(dev), None, [0.25], (7, 7))[0]
    torch.cuda.synchronize()
    gref = C.roi_align_bwd(dout, feat.shape, rois, (7, 7), 0.25)
    print("bwd err:", np.abs(g.cpu().numpy() - gref).max())
