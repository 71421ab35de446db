"""HBM ceilings of this device: write-only, read-only, copy (own kernels, cudaMemset, torch copy_)."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from basedet_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
GB = 1 << 30
a = torch.empty(2 * GB, dtype=torch.uint8, device=dev); b = torch.empty(2 * GB, dtype=torch.uint8, device=dev)
a.zero_(); b.zero_()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
res = {}
nbytes = 2 * GB
for cps in (4, 8, 16, 32):
    res["write_v4_cps%d_GBs" % cps] = nbytes / t(lambda: _lib.check(lib.bdet_bw_probe(a.data_ptr(), None, nbytes, 0, cps, st))) / 1e6
    res["read_v4_cps%d_GBs" % cps] = nbytes / t(lambda: _lib.check(lib.bdet_bw_probe(b.data_ptr(), a.data_ptr(), nbytes, 1, cps, st))) / 1e6
    res["copy_v4_cps%d_GBs(r+w)" % cps] = 2 * nbytes / t(lambda: _lib.check(lib.bdet_bw_probe(b.data_ptr(), a.data_ptr(), nbytes, 2, cps, st))) / 1e6
res["cudaMemset_GBs"] = nbytes / t(lambda: _lib.check(lib.bdet_bw_probe(a.data_ptr(), None, nbytes, 3, 0, st))) / 1e6
res["torch_zero_GBs"] = nbytes / t(lambda: a.zero_()) / 1e6
res["torch_copy_GBs(r+w)"] = 2 * nbytes / t(lambda: b.copy_(a)) / 1e6
af = a.view(torch.float32)
res["torch_sum_read_GBs"] = nbytes / t(lambda: af.sum()) / 1e6
# smaller write (768 MB, the IoU matrix size of config 2)
nb = 768 << 20
res["write_v4_768MB_GBs"] = nb / t(lambda: _lib.check(lib.bdet_bw_probe(a.data_ptr(), None, nb, 0, 8, st))) / 1e6
res["read_v4_768MB_GBs"] = nb / t(lambda: _lib.check(lib.bdet_bw_probe(b.data_ptr(), a.data_ptr(), nb, 1, 8, st))) / 1e6
print(json.dumps(res, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bw_probe.json", "w"), indent=1)
