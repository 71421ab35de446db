import torch, triton, triton.language as tl
from triton.tools.tensor_descriptor import TensorDescriptor
@triton.jit
def k(desc, out_ptr, BM: tl.constexpr, BN: tl.constexpr):
    t = desc.load([0, 0])
    offs = tl.arange(0, BM)[:, None] * BN + tl.arange(0, BN)[None, :]
    tl.store(out_ptr + offs, t)
x = torch.arange(64 * 64, dtype=torch.float32, device="cuda").reshape(64, 64)
out = torch.empty((16, 32), dtype=torch.float32, device="cuda")
d = TensorDescriptor(x, [64, 64], [64, 1], [16, 32])
k[(1,)](d, out, 16, 32)
torch.cuda.synchronize()
print("triton TMA ok:", torch.equal(out, x[:16, :32]))
