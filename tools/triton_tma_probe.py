# diagnostic only (not product): does a TMA tensor load work at all on this box?
import torch, triton, triton.language as tl
from triton.tools.tensor_descriptor import TensorDescriptor

@triton.jit
def k(desc, out_ptr, BH: tl.constexpr, BW: tl.constexpr):
    t = desc.load([3, 5])
    offs = tl.arange(0, BH)[:, None] * BW + tl.arange(0, BW)[None, :]
    tl.store(out_ptr + offs, t)

x = torch.arange(256 * 128, device="cuda", dtype=torch.float32).reshape(256, 128)
out = torch.empty(32 * 32, device="cuda", dtype=torch.float32)
desc = TensorDescriptor.from_tensor(x, [32, 32])
kk = k[(1,)](desc, out, 32, 32)
torch.cuda.synchronize()
print("triton TMA ok:", out[0].item(), "expect", 3 * 128 + 5)
ptx = kk.asm["ptx"]
print([l.strip() for l in ptx.splitlines() if "cp.async.bulk.tensor" in l][:2])
sass = kk.asm.get("sass", "")
print([l.strip()[:80] for l in sass.splitlines() if "UTMALDG" in l][:2])
