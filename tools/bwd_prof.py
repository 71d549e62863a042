"""Driver for ncu: flow_warp backward at the training shape [8,64,256,256], channels_last
(warp_bwd_vec4_kernel) and NCHW (warp_bwd_strided_kernel), and the GC / stage-B backward."""
import math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
n, c, h, w = 8, 64, 256, 256
f = torch.nn.functional.avg_pool2d(torch.randn(n, 2, h, w, device=dev, generator=g), 15, 1, 7, count_include_pad=False)
flow = (f / f.std() * 3.0).contiguous().requires_grad_(True)
for mf in (torch.channels_last, torch.contiguous_format):
    im = torch.randn(n, c, h, w, device=dev, generator=g).contiguous(memory_format=mf).requires_grad_(True)
    go = torch.randn(n, c, h, w, device=dev, generator=g).contiguous(memory_format=mf)
    for _ in range(3):
        out = dvc.flow_warp(im, flow)
        torch.autograd.grad((out,), (im, flow), (go,))
gc = dvc.GaussianConditional(None).to(dev).train()
mu = (torch.randn(8, 96, 16, 16, device=dev, generator=g) * 3).requires_grad_(True)
sg = torch.exp(torch.empty(8, 96, 16, 16, device=dev).uniform_(math.log(0.05), math.log(32), generator=g)).requires_grad_(True)
y = (mu.detach() + sg.detach() * torch.randn(8, 96, 16, 16, device=dev, generator=g)).requires_grad_(True)
gl = torch.randn(8, 96, 16, 16, device=dev, generator=g)
for _ in range(3):
    torch.autograd.grad((gc(y, sg, mu)[1],), (y, sg, mu), (gl,))
torch.cuda.synchronize()
print("ok")
