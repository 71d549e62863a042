import sys, torch
sys.path.insert(0, '/root/repo')
import deepvideocodec_b200 as dvc
from oracle import dmc_ref
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(1)
im = torch.randn(1, 16, 64, 128, device=dev, generator=g)
flow = torch.randn(1, 2, 64, 128, device=dev, generator=g) * 0.5
out = dvc.flow_warp(im, flow)
torch.cuda.synchronize()
print("ok", torch.equal(out, dmc_ref.flow_warp(im, flow)))
