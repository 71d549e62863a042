import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
h, w, c = 1088, 1920, 64
f = torch.randn(1, 2, h, w, device=dev, generator=g)
f = torch.nn.functional.avg_pool2d(f, 31, stride=1, padding=15, count_include_pad=False)
flow = (f / f.std() * 4.0).contiguous()
ims = [torch.randn(1, c, h, w, device=dev, generator=g) for _ in range(3)]
with torch.no_grad():
    for i in range(6):
        dvc.flow_warp(ims[i % 3], flow)
torch.cuda.synchronize()
print("ok")
