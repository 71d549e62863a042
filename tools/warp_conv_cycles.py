"""In-kernel cycle accounting of the fused warp + conv kernel (CTA 0): who waits for whom.
Needs a profiling build:  nvcc ... -DWC_PROFILE -o deepvideocodec_b200/libdvc_prof.so csrc/dvc_*.cu"""
import os, sys, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["DVC_B200_LIB"] = os.path.join(ROOT, "deepvideocodec_b200", "libdvc_prof.so")
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
from deepvideocodec_b200 import layers, _native as nat
dev = torch.device("cuda:0")
H, W = 1088, 1920
cl = torch.channels_last
torch.manual_seed(0)
feat = torch.randn(1, 64, H, W, device=dev).contiguous(memory_format=cl)
extra = torch.randn(1, 64, H, W, device=dev).contiguous(memory_format=cl)
lp = torch.nn.functional.avg_pool2d(torch.randn(1, 2, H, W, device=dev), 31, 1, 15)
flow = lp / lp.std() * 4.0
weight = torch.randn(64, 128, 3, 3, device=dev) * 0.05
bias = torch.randn(64, device=dev)
fn = nat.lib().dvc_debug_warp_conv_profile
buf = (ctypes.c_ulonglong * 12)()
with torch.no_grad():
    for _ in range(3):
        layers.warp_conv3x3(feat, flow, weight, bias, extra)
    fn(buf, 1)
    layers.warp_conv3x3(feat, flow, weight, bias, extra)
    fn(buf, 1)
names = ["mma wait-full (extra slices)", "mma wait-full (warped slices)", "mma total", "producer wait-empty (extra)",
         "producer wait-empty (warped)", "producer fill warped (acquire->arrive)", "producer total", "taps phase",
         "extra: STS + LDG issue", "extra: fence.proxy.async", "extra: syncwarp + arrive", "-"]
tot = buf[2]
for n, v in zip(names, buf):
    print(f"{n:42s} {v:12d} clk  {100.0 * v / max(tot, 1):5.1f} % of the MMA role's time")
