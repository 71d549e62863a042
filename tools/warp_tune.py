"""Time the 1080p motion-compensation launch (warp_multi) alone, CUDA events."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
from deepvideocodec_b200.pipeline import PFramePath, synthetic_pframe_inputs, pframe_algorithmic_bytes
dev = torch.device("cuda:0")
regime = sys.argv[1] if len(sys.argv) > 1 else "smooth"
torch.manual_seed(1234)
ebs = {"motion": dvc.EntropyBottleneck(64).to(dev).eval(), "frame": dvc.EntropyBottleneck(64).to(dev).eval()}
paths = [PFramePath(synthetic_pframe_inputs(1088, 1920, dev, 1234 + s, regime=regime), ebs) for s in range(4)]
alg = pframe_algorithmic_bytes(1088, 1920)["warp_multi"]
s = torch.cuda.current_stream().cuda_stream
def run(p):
    fn, name, args = p._warp_call
    rc = fn(*args, s); assert rc == 0
for i in range(8): run(paths[i % 4])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
N = 200
for i in range(N): run(paths[i % 4])
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / N
print(f"DVC_WARP_PREFETCH={os.environ.get('DVC_WARP_PREFETCH','default')} regime={regime} warp_multi {ms*1e3:.1f} us  {alg/ms/1e6:.0f} GB/s")
