"""Where does the step time go?  warp only / entropy only / serial / concurrent."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
from deepvideocodec_b200.pipeline import PFramePath, synthetic_pframe_inputs
dev = torch.device("cuda:0")
torch.manual_seed(1234)
ebs = {"motion": dvc.EntropyBottleneck(64).to(dev).eval(), "frame": dvc.EntropyBottleneck(64).to(dev).eval()}
paths = [PFramePath(synthetic_pframe_inputs(1088, 1920, dev, 1234 + s), ebs) for s in range(4)]
s = torch.cuda.current_stream().cuda_stream

def warp_only(p):
    fn, name, args = p._warp_call; assert fn(*args, s) == 0
def ent_only(p):
    for fn, name, args in p._ent_calls: assert fn(*args, s) == 0
def t(fn, N=300):
    for i in range(10): fn(paths[i % 4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(N): fn(paths[i % 4])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / N * 1e3
print("warp only        %.1f us" % t(warp_only))
print("entropy only     %.1f us" % t(ent_only))
print("serial           %.1f us" % t(lambda p: p.launch(concurrent=False)))
print("concurrent       %.1f us" % t(lambda p: p.launch(concurrent=True)))
g = torch.cuda.CUDAGraph()
graphs = []
for p in paths:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        p.launch(concurrent=True)
    graphs.append(g)
print("concurrent graph %.1f us" % t(lambda p: graphs[paths.index(p)].replay()))
