"""Where a `rans_decode` call spends its time: host side (wall clock of the call with an idle GPU)
against device side (CUDA events around the call), per stream of a 1080p frame model
(`[1,48,68,120]`, sigma in [0.5, 32], product-default partition)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc  # noqa: E402
from deepvideocodec_b200 import coder  # noqa: E402

dev = torch.device("cuda:0")
gc = dvc.GaussianConditional(None)
gc.update_scale_table(np.exp(np.linspace(np.log(0.11), np.log(256), 64)).tolist())
gc = gc.to(dev).eval()
shape = (1, 48, 68, 120)
g = torch.Generator(device=dev).manual_seed(1)
scales = torch.exp(torch.empty(shape, device=dev).uniform_(np.log(0.5), np.log(32.0), generator=g))
x = torch.round(torch.randn(shape, device=dev, generator=g) * scales)
kw = dict(scales=scales, scale_table=gc.scale_table)
est = len(coder.rans_encode(gc._tables(), x=x, stream_symbols=0, **kw)[0])
s = coder.rans_encode(gc._tables(), x=x, est_bytes=est, **kw)
print("container", coder.container_of(s[0], x[0].numel()), len(s[0]), "bytes")


def call(sts):
    return coder.rans_decode(s, gc._tables(), shape, want_symbols=True, statuses=sts, **kw)


for _ in range(3):
    call([])
torch.cuda.synchronize()
host, gpu = [], []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    out = call([])
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    host.append(1e3 * (t1 - t0))
    gpu.append(a.elapsed_time(b))
assert torch.equal(out, x.int())
print(f"host {np.median(host):.3f} ms   device {np.median(gpu):.3f} ms per call")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    call([])
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
