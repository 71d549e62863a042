"""One encode + decode of a 1080p y tensor pair through the product defaults, for ncu
(`ncu --set full -k regex:ilv ...` / `--metrics gpu__time_duration.sum`).  Two regimes:
low rate with symbols drawn from the model (DVS3) and sigma in [0.5, 32] (DVC3)."""
import os
import sys

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
shape = (2, 48, 68, 120)                      # both checkerboard passes of the frame model
g = torch.Generator(device=dev).manual_seed(1)
for regime, lo, hi, floor in (("low", 0.05, 2.0, 0.97), ("high", 0.5, 32.0, 0.0)):
    scales = torch.exp(torch.empty(shape, device=dev).uniform_(np.log(lo), np.log(hi), generator=g))
    scales[torch.rand(shape, device=dev, generator=g) < floor] = 0.05
    x = torch.round(torch.randn(shape, device=dev, generator=g) * scales.clamp_min(0.11))
    kw = dict(x=x, scales=scales, scale_table=gc.scale_table)
    est = len(coder.rans_encode(gc._tables(), stream_symbols=0, **kw)[0]) if regime == "high" else 8000
    torch.cuda.synchronize()
    for _ in range(2):
        s = coder.rans_encode(gc._tables(), est_bytes=est, **kw)
        out = coder.rans_decode(s, gc._tables(), shape, scales=scales, scale_table=gc.scale_table)
    torch.cuda.synchronize()
    assert torch.equal(out, x)
    print(regime, coder.container_of(s[0], x[0].numel()), [len(b) for b in s], flush=True)
