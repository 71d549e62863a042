"""Small driver for ncu: one Gaussian-conditional plane of a 1080p frame model
(48 x 68 x 120 symbols, sigma in [0.05, 32]) encoded and decoded a few times at 4 096-symbol
sub-streams.  `ncu --set full --import-source on -k regex:rans_decode|rans_encode`."""
import math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
from deepvideocodec_b200 import coder
dev = torch.device("cuda:0")
gc = dvc.GaussianConditional(None)
gc.update_scale_table(np.exp(np.linspace(np.log(0.11), np.log(256), 64)).tolist())
gc = gc.to(dev)
g = torch.Generator(device=dev).manual_seed(1)
shape = (1, 48, 68, 120)
sg = torch.exp(torch.empty(shape, device=dev).uniform_(math.log(0.05), math.log(32), generator=g))
x = torch.round(sg * torch.randn(shape, device=dev, generator=g))
tables = gc._tables()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for _ in range(3):
    s = coder.rans_encode(tables, x=x, scales=sg, scale_table=gc.scale_table, scale_bound=0.11,
                          stream_symbols=S)
    out = coder.rans_decode(s, tables, shape, scales=sg, scale_table=gc.scale_table, scale_bound=0.11,
                            device=dev)
torch.cuda.synchronize()
assert torch.equal(out, x)
print("ok", len(s[0]), "bytes,", len(s[0]) * 8 / x.numel(), "bits/symbol")
