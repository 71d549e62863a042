"""Cycle accounting of the lane-interleaved coder's chain warp (build with -DDVC_ILV_PROF:
`python tools/coder_prof.py --build` writes tools/_ab/libdvc_prof.so; run on the GPU with
DVC_B200_LIB pointing at it).  One encode + decode per regime, the kernels print their own
clock64() split: waiting for the helpers / building lists / walking the chains."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_ab", "libdvc_prof.so")

if "--build" in sys.argv:
    from deepvideocodec_b200 import _native as nat
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = nat.nvcc_command(out_path=OUT, extra=("-DDVC_ILV_PROF",))
    subprocess.run(cmd, check=True)
    print("built", OUT)
    sys.exit(0)

os.environ.setdefault("DVC_B200_LIB", OUT)
import torch  # noqa: E402
import deepvideocodec_b200 as dvc  # noqa: E402
from deepvideocodec_b200 import coder  # noqa: E402
import numpy as np  # noqa: E402

dev = torch.device("cuda:0")
gc = dvc.GaussianConditional(None)
gc.update_scale_table(np.exp(np.linspace(np.log(0.11), np.log(256), 64)).tolist())
gc = gc.to(dev).eval()
shape = (1, 32, 68, 120)
g = torch.Generator(device=dev).manual_seed(1)
for regime, lo, hi, floor in (("high", 0.5, 32.0, 0.0), ("low", 0.05, 2.0, 0.97), ("mid", 20.0, 64.0, 0.0), ("wide", 64.0, 256.0, 0.0)):
    scales = torch.exp(torch.empty(shape, device=dev).uniform_(np.log(lo), np.log(hi), generator=g))
    scales[torch.rand(shape, device=dev, generator=g) < floor] = 0.05
    x = torch.round(torch.randn(shape, device=dev, generator=g) * scales.clamp_min(0.11))
    for skip in ((False,) if regime in ("mid", "wide") else (False, True)):
        print(f"--- {regime} skip={skip}", flush=True)
        s = coder.rans_encode(gc._tables(), x=x, scales=scales, scale_table=gc.scale_table,
                              stream_symbols=131072, lanes=32, skip=skip)
        torch.cuda.synchronize()
        out = coder.rans_decode(s, gc._tables(), shape, scales=scales, scale_table=gc.scale_table)
        torch.cuda.synchronize()
        assert torch.equal(out, x)
        print(f"bytes {len(s[0])}", flush=True)
