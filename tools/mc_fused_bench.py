"""Row f3 at module level: DMC.motion_compensation (4 warps + MultiScaleContextFusion,
video_model.py:37-66, 497-506) at 1080p --
  stock     : the oracle restatement of the reference path in eager PyTorch (NCHW, cuDNN TF32)
  unfused   : this package's one-launch warps + the same fusion net on cuDNN
  fused     : every context warp fused into its conv{1,2,3}_out (tcgen05), rest on cuDNN
Reports times and the agreement of the outputs."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
from deepvideocodec_b200 import patch as _  # noqa
import importlib
P = importlib.import_module("deepvideocodec_b200.patch")
from oracle import dmc_ref

dev = torch.device("cuda:0")
H, W = 1088, 1920
torch.manual_seed(0)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
cl = torch.channels_last


import copy
_net0 = dmc_ref.MultiScaleContextFusionRef().to(dev).eval()


def make(fmt):
    net = copy.deepcopy(_net0).to(memory_format=fmt)      # same parameters in both layouts
    g = torch.Generator(device="cpu").manual_seed(1)
    x_ref = torch.rand(1, 3, H, W, generator=g).to(dev)
    feats = [torch.randn(1, 64, H >> k, W >> k, generator=g).to(dev).contiguous(memory_format=fmt) for k in range(3)]
    lp = torch.nn.functional.avg_pool2d(torch.randn(1, 2, H, W, generator=g).to(dev), 31, 1, 15)
    mv = (lp / lp.std() * 4.0).contiguous()

    class Stub:
        context_fusion_net = net

        def multi_scale_feature_extractor(self, dpb):
            return feats
    return Stub(), x_ref, feats, mv, net


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


res = {"H": H, "W": W}
with torch.no_grad():
    stub, x_ref, feats, mv, net = make(torch.contiguous_format)
    res["stock_eager_nchw_us"] = timeit(lambda: dmc_ref.motion_compensation(x_ref, *feats, mv, net))
    ref = dmc_ref.motion_compensation(x_ref, *feats, mv, net)
    res["unfused_nchw_us"] = timeit(lambda: P._motion_compensation(stub, mv, {"x_ref": x_ref}))
    del stub, feats, net
    torch.cuda.empty_cache()
    stub, x_ref, feats, mv, net = make(cl)
    res["unfused_nhwc_us"] = timeit(lambda: P._motion_compensation(stub, mv, {"x_ref": x_ref}))
    res["fused_nhwc_us"] = timeit(lambda: dvc.motion_compensation_fused(stub, mv, {"x_ref": x_ref}))
    out = dvc.motion_compensation_fused(stub, mv, {"x_ref": x_ref})
    res["warpframe_bit_exact"] = bool(torch.equal(out[3], ref[3]))
    for k in range(3):
        res[f"context{k + 1}_max_abs_vs_stock_tf32"] = float((out[k] - ref[k]).abs().max())
        res[f"context{k + 1}_scale"] = float(ref[k].abs().max())
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mc_fused_bench.json"), "w"), indent=1)
