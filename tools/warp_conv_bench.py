"""Row f3 benchmark: fused warp + conv1_out (tcgen05) vs flow_warp kernel + cuDNN conv.
1080p, Cf = Ce = 64 (MultiScaleContextFusion.conv1_out, video_model.py:46,61)."""
import os, sys, json
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
from deepvideocodec_b200 import layers

dev = torch.device("cuda:0")
H, W = int(os.environ.get("H", 1088)), int(os.environ.get("W", 1920))
CE = int(os.environ.get("CE", 64))
DEBUG = int(os.environ.get("DEBUG", 0))
FLOW = os.environ.get("FLOW", "smooth")
torch.manual_seed(0)
NSETS = 3
cl = torch.channels_last
feats = [torch.randn(1, 64, H, W, device=dev).contiguous(memory_format=cl) for _ in range(NSETS)]
extras = [torch.randn(1, CE, H, W, device=dev).contiguous(memory_format=cl) for _ in range(NSETS)] if CE else [None] * NSETS
lp = F.avg_pool2d(torch.randn(1, 2, H, W, device=dev), 31, 1, 15)
flow = lp / lp.std() * 4.0
if FLOW == "iid":
    flow = torch.randn(1, 2, H, W, device=dev) * 16.0
elif FLOW == "rigid":
    flow = torch.zeros(1, 2, H, W, device=dev) + 2.3
weight = torch.randn(64, CE + 64, 3, 3, device=dev) * 0.05
w_cl = weight.contiguous(memory_format=cl)
bias = torch.randn(64, device=dev)
if os.environ.get("OLD_ABI"):      # A/B against a build that predates the Ce/Cf-aware packing
    import ctypes
    from deepvideocodec_b200 import _native as nat
    fn = nat.lib().dvc_conv3x3_pack_weights
    fn.argtypes = [ctypes.c_void_p, nat._P4, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    packed = torch.empty((CE + 64) * 9 * 64, device=dev)
    assert fn(weight.data_ptr(), nat.st4(weight), 64, CE + 64, packed.data_ptr(), nat.stream_of(weight)) == 0
else:
    packed = layers.pack_conv3x3_weight(weight, CE)


def timeit(fn, n=30, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def fused(i):
    return layers.warp_conv3x3(feats[i % NSETS], flow, weight, bias, extras[i % NSETS], packed=packed, _debug=DEBUG)


def unfused_nhwc(i):
    ctx = layers.flow_warp(feats[i % NSETS], flow)
    x = ctx if CE == 0 else torch.cat((extras[i % NSETS], ctx), 1)
    return ctx, F.conv2d(x, w_cl, bias, padding=1)


feats_nchw = [f.contiguous() for f in feats[:2]]
extras_nchw = [e.contiguous() if e is not None else None for e in extras[:2]]


def unfused_nchw(i):     # the reference's memory format
    ctx = layers.flow_warp(feats_nchw[i % 2], flow)
    x = ctx if CE == 0 else torch.cat((extras_nchw[i % 2], ctx), 1)
    return ctx, F.conv2d(x, weight, bias, padding=1)


def conv_only_nhwc(i, xs=[None]):
    if xs[0] is None:
        xs[0] = torch.randn(1, CE + 64, H, W, device=dev).contiguous(memory_format=cl)
    return F.conv2d(xs[0], w_cl, bias, padding=1)


res = {"H": H, "W": W, "Ce": CE, "Cf": 64, "debug": DEBUG, "flow": FLOW}
with torch.no_grad():
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    res["fused_us"] = timeit(fused)
    if not os.environ.get("FUSED_ONLY"):
        res["warp_plus_cudnn_nhwc_tf32_us"] = timeit(unfused_nhwc)
        res["cudnn_conv_only_nhwc_tf32_us"] = timeit(conv_only_nhwc)
        res["warp_plus_cudnn_nchw_tf32_us"] = timeit(unfused_nchw, n=10, warm=3)
    flops = 2.0 * H * W * 64 * (CE + 64) * 9
    res["fused_tflops"] = flops / res["fused_us"] / 1e6
    # algorithmic HBM bytes: read feat, extra, flow; write ctx, conv
    byts = 4.0 * H * W * (64 + CE + 2 + 64 + 64)
    res["fused_hbm_gbs_algorithmic"] = byts / res["fused_us"] / 1e3
    c1, v1 = fused(0)
    c2, v2 = unfused_nhwc(0)
    res["warp_bit_exact"] = bool(torch.equal(c1, c2))
    res["conv_vs_cudnn_tf32_max_abs"] = float((v1 - v2).abs().max())
    res["conv_scale"] = float(v2.abs().max())
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"warp_conv_bench_{H}x{W}_ce{CE}_{FLOW}_d{DEBUG}.json"), "w"), indent=1)
