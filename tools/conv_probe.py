"""Row f3 groundwork: how long does cuDNN take for the 3x3 convs that consume the
warped contexts (video_model.py:37-65 MultiScaleContextFusion), and what does a
TF32 GEMM of the same flop count reach?  Decides whether a fused warp + conv
implicit GEMM can win.  Prints JSON to stdout."""
import json
import os
import sys

import torch
import torch.nn.functional as F

dev = torch.device("cuda:0")
torch.manual_seed(0)
H, W = 1088, 1920
out = {}


def timeit(fn, n=20, warm=5):
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


cases = {
    "conv1_out 128->64 @1088x1920": (128, 64, H, W),
    "conv2_out 128->64 @544x960": (128, 64, H // 2, W // 2),
    "conv3_out 64->64 @272x480": (64, 64, H // 4, W // 4),
    "conv3_up 64->256 @272x480": (64, 256, H // 4, W // 4),
    "resblock conv 64->64 @1088x1920": (64, 64, H, W),
}
for name, (ci, co, h, w) in cases.items():
    wgt = torch.randn(co, ci, 3, 3, device=dev) * 0.05
    bias = torch.randn(co, device=dev)
    flops = 2.0 * h * w * co * ci * 9
    for fmt_name, fmt in (("nchw", torch.contiguous_format), ("nhwc", torch.channels_last)):
        xs = [torch.randn(1, ci, h, w, device=dev).contiguous(memory_format=fmt) for _ in range(3)]
        wf = wgt.contiguous(memory_format=fmt)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            i = [0]

            def fn():
                i[0] += 1
                return F.conv2d(xs[i[0] % 3], wf, bias, padding=1)
            try:
                us = timeit(fn)
            except Exception as e:  # noqa
                out[f"{name} {fmt_name} tf32={tf32}"] = str(e)[:100]
                continue
            out[f"{name} {fmt_name} tf32={tf32}"] = {"us": round(us, 1), "tflops": round(flops / us / 1e6, 1)}
        del xs

# TF32 / fp32 GEMM of the conv1_out shape as an implicit GEMM: M = pixels, N = 64, K = 1152
torch.backends.cuda.matmul.allow_tf32 = True
a = torch.randn(H * W, 1152, device=dev)
b = torch.randn(1152, 64, device=dev)
us = timeit(lambda: a @ b)
out["matmul tf32 [2088960x1152]x[1152x64]"] = {"us": round(us, 1), "tflops": round(2.0 * H * W * 1152 * 64 / us / 1e6, 1)}
a2 = torch.randn(8192, 8192, device=dev)
b2 = torch.randn(8192, 8192, device=dev)
us = timeit(lambda: a2 @ b2)
out["matmul tf32 8192^3"] = {"us": round(us, 1), "tflops": round(2.0 * 8192 ** 3 / us / 1e6, 1)}
a3 = a.bfloat16()
b3 = b.bfloat16()
us = timeit(lambda: a3 @ b3)
out["matmul bf16 [2088960x1152]x[1152x64]"] = {"us": round(us, 1), "tflops": round(2.0 * H * W * 1152 * 64 / us / 1e6, 1)}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "conv_probe.json"), "w") as f:
    json.dump(out, f, indent=1)
