"""Debug/parity probe of the fused warp + 3x3 conv kernel (row f3)."""
import os, sys, json, time
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
from deepvideocodec_b200 import layers

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_rna(x):
    i = x.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def run(h, w, ce, cf, debug=0, seed=0, n=1):
    g = torch.Generator(device="cpu").manual_seed(seed)
    feat = torch.randn(n, cf, h, w, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    extra = torch.randn(n, ce, h, w, generator=g).to(dev).contiguous(memory_format=torch.channels_last) if ce else None
    flow = (torch.randn(n, 2, h, w, generator=g) * 3).to(dev)
    weight = (torch.randn(64, ce + cf, 3, 3, generator=g) * 0.05).to(dev)
    bias = torch.randn(64, generator=g).to(dev)
    ctx, conv = layers.warp_conv3x3(feat, flow, weight, bias, extra, _debug=debug)
    torch.cuda.synchronize()
    ref_ctx = layers.flow_warp(feat, flow)
    x = ref_ctx if extra is None else torch.cat((extra, ref_ctx), 1)
    res = {"shape": [n, ce, cf, h, w], "debug": debug,
           "warp_bit_exact": bool(torch.equal(ctx, ref_ctx)),
           "warp_max_abs": float((ctx - ref_ctx).abs().max())}
    ref32 = F.conv2d(x, weight, bias, padding=1)
    scale = float(ref32.abs().max())
    res["conv_vs_fp32_max_abs"] = float((conv - ref32).abs().max())
    res["scale"] = scale
    for name, fn in (("trunc", tf32_trunc), ("rna", tf32_rna)):
        r64 = F.conv2d(fn(x).double(), fn(weight).double(), bias.double(), padding=1)
        res[f"conv_vs_{name}_fp64_max_abs"] = float((conv.double() - r64).abs().max())
    return res


if __name__ == "__main__":
    out = []
    cases = [(8, 128, 0, 16), (8, 128, 16, 16), (12, 200, 64, 64), (68, 120, 0, 64), (272, 480, 64, 64)]
    for h, w, ce, cf in cases:
        for dbg in (0,):
            try:
                r = run(h, w, ce, cf, dbg)
            except Exception as e:
                r = {"shape": [h, w, ce, cf], "error": str(e)[:300]}
            print(json.dumps(r), flush=True)
            out.append(r)
            if "error" in r:
                break
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "warp_conv_check.json"), "w"), indent=1)
