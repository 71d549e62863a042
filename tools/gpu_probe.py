"""First-contact diagnostics on a B200: stage-by-stage numerics of the warp
replay, error distributions of the likelihood kernels, and quick CUDA-event
timings.  Writes gpurun_out/probe.json.  Not a test, not a benchmark."""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc  # noqa: E402
from oracle import dmc_ref  # noqa: E402

dev = torch.device("cuda:0")
out = {"gpu": torch.cuda.get_device_name(0)}


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


# 1. linspace closed form vs torch.linspace on CUDA
bad = {}
for S in (68, 120, 136, 240, 256, 272, 480, 544, 960, 1088, 1920, 2160, 2176, 3840, 17, 23):
    ref = torch.linspace(-1.0, 1.0, S, device=dev)
    step = torch.tensor(2.0, dtype=torch.float32) / torch.tensor(float(S - 1), dtype=torch.float32)
    j = torch.arange(S, dtype=torch.float64)
    lo = (step.double() * j - 1.0).float()          # single rounding == fma
    hi = (1.0 - step.double() * (S - 1 - j)).float()
    mine = torch.where(j < S // 2, lo, hi).to(dev)
    bad[S] = int((mine != ref).sum())
out["linspace_mismatches"] = bad

# 2. division by python scalar on CUDA == multiply by fp32 reciprocal?
x = torch.randn(1 << 20, device=dev) * 30
d = 959.5
q = x / d
recip = x * torch.tensor(1.0, dtype=torch.float32).div(torch.tensor(d, dtype=torch.float32)).item()
true = (x.double() / d).float()
out["div_scalar_equals_recip_mul"] = bool(torch.equal(q, recip))
out["div_scalar_equals_true_div"] = bool(torch.equal(q, true))

# 3. warp parity statistics at 1080p
def smooth_flow(h, w, sigma, g):
    f = torch.randn(1, 2, h, w, device=dev, generator=g)
    f = torch.nn.functional.avg_pool2d(f, 31, stride=1, padding=15, count_include_pad=False)
    return (f / f.std() * sigma).contiguous()

g = torch.Generator(device=dev).manual_seed(1)
H, W = 1088, 1920
stats = {}
for c, fmt, name in ((3, torch.contiguous_format, "c3_nchw"), (64, torch.contiguous_format, "c64_nchw"),
                     (64, torch.channels_last, "c64_nhwc")):
    im = torch.randn(1, c, H, W, device=dev, generator=g).contiguous(memory_format=fmt)
    for regime, flow in (("smooth", smooth_flow(H, W, 4.0, g)),
                         ("wild", torch.randn(1, 2, H, W, device=dev, generator=g) * 16)):
        ref = dmc_ref.flow_warp(im, flow)
        o = dvc.flow_warp(im, flow)
        diff = (o - ref).abs()
        alg_bytes = 4 * H * W * (2 * c + 2)
        t = timeit(lambda: dvc.flow_warp(im, flow))
        te = timeit(lambda: dmc_ref.flow_warp(im, flow), iters=5, warm=2)
        stats[f"{name}_{regime}"] = {
            "max_abs": diff.max().item(), "frac_bitexact": (o == ref).float().mean().item(),
            "ms": t, "GBps": alg_bytes / t / 1e6, "eager_ms": te}
        del ref, o, diff
    del im
out["warp"] = stats

# 4. likelihood error distributions
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_entropy import _oracle_entropy_models, _latents  # noqa: E402
oem = _oracle_entropy_models()
lik = {}
for lo, hi in ((0.05, 2.0), (2.0, 32.0), (32.0, 256.0)):
    mu = torch.randn(1, 96, 68, 120, device=dev, generator=g) * 3
    sg = torch.exp(torch.empty(1, 96, 68, 120, device=dev).uniform_(math.log(lo), math.log(hi), generator=g))
    y = mu + sg * torch.randn(1, 96, 68, 120, device=dev, generator=g)
    with torch.no_grad():
        _, r = oem.GaussianConditional(None).to(dev).eval()(y, sg, mu)
        _, o = dvc.GaussianConditional(None).to(dev).eval()(y, sg, mu)
    rel = ((o - r).abs() / r)
    lik[f"gc_scales_{lo}_{hi}"] = {"max_rel": rel.max().item(), "frac_exact": (o == r).float().mean().item(),
                                    "frac_gt_1e-5": (rel > 1e-5).float().mean().item()}
for spread in (1.0, 10.0, 40.0):
    torch.manual_seed(3)
    rm = oem.EntropyBottleneck(64).to(dev).eval()
    with torch.no_grad():
        for name, p in rm.named_parameters():
            if name.startswith("_factor"):
                p.uniform_(-0.8, 0.8)
    m = dvc.EntropyBottleneck(64).to(dev).eval()
    m.load_state_dict(rm.state_dict())
    z = torch.randn(1, 64, 17, 30, device=dev, generator=g) * spread
    with torch.no_grad():
        _, r = rm(z)
        _, o = m(z)
    rel = ((o - r).abs() / r)
    lik[f"eb_spread_{spread}"] = {"max_rel": rel.max().item(), "frac_exact": (o == r).float().mean().item(),
                                  "frac_gt_1e-5": (rel > 1e-5).float().mean().item()}
out["likelihood"] = lik

# 5. micro timings of the entropy kernels at 1080p latent sizes
y, mu, sg = _latents(1, 96, 68, 120, dev, g)
prior = torch.randn(1, 192, 68, 120, device=dev, generator=g)
gc = dvc.GaussianConditional(None).to(dev).eval()
eb = dvc.EntropyBottleneck(64).to(dev).eval()
z = torch.randn(1, 64, 17, 30, device=dev, generator=g) * 10
with torch.no_grad():
    out["entropy_ms"] = {
        "stage_a": timeit(lambda: dvc.dual_prior_stage_a(y, mu, sg)),
        "stage_b_gc": timeit(lambda: dvc.dual_prior_stage_b_gc(y, mu, sg, prior, gc, False)),
        "eb": timeit(lambda: dvc.entropy_models.eb_forward(eb, z, want_outputs=False, want_zhat=True)),
        "oracle_gc_eager": timeit(lambda: oem.GaussianConditional(None).to(dev).eval()(y, sg, mu), 5, 2),
    }
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
