"""BASELINE.json configs[4]: microbench sweep -- flow_warp at 3840x2160 (and the
x64-padded 3840x2176), C in {3, 64}, NCHW and channels_last; Gaussian-conditional
likelihood over 192-ch latents at H/16 (136x240); entropy bottleneck over 128-ch
hyper-latents at H/64 (34x60).  Also the eager PyTorch-CUDA reference ops on the
same tensors (the GPU baseline the reference would run).  Prints one JSON object.

GB/s are ALGORITHMIC bytes / CUDA-event time; inputs are rotated
through several buffer sets so that the aggregate exceeds L2 for the big cases
(small cases are L2 resident by nature and flagged)."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepvideocodec_b200 as dvc  # noqa: E402
from oracle import dmc_ref  # noqa: E402
from test_gpu_entropy import _oracle_entropy_models  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6531.9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def timeit(fns, iters=40, warm=6):
    k = len(fns)
    for i in range(warm):
        fns[i % k]()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fns[i % k]()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def smooth_flow(h, w, g):
    f = torch.randn(1, 2, h, w, device=dev, generator=g)
    f = torch.nn.functional.avg_pool2d(f, 31, stride=1, padding=15, count_include_pad=False)
    return (f / f.std() * 4.0).contiguous()


out = {"gpu": torch.cuda.get_device_name(0), "peak_gbs": PEAK, "cases": []}
g = torch.Generator(device=dev).manual_seed(5)
with torch.no_grad():
    for (h, w) in ((2160, 3840), (2176, 3840), (1088, 1920)):
        for c, fmt in ((3, "nchw"), (64, "nhwc"), (64, "nchw")):
            nset = 3 if c == 64 else 6
            mf = torch.channels_last if fmt == "nhwc" else torch.contiguous_format
            ims = [torch.randn(1, c, h, w, device=dev, generator=g).contiguous(memory_format=mf)
                   for _ in range(nset)]
            flows = [smooth_flow(h, w, g) for _ in range(nset)]
            alg = 4 * h * w * (2 * c + 2)
            t = timeit([(lambda i=i: dvc.flow_warp(ims[i], flows[i])) for i in range(nset)])
            te = timeit([(lambda i=i: dmc_ref.flow_warp(ims[i], flows[i])) for i in range(nset)], 10, 3)
            ok = torch.equal(dvc.flow_warp(ims[0], flows[0]), dmc_ref.flow_warp(ims[0], flows[0]))
            out["cases"].append({"op": "flow_warp", "shape": [1, c, h, w], "layout": fmt,
                                 "algorithmic_MB": alg / 1e6, "us": t * 1e3, "GBps": alg / t / 1e6,
                                 "frac_of_measured_peak": alg / t / 1e6 / PEAK,
                                 "eager_cuda_us": te * 1e3, "bit_identical_to_eager": bool(ok)})
            del ims, flows
            torch.cuda.empty_cache()
    # SpyNet's four 3-channel warps of a 1080p P-frame (layers.py:261; SURVEY.md 8d "optional second
    # figure", 88.78 MB): per-op launches (what the patched ME_Spynet does) and one warp_multi launch
    sp = [(136, 240), (272, 480), (544, 960), (1088, 1920)]
    sets = [[(torch.rand(1, 3, h, w, device=dev, generator=g), smooth_flow(h, w, g)) for h, w in sp]
            for _ in range(6)]
    alg = sum(4 * h * w * 8 for h, w in sp)
    t = timeit([(lambda s=s: [dvc.flow_warp(im, fl) for im, fl in s]) for s in sets])
    tm = timeit([(lambda s=s: dvc.warp_multi(s)) for s in sets])
    te = timeit([(lambda s=s: [dmc_ref.flow_warp(im, fl) for im, fl in s]) for s in sets], 10, 3)
    out["cases"].append({"op": "SpyNet warps x4 (3 ch, 136x240 .. 1088x1920)", "shape": [4, 3, 1088, 1920],
                         "layout": "nchw", "algorithmic_MB": alg / 1e6, "us": t * 1e3,
                         "us_one_launch": tm * 1e3, "GBps": alg / tm / 1e6,
                         "frac_of_measured_peak": alg / tm / 1e6 / PEAK, "eager_cuda_us": te * 1e3,
                         "bit_identical_to_eager": bool(all(
                             torch.equal(a, dmc_ref.flow_warp(im, fl))
                             for a, (im, fl) in zip(dvc.warp_multi(sets[0]), sets[0])))})
    del sets
    # Gaussian conditional, 192 ch x 136 x 240 (module boundary: 20 B/element)
    oem = _oracle_entropy_models()
    n_el = 192 * 136 * 240
    sets = []
    for _ in range(4):
        mu = torch.randn(1, 192, 136, 240, device=dev, generator=g) * 3
        sg = torch.exp(torch.empty(1, 192, 136, 240, device=dev).uniform_(
            math.log(0.05), math.log(32), generator=g))
        sets.append((mu + sg * torch.randn(1, 192, 136, 240, device=dev, generator=g), sg, mu))
    gc = dvc.GaussianConditional(None).to(dev).eval()
    gc_ref = oem.GaussianConditional(None).to(dev).eval()
    t = timeit([(lambda s=s: gc(*s)) for s in sets])
    te = timeit([(lambda s=s: gc_ref(*s)) for s in sets], 10, 3)
    out["cases"].append({"op": "GaussianConditional.forward", "shape": [1, 192, 136, 240],
                         "algorithmic_MB": 20 * n_el / 1e6, "us": t * 1e3,
                         "GBps": 20 * n_el / t / 1e6, "frac_of_measured_peak": 20 * n_el / t / 1e6 / PEAK,
                         "eager_cuda_us": te * 1e3,
                         "bit_identical_to_eager": bool(torch.equal(gc(*sets[0])[1], gc_ref(*sets[0])[1]))})
    # fused dual prior stage A + stage B/GC (40 B/element)
    priors = [torch.randn(1, 384, 136, 240, device=dev, generator=g) for _ in range(4)]
    t = timeit([(lambda i=i: (dvc.dual_prior_stage_a(*sets[i]),
                              dvc.dual_prior_stage_b_gc(sets[i][0], sets[i][2], sets[i][1], priors[i], gc, False)))
                for i in range(4)])
    out["cases"].append({"op": "dual_prior stage A + stage B/GC (2 launches)", "shape": [1, 192, 136, 240],
                         "algorithmic_MB": 40 * n_el / 1e6, "us": t * 1e3, "GBps": 40 * n_el / t / 1e6,
                         "frac_of_measured_peak": 40 * n_el / t / 1e6 / PEAK})
    del sets, priors
    # entropy bottleneck, 128 ch x 34 x 60 (12 B/element) -- 3 MB: L2 resident, latency bound
    n_el = 128 * 34 * 60
    eb = dvc.EntropyBottleneck(128).to(dev).eval()
    eb_ref = oem.EntropyBottleneck(128).to(dev).eval()
    eb_ref.load_state_dict(eb.state_dict())
    zs = [torch.randn(1, 128, 34, 60, device=dev, generator=g) * 10 for _ in range(4)]
    t = timeit([(lambda z=z: eb(z)) for z in zs])
    te = timeit([(lambda z=z: eb_ref(z)) for z in zs], 10, 3)
    out["cases"].append({"op": "EntropyBottleneck.forward", "shape": [1, 128, 34, 60],
                         "algorithmic_MB": 12 * n_el / 1e6, "us": t * 1e3, "GBps": 12 * n_el / t / 1e6,
                         "frac_of_measured_peak": 12 * n_el / t / 1e6 / PEAK, "eager_cuda_us": te * 1e3,
                         "note": "3 MB problem: L2 resident and latency bound, no roofline claim",
                         "bit_identical_to_eager": bool(torch.equal(eb(zs[0])[1], eb_ref(zs[0])[1]))})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)
for c in out["cases"]:
    print(json.dumps(c))
