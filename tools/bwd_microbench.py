#!/usr/bin/env python
"""Backward kernels of the hot path at BASELINE.json configs[3] shapes (training:
[8,3,256,256] crops, 64-channel features at 256/128/64, latents 16x16, hyper-
latents 4x4) and at 1080p: device time (CUDA events), ALGORITHMIC bytes and the
fraction of the measured HBM peak, next to the same op's eager autograd
("no roofline, no algorithmic-byte accounting for any
backward kernel").

Algorithmic bytes (fp32; every operand read once, every result written once):
  warp bwd        4*N*H*W*(3C + 4)   read grad_out C, im C, flow 2; write grad_im C, grad_flow 2
  down2 bwd       4*N*C*(H*W/4 + H*W)
  GC bwd          4*E*8              read y, mu, sigma, noise, grad_lik; write 3 grads
  stage B bwd     4*E*13             read y, mu, sigma, 2 prior planes, noise, g_lik, g_yhat;
                                     write g_y, g_mu, g_sigma, 2 g_prior
  stage A bwd     4*E*6              read 3 grad planes; write 3
  EB bwd          4*E*4              read z, noise, g_lik; write g_z (+ 58 parameter sums / channel)
The 256x256 cases are latency-bound (a launch moves 1-100 MB); the fraction is
reported for completeness, the microseconds are what matters there.
Output: gpurun_out/bwd_microbench.json
"""
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


def time_bwd(make, iters=20, warm=4):
    """make() -> (outputs tuple, inputs tuple, grad_outputs tuple); times ONLY the backward."""
    tot = 0.0
    for i in range(warm + iters):
        outs, ins, gos = make()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        torch.autograd.grad(outs, ins, gos, allow_unused=True)
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            tot += a.elapsed_time(b)
    return tot / iters * 1e3       # us


def row(name, shape, alg_bytes, us, us_eager, note=None):
    r = {"op": name, "shape": list(shape), "algorithmic_MB": alg_bytes / 1e6, "us": us,
         "GBps": alg_bytes / us / 1e3, "frac_of_measured_peak": alg_bytes / us / 1e3 / PEAK,
         "eager_autograd_us": us_eager}
    if note:
        r["note"] = note
    print(json.dumps(r), flush=True)
    return r


def main():
    g = torch.Generator(device=dev).manual_seed(3)
    out = {"gpu": torch.cuda.get_device_name(0), "peak_gbs": PEAK, "rows": []}
    oem = _oracle_entropy_models()

    def randn(*s):
        return torch.randn(*s, device=dev, generator=g)

    # ---- warp backward ------------------------------------------------------------------
    for (n, c, h, w, fmt) in ((8, 64, 256, 256, "nchw"), (8, 64, 256, 256, "nhwc"),
                              (8, 64, 128, 128, "nchw"), (8, 3, 256, 256, "nchw"),
                              (1, 64, 1088, 1920, "nhwc"), (1, 64, 1088, 1920, "nchw")):
        mf = torch.channels_last if fmt == "nhwc" else torch.contiguous_format
        im = randn(n, c, h, w).contiguous(memory_format=mf).requires_grad_(True)
        f = torch.nn.functional.avg_pool2d(randn(n, 2, h, w), 15, 1, 7, count_include_pad=False)
        flow = (f / f.std() * 3.0).contiguous().requires_grad_(True)
        go = randn(n, c, h, w).contiguous(memory_format=mf)

        def ours():
            return (dvc.flow_warp(im, flow),), (im, flow), (go,)

        def eager():
            return (dmc_ref.flow_warp(im, flow),), (im, flow), (go,)
        alg = 4 * n * h * w * (3 * c + 4)
        out["rows"].append(row("flow_warp backward", (n, c, h, w, fmt), alg, time_bwd(ours),
                               time_bwd(eager, 6, 2)))
        del im, flow, go
    if os.environ.get("ONLY_WARP"):
        return
    # ---- flow pyramid backward ------------------------------------------------------------
    mv = randn(8, 2, 256, 256).requires_grad_(True)
    go = randn(8, 2, 128, 128)
    out["rows"].append(row(
        "bilineardownsacling backward", (8, 2, 256, 256), 4 * 8 * 2 * (256 * 256 // 4 + 256 * 256),
        time_bwd(lambda: ((dvc.bilineardownsacling(mv),), (mv,), (go,))),
        time_bwd(lambda: ((dmc_ref.bilinear_down2(mv),), (mv,), (go,)), 6, 2)))
    # ---- likelihood kernels ----------------------------------------------------------------
    for (n, c, h, w) in ((8, 96, 16, 16), (1, 96, 68, 120), (1, 192, 136, 240)):
        E = n * c * h * w
        mu = (randn(n, c, h, w) * 3).requires_grad_(True)
        sg = torch.exp(torch.empty(n, c, h, w, device=dev).uniform_(
            math.log(0.05), math.log(32), generator=g)).requires_grad_(True)
        y = (mu.detach() + sg.detach() * randn(n, c, h, w)).requires_grad_(True)
        prior = randn(n, 2 * c, h, w).abs().add_(0.2).requires_grad_(True)
        gl = randn(n, c, h, w)
        gc = dvc.GaussianConditional(None).to(dev).train()
        gc_ref = oem.GaussianConditional(None).to(dev).train()
        out["rows"].append(row(
            "GaussianConditional backward (noise)", (n, c, h, w), 4 * E * 8,
            time_bwd(lambda: ((gc(y, sg, mu)[1],), (y, sg, mu), (gl,))),
            time_bwd(lambda: ((gc_ref(y, sg, mu)[1],), (y, sg, mu), (gl,)), 6, 2)))

        def stage_b():
            yh, _, _, lik, _ = dvc.dual_prior_stage_b_gc(y, mu, sg, prior, gc, training=True)
            return (yh, lik), (y, mu, sg, prior), (gl, gl)
        out["rows"].append(row("dual prior stage B + GC backward", (n, c, h, w), 4 * E * 13,
                               time_bwd(stage_b), None))
        gp = randn(n, 3 * c, h, w)
        out["rows"].append(row(
            "dual prior stage A backward", (n, c, h, w), 4 * E * 6,
            time_bwd(lambda: ((dvc.dual_prior_stage_a(y, mu, sg),), (y, mu, sg), (gp,))), None))
    for (n, c, h, w) in ((8, 64, 4, 4), (1, 64, 17, 30), (1, 128, 34, 60)):
        E = n * c * h * w
        eb = dvc.EntropyBottleneck(c).to(dev).train()
        eb_ref = oem.EntropyBottleneck(c).to(dev).train()
        eb_ref.load_state_dict(eb.state_dict())
        z = (randn(n, c, h, w) * 10).requires_grad_(True)
        gl = randn(n, c, h, w)
        ps = [p for p in eb.parameters()]
        ps_r = [p for p in eb_ref.parameters()]
        out["rows"].append(row(
            "EntropyBottleneck backward (noise, incl. parameter grads)", (n, c, h, w), 4 * E * 4,
            time_bwd(lambda: ((eb(z)[1],), (z, *ps), (gl,))),
            time_bwd(lambda: ((eb_ref(z)[1],), (z, *ps_r), (gl,)), 6, 2),
            note="58 per-channel parameter sums reduced in shared memory, no global atomics"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bwd_microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
