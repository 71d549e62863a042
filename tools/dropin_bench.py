#!/usr/bin/env python
"""Whole-P-frame time of the reference's own ``DMC`` on one B200: stock eager
vs ``deepvideocodec_b200.patch``.

Arms, all on identical weights and frames at 1088x1920 (and 256x256):

  stock_eager        unmodified dmc/models over eager PyTorch ops (the reference as written)
  patched            same files + dvc.patch(models)            (kernels bound)
  patched_graph      patched ``forward_inter`` captured in ONE CUDA graph
  patched_fused      + fuse_warp_conv=True (row f3; TF32 tensor-core conv, opt-in)

each in the reference's NCHW layout and in channels_last.  Timed region = one
``DMC.forward_inter`` call with a populated dpb (video_model.py:556-579),
CUDA events, cuDNN in its default configuration (TF32 convs allowed -- what a
user of the reference gets).  Kernel launches per P-frame are counted with the
torch profiler.  A parity summary (deterministic fp32 convs) is recorded next to
the times.  Output: gpurun_out/dropin_bench.json (copied to profiles/r02_dropin.json).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepvideocodec_b200 as dvc  # noqa: E402
import dropin_util as du  # noqa: E402


def timeit(fn, n, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def count_launches(fn):
    from torch.profiler import ProfilerActivity, profile
    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    names = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            names[ev.name] = names.get(ev.name, 0) + 1
    total = sum(names.values())
    ours = sum(v for k, v in names.items() if "dvc::" in k or "wc::" in k or "dvc_" in k)
    return total, ours


def populated_dpb(model, fr):
    """dpb after the first P-frame (so feature_ref / y_ref / y_mv_ref are set)."""
    dpb = {"x_ref": fr[0], "feature_ref": None, "y_ref": None, "y_mv_ref": None}
    x_rec, _, ctx = model.forward_inter(fr[1], dpb)
    return {"x_ref": x_rec, "feature_ref": ctx["feature_ref"], "y_ref": ctx["y_ref"],
            "y_mv_ref": ctx["y_mv_ref"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1088x1920,256x256")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "dropin_bench.json"))
    args = ap.parse_args()
    if not du.reference_available():
        raise SystemExit("reference not staged: run tools/stage_reference.py first")
    dev = torch.device("cuda:0")
    dvc.lib()
    res = {"device": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "weights": "reference init (seed 0), conv weights x 0.7 (O(1) activations)",
           "timed_region": "DMC.forward_inter with populated dpb (video_model.py:556-579), batch 1",
           "cudnn": "default (allow_tf32=True, benchmark=False)", "arms": {}}
    patch_mod = sys.modules["deepvideocodec_b200.patch"]
    for size in args.sizes.split(","):
        h, w = (int(v) for v in size.split("x"))
        for cl in (False, True):
            tag = f"{h}x{w}/{'channels_last' if cl else 'nchw'}"
            stock, patched = du.build_pair(dev, seed=0, channels_last=cl, weight_scale=0.7)
            stock.eval(), patched.eval()
            fr = du.frames(3, 1, h, w, dev, seed=1, channels_last=cl)
            arm = {}
            with torch.no_grad():
                # ---- parity (deterministic fp32 convs) -------------------------------
                with du.deterministic_convs():
                    out_s, lat_s = du.run_forward(stock, fr)
                    out_p, lat_p = du.run_forward(patched, fr)
                    rep = du.compare_forward(out_s, lat_s, out_p, lat_p, h * w * 2)
                arm["parity"] = {
                    "x_hat_max_abs": max(f["x_hat_max_abs"] for f in rep["frames"]),
                    "x_hat_scale": max(f["x_hat_scale"] for f in rep["frames"]),
                    "y_hat_bit_exact": all(f["motion.y_hat_equal"] and f["frame.y_hat_equal"]
                                           for f in rep["frames"]),
                    "lik_max_rel": max(f[f"{l}.{k}_lik_max_rel"] for f in rep["frames"]
                                       for l in ("motion", "frame") for k in ("y", "z")),
                    "bpp_max_rel": rep["bpp_max_rel"],
                    "bits_per_frame_stock": rep["bits_per_frame_stock"],
                }
                del out_s, out_p, lat_s, lat_p
                # ---- times -----------------------------------------------------------
                for name, model in (("stock_eager", stock), ("patched", patched)):
                    dpb = populated_dpb(model, fr)
                    fn = lambda m=model, d=dpb: m.forward_inter(fr[2], d)  # noqa: E731
                    arm[f"{name}_ms"] = timeit(fn, args.iters, args.warm)
                    arm[f"{name}_launches"], arm[f"{name}_own_kernels"] = count_launches(fn)
                # ---- one CUDA graph for the whole patched P-frame (row f4) -------------
                dpb = populated_dpb(patched, fr)
                try:
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            patched.forward_inter(fr[2], dpb)
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        g_out = patched.forward_inter(fr[2], dpb)
                    graph.replay()
                    torch.cuda.synchronize()
                    eager_out = patched.forward_inter(fr[2], dpb)
                    arm["patched_graph_x_hat_equal"] = bool(torch.equal(g_out[0], eager_out[0]))
                    arm["patched_graph_ms"] = timeit(graph.replay, args.iters, args.warm)
                    del graph, g_out, eager_out
                except Exception as e:  # noqa: BLE001
                    arm["patched_graph_error"] = repr(e)[:400]
                # ---- the shipped form of it: dvc.GraphedInter (copies in, clones out) -----------
                try:
                    gi = dvc.GraphedInter(patched)
                    dpb = populated_dpb(patched, fr)
                    fn = lambda d=dpb: gi(fr[2], d)  # noqa: E731
                    ref_out = patched.forward_inter(fr[2], dpb)
                    arm["graphed_inter_x_hat_equal"] = bool(torch.equal(fn()[0], ref_out[0]))
                    arm["graphed_inter_ms"] = timeit(fn, args.iters, args.warm)
                    del gi, ref_out
                except Exception as e:  # noqa: BLE001
                    arm["graphed_inter_error"] = repr(e)[:400]
                # ---- warp fused into its 3x3 conv (row f3, opt-in) ----------------------
                try:
                    dvc.unpatch()
                    dvc.patch(sys.modules["dvc_ref_patched"], fuse_warp_conv=True)
                    dpb = populated_dpb(patched, fr)
                    fn = lambda d=dpb: patched.forward_inter(fr[2], d)  # noqa: E731
                    arm["patched_fused_ms"] = timeit(fn, args.iters, args.warm)
                except Exception as e:  # noqa: BLE001
                    arm["patched_fused_error"] = repr(e)[:400]
                finally:
                    dvc.unpatch()
                    dvc.patch(sys.modules["dvc_ref_patched"])
            arm["speedup_patched"] = arm["stock_eager_ms"] / arm["patched_ms"]
            if "patched_graph_ms" in arm:
                arm["speedup_patched_graph"] = arm["stock_eager_ms"] / arm["patched_graph_ms"]
            if "graphed_inter_ms" in arm:
                arm["speedup_graphed_inter"] = arm["stock_eager_ms"] / arm["graphed_inter_ms"]
            res["arms"][tag] = arm
            print(tag, json.dumps(arm), flush=True)
            torch.cuda.empty_cache()
    assert patch_mod._saved
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
