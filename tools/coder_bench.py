"""Entropy-coder bench at the 1080p latent sizes of
BASELINE.json configs[1]: the six bit streams of one P-frame
(motion / frame: two checkerboard planes of y + the hyper-latent z).

For each sub-stream length S: CUDA-event time of the encode launches (kernel
only, streams stay on the device), of the decode launches, the coded size and
its container overhead; next to it the plain-C oracle coder (one host thread,
the arithmetic CompressAI runs on the CPU; its Python list marshalling is NOT
included, so this flatters the CPU side) on the same symbols.  Two regimes:
"synthetic" (SURVEY.md 8d latents, sigma in [0.05, 32]: ~3 bits/symbol) and
"sparse" (97 % of the scales at the 0.11 floor and a spatial prior that predicts
y: what a trained codec at low rate produces), because the overhead of sub-streams only shows
against small payloads.  Prints one JSON object."""
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc  # noqa: E402
from deepvideocodec_b200 import coder  # noqa: E402
from deepvideocodec_b200.context import dual_prior_stage_b_gc, stacked_planes  # noqa: E402
from deepvideocodec_b200.pipeline import synthetic_pframe_inputs  # noqa: E402

dev = torch.device("cuda:0")


def scale_table():
    return np.exp(np.linspace(np.log(0.11), np.log(256), 64)).tolist()


def ev_time(fn, iters=20, warm=3):
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


def planes_of(inp, gc, sparse):
    """The tensors the six encoder calls of one P-frame see."""
    jobs, pairs = [], []
    for label in ("motion", "frame"):
        y, mu, sg, prior = (inp[f"{label}.{k}"] for k in ("y", "means", "scales", "prior"))
        if sparse:
            g = torch.Generator(device=dev).manual_seed(3)
            keep = torch.rand(sg.shape, device=dev, generator=g) < 0.03
            sg = torch.where(keep, sg, torch.full_like(sg, 0.05))
            y = mu + sg * torch.randn(sg.shape, device=dev, generator=g)
            c = y.size(1)
            ps = prior.clone()
            ps[:, :c // 2] = mu[:, :c // 2]          # a trained spatial prior predicts y
            ps[:, c:3 * c // 2] = mu[:, c // 2:]
            for lo in (c // 2, 3 * c // 2):
                blk = ps[:, lo:lo + c // 2]
                kp = torch.rand(blk.shape, device=dev, generator=g) < 0.03
                ps[:, lo:lo + c // 2] = torch.where(kp, blk, torch.full_like(blk, 0.05))
            prior = ps
        _, _, _, _, planes = dual_prior_stage_b_gc(
            y, mu, sg, prior, gc, training=False, compress=True)
        q0, q1, s0, s1 = planes
        jobs += [(f"{label}.y0", q0, s0), (f"{label}.y1", q1, s1)]
        pairs.append(stacked_planes(planes))
    return jobs, pairs


def main():
    from oracle import rans
    torch.manual_seed(0)
    gc = dvc.GaussianConditional(None)
    gc.update_scale_table(scale_table())
    cdf, size, off = (gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy())
    gc = gc.to(dev).eval()
    eb = dvc.EntropyBottleneck(64)
    eb.update()
    ecdf, esize, eoff = (eb._quantized_cdf.numpy(), eb._cdf_length.numpy(), eb._offset.numpy())
    eb = eb.to(dev).eval()
    med = eb._get_medians().detach().reshape(1, -1, 1, 1)
    inp = synthetic_pframe_inputs(1088, 1920, dev, seed=7)
    out = {"gpu": torch.cuda.get_device_name(0), "regimes": {}}
    with torch.no_grad():
        for regime in ("synthetic", "sparse"):
            jobs, pairs = planes_of(inp, gc, regime == "sparse")
            zs = [inp["motion.z"], inp["frame.z"]]
            if regime == "sparse":
                zs = [z * 0.1 for z in zs]
            n_sym = sum(q.numel() for _, q, _ in jobs) + sum(z.numel() for z in zs)
            res = {"symbols_per_frame": n_sym, "sweep": []}
            # ---- CPU: plain-C coder, one thread, symbols/indexes already as int32 arrays
            cpu_in = []
            for _, q, s in jobs:
                cpu_in.append((q.int().cpu().numpy().reshape(-1),
                               gc.build_indexes(s).cpu().numpy().reshape(-1), cdf, size, off))
            for z in zs:
                sym = torch.round(z - med).int().cpu().numpy().reshape(-1)
                idx = np.repeat(np.arange(64, dtype=np.int32), z.size(2) * z.size(3))
                cpu_in.append((sym, idx, ecdf, esize, eoff))
            t0 = time.perf_counter()
            stock = [rans.encode_with_indexes(*a) for a in cpu_in]
            t_enc = time.perf_counter() - t0
            t0 = time.perf_counter()
            for s, a in zip(stock, cpu_in):
                rans.decode_with_indexes(s, a[1], *a[2:])
            t_dec = time.perf_counter() - t0
            res["cpu_c_oracle"] = {"encode_ms": 1e3 * t_enc, "decode_ms": 1e3 * t_dec,
                                   "bytes": sum(len(s) for s in stock), "threads": 1}
            res["bits_per_symbol"] = 8 * sum(len(s) for s in stock) / n_sym
            for S in (256, 1024, 4096, 16384, 65536, "auto"):
                def enc(S=S):
                    # both checkerboard passes of a model = one launch (as the product does);
                    # "auto" = the product's defaults: per-tensor sub-stream length
                    # (coder.auto_stream_symbols) and the z coder on the side stream
                    auto = S == "auto"
                    ps = [coder.rans_encode_async(eb._tables(), x=z, means=med.expand_as(z),
                                                  stream_symbols=None if auto else S, overlap=auto)
                          for z in zs]
                    ps = [coder.rans_encode_async(gc._tables(), x=q, scales=s,
                                                  scale_table=gc.scale_table,
                                                  stream_symbols=None if auto else S)
                          for q, s in pairs] + ps
                    for p in ps:
                        if p.done is not None:
                            torch.cuda.current_stream().wait_event(p.done)
                    return ps
                t_e = ev_time(enc)
                t0 = time.perf_counter()
                strings = coder.collect(enc())
                t_e2e = time.perf_counter() - t0
                nbytes = sum(len(b) for s in strings for b in s)
                strings = [[strings[0][0]], [strings[0][1]], [strings[1][0]], [strings[1][1]],
                           strings[2], strings[3]]
                # decode: time the launches only (streams pre-staged on the device is not
                # offered by the Python face, so this includes the small H2D of the strings)
                def dec():
                    for (name, q, s), st in zip(jobs, strings[:4]):
                        coder.rans_decode(st, gc._tables(), q.shape, scales=s,
                                          scale_table=gc.scale_table, want_symbols=True)
                    for z, st in zip(zs, strings[4:]):
                        coder.rans_decode(st, eb._tables(), z.shape, means=med)
                t_d = ev_time(dec, iters=10)
                res["sweep"].append({
                    "S": S, "encode_launch_ms": t_e, "encode_with_d2h_ms": 1e3 * t_e2e,
                    "decode_with_h2d_ms": t_d, "bytes": nbytes,
                    "overhead_vs_stock_pct": 100.0 * (nbytes - res["cpu_c_oracle"]["bytes"]) /
                    res["cpu_c_oracle"]["bytes"],
                    "Msym_per_s_encode": n_sym / t_e / 1e3})
            out["regimes"][regime] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
