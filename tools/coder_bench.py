"""Entropy-coder bench at the 1080p latent sizes of
BASELINE.json configs[1]: the six bit streams of one P-frame
(motion / frame: two checkerboard planes of y + the hyper-latent z).

For each sub-stream length S: CUDA-event time of the encode launches (kernel
only, streams stay on the device), of the decode launches, the coded size and
its container overhead; next to it the plain-C oracle coder (one host thread,
the arithmetic CompressAI runs on the CPU; its Python list marshalling is NOT
included, so this flatters the CPU side) on the same symbols.  Three regimes ("matched" = the sparse scales with symbols drawn from the model itself):
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
        for regime in ("synthetic", "sparse", "matched"):
            jobs, pairs = planes_of(inp, gc, regime != "synthetic")
            if regime == "matched":
                # symbols drawn from the model the coder uses for them (what a trained codec
                # produces): q = round(N(0, max(scale, 0.11))) on the sparse regime's scale planes
                g = torch.Generator(device=dev).manual_seed(5)
                for q, sc in pairs:
                    q.copy_(torch.round(torch.randn(q.shape, device=dev, generator=g) *
                                        sc.clamp_min(0.11)))
                jobs = [(f"{lab}.y{k}", pairs[m][0][k:k + 1], pairs[m][1][k:k + 1])
                        for m, lab in enumerate(("motion", "frame")) for k in (0, 1)]
            zs = [inp["motion.z"], inp["frame.z"]]
            if regime != "synthetic":
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
            # per-tensor payload estimates (what the likelihood kernel's sum ln p gives the
            # product for free): here the stock stream sizes of the same symbols, per pass
            est_pairs = [(len(stock[0]) + len(stock[1])) / 2.0, (len(stock[2]) + len(stock[3])) / 2.0]
            rows = [dict(label=f"dvc1_S{S}", S=S, lanes=1, skip=False, est=False)
                    for S in (256, 1024, 4096, 16384)]
            rows += [dict(label="dvc1_payload", S=None, lanes=1, skip=False, est=True)]
            rows += [dict(label=f"dvc3_S{S}", S=S, lanes=32, skip=False, est=False)
                     for S in (8192, 32768, 131072)]
            rows += [dict(label=f"dvs3_S{S}", S=S, lanes=32, skip=True, est=False)
                     for S in (8192, 32768, 131072)]
            rows += [dict(label="dvc3_payload", S=None, lanes=32, skip=False, est=True),
                     dict(label="dvs3_payload_forced", S=None, lanes=32, skip=True, est=True),
                     dict(label="adaptive_payload (product default)", S=None, lanes=32, skip=None, est=True)]
            for row in rows:
                def enc(row=row):
                    # both checkerboard passes of a model = one launch (as the product does);
                    # S = None: the product's policy (coder.auto_stream_symbols, payload-driven
                    # for the y tensors) and the z coder on the side stream
                    auto = row["S"] is None
                    ps = [coder.rans_encode_async(eb._tables(), x=z, means=med.expand_as(z),
                                                  stream_symbols=row["S"], overlap=auto,
                                                  lanes=row["lanes"], skip=row["skip"])
                          for z in zs]
                    ps = [coder.rans_encode_async(gc._tables(), x=q, scales=s,
                                                  scale_table=gc.scale_table,
                                                  stream_symbols=row["S"], lanes=row["lanes"],
                                                  skip=row["skip"],
                                                  est_bytes=est_pairs[k] if row["est"] else None)
                          for k, (q, s) in enumerate(pairs)] + ps
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
                # decode: the six launches one after the other, as a decoder has to run them
                # (streams pre-staged on the device is not offered by the Python face, so this
                # includes the small H2D of the strings)
                def dec(check=True):
                    outs, sts = [], []
                    for (name, q, s), st in zip(jobs, strings[:4]):
                        outs.append(coder.rans_decode(st, gc._tables(), q.shape, scales=s,
                                                      scale_table=gc.scale_table, want_symbols=True,
                                                      statuses=sts))
                    for z, st in zip(zs, strings[4:]):
                        outs.append(coder.rans_decode(st, eb._tables(), z.shape, means=med,
                                                      statuses=sts))
                    if check:
                        coder.check_decode_status(sts)   # one host read, as the context models do
                    return outs
                outs = dec()
                ok = all(torch.equal(o, q.int()) for o, (_, q, _) in zip(outs[:4], jobs)) and \
                    all(torch.equal(o, torch.round(z - med) + med) for o, z in zip(outs[4:], zs))
                t_d = ev_time(dec, iters=10)
                # host side of the six decode calls alone (wall clock, nothing waited for)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    dec(check=False)
                t_host = (time.perf_counter() - t0) / 5
                torch.cuda.synchronize()
                res["sweep"].append({
                    "layout": row["label"],
                    "container": [list(coder.container_of(st[0], q.numel())[1:])
                                  for (_, q, _), st in zip(jobs, strings[:4])],
                    "stream_symbols_y": [coder.stream_symbols_of(st[0], q.numel())
                                         for (_, q, _), st in zip(jobs, strings[:4])],
                    "round_trip_exact": bool(ok),
                    "encode_launch_ms": t_e, "encode_with_d2h_ms": 1e3 * t_e2e,
                    "decode_with_h2d_ms": t_d, "decode_host_ms": 1e3 * t_host, "bytes": nbytes,
                    "overhead_vs_stock_pct": 100.0 * (nbytes - res["cpu_c_oracle"]["bytes"]) /
                    res["cpu_c_oracle"]["bytes"],
                    "Msym_per_s_encode": n_sym / t_e / 1e3, "Msym_per_s_decode": n_sym / t_d / 1e3})
                r = res["sweep"][-1]
                print(f"  {regime:10s} {r['layout']:32s} S={r['stream_symbols_y'][0]:<7d} "
                      f"{''.join('S' if c[1] else '-' for c in r['container'])} enc {r['encode_launch_ms']:7.3f} ms  dec {r['decode_with_h2d_ms']:7.3f} ms (host {r['decode_host_ms']:5.2f})  "
                      f"{r['bytes']:8d} B ({r['overhead_vs_stock_pct']:+.2f} %)  ok={ok}",
                      file=sys.stderr)
            out["regimes"][regime] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
