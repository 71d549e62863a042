#!/usr/bin/env python
"""Real-bit-stream P-frame (dmc/test.py:185-196) of the reference's own ``DMC`` on one
B200: ``encode_inter`` / ``decode_inter`` of the unmodified reference over the CPU coder
(CompressAI's arithmetic, plain-C restatement -- what stock ``test.py --write_stream``
runs) vs the patched model with the GPU coder in each container layout.

Same weights (reference init x 0.7), same frames, 1088x1920 and 256x256, batch 1.
Wall-clock per call with a device synchronisation on both sides (the calls end in
host-side ``bytes`` / start from them).  Also records the coded bytes and that every arm
decodes to the same reconstruction.  Output: one JSON object (profiles/r02_codec_bench.json).
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dropin_util as du  # noqa: E402
from deepvideocodec_b200 import coder  # noqa: E402


def wall(fn, n, warm):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n, out


def n_bytes(enc):
    return sum(len(s) for key in ("motion", "frame") for grp in enc["strings"][key] for s in grp)


def main():
    dev = torch.device("cuda:0")
    if not du.reference_available():
        print(json.dumps({"unavailable": "baseline/_ref not staged (tools/stage_reference.py)"}))
        return
    torch.backends.cudnn.deterministic = True         # test.py:26; encoder and decoder must agree
    stock, patched = du.build_pair(dev, seed=0, weight_scale=0.7)
    stock.eval(), patched.eval()
    stock.update(force=True), patched.update(force=True)
    out = {"device": torch.cuda.get_device_name(0), "weights": "reference init (seed 0) x 0.7",
           "timed": "wall clock per call, device synchronised on both sides, batch 1", "sizes": {}}
    layouts = [("gpu_raw_stock_stream", dict(PINNED_STREAM_SYMBOLS=0)),
               ("gpu_dvc1_payload", dict(DEFAULT_LANES=1)),
               ("gpu_default (DVC3/DVS3, adaptive)", dict())]
    for h, w in ((256, 256), (1088, 1920)):
        f0, f1 = du.frames(2, 1, h, w, dev, seed=3)
        dpb = {"x_ref": f0, "feature_ref": None, "y_ref": None, "y_mv_ref": None}
        res = {}
        with torch.no_grad():
            x_fwd = patched([f0, f1])["x_hat"][0]
            iters = 10 if h <= 256 else 5
            # ---- stock: the reference's code over the CPU coder ---------------------------
            t_e, enc_s = wall(lambda: stock.encode_inter(f1, dpb), max(2, iters // 2), 1)
            t_d, dec_s = wall(lambda: stock.decode_inter(enc_s["strings"], enc_s["shape"], dpb),
                              max(2, iters // 2), 1)
            res["stock_cpu_coder"] = {"encode_ms": t_e, "decode_ms": t_d, "bytes": n_bytes(enc_s),
                                      "x_hat_max_abs_vs_forward": float((dec_s[0] - x_fwd).abs().max())}
            # ---- patched: GPU coder, each container ------------------------------------------
            for name, knobs in layouts:
                saved = {k: getattr(coder, k) for k in knobs}
                for k, v in knobs.items():
                    setattr(coder, k, v)
                try:
                    t_e, enc_p = wall(lambda: patched.encode_inter(f1, dpb), iters, 2)
                    t_d, dec_p = wall(lambda: patched.decode_inter(enc_p["strings"], enc_p["shape"], dpb),
                                      iters, 2)
                finally:
                    for k, v in saved.items():
                        setattr(coder, k, v)
                kinds = sorted({bytes(s[:4]).decode() if bytes(s[:4]) in (b"DVC1", b"DVC3", b"DVS3")
                                else "raw" for key in ("motion", "frame")
                                for grp in enc_p["strings"][key] for s in grp})
                res[name] = {"encode_ms": t_e, "decode_ms": t_d, "bytes": n_bytes(enc_p),
                             "containers": kinds,
                             "x_hat_max_abs_vs_forward": float((dec_p[0] - x_fwd).abs().max()),
                             "identical_to_stock_bytes": [bytes(x) for g in enc_p["strings"]["frame"] for x in g] ==
                             [bytes(x) for g in enc_s["strings"]["frame"] for x in g]}
            # the network half of the two calls, for scale: forward_inter does the same convs
            t_f, _ = wall(lambda: patched.forward_inter(f1, dpb), iters, 2)
            res["patched_forward_inter_ms"] = t_f
        out["sizes"][f"{h}x{w}"] = res
        print(f"{h}x{w}", json.dumps(res, indent=1), file=sys.stderr)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
