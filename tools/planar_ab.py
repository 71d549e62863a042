"""NCHW flow_warp: smem-staged planar kernel vs the strided path (DVC_WARP_PLANAR=0), same box."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
import deepvideocodec_b200 as dvc
from oracle import dmc_ref
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
def smooth(h, w, sigma=4.0, k=31):
    f = torch.randn(1, 2, h, w, device=dev, generator=g)
    f = torch.nn.functional.avg_pool2d(f, k, stride=1, padding=k // 2, count_include_pad=False)
    return (f / f.std() * sigma).contiguous()
def t(fn, n=30, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
res = {}
with torch.no_grad():
    for (c, h, w) in ((64, 1088, 1920), (64, 544, 960), (64, 2160, 3840)):
        ims = [torch.randn(1, c, h, w, device=dev, generator=g) for _ in range(3)]
        for name, flow in (("smooth4", smooth(h, w)), ("gentle", smooth(h, w, 2.0, 127)), ("iid16", torch.randn(1, 2, h, w, device=dev, generator=g) * 16)):
            us = t(lambda i: dvc.flow_warp(ims[i %% 3], flow))
            ok = bool(torch.equal(dvc.flow_warp(ims[0], flow), dmc_ref.flow_warp(ims[0], flow))) if h <= 1088 else None
            mb = 4.0 * h * w * (2 * c + 2) / 1e6
            res["%%dx%%dx%%d %%s" %% (c, h, w, name)] = {"us": round(us, 1), "GBps": round(mb / us * 1e3, 0), "bit_identical_to_eager": ok}
        del ims
print(json.dumps(res))
''' % ROOT
out = {}
# usage: planar_ab.py [libA.so libB.so ...]  (A/B builds next to libdvc_b200.so); default: shipped
# library, planar vs strided
# an argument may carry environment settings: libX.so:VAR=1:VAR2=0
variants = [("1", a) for a in sys.argv[1:]] or [("1", None), ("0", None)]
for planar, spec in variants:
    lib, extra = None, {}
    if spec:
        parts = spec.split(":")
        lib = parts[0] or None
        extra = dict(kv.split("=", 1) for kv in parts[1:])
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DVC_WARP_PLANAR=planar, **extra, **({"DVC_B200_LIB": os.path.join(ROOT, "deepvideocodec_b200", lib)} if lib else {})), capture_output=True, text=True)
    key = ("planar" if planar == "1" else "strided") + (spec or "")
    try:
        out[key] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        out[key] = r.stderr[-800:]
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "planar_ab.json"), "w"), indent=1)
