"""Gaussian-conditional / stage-B kernel: dense-NCHW path vs the strided path (DVC_GC_DENSE=0),
same tensors, CUDA events; config-5 size (192 x 136 x 240) and the 1080p frame model (96 x 68 x 120)."""
import json, math, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import json, math, sys, torch
sys.path.insert(0, %r)
import deepvideocodec_b200 as dvc
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
def t(fns, n=60, warm=8):
    for i in range(warm): fns[i %% len(fns)]()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fns[i %% len(fns)]()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
res = {}
gc = dvc.GaussianConditional(None).to(dev).eval()
with torch.no_grad():
    for (c, h, w) in ((192, 136, 240), (96, 68, 120)):
        sets = []
        for _ in range(4):
            mu = torch.randn(1, c, h, w, device=dev, generator=g) * 3
            sg = torch.exp(torch.empty(1, c, h, w, device=dev).uniform_(math.log(0.05), math.log(32), generator=g))
            y = mu + sg * torch.randn(1, c, h, w, device=dev, generator=g)
            pr = torch.randn(1, 2 * c, h, w, device=dev, generator=g).abs() + 0.2
            sets.append((y, sg, mu, pr))
        E = c * h * w
        us = t([(lambda s=s: gc(s[0], s[1], s[2])) for s in sets])
        res["gc_module %%dx%%dx%%d" %% (c, h, w)] = {"us": round(us, 2), "GBps": round(20 * E / us / 1e3, 1)}
        us = t([(lambda s=s: dvc.dual_prior_stage_b_gc(s[0], s[2], s[1], s[3], gc, False)) for s in sets])
        res["stage_b_gc %%dx%%dx%%d" %% (c, h, w)] = {"us": round(us, 2), "GBps": round(28 * E / us / 1e3, 1)}
print(json.dumps(res))
''' % ROOT
out = {}
for dense in ("1", "0"):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, DVC_GC_DENSE=dense), capture_output=True, text=True)
    try:
        out["dense" if dense == "1" else "strided"] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        out[dense] = r.stderr[-800:]
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gc_ab.json"), "w"), indent=1)
