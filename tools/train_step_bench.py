"""BASELINE.json configs[3] (training shapes): forward + backward of the hot-path ops
of ONE P-frame at batch 8, 256x256 crops, noise-quantisation likelihoods --
ours vs the same graph in PyTorch-CUDA eager (oracle ops), with loss and
gradient parity on shared noise (same torch seed).  Convs are outside the path:
a single 3x3 conv stands in for the spatial prior so gradients flow through both
dual-prior stages.  Prints one JSON object."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deepvideocodec_b200 as dvc  # noqa: E402
from deepvideocodec_b200 import context as ctxmod  # noqa: E402
from oracle import dmc_ref  # noqa: E402
from test_gpu_entropy import _oracle_entropy_models  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
N, H, W = 8, 256, 256
oem = _oracle_entropy_models()
g = torch.Generator(device=dev).manual_seed(11)
cl = torch.channels_last


def randn(*s):
    return torch.randn(*s, device=dev, generator=g)


x_ref = torch.rand(N, 3, H, W, device=dev, generator=g)
feats = [randn(N, 64, H >> k, W >> k).contiguous(memory_format=cl) for k in range(3)]
f = torch.nn.functional.avg_pool2d(randn(N, 2, H, W), 15, stride=1, padding=7, count_include_pad=False)
mv = (f / f.std() * 2.0).contiguous()
lat = {}
for name, c in (("motion", 64), ("frame", 96)):
    mu = randn(N, c, H // 16, W // 16) * 3
    sg = torch.exp(torch.empty(N, c, H // 16, W // 16, device=dev).uniform_(math.log(0.05), math.log(32), generator=g))
    lat[name] = (mu + sg * randn(N, c, H // 16, W // 16), mu, sg, randn(N, 64, H // 64, W // 64) * 5)
torch.manual_seed(5)
convs = {k: torch.nn.Conv2d(3 * c, 2 * c, 3, padding=1).to(dev) for k, c in (("motion", 64), ("frame", 96))}
torch.manual_seed(6)
eb_ref = {k: oem.EntropyBottleneck(64).to(dev).train() for k in ("motion", "frame")}
eb_mine = {k: dvc.EntropyBottleneck(64).to(dev).train() for k in ("motion", "frame")}
for k in eb_ref:
    eb_mine[k].load_state_dict(eb_ref[k].state_dict())
gc_ref = oem.GaussianConditional(None).to(dev).train()
gc_mine = dvc.GaussianConditional(None).to(dev).train()
num_pixels = H * W * 6


class Holder:
    pass


def step(impl, leaves):
    x, f1, f2, f3, m = leaves["x_ref"], leaves["f1"], leaves["f2"], leaves["f3"], leaves["mv"]
    liks = {}
    if impl == "ref":
        outs = dmc_ref.motion_compensation_warps(x, f1, f2, f3, m)
        for k in ("motion", "frame"):
            y, mu, sg, z = leaves[k]
            _, z_lik = eb_ref[k](z)
            z_hat = dmc_ref.quantize_hyper(z, eb_ref[k]._get_medians())
            y_hat, mh, sh = dmc_ref.dual_prior(y, mu, sg, convs[k])
            _, y_lik = gc_ref(y, sh, mh)
            liks[k] = {"y": y_lik, "z": z_lik, "_yh": y_hat, "_zh": z_hat}
        collect = dmc_ref.collect_likelihoods_list
    else:
        outs = dvc.motion_compensation_warps(x, f1, f2, f3, m)
        for k in ("motion", "frame"):
            y, mu, sg, z = leaves[k]
            _, z_hat, z_lik = dvc.entropy_models.eb_forward(eb_mine[k], z, want_outputs=False, want_zhat=True)
            hold = Holder()
            hold.y_spatial_prior = convs[k]
            hold.gaussian_conditional = gc_mine
            y_hat, lk = ctxmod._context_tail(hold, y, mu, sg, z_lik)
            liks[k] = {"y": lk["y"], "z": lk["z"], "_yh": y_hat, "_zh": z_hat}
        collect = dvc.collect_likelihoods_list
    aux = sum((o * o).mean() for o in outs) + sum((v["_yh"] ** 2).mean() + (v["_zh"] ** 2).mean() for v in liks.values())
    bpp, _ = collect([{k: {"y": v["y"], "z": v["z"]} for k, v in liks.items()}], num_pixels)
    loss = bpp.mean() + 1e-3 * aux
    loss.backward()
    return loss.detach()


def make_leaves():
    lv = {"x_ref": x_ref.clone().requires_grad_(True), "mv": mv.clone().requires_grad_(True)}
    for i, ft in enumerate(feats):
        lv[f"f{i + 1}"] = ft.clone().requires_grad_(True)
    for k, t in lat.items():
        lv[k] = tuple(v.clone().requires_grad_(True) for v in t)
    return lv


def clear():
    for c in convs.values():
        c.zero_grad()
    for d in (eb_ref, eb_mine):
        for m in d.values():
            m.zero_grad()


res = {}
grads = {}
for impl in ("ref", "dvc"):
    clear()
    lv = make_leaves()
    torch.manual_seed(99)
    loss = step(impl, lv)
    grads[impl] = {"mv": lv["mv"].grad, "f1": lv["f1"].grad, "x_ref": lv["x_ref"].grad,
                   "y_frame": lv["frame"][0].grad, "scales_frame": lv["frame"][2].grad,
                   "z_motion": lv["motion"][3].grad,
                   "conv_w": convs["frame"].weight.grad.clone(),
                   "eb_matrix1": (eb_ref if impl == "ref" else eb_mine)["motion"]._matrix1.grad.clone()}
    res[impl + "_loss"] = float(loss)
    for _ in range(3):
        clear(); step(impl, make_leaves())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    lvs = [make_leaves() for _ in range(iters)]
    torch.cuda.synchronize()
    a.record()
    for i in range(iters):
        step(impl, lvs[i])
    b.record()
    torch.cuda.synchronize()
    res[impl + "_ms"] = a.elapsed_time(b) / iters
res["loss_rel_err"] = abs(res["dvc_loss"] - res["ref_loss"]) / abs(res["ref_loss"])
res["grad_rel_err"] = {k: float((grads["dvc"][k] - grads["ref"][k]).abs().max() / grads["ref"][k].abs().max())
                       for k in grads["ref"]}
res["speedup_fwd_bwd"] = res["ref_ms"] / res["dvc_ms"]
res["config"] = "batch 8 x 256x256, one P-frame of hot-path ops fwd+bwd, noise likelihoods (train mode)"
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "train_step.json"), "w"), indent=1)
