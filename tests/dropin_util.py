"""Shared by tests/test_gpu_dropin.py and tools/dropin_bench.py: run the
reference's own ``DMC`` twice on one device -- stock (unmodified reference over
eager PyTorch ops) and patched (same files, ``deepvideocodec_b200.patch``
applied) -- on identical weights and inputs, and compare what the reference's
callers consume (``dmc/train.py:162-211``, ``dmc/test.py:185-196``).
TEST INFRASTRUCTURE (imports ``oracle/``)."""
import contextlib
import math

import torch

from oracle.load_reference import (load_reference_train_fn, load_stock_and_patched,
                                   reference_available)

__all__ = ["reference_available", "build_pair", "frames", "deterministic_convs", "run_forward",
           "compare_forward", "rel_err", "stock_collect", "capture_latents"]

_pair = {}


@contextlib.contextmanager
def deterministic_convs():
    """Convolutions must not add noise of their own (SURVEY.md 4 "Module
    drop-in"): fp32 (no TF32), deterministic algorithms, no autotuning."""
    cd = torch.backends.cudnn
    saved = (cd.allow_tf32, cd.deterministic, cd.benchmark, torch.backends.cuda.matmul.allow_tf32)
    cd.allow_tf32, cd.deterministic, cd.benchmark = False, True, False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        cd.allow_tf32, cd.deterministic, cd.benchmark, torch.backends.cuda.matmul.allow_tf32 = saved


def build_pair(device, seed=0, channels_last=False, weight_scale=1.0):
    """(stock DMC, patched DMC) with identical parameters (random init, seed
    ``seed``: the only weights that exist -- test.py:113 points at a private
    checkpoint).

    ``weight_scale``: the reference's own init (xavier, gain sqrt 2,
    video_model.py:508-513) is numerically degenerate -- activations grow from
    frame to frame (|x_hat| ~ 1e5 after one P-frame, ~1e9 after two, > 50 % of
    the likelihoods on the 1e-9 floor).  Scaling every conv weight by 0.7 after
    that init gives O(1) activations and latents (a trained model's regime);
    both regimes are tested."""
    key = (str(device), seed, channels_last, weight_scale)
    if key in _pair:
        return _pair[key]
    stock_pkg, patched_pkg = load_stock_and_patched()
    torch.manual_seed(seed)
    stock = stock_pkg.DMC()
    if weight_scale != 1.0:
        with torch.no_grad():
            for mod in stock.modules():
                if isinstance(mod, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                    mod.weight.mul_(weight_scale)
    patched = patched_pkg.DMC()
    patched.load_state_dict(stock.state_dict())
    stock, patched = stock.to(device), patched.to(device)
    if channels_last:
        stock = stock.to(memory_format=torch.channels_last)
        patched = patched.to(memory_format=torch.channels_last)
    _pair.clear()                       # one pair resident at a time
    _pair[key] = (stock, patched)
    return stock, patched


def frames(n_frames, batch, h, w, device, seed=0, channels_last=False):
    """Synthetic video: a smooth random image translated by a couple of pixels
    per frame plus noise, so SpyNet sees real motion (config 1 / 2 use
    ``torch.rand`` frames; pure noise has no motion to estimate)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(batch, 3, h // 8 + 8, w // 8 + 8, generator=g)
    big = torch.nn.functional.interpolate(base, size=(h + 64, w + 64), mode="bicubic",
                                          align_corners=False).clamp(0, 1)
    out = []
    for t in range(n_frames):
        dx, dy = 3 * t, 2 * t
        f = big[:, :, 16 + dy:16 + dy + h, 16 + dx:16 + dx + w]
        f = (f + 0.02 * torch.randn(f.shape, generator=g)).clamp(0, 1).contiguous()
        f = f.to(device)
        if channels_last:
            f = f.contiguous(memory_format=torch.channels_last)
        out.append(f)
    return out


def capture_latents(model):
    """Forward hooks recording the quantised latents ``y_hat`` of both context
    models (first return value of their ``forward``)."""
    store = {"motion": [], "frame": [], "motion.z": [], "frame.z": []}
    hooks = [
        model.motion_context_model.register_forward_hook(
            lambda m, i, o: store["motion"].append(o[0].detach())),
        model.frame_context_model.register_forward_hook(
            lambda m, i, o: store["frame"].append(o[0].detach())),
    ]
    # hyper-latents z: input of hyper_decoder is z_hat; z itself is the hyper-encoder output
    hooks += [
        model.motion_context_model.hyper_encoder.register_forward_hook(
            lambda m, i, o: store["motion.z"].append(o.detach())),
        model.frame_context_model.hyper_encoder.register_forward_hook(
            lambda m, i, o: store["frame.z"].append(o.detach())),
    ]
    return store, hooks


def eb_likelihood_fp64(eb, z):
    """The factorised-bottleneck likelihood of ``z`` evaluated in fp64 with the
    eager restatement (the value both fp32 implementations approximate)."""
    import copy
    eb64 = copy.deepcopy(eb).double().eval()
    with torch.no_grad():
        return eb64(z.double())[1]


def run_forward(model, fr, seed=None, grad=False):
    store, hooks = capture_latents(model)
    if seed is not None:
        torch.manual_seed(seed)           # training-mode noise comes from torch's generator
    try:
        with contextlib.nullcontext() if grad else torch.no_grad():
            out = model(list(fr))
    finally:
        for h in hooks:
            h.remove()
    return out, store


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs() / b.abs().clamp_min(1e-300)).max().item()


_stock_collect = None


def stock_collect():
    """The reference's own ``collect_likelihoods_list`` (train.py:74-93)."""
    global _stock_collect
    if _stock_collect is None:
        _stock_collect = load_reference_train_fn("collect_likelihoods_list")
    return _stock_collect


def compare_forward(out_s, lat_s, out_p, lat_p, num_pixels, stock=None):
    """Error summary of one stock-vs-patched ``DMC.forward``.  With ``stock``
    (eval mode only) the z likelihoods of both arms are also compared with the
    fp64 evaluation of the same formula: eager's own fp32 rounding order in the
    bottleneck's last matmul changes with the tensor size
    (profiles/r02_eb_probe.json), so at small sizes neither arm is "the" fp32
    result and the distance to fp64 is the meaningful yardstick."""
    import deepvideocodec_b200 as dvc
    rep = {"frames": []}
    for i, (xs, xp) in enumerate(zip(out_s["x_hat"], out_p["x_hat"])):
        fr = {"x_hat_max_abs": (xs - xp).abs().max().item(),
              "x_hat_scale": xs.abs().max().item(),
              "x_hat_finite": bool(torch.isfinite(xs).all() and torch.isfinite(xp).all())}
        for label in ("motion", "frame"):
            a, b = lat_s[label][i], lat_p[label][i]
            fr[f"{label}.y_hat_equal"] = bool(torch.equal(a, b))
            fr[f"{label}.y_hat_mismatch"] = int((a != b).sum().item())
            for field in ("y", "z"):
                ls = out_s["likelihoods"][i][label][field]
                lp = out_p["likelihoods"][i][label][field]
                fr[f"{label}.{field}_lik_max_rel"] = rel_err(lp, ls)
                fr[f"{label}.{field}_floor_frac"] = (ls <= 1e-9).float().mean().item()
            if stock is not None:
                cm = stock.motion_context_model if label == "motion" else stock.frame_context_model
                z = lat_s[f"{label}.z"][i]
                fr[f"{label}.z_equal_inputs"] = bool(torch.equal(z, lat_p[f"{label}.z"][i]))
                l64 = eb_likelihood_fp64(cm.entropy_bottleneck, z)
                fr[f"{label}.z_lik_stock_vs_fp64"] = rel_err(out_s["likelihoods"][i][label]["z"], l64)
                fr[f"{label}.z_lik_patched_vs_fp64"] = rel_err(out_p["likelihoods"][i][label]["z"], l64)
        rep["frames"].append(fr)
    bs, ds = stock_collect()(out_s["likelihoods"], num_pixels)
    bp, dp = dvc.collect_likelihoods_list(out_p["likelihoods"], num_pixels)
    rep["bpp_stock"] = [float(v) for v in bs.reshape(-1).tolist()]
    rep["bpp_patched"] = [float(v) for v in bp.reshape(-1).tolist()]
    rep["bpp_max_rel"] = rel_err(bp, bs)
    rep["detail_keys_equal"] = list(ds) == list(dp)
    rep["detail_max_rel"] = max(
        abs(float(dp[k]) - float(ds[k])) / max(abs(float(ds[k])), 1e-300) for k in ds)
    rep["bits_per_frame_stock"] = [
        float(ds[f"bpp_loss.{i}"]) * num_pixels for i in range(len(out_s["x_hat"]))]
    assert math.isfinite(rep["bpp_max_rel"])
    return rep
