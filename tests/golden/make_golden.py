"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What executes:
* ``flow_warp``, ``bilineardownsacling``            -> /root/reference/dmc/models/layers.py
* ``quantize_ste``                                  -> /root/reference/dmc/models/utils.py
* ``MotionContextModel.forward_dual_prior/get_mask``-> /root/reference/dmc/models/video_model.py
* ``collect_likelihoods_list``                      -> /root/reference/dmc/train.py:74-93
* ``GaussianConditional`` / ``EntropyBottleneck``   -> oracle/compressai shim (the
  real package is absent: these two vectors are *regression* vectors of the
  restatement, not reference truth -- "parity unpinned", see oracle/__init__.py)

All inputs are seeded; files are small (a few hundred KB in total).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.load_reference import load_reference_models, load_reference_train_fn  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def main():
    torch.set_num_threads(1)
    models = load_reference_models()
    layers = sys.modules["models.layers"]
    utils = sys.modules["models.utils"]
    vm = sys.modules["models.video_model"]
    collect = load_reference_train_fn("collect_likelihoods_list")

    # ---------------------------------------------------------------- warp
    g = torch.Generator().manual_seed(20261018)
    cases = {}
    for name, (n, c, h, w, amp) in {
        "rgb_small": (2, 3, 24, 40, 3.0),
        "c8_even": (1, 8, 32, 48, 6.0),
        "c5_odd": (1, 5, 17, 23, 2.5),
        "c64_oob": (1, 64, 8, 12, 40.0),       # most taps clamp to the border
    }.items():
        im = torch.randn(n, c, h, w, generator=g)
        flow = torch.randn(n, 2, h, w, generator=g) * amp
        layers.backward_grid[-1].clear()       # reference cache is keyed by shape only
        out = layers.flow_warp(im, flow)
        cases[name] = (im, flow, out)
    # exact-integer and zero flows
    im = torch.randn(1, 4, 16, 20, generator=g)
    flow = torch.zeros(1, 2, 16, 20)
    flow[:, 0] = 3.0
    flow[:, 1] = -2.0
    layers.backward_grid[-1].clear()
    cases["integer_shift"] = (im, flow, layers.flow_warp(im, flow))
    np.savez_compressed(
        os.path.join(HERE, "warp.npz"),
        **{f"{k}.{f}": _np(v) for k, t in cases.items() for f, v in zip(("im", "flow", "out"), t)})

    # ------------------------------------------------------------- pyramid
    pyr = {}
    for name, (n, h, w) in {"even": (2, 32, 48), "odd": (1, 19, 27), "by4": (1, 64, 128)}.items():
        mv = torch.randn(n, 2, h, w, generator=g) * 5
        mv2 = layers.bilineardownsacling(mv) / 2
        mv3 = layers.bilineardownsacling(mv2) / 2
        pyr[name] = (mv, mv2, mv3)
    np.savez_compressed(
        os.path.join(HERE, "pyramid.npz"),
        **{f"{k}.{f}": _np(v) for k, t in pyr.items() for f, v in zip(("mv", "mv2", "mv3"), t)})

    # ------------------------------------------------------------ quantise
    x = torch.cat([torch.arange(-6, 7).float() / 2,                 # exact ties
                   torch.randn(200, generator=g) * 4,
                   torch.tensor([1e-8, -1e-8, 8388607.5, -8388608.5, 0.49999997, -0.49999997])])
    np.savez_compressed(os.path.join(HERE, "quantize.npz"), x=_np(x), q=_np(utils.quantize_ste(x)))

    # ---------------------------------------------------------- dual prior
    torch.manual_seed(7)
    ch = 8
    mcm = vm.MotionContextModel(ch_mv=ch).eval()
    n, h, w = 2, 6, 10
    y = torch.randn(n, ch, h, w, generator=g) * 4
    means = torch.randn(n, ch, h, w, generator=g) * 3
    scales = torch.exp(torch.empty(n, ch, h, w).uniform_(np.log(0.05), np.log(32), generator=g))
    scales[0, 0, 0, :4] = torch.tensor([0.0, -1.0, 0.11, 0.109999])   # raw conv outputs can be <= 0
    m0, m1 = mcm.get_mask(h, w, y.device)
    with torch.no_grad():
        y0, y1 = y.chunk(2, 1)
        mu0, mu1 = means.chunk(2, 1)
        s0, s1 = scales.chunk(2, 1)
        a00 = mcm.process_with_mask(y0, mu0, s0, m0)
        a11 = mcm.process_with_mask(y1, mu1, s1, m1)
        params = torch.cat((a00[1], a11[1], means, scales), dim=1)
        prior_out = mcm.y_spatial_prior(params)
        y_hat, means_hat, scales_hat = mcm.forward_dual_prior(y, means, scales)
        c = mcm.forward_dual_prior(y, means, scales, mode="compress")
        _, y_lik = mcm.gaussian_conditional(y, scales_hat, means_hat)      # shim (unpinned)
    np.savez_compressed(
        os.path.join(HERE, "dual_prior.npz"),
        y=_np(y), means=_np(means), scales=_np(scales), mask0=_np(m0), mask1=_np(m1),
        params=_np(params), prior_out=_np(prior_out),
        y_hat=_np(y_hat), means_hat=_np(means_hat), scales_hat=_np(scales_hat),
        c_y_hat=_np(c[0]), c_q_w0=_np(c[1]), c_q_w1=_np(c[2]), c_s_w0=_np(c[3]), c_s_w1=_np(c[4]),
        y_lik_shim=_np(y_lik))

    # ------------------------------------------ entropy models (shim, unpinned)
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional
    torch.manual_seed(11)
    eb = EntropyBottleneck(6).eval()
    with torch.no_grad():       # move parameters away from the init point
        for name, p in eb.named_parameters():
            if name.startswith("_factor"):
                p.uniform_(-0.8, 0.8)
            elif name.startswith("_matrix"):
                p.add_(torch.randn_like(p) * 0.3)
            elif name == "quantiles":
                p[:, 0, 1] = torch.randn(6) * 2
    z = torch.randn(2, 6, 5, 7, generator=g) * 10
    with torch.no_grad():
        z_out, z_lik = eb(z)
        z_hat = utils.quantize_ste(z - eb._get_medians()) + eb._get_medians()
        aux = eb.loss()
    gc = GaussianConditional(None).eval()
    gy = torch.randn(2, 6, 8, 10, generator=g) * 5
    gmu = torch.randn(2, 6, 8, 10, generator=g) * 3
    gs = torch.exp(torch.empty(2, 6, 8, 10).uniform_(np.log(0.05), np.log(64), generator=g))
    with torch.no_grad():
        gy_out, gy_lik = gc(gy, gs, gmu)
    np.savez_compressed(
        os.path.join(HERE, "entropy_shim.npz"),
        **{f"eb.{k}": _np(v) for k, v in eb.state_dict().items() if v.numel()},
        z=_np(z), z_out=_np(z_out), z_lik=_np(z_lik), z_hat=_np(z_hat), aux_loss=_np(aux),
        gy=_np(gy), gmu=_np(gmu), gs=_np(gs), gy_out=_np(gy_out), gy_lik=_np(gy_lik))

    # ---------------------------------------------------------------- rate
    liks = []
    for i in range(2):
        liks.append({
            "motion": {"y": torch.rand(2, 8, 6, 10, generator=g).clamp_min(1e-9),
                       "z": torch.rand(2, 4, 2, 3, generator=g).clamp_min(1e-9)},
            "frame": {"y": torch.rand(2, 12, 6, 10, generator=g).clamp_min(1e-9),
                      "z": torch.rand(2, 4, 2, 3, generator=g).clamp_min(1e-9)}})
    num_pixels = 96 * 160 * 2
    bpp, info = collect(liks, num_pixels)
    flat = {}
    for i, fr in enumerate(liks):
        for label, d in fr.items():
            for field, v in d.items():
                flat[f"lik.{i}.{label}.{field}"] = _np(v)
    np.savez_compressed(os.path.join(HERE, "rate.npz"), num_pixels=np.int64(num_pixels),
                        bpp_loss=_np(bpp), **flat,
                        **{f"info.{k}": _np(torch.as_tensor(v)) for k, v in info.items()})
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
