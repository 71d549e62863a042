"""CPU: the oracle restatement vs the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import dmc_ref

ORACLE_DIR = os.path.dirname(os.path.abspath(dmc_ref.__file__))


def _oem():
    import importlib
    import importlib.util
    name = "oracle_compressai"
    if name + ".entropy_models" not in sys.modules:
        pkg_dir = os.path.join(ORACLE_DIR, "compressai")
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules[name] = pkg
        spec.loader.exec_module(pkg)
    return importlib.import_module(name + ".entropy_models")


def _t(z, k):
    return torch.from_numpy(z[k])


def test_warp_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "warp.npz"))
    names = sorted({k.split(".")[0] for k in z.files})
    assert set(names) >= {"rgb_small", "c8_even", "c5_odd", "c64_oob", "integer_shift"}
    for name in names:
        out = dmc_ref.flow_warp(_t(z, f"{name}.im"), _t(z, f"{name}.flow"))
        assert torch.equal(out, _t(z, f"{name}.out")), name     # same ATen CPU kernels: bit exact


def test_pyramid_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "pyramid.npz"))
    for name in sorted({k.split(".")[0] for k in z.files}):
        mv2, mv3 = dmc_ref.flow_pyramid(_t(z, f"{name}.mv"))
        assert torch.equal(mv2, _t(z, f"{name}.mv2")) and torch.equal(mv3, _t(z, f"{name}.mv3"))


def test_even_pyramid_is_2x2_mean(golden_dir):
    """for even sizes the bilinear 2x downscale is a 2x2 mean.
    Measured here: ATen-CPU accumulates ((a+b)+c)+d, ATen-CUDA (the device the
    kernels replay, checked bit-exactly in tests/test_gpu_warp.py) pairs
    (a+b)+(c+d); the two differ by at most one ulp of the sum."""
    z = np.load(os.path.join(golden_dir, "pyramid.npz"))
    mv = _t(z, "even.mv")
    a, b = mv[:, :, 0::2, 0::2], mv[:, :, 0::2, 1::2]
    c, d = mv[:, :, 1::2, 0::2], mv[:, :, 1::2, 1::2]
    assert torch.equal((((a + b) + c) + d) * 0.25 * 0.5, _t(z, "even.mv2"))
    cuda_order = ((a + b) + (c + d)) * 0.25 * 0.5
    assert (cuda_order - _t(z, "even.mv2")).abs().max().item() <= 5e-7


def test_quantize_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "quantize.npz"))
    q = dmc_ref.quantize_ste(_t(z, "x"))
    assert torch.equal(q, _t(z, "q"))
    # ties go to the even integer: -3 -2.5 ... 3 -> -3 -2 -2 -2 -1 0 0 0 1 2 2 2 3
    assert torch.equal(dmc_ref.quantize_ste(torch.arange(-6, 9).float() / 2 - 0.0)[:13],
                       torch.round(torch.arange(-6, 7).float() / 2))
    assert torch.equal(torch.round(torch.tensor([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5])),
                       torch.tensor([-2., -2., -0., 0., 2., 2.]))


def test_dual_prior_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "dual_prior.npz"))
    y, mu, sg = _t(z, "y"), _t(z, "means"), _t(z, "scales")
    m0, m1 = dmc_ref.checkerboard_masks(y.size(2), y.size(3), y.device)
    assert torch.equal(m0, _t(z, "mask0")) and torch.equal(m1, _t(z, "mask1"))
    params, _, _ = dmc_ref.dual_prior_stage_a(y, mu, sg)
    assert torch.equal(params, _t(z, "params"))
    prior = _t(z, "prior_out")
    y_hat, mh, sh = dmc_ref.dual_prior(y, mu, sg, lambda p: prior)
    assert torch.equal(y_hat, _t(z, "y_hat"))
    assert torch.equal(mh, _t(z, "means_hat")) and torch.equal(sh, _t(z, "scales_hat"))
    c = dmc_ref.dual_prior(y, mu, sg, lambda p: prior, mode="compress")
    for got, key in zip(c, ("c_y_hat", "c_q_w0", "c_q_w1", "c_s_w0", "c_s_w1")):
        assert torch.equal(got, _t(z, key)), key
    _, lik = _oem().GaussianConditional(None).eval()(y, sh, mh)
    assert torch.equal(lik, _t(z, "y_lik_shim"))


def test_entropy_shim_regression(golden_dir):
    """Regression vectors of the CompressAI restatement (parity unpinned)."""
    oem = _oem()
    z = np.load(os.path.join(golden_dir, "entropy_shim.npz"))
    eb = oem.EntropyBottleneck(6).eval()
    sd = {k[len("eb."):]: _t(z, k) for k in z.files if k.startswith("eb.")}
    eb.load_state_dict(sd, strict=False)
    with torch.no_grad():
        out, lik = eb(_t(z, "z"))
        assert torch.equal(out, _t(z, "z_out"))
        assert torch.allclose(lik, _t(z, "z_lik"), rtol=1e-6, atol=0)
        assert torch.equal(dmc_ref.quantize_hyper(_t(z, "z"), eb._get_medians()), _t(z, "z_hat"))
        assert torch.allclose(eb.loss(), _t(z, "aux_loss"), rtol=1e-6)
        out, lik = oem.GaussianConditional(None).eval()(_t(z, "gy"), _t(z, "gs"), _t(z, "gmu"))
    assert torch.equal(out, _t(z, "gy_out"))
    assert torch.equal(lik, _t(z, "gy_lik"))


def test_rate_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "rate.npz"))
    liks = [{label: {f: _t(z, f"lik.{i}.{label}.{f}") for f in ("y", "z")}
             for label in ("motion", "frame")} for i in range(2)]
    bpp, info = dmc_ref.collect_likelihoods_list(liks, int(z["num_pixels"]))
    assert torch.equal(bpp, _t(z, "bpp_loss"))
    keys = [k[len("info."):] for k in z.files if k.startswith("info.")]
    assert sorted(info) == sorted(keys)
    for k in keys:
        assert torch.equal(torch.as_tensor(info[k]), _t(z, "info." + k)), k


def test_model_capture_pins_the_oracle(golden_dir):
    """Model-driven regime: every hot-path tensor recorded while the
    stock reference DMC (random init) coded three 64x64 frames on the CPU
    (tests/golden/make_golden_model.py).  The oracle restatement must reproduce
    each of them bit for bit from the recorded inputs."""
    z = np.load(os.path.join(golden_dir, "model_capture.npz"))
    oem = _oem()
    gc = oem.GaussianConditional(None).eval()
    n_checked = 0
    for f in range(2):
        for k in range(4):                                   # x_ref, feature 1..3
            out = dmc_ref.flow_warp(_t(z, f"f{f}.warp{k}.im"), _t(z, f"f{f}.warp{k}.flow"))
            assert torch.equal(out, _t(z, f"f{f}.warp{k}.out")), (f, k)
        for k in range(2):
            assert torch.equal(dmc_ref.bilinear_down2(_t(z, f"f{f}.down{k}.in")),
                               _t(z, f"f{f}.down{k}.out")), (f, k)
        # the recorded pyramid is what the recorded warps used (video_model.py:499-504)
        mv2, mv3 = dmc_ref.flow_pyramid(_t(z, f"f{f}.down0.in"))
        assert torch.equal(mv2, _t(z, f"f{f}.warp2.flow")) and torch.equal(mv3, _t(z, f"f{f}.warp3.flow"))
        for label in ("motion", "frame"):
            p = f"f{f}.{label}."
            y, mu, sg, prior = (_t(z, p + s) for s in ("y", "means", "scales", "prior"))
            y_hat, mh, sh = dmc_ref.dual_prior(y, mu, sg, lambda _params: prior)
            assert torch.equal(y_hat, _t(z, p + "y_hat"))
            assert torch.equal(mh, _t(z, p + "means_hat")) and torch.equal(sh, _t(z, p + "scales_hat"))
            with torch.no_grad():
                _, lik = gc(y, sh, means=mh)
            assert torch.equal(lik, _t(z, p + "y_lik"))
            eb = oem.EntropyBottleneck(y.size(1) if label == "motion" else 64).eval()
            eb.load_state_dict({k[len(f"eb.{label}."):]: _t(z, k) for k in z.files
                                if k.startswith(f"eb.{label}.")})
            with torch.no_grad():
                z_out, z_lik = eb(_t(z, p + "z"))
                z_hat = dmc_ref.quantize_hyper(_t(z, p + "z"), eb._get_medians())
            assert torch.equal(z_out, _t(z, p + "z_out")) and torch.equal(z_lik, _t(z, p + "z_lik"))
            assert torch.equal(z_hat, _t(z, p + "z_hat"))
            n_checked += 1
    assert n_checked == 4
    liks = [{label: {"y": _t(z, f"f{f}.{label}.y_lik"), "z": _t(z, f"f{f}.{label}.z_lik")}
             for label in ("motion", "frame")} for f in range(2)]
    bpp, info = dmc_ref.collect_likelihoods_list(liks, int(z["num_pixels"]))
    assert torch.equal(bpp, _t(z, "bpp_loss"))
    for k, v in info.items():
        assert float(v) == float(z["info." + k]), k
    # the regime this fixture exists for: most likelihoods sit on the 1e-9 floor
    floor = np.mean([np.mean(z[f"f{f}.{m}.y_lik"] <= 1e-9) for f in range(2) for m in ("motion", "frame")])
    assert floor > 0.5
