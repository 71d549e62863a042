"""GPU: the reference's OWN ``DMC`` (unmodified files, staged by
``tools/stage_reference.py`` into the git-ignored ``baseline/_ref/``) run
twice on the B200 -- stock over eager PyTorch ops, and with
``deepvideocodec_b200.patch`` applied -- on identical weights and inputs
(SURVEY.md 4 "Module drop-in").

What the reference's callers consume is compared at north_star's tolerances:
``x_hat`` (1e-5 abs), the four likelihood tensors per frame (1e-5 rel), rounded
latents (bit exact), bits per frame (1e-4 rel), ``aux_loss``, the training
loss and gradients, and ``encode_inter`` -> ``decode_inter`` (dmc/test.py:185-196).

Follows ``video_model.py:515-614``, ``train.py:162-211, 285-346``.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dropin_util as du  # noqa: E402

pytestmark = [pytest.mark.gpu]

LOUD = ("reference sources not staged: run `python tools/stage_reference.py` in the build "
        "container (copies /root/reference/dmc into the git-ignored baseline/_ref/) -- "
        "THE DROP-IN TESTS DID NOT RUN")


@pytest.fixture(scope="module", autouse=True)
def _conv_settings():
    with du.deterministic_convs():
        yield


TAMED = 0.7      # conv weights x 0.7 after the reference's init: O(1) activations (dropin_util)


def _pair(cuda_dev, channels_last=False, weight_scale=TAMED):
    if not du.reference_available():
        pytest.skip(LOUD)
    return du.build_pair(cuda_dev, seed=0, channels_last=channels_last,
                         weight_scale=weight_scale)


def _check_forward(rep, n_frames, small_frame=False):
    assert len(rep["frames"]) == n_frames
    for fr in rep["frames"]:
        assert fr["x_hat_finite"]
        assert fr["motion.y_hat_equal"] and fr["frame.y_hat_equal"], fr      # bit exact
        # 1e-5 abs for O(1) frames; relative to the frame's scale where the reference's
        # own init makes |x_hat| explode
        assert fr["x_hat_max_abs"] <= 1e-5 * max(1.0, fr["x_hat_scale"]), fr
        for label in ("motion", "frame"):
            assert fr[f"{label}.y_lik_max_rel"] <= 1e-5, fr                  # 1e-5 rel
            # Entropy bottleneck: bit-identical to eager wherever eager's last matmul is the
            # ascending FMA chain (every 1080p shape).  At <= ~21k columns x channels cuBLAS
            # switches to other reduction orders (profiles/r02_eb_probe.json), so at 256x256
            # eager is itself size-dependent in the last bits: there the patched likelihood
            # must be within 5e-5 of eager AND both must sit inside the fp32 conditioning band
            # around the fp64 evaluation of the same formula (SURVEY.md 7 measured
            # 3.4e-6 ... 3.6e-5 for eager fp32 vs fp64: the tails are differences of sigmoids).
            z_rel = fr[f"{label}.z_lik_max_rel"]
            if z_rel > 1e-5:
                assert small_frame and z_rel <= 5e-5, fr
                if f"{label}.z_lik_stock_vs_fp64" in fr:        # eval mode: fp64 yardstick
                    assert fr[f"{label}.z_equal_inputs"], fr
                    assert fr[f"{label}.z_lik_patched_vs_fp64"] <= 5e-5, fr
                    assert fr[f"{label}.z_lik_stock_vs_fp64"] <= 5e-5, fr
    assert rep["detail_keys_equal"]
    assert rep["bpp_max_rel"] <= 1e-4 and rep["detail_max_rel"] <= 1e-4, rep  # 1e-4 rel


@pytest.mark.parametrize("weight_scale", [TAMED, 1.0], ids=["tamed", "stock_init"])
@pytest.mark.parametrize("channels_last", [False, True], ids=["nchw", "channels_last"])
@pytest.mark.parametrize("hw", [(256, 256), (1088, 1920)], ids=["256x256", "1088x1920"])
def test_forward_eval_stock_vs_patched(cuda_dev, hw, channels_last, weight_scale):
    """``DMC.forward`` on a 3-frame GOP (two P-frames: the second one runs with a
    populated dpb, video_model.py:543-549), eval mode (round quantisation)."""
    stock, patched = _pair(cuda_dev, channels_last, weight_scale)
    stock.eval(), patched.eval()
    h, w = hw
    fr = du.frames(3, 1, h, w, cuda_dev, seed=1, channels_last=channels_last)
    out_s, lat_s = du.run_forward(stock, fr)
    out_p, lat_p = du.run_forward(patched, fr)
    rep = du.compare_forward(out_s, lat_s, out_p, lat_p, num_pixels=h * w * 2, stock=stock)
    _check_forward(rep, 2, small_frame=h * w <= 256 * 256)


def test_forward_train_mode_loss_and_grads(cuda_dev):
    """Config 4 shape class: [B,3,256,256], 3 frames, ``net.train()`` ->
    noise-quantised likelihoods drawn from torch's generator in the reference's
    order; loss as ``RateDistortionLoss`` computes it (train.py:162-211) and its
    gradients through every hot-path op (train.py:301)."""
    import deepvideocodec_b200 as dvc
    stock, patched = _pair(cuda_dev)
    stock.train(), patched.train()
    h = w = 256
    fr = du.frames(3, 2, h, w, cuda_dev, seed=2)
    num_pixels = h * w * 2
    res = {}
    for name, model, collect in (("stock", stock, du.stock_collect()),
                                 ("patched", patched, dvc.collect_likelihoods_list)):
        model.zero_grad(set_to_none=True)
        out, lat = du.run_forward(model, fr, seed=1234, grad=True)
        bpp, _ = collect(out["likelihoods"], num_pixels)
        mse = sum(((x - t) ** 2).mean() for x, t in zip(out["x_hat"], fr[1:])) / 2
        loss = 1e-2 * mse + bpp.mean()
        loss.backward()
        aux = sum(model.aux_loss())
        res[name] = (out, lat, loss.detach(), aux.detach(), bpp.detach(),
                     {n: p.grad.detach().clone() for n, p in model.named_parameters()
                      if p.grad is not None})
    (out_s, lat_s, loss_s, aux_s, bpp_s, g_s), (out_p, lat_p, loss_p, aux_p, bpp_p, g_p) = \
        res["stock"], res["patched"]
    with torch.no_grad():
        rep = du.compare_forward(out_s, lat_s, out_p, lat_p, num_pixels)
    _check_forward(rep, 2, small_frame=True)      # noise-quantised z: no fp64 leg
    assert du.rel_err(loss_p, loss_s) <= 1e-4                    # north_star: loss 1e-4 rel
    assert torch.equal(aux_p, aux_s)
    assert set(g_s) == set(g_p)
    worst = 0.0
    for n in g_s:
        scale = g_s[n].abs().max().item()
        if scale == 0.0:
            assert g_p[n].abs().max().item() == 0.0, n
            continue
        worst = max(worst, (g_p[n] - g_s[n]).abs().max().item() / scale)
    # gradients: 1e-4 of each parameter's gradient scale; the warp's input gradient is an
    # atomic scatter in both implementations (run-to-run order noise, ~1e-6)
    assert worst <= 1e-4, worst


@pytest.mark.parametrize("hw", [(256, 256), (1088, 1920)], ids=["256x256", "1088x1920"])
def test_encode_decode_inter_round_trip(cuda_dev, hw):
    """dmc/test.py:185-196: ``encode_inter`` -> ``decode_inter`` with real bit
    streams.  Stock = CompressAI-style CPU coder (oracle restatement), patched =
    GPU rANS.  Raw-stream mode must be byte-identical to the stock strings;
    every mode must decode to the reconstruction ``DMC.forward`` produced."""
    from deepvideocodec_b200 import coder
    stock, patched = _pair(cuda_dev)
    stock.eval(), patched.eval()
    stock.update(force=True), patched.update(force=True)
    h, w = hw
    f0, f1 = du.frames(2, 1, h, w, cuda_dev, seed=3)
    dpb = {"x_ref": f0, "feature_ref": None, "y_ref": None, "y_mv_ref": None}
    with torch.no_grad():
        x_fwd = patched([f0, f1])["x_hat"][0]
        enc_p = patched.encode_inter(f1, dpb)
        x_dec, new_dpb = patched.decode_inter(enc_p["strings"], enc_p["shape"], dpb)
        assert (x_dec - x_fwd).abs().max().item() <= 1e-5
        assert set(new_dpb) == {"x_ref", "feature_ref", "y_ref", "y_mv_ref"}
        saved = coder.PINNED_STREAM_SYMBOLS
        coder.PINNED_STREAM_SYMBOLS = 0           # one raw stock rans64 stream per sample
        try:
            enc_raw = patched.encode_inter(f1, dpb)
        finally:
            coder.PINNED_STREAM_SYMBOLS = saved
        if hw == (256, 256):                      # the CPU coder is slow: small size only
            enc_s = stock.encode_inter(f1, dpb)
            for key in ("motion", "frame"):
                assert list(enc_s["shape"][key]) == list(enc_raw["shape"][key])
                for a, b in zip(enc_s["strings"][key], enc_raw["strings"][key]):
                    assert [bytes(x) for x in a] == [bytes(x) for x in b], key
            x_dec_s, _ = stock.decode_inter(enc_raw["strings"], enc_raw["shape"], dpb)
            assert (x_dec_s - x_fwd).abs().max().item() <= 1e-5      # stock decodes our bytes
        x_dec_raw, _ = patched.decode_inter(enc_raw["strings"], enc_raw["shape"], dpb)
        assert torch.equal(x_dec_raw, x_dec)
        # bytes written vs the likelihood estimate of the same frame
        bits_est = sum(float(-torch.log2(v).sum()) for lab in patched([f0, f1])["likelihoods"][0].values()
                       for v in lab.values())
        n_bytes = sum(len(s) for key in ("motion", "frame") for grp in enc_raw["strings"][key]
                      for s in grp)
        # sanity only: with random-init weights > 50 % of the symbols sit on the 1e-9
        # likelihood floor (30 bits each in the estimate, escape + bypass code in the stream)
        assert 0.3 * bits_est <= n_bytes * 8 <= 3.0 * bits_est + 4096, (n_bytes * 8, bits_est)
