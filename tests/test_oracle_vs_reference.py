"""CPU, build container only: the oracle restatement against the UNMODIFIED
reference executed from /root/reference (skipped where the reference does not
exist, e.g. on the GPU box)."""
import sys

import pytest
import torch

from oracle import dmc_ref
from oracle.load_reference import load_reference_models, load_reference_train_fn, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    models = load_reference_models()
    return {"models": models, "layers": sys.modules["models.layers"],
            "utils": sys.modules["models.utils"], "vm": sys.modules["models.video_model"]}


@pytest.mark.parametrize("shape", [(1, 3, 32, 48), (2, 64, 16, 24), (1, 7, 19, 23)])
def test_flow_warp(ref, shape):
    g = torch.Generator().manual_seed(1)
    n, c, h, w = shape
    im = torch.randn(n, c, h, w, generator=g)
    flow = torch.randn(n, 2, h, w, generator=g) * 5
    ref["layers"].backward_grid[-1].clear()
    assert torch.equal(dmc_ref.flow_warp(im, flow), ref["layers"].flow_warp(im, flow))


def test_pyramid_and_quantize(ref):
    g = torch.Generator().manual_seed(2)
    mv = torch.randn(1, 2, 64, 96, generator=g) * 4
    mv2 = ref["layers"].bilineardownsacling(mv) / 2
    mv3 = ref["layers"].bilineardownsacling(mv2) / 2
    o2, o3 = dmc_ref.flow_pyramid(mv)
    assert torch.equal(o2, mv2) and torch.equal(o3, mv3)
    x = torch.randn(1000, generator=g) * 10
    assert torch.equal(dmc_ref.quantize_ste(x), ref["utils"].quantize_ste(x))


@pytest.mark.parametrize("which", ["motion", "frame"])
def test_context_models(ref, which):
    """Stock Motion/FrameContextModel.forward (conv nets included) vs the oracle's
    non-conv restatement fed with the stock model's own conv outputs."""
    vm = ref["vm"]
    torch.manual_seed(3)
    if which == "motion":
        model = vm.MotionContextModel(ch_mv=8).eval()
        c = 8
    else:
        model = vm.FrameContextModel(N=8, M=12).eval()
        c = 12
    g = torch.Generator().manual_seed(4)
    y = torch.randn(2, c, 16, 16, generator=g) * 3
    with torch.no_grad():
        z = model.hyper_encoder(y)
        if which == "motion":
            y_hat_ref, lik_ref = model(y, None)
            params = model.hyper_decoder(dmc_ref.quantize_hyper(z, model.entropy_bottleneck._get_medians()))
            means, scales = model.y_prior_fusion(torch.cat((params, torch.zeros_like(y)), 1)).chunk(2, 1)
        else:
            context = torch.randn(2, 8, 64, 64, generator=g)
            y_hat_ref, lik_ref = model(y, None, context)
            params = model.hyper_decoder(dmc_ref.quantize_hyper(z, model.entropy_bottleneck._get_medians()))
            temporal = model.temporal_prior_encoder(context)
            means, scales = model.y_prior_fusion(
                torch.cat((temporal, params, torch.zeros_like(y)), 1)).chunk(2, 1)
        y_hat, z_hat, lik = dmc_ref.context_model_forward(
            y, z, means, scales, model.y_spatial_prior, model.entropy_bottleneck,
            model.gaussian_conditional)
        comp_ref = model.forward_dual_prior(y, means, scales, mode="compress")
        comp = dmc_ref.dual_prior(y, means, scales, model.y_spatial_prior, mode="compress")
    assert torch.equal(y_hat, y_hat_ref)
    assert torch.equal(lik["y"], lik_ref["y"]) and torch.equal(lik["z"], lik_ref["z"])
    for a, b in zip(comp, comp_ref):
        assert torch.equal(a, b)


@pytest.mark.parametrize("which", ["motion", "frame"])
def test_context_models_real_bitstreams(ref, which):
    """Stock ``compress`` / ``decompress`` of both context models (the
    reference's own code, video_model.py:236-291 / :408-466, running over the
    oracle's CompressAI surface) vs the oracle's non-conv restatement: same
    strings, same y_hat, and decode(encode) reproduces the encoder's y_hat."""
    vm = ref["vm"]
    torch.manual_seed(5)
    if which == "motion":
        model = vm.MotionContextModel(ch_mv=8).eval()
        c = 8
    else:
        model = vm.FrameContextModel(N=8, M=12).eval()
        c = 12
    model.gaussian_conditional.update_scale_table(
        sys.modules["models.base_model"].get_scale_table(), force=True)
    model.entropy_bottleneck.update(force=True)
    g = torch.Generator().manual_seed(6)
    y = torch.randn(2, c, 16, 16, generator=g) * 4
    context = torch.randn(2, 8, 64, 64, generator=g)
    y_ref = torch.randn(2, c, 16, 16, generator=g)

    def prior_fusion(z_hat):
        params = model.hyper_decoder(z_hat)
        if which == "motion":
            cat = torch.cat((params, y_ref), 1)
        else:
            cat = torch.cat((model.temporal_prior_encoder(context), params, y_ref), 1)
        return model.y_prior_fusion(cat).chunk(2, 1)

    extra = () if which == "motion" else (context,)
    with torch.no_grad():
        y_hat_ref, out_ref = model.compress(y, y_ref, *extra)
        dec_ref = model.decompress(out_ref["strings"], out_ref["shape"], y_ref, *extra)
        z = model.hyper_encoder(y)
        y_hat, out = dmc_ref.context_model_compress(
            y, z, prior_fusion, model.y_spatial_prior, model.entropy_bottleneck,
            model.gaussian_conditional)
        dec = dmc_ref.context_model_decompress(
            out["strings"], out["shape"], prior_fusion, model.y_spatial_prior,
            model.entropy_bottleneck, model.gaussian_conditional)
    assert out["strings"] == out_ref["strings"] and tuple(out["shape"]) == tuple(out_ref["shape"])
    assert torch.equal(y_hat, y_hat_ref)
    assert torch.equal(dec, dec_ref)
    assert torch.equal(dec_ref, y_hat_ref)          # the decoder reconstructs the encoder's latents
    assert all(len(s) == 2 and isinstance(s[0], bytes) for s in out["strings"])


def test_rate(ref):
    collect = load_reference_train_fn("collect_likelihoods_list")
    g = torch.Generator().manual_seed(5)
    liks = [{"motion": {"y": torch.rand(2, 4, 6, 6, generator=g) + 1e-9, "z": torch.rand(2, 4, 2, 2, generator=g) + 1e-9},
             "frame": {"y": torch.rand(2, 6, 6, 6, generator=g) + 1e-9, "z": torch.rand(2, 4, 2, 2, generator=g) + 1e-9}}
            for _ in range(3)]
    a, ia = dmc_ref.collect_likelihoods_list(liks, 96 * 96 * 3)
    b, ib = collect(liks, 96 * 96 * 3)
    assert torch.equal(a, b)
    assert list(ia) == list(ib)
    for k in ib:
        assert torch.equal(torch.as_tensor(ia[k]), torch.as_tensor(ib[k])), k


def test_stock_dmc_forward_regression(ref):
    """regression values of the shim under the stock model."""
    torch.manual_seed(0)
    net = ref["models"].DMC()
    assert sum(p.numel() for p in net.parameters()) == 16884403
    assert len(net.state_dict()) == 438
    aux = [float(a) for a in net.aux_loss()]
    assert abs(aux[0] - 2641.55) < 0.5 and abs(aux[1] - 2636.41) < 0.5


def test_context_fusion_net_restatement(ref):
    """oracle MultiScaleContextFusionRef == stock MultiScaleContextFusion
    (video_model.py:37-66) under the stock module's own state_dict; it anchors
    the fused warp + conv drop-in (tests/test_gpu_warp_conv.py)."""
    vm = ref["vm"]
    torch.manual_seed(5)
    stock = vm.MultiScaleContextFusion().eval()
    mine = dmc_ref.MultiScaleContextFusionRef().eval()
    missing, unexpected = mine.load_state_dict(stock.state_dict(), strict=True)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(6)
    c1 = torch.randn(1, 64, 16, 32, generator=g)
    c2 = torch.randn(1, 64, 8, 16, generator=g)
    c3 = torch.randn(1, 64, 4, 8, generator=g)
    with torch.no_grad():
        a = stock(c1, c2, c3)
        b = mine(c1, c2, c3)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
