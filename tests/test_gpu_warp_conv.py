"""GPU parity of SURVEY.md row f3: flow_warp fused into the 3x3 conv that
consumes it (tcgen05 TF32 implicit GEMM, dvc_warp_conv3x3_fwd).

* the warped context must be BIT-IDENTICAL to dvc.flow_warp (which is itself
  pinned to the reference, tests/test_gpu_warp.py);
* the convolution is checked against an fp64 convolution of TF32-TRUNCATED
  operands (what the tensor core computes: the low 13 mantissa bits of both
  operands are dropped, products are exact, accumulation is fp32) to 2e-5 of
  the output scale, and against the plain fp32 convolution to TF32 tolerance
  (5e-3 of the scale; cuDNN's own TF32 path -- PyTorch's default -- differs
  from fp32 by the same amount).
"""
import zlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TF32_EXACT_RTOL = 2e-5     # vs fp64 conv of truncated operands, relative to max |out|
TF32_RTOL = 5e-3           # vs fp32 conv, relative to max |out|


def _tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _inputs(dev, n, ce, cf, h, w, seed, flow_scale=3.0, level=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    cl = torch.channels_last
    feat = torch.randn(n, cf, h, w, generator=g).to(dev).contiguous(memory_format=cl)
    extra = torch.randn(n, ce, h, w, generator=g).to(dev).contiguous(memory_format=cl) if ce else None
    flow = (torch.randn(n, 2, h << level, w << level, generator=g) * flow_scale).to(dev)
    weight = (torch.randn(64, ce + cf, 3, 3, generator=g) * 0.05).to(dev)
    bias = torch.randn(64, generator=g).to(dev)
    return feat, extra, flow, weight, bias


def _reference(feat, extra, flow, weight, bias, level=0):
    import deepvideocodec_b200 as dvc
    fl = flow
    for _ in range(level):
        fl = dvc.bilineardownsacling(fl) / 2          # video_model.py:499-500
    ctx = dvc.flow_warp(feat, fl)
    x = ctx if extra is None else torch.cat((extra, ctx), 1)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref32 = F.conv2d(x, weight, bias, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    ref64 = F.conv2d(_tf32_trunc(x).double(), _tf32_trunc(weight).double(), bias.double(), padding=1)
    return ctx, ref32, ref64


@pytest.mark.parametrize("shape", [
    (1, 0, 16, 8, 128),      # one tile, no extra
    (1, 16, 16, 8, 128),
    (2, 64, 64, 12, 200),    # ragged width (partial strip), batch 2
    (1, 0, 64, 68, 120),     # conv3_out at 1080p/4 ... latents
    (1, 64, 64, 7, 130),     # ragged height and width
    (1, 64, 64, 272, 480),   # conv3_out / conv2_out scale of a 1080p frame
    (1, 32, 48, 20, 260),    # unequal channel split
])
def test_warp_conv_matches(cuda_dev, shape):
    from deepvideocodec_b200 import layers
    n, ce, cf, h, w = shape
    feat, extra, flow, weight, bias = _inputs(cuda_dev, n, ce, cf, h, w, zlib.crc32(repr(shape).encode()))
    with torch.no_grad():
        ctx, conv = layers.warp_conv3x3(feat, flow, weight, bias, extra)
        ref_ctx, ref32, ref64 = _reference(feat, extra, flow, weight, bias)
    assert conv.shape == (n, 64, h, w) and conv.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(ctx, ref_ctx), "warped context differs from dvc.flow_warp"
    scale = ref32.abs().max().item()
    assert (conv.double() - ref64).abs().max().item() <= TF32_EXACT_RTOL * scale
    assert (conv - ref32).abs().max().item() <= TF32_RTOL * scale


@pytest.mark.parametrize("level", [1, 2])
def test_warp_conv_in_kernel_flow_pyramid(cuda_dev, level):
    """flow_downscale = k: the kernel reduces the full-resolution motion field
    with the reference's pyramid arithmetic (video_model.py:499-500)."""
    from deepvideocodec_b200 import layers
    h, w = 24, 136
    feat, extra, flow, weight, bias = _inputs(cuda_dev, 1, 64, 64, h, w, 77 + level, level=level)
    with torch.no_grad():
        ctx, conv = layers.warp_conv3x3(feat, flow, weight, bias, extra, flow_downscale=level)
        ref_ctx, ref32, ref64 = _reference(feat, extra, flow, weight, bias, level=level)
    assert torch.equal(ctx, ref_ctx)
    scale = ref32.abs().max().item()
    assert (conv.double() - ref64).abs().max().item() <= TF32_EXACT_RTOL * scale


def test_warp_conv_extreme_flow_and_no_warp_output(cuda_dev):
    """Border clamping (flow far outside the image), NaN flow -> coordinate 0,
    and want_warp=False."""
    from deepvideocodec_b200 import layers
    feat, extra, flow, weight, bias = _inputs(cuda_dev, 1, 64, 64, 16, 140, 5, flow_scale=300.0)
    flow[0, 0, 3, 7] = float("nan")
    with torch.no_grad():
        ctx, conv = layers.warp_conv3x3(feat, flow, weight, bias, extra)
        none_ctx, conv2 = layers.warp_conv3x3(feat, flow, weight, bias, extra, want_warp=False)
        ref_ctx, ref32, ref64 = _reference(feat, extra, flow, weight, bias)
    assert none_ctx is None and torch.equal(conv, conv2)
    assert torch.equal(ctx, ref_ctx)
    scale = ref32.abs().max().item()
    assert (conv.double() - ref64).abs().max().item() <= TF32_EXACT_RTOL * scale


def test_warp_conv_1080p_linearity_and_cudnn(cuda_dev):
    """Full BASELINE size (1088x1920, conv1_out 128 -> 64): bit-identical warp,
    agreement with cuDNN's own TF32 convolution, and two size-independent
    properties: homogeneity in the weights by a power of two (exact in any
    binary floating point) and run-to-run determinism."""
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200 import layers
    feat, extra, flow, weight, bias = _inputs(cuda_dev, 1, 64, 64, 1088, 1920, 11)
    zero_bias = torch.zeros_like(bias)
    with torch.no_grad():
        ctx, conv = layers.warp_conv3x3(feat, flow, weight, zero_bias, extra)
        ctx_b, conv_b = layers.warp_conv3x3(feat, flow, weight, zero_bias, extra)
        _, conv4 = layers.warp_conv3x3(feat, flow, weight * 4.0, zero_bias, extra, want_warp=False)
        ref_ctx = dvc.flow_warp(feat, flow)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = True
        try:
            cudnn = F.conv2d(torch.cat((extra, ref_ctx), 1), weight, None, padding=1)
        finally:
            torch.backends.cudnn.allow_tf32 = prev
    assert torch.equal(ctx, ref_ctx)
    assert torch.equal(conv, conv_b) and torch.equal(ctx, ctx_b), "not deterministic"
    assert torch.equal(conv4, conv * 4.0), "not homogeneous in the weights"
    scale = cudnn.abs().max().item()
    assert (conv - cudnn).abs().max().item() <= TF32_RTOL * scale


def test_weight_cache_tracks_parameter_updates(cuda_dev):
    from deepvideocodec_b200 import layers
    feat, extra, flow, weight, bias = _inputs(cuda_dev, 1, 16, 16, 8, 128, 3)
    with torch.no_grad():
        _, a = layers.warp_conv3x3(feat, flow, weight, bias, extra)
        weight.mul_(2.0)                      # in-place update bumps the version counter
        _, b = layers.warp_conv3x3(feat, flow, weight, bias, extra)
        _, ref32, ref64 = _reference(feat, extra, flow, weight, bias)
    assert not torch.equal(a, b)
    assert (b.double() - ref64).abs().max().item() <= TF32_EXACT_RTOL * ref32.abs().max().item()


def test_warp_conv_argument_errors(cuda_dev):
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200 import layers
    feat, extra, flow, weight, bias = _inputs(cuda_dev, 1, 16, 16, 8, 128, 3)
    with torch.no_grad():
        with pytest.raises(dvc.DvcError):
            layers.warp_conv3x3(feat, flow[:, :, :4], weight, bias, extra)        # flow size
        with pytest.raises(dvc.DvcError):
            layers.warp_conv3x3(feat, flow, weight[:, :16], bias, extra)          # channel count
        with pytest.raises(dvc.DvcError):
            layers.warp_conv3x3(feat, flow, weight[:32], bias[:32], extra)        # Co != 64
        with pytest.raises(dvc.DvcError):
            layers.warp_conv3x3(feat.cpu(), flow, weight, bias, extra)            # no CPU fallback
    w = weight.clone().requires_grad_(True)
    with pytest.raises(dvc.DvcError):
        layers.warp_conv3x3(feat, flow, w, bias, extra)                            # inference only


def test_fused_motion_compensation_matches_reference_structure(cuda_dev):
    """dvc.motion_compensation_fused vs the oracle's DMC.motion_compensation
    (warps + MultiScaleContextFusion, video_model.py:37-66, 497-506) with shared
    parameters: warpframe bit-identical, contexts within TF32 tolerance."""
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    torch.manual_seed(0)
    net = dmc_ref.MultiScaleContextFusionRef().to(cuda_dev).eval()
    h, w = 64, 256
    g = torch.Generator(device="cpu").manual_seed(9)
    x_ref = torch.rand(1, 3, h, w, generator=g).to(cuda_dev)
    feats = [torch.randn(1, 64, h >> k, w >> k, generator=g).to(cuda_dev) for k in range(3)]
    mv = (torch.randn(1, 2, h, w, generator=g) * 2).to(cuda_dev)

    class Stub:          # the attributes DMC.motion_compensation touches
        context_fusion_net = net

        def multi_scale_feature_extractor(self, dpb):
            return feats
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            out = dvc.motion_compensation_fused(Stub(), mv, {"x_ref": x_ref})
            ref = dmc_ref.motion_compensation(x_ref, *feats, mv, net)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert torch.equal(out[3], ref[3])
    for k in range(3):
        assert out[k].shape == ref[k].shape
        scale = ref[k].abs().max().item()
        assert (out[k] - ref[k]).abs().max().item() <= 2e-2 * scale, k
