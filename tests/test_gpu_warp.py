"""GPU parity: warp / pyramid kernels vs the oracle executed by PyTorch-CUDA
eager on the same device and vs the committed golden vectors
generated from the reference on CPU.  Tolerance: 1e-5 absolute (north_star)."""
import os
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WARP_ATOL = 1e-5


def _smooth_flow(n, h, w, sigma_px, dev, gen):
    f = torch.randn(n, 2, h, w, device=dev, generator=gen)
    k = 15
    f = torch.nn.functional.avg_pool2d(f, k, stride=1, padding=k // 2, count_include_pad=False)
    f = f / f.std() * sigma_px
    return f.contiguous()


@pytest.mark.parametrize("shape", [(1, 3, 64, 96), (2, 64, 40, 56), (1, 8, 17, 23), (1, 96, 68, 120),
                                   (1, 5, 33, 47), (3, 4, 16, 16)])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("regime", ["smooth", "wild"])
def test_flow_warp_matches_cuda_eager(cuda_dev, shape, layout, regime):
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    n, c, h, w = shape
    g = torch.Generator(device=cuda_dev).manual_seed(zlib.crc32(repr((shape, layout, regime)).encode()))
    im = torch.randn(n, c, h, w, device=cuda_dev, generator=g)
    flow = _smooth_flow(n, h, w, 4.0, cuda_dev, g) if regime == "smooth" else \
        torch.randn(n, 2, h, w, device=cuda_dev, generator=g) * 16
    if layout == "nhwc":
        im = im.contiguous(memory_format=torch.channels_last)
        flow = flow.contiguous(memory_format=torch.channels_last)
    ref = dmc_ref.flow_warp(im, flow)
    out = dvc.flow_warp(im, flow)
    assert out.shape == ref.shape
    assert out.stride() == im.stride()
    err = (out - ref).abs().max().item()
    assert err <= WARP_ATOL, f"max abs err {err}"


def test_flow_warp_1080p_tolerance_and_bitexact_fraction(cuda_dev):
    """Full BASELINE size (1088x1920): the op-for-op replay must stay inside
    1e-5 where a 'natural' x+flow kernel is off by 2e-4."""
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    g = torch.Generator(device=cuda_dev).manual_seed(5)
    for c, fmt in ((3, torch.contiguous_format), (64, torch.channels_last), (64, torch.contiguous_format)):
        im = torch.randn(1, c, 1088, 1920, device=cuda_dev, generator=g).contiguous(memory_format=fmt)
        flow = _smooth_flow(1, 1088, 1920, 4.0, cuda_dev, g)
        ref = dmc_ref.flow_warp(im, flow)
        out = dvc.flow_warp(im, flow)
        diff = (out - ref).abs()
        assert diff.max().item() <= WARP_ATOL, (c, diff.max().item())
        exact = (out == ref).float().mean().item()
        assert exact > 0.999, f"only {exact:.4f} of outputs bit-identical to CUDA eager"


def test_flow_warp_properties_full_size(cuda_dev):
    """Size-independent properties at 1080p: zero flow is the identity, integer
    flow is a pure shift with border clamp, linear in the image."""
    import deepvideocodec_b200 as dvc
    g = torch.Generator(device=cuda_dev).manual_seed(9)
    im = torch.randn(1, 64, 1088, 1920, device=cuda_dev, generator=g).contiguous(
        memory_format=torch.channels_last)
    zero = torch.zeros(1, 2, 1088, 1920, device=cuda_dev)
    out = dvc.flow_warp(im, zero)
    from oracle import dmc_ref
    assert torch.equal(out, dmc_ref.flow_warp(im, zero))
    # identity only up to the reference's own fp32 coordinate noise (~1e-4 px at
    # x~1900) times the white-noise image gradient
    assert (out - im).abs().mean().item() <= 1e-3
    flow = zero.clone()
    flow[:, 0] = 5.0
    flow[:, 1] = -3.0
    out = dvc.flow_warp(im, flow)
    ys = (torch.arange(1088, device=cuda_dev) - 3).clamp(0, 1087)
    xs = (torch.arange(1920, device=cuda_dev) + 5).clamp(0, 1919)
    expect = im[:, :, ys][:, :, :, xs]
    assert torch.equal(out, dmc_ref.flow_warp(im, flow))
    assert (out - expect).abs().mean().item() <= 1e-3
    f2 = torch.randn(1, 2, 1088, 1920, device=cuda_dev, generator=g) * 3
    a = dvc.flow_warp(im, f2)
    b = dvc.flow_warp(im * 2.0, f2)
    assert torch.equal(b, a * 2.0)        # power-of-two scaling commutes with rounding


def test_flow_warp_border_and_nan_flow(cuda_dev):
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    im = torch.randn(1, 4, 12, 20, device=cuda_dev)
    flow = torch.full((1, 2, 12, 20), 1e6, device=cuda_dev)
    flow[0, 0, 0, 0] = float("nan")
    flow[0, 1, 3, 3] = -1e9
    ref = dmc_ref.flow_warp(im, flow)
    out = dvc.flow_warp(im, flow)
    assert torch.equal(out, ref)


def test_flow_warp_golden_cpu_reference(cuda_dev, golden_dir):
    """Golden vectors produced by the reference itself on CPU (true division of
    the flow): matched with ieee_div=True; the default (CUDA-eager) mode must
    also stay within tolerance at these small sizes."""
    import deepvideocodec_b200 as dvc
    z = np.load(os.path.join(golden_dir, "warp.npz"))
    names = sorted({k.split(".")[0] for k in z.files})
    assert names
    for name in names:
        im = torch.from_numpy(z[f"{name}.im"]).to(cuda_dev)
        flow = torch.from_numpy(z[f"{name}.flow"]).to(cuda_dev)
        ref = torch.from_numpy(z[f"{name}.out"]).to(cuda_dev)
        # CPU eager divides the flow (ieee_div=True replays that: 1e-5 gate);
        # the default replays CUDA eager's reciprocal multiply, one ulp of the
        # source coordinate away from the CPU result -> looser bound.
        for ieee, tol in ((True, WARP_ATOL), (False, 1e-4)):
            out = dvc.flow_warp(im, flow, ieee_div=ieee)
            err = (out - ref).abs().max().item()
            assert err <= tol, (name, ieee, err)
        out = dvc.flow_warp(im.contiguous(memory_format=torch.channels_last), flow, ieee_div=True)
        assert (out - ref).abs().max().item() <= WARP_ATOL, name


@pytest.mark.parametrize("shape", [(1, 2, 64, 96), (2, 2, 1088, 1920), (1, 2, 19, 27), (1, 6, 32, 48)])
def test_bilinear_down_and_pyramid(cuda_dev, shape):
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    g = torch.Generator(device=cuda_dev).manual_seed(3)
    mv = torch.randn(*shape, device=cuda_dev, generator=g) * 5
    ref = dmc_ref.bilinear_down2(mv)
    out = dvc.bilineardownsacling(mv)
    even = shape[2] % 2 == 0 and shape[3] % 2 == 0
    if even:
        assert torch.equal(out, ref)
    else:
        assert (out - ref).abs().max().item() <= 1e-5
    if shape[1] == 2:
        r2, r3 = dmc_ref.flow_pyramid(mv)
        o2, o3 = dvc.flow_pyramid(mv)
        if shape[2] % 4 == 0 and shape[3] % 4 == 0:
            assert torch.equal(o2, r2) and torch.equal(o3, r3)
        else:
            assert (o2 - r2).abs().max().item() <= 1e-5 and (o3 - r3).abs().max().item() <= 1e-5


def test_pyramid_golden(cuda_dev, golden_dir):
    import deepvideocodec_b200 as dvc
    z = np.load(os.path.join(golden_dir, "pyramid.npz"))
    for name in sorted({k.split(".")[0] for k in z.files}):
        mv = torch.from_numpy(z[f"{name}.mv"]).to(cuda_dev)
        o2, o3 = dvc.flow_pyramid(mv)
        for o, key in ((o2, "mv2"), (o3, "mv3")):
            ref = torch.from_numpy(z[f"{name}.{key}"]).to(cuda_dev)
            assert (o - ref).abs().max().item() <= 1e-5, (name, key)


def test_motion_compensation_warps_match(cuda_dev):
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    g = torch.Generator(device=cuda_dev).manual_seed(17)
    h, w = 128, 192
    x_ref = torch.rand(1, 3, h, w, device=cuda_dev, generator=g)
    cl = torch.channels_last
    f1 = torch.randn(1, 64, h, w, device=cuda_dev, generator=g).contiguous(memory_format=cl)
    f2 = torch.randn(1, 64, h // 2, w // 2, device=cuda_dev, generator=g).contiguous(memory_format=cl)
    f3 = torch.randn(1, 64, h // 4, w // 4, device=cuda_dev, generator=g).contiguous(memory_format=cl)
    mv = _smooth_flow(1, h, w, 4.0, cuda_dev, g)
    ref = dmc_ref.motion_compensation_warps(x_ref, f1, f2, f3, mv)
    out = dvc.motion_compensation_warps(x_ref, f1, f2, f3, mv)
    assert len(out) == len(ref) == 4
    for o, r in zip(out, ref):
        assert torch.equal(o, r)          # one launch, pyramid evaluated in-kernel
    # same thing through the materialised pyramid, NCHW features
    mv2, mv3 = dvc.flow_pyramid(mv)
    assert torch.equal(dvc.flow_warp(f2.contiguous(), mv2), ref[1])
    assert torch.equal(dvc.flow_warp(f3.contiguous(), mv3), ref[2])
    outs = dvc.warp_multi([(f2.contiguous(), mv, 1), (f3.contiguous(), mv, 2)])
    assert torch.equal(outs[0], ref[1]) and torch.equal(outs[1], ref[2])


def test_cpu_tensor_is_rejected(cuda_dev):
    import deepvideocodec_b200 as dvc
    with pytest.raises(dvc.DvcError):
        dvc.flow_warp(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))


@pytest.mark.parametrize("shape", [(1, 64, 68, 120), (2, 16, 37, 132), (1, 8, 8, 64), (3, 64, 9, 68),
                                   (1, 96, 130, 260)])
@pytest.mark.parametrize("regime", ["gentle", "rough", "wild", "huge", "mixed"])
def test_planar_nchw_kernel_bit_identical(cuda_dev, shape, regime):
    """The shared-memory staged NCHW path (warp_planar_kernel: W % 4 == 0, C >= 8)
    and its complementary gather launch, over coherent, incoherent and mixed flows,
    ragged tiles, batches, and a non-dense (sliced) input: bit-identical to the
    oracle on CUDA and to the strided path."""
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    n, c, h, w = shape
    g = torch.Generator(device=cuda_dev).manual_seed(zlib.crc32(repr((shape, regime)).encode()))
    im = torch.randn(n, c, h, w, device=cuda_dev, generator=g)
    if regime == "gentle":
        flow = _smooth_flow(n, h, w, 1.5, cuda_dev, g)
    elif regime == "rough":
        flow = _smooth_flow(n, h, w, 6.0, cuda_dev, g)
    elif regime == "wild":
        flow = torch.randn(n, 2, h, w, device=cuda_dev, generator=g) * 16
    elif regime == "huge":       # far outside the image: border clamping on every side
        flow = torch.randn(n, 2, h, w, device=cuda_dev, generator=g) * 500
    else:                        # coherent left half, incoherent right half: both launches work
        flow = _smooth_flow(n, h, w, 2.0, cuda_dev, g)
        flow[..., w // 2:] = torch.randn(n, 2, h, w - w // 2, device=cuda_dev, generator=g) * 20
    ref = dmc_ref.flow_warp(im, flow)
    out = dvc.flow_warp(im, flow)
    assert torch.equal(out, ref)
    # a view with padded rows (h-stride > W, still 16-byte aligned) takes the same path
    wide = torch.randn(n, c, h, w + 8, device=cuda_dev, generator=g)
    view = wide[..., 4:4 + w]
    assert torch.equal(dvc.flow_warp(view, flow), dmc_ref.flow_warp(view.contiguous(), flow))


@pytest.mark.parametrize("hw", [(2160, 3840), (2176, 3840)], ids=["4k", "4k_padded"])
def test_flow_warp_4k_config5(cuda_dev, hw):
    """BASELINE.json configs[4]: flow_warp at 3840x2160 (and the x64-padded 3840x2176),
    C in {3, 64}, NCHW and channels_last: bit-identical to CUDA eager, and one
    size-independent property -- a constant integer flow is an exact shift."""
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    h, w = hw
    g = torch.Generator(device=cuda_dev).manual_seed(h)
    flow = _smooth_flow(1, h, w, 4.0, cuda_dev, g)
    for c, fmt in ((3, torch.contiguous_format), (64, torch.channels_last), (64, torch.contiguous_format)):
        im = torch.randn(1, c, h, w, device=cuda_dev, generator=g).contiguous(memory_format=fmt)
        out = dvc.flow_warp(im, flow)
        ref = dmc_ref.flow_warp(im, flow)
        assert (out - ref).abs().max().item() <= WARP_ATOL, (c, fmt)
        assert (out == ref).float().mean().item() >= 0.999
        del ref
        shift = torch.zeros(1, 2, h, w, device=cuda_dev)
        shift[:, 0] = 3.0
        shift[:, 1] = -2.0
        moved = dvc.flow_warp(im, shift)
        # interior: out[h, w] = im[h - 2, w + 3] up to the reference's own coordinate rounding
        a = moved[:, :, 8:-8, 8:-8]
        b = im[:, :, 6:-10, 11:-5]
        assert (a - b).abs().max().item() <= 2e-3 * im.abs().max().item()
        del im, out, moved


def test_motion_compensation_nchw_one_launch_pair(cuda_dev):
    """The three NCHW context scales of a P-frame go through ONE staged launch and ONE
    complement launch (batched planar tasks): same bits as three separate warps."""
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    g = torch.Generator(device=cuda_dev).manual_seed(91)
    h, w = 272, 480
    x_ref = torch.rand(1, 3, h, w, device=cuda_dev, generator=g)
    f1 = torch.randn(1, 64, h, w, device=cuda_dev, generator=g)
    f2 = torch.randn(1, 64, h // 2, w // 2, device=cuda_dev, generator=g)
    f3 = torch.randn(1, 64, h // 4, w // 4, device=cuda_dev, generator=g)
    for mv in (_smooth_flow(1, h, w, 4.0, cuda_dev, g),
               torch.randn(1, 2, h, w, device=cuda_dev, generator=g) * 16):
        got = dvc.motion_compensation_warps(x_ref, f1, f2, f3, mv)
        want = dmc_ref.motion_compensation_warps(x_ref, f1, f2, f3, mv)
        for a, b in zip(got, want):
            assert torch.equal(a, b)
