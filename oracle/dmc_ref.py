"""Eager-torch restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Every function names the reference lines it follows (paths relative to
``/root/reference``).  The functions are device agnostic: run on CPU they are
the golden-vector oracle and the CPU baseline ("port"); run on ``cuda`` they
are *the reference executed by PyTorch-CUDA eager on the same device*, which is
the parity oracle SURVEY.md section 8c prescribes for the 1e-5 tolerances
(fp32 coordinate and cancellation noise make a CPU/fp64 oracle unusable at
1080p -- SURVEY.md section 0 facts 2-3).

Pinned against the reference itself by ``tests/golden/make_golden.py`` ->
``tests/golden/*.npz`` and ``tests/test_oracle_vs_reference.py``.
"""
import math
from collections import defaultdict

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------
# piece 1: warp  (dmc/models/layers.py:175-198)
# --------------------------------------------------------------------------
def base_grid(n, h, w, device, dtype):
    """Normalised sampling lattice of ``torch_warp`` (layers.py:178-183).

    The reference caches it per device/shape; the values only depend on
    ``torch.linspace`` so no cache is reproduced here."""
    xs = torch.linspace(-1.0, 1.0, w, device=device, dtype=dtype)
    ys = torch.linspace(-1.0, 1.0, h, device=device, dtype=dtype)
    gx = xs.view(1, 1, 1, w).expand(n, -1, h, -1)
    gy = ys.view(1, 1, h, 1).expand(n, -1, -1, w)
    return torch.cat([gx, gy], 1)


def flow_warp(im, flow):
    """``flow_warp`` / ``torch_warp`` (layers.py:175-198): backward warp of
    ``im[N,C,H,W]`` by a pixel-unit flow ``[N,2,H,W]``; bilinear, border
    padding, ``align_corners=True``."""
    n, _, h, w = flow.shape
    half_w = (im.size(3) - 1.0) / 2.0          # python floats, as in layers.py:185-186
    half_h = (im.size(2) - 1.0) / 2.0
    flow_n = torch.cat([flow[:, 0:1] / half_w, flow[:, 1:2] / half_h], 1)
    grid = base_grid(n, h, w, im.device, im.dtype) + flow_n
    return F.grid_sample(im, grid.permute(0, 2, 3, 1), mode="bilinear",
                         padding_mode="border", align_corners=True)


# --------------------------------------------------------------------------
# flow pyramid  (layers.py:201-206, video_model.py:499-500)
# --------------------------------------------------------------------------
def bilinear_down2(x):
    """``bilineardownsacling`` (layers.py:201-206)."""
    return F.interpolate(x, (x.size(2) // 2, x.size(3) // 2), mode="bilinear",
                         align_corners=False)


def flow_pyramid(mv):
    """``mv2``/``mv3`` of ``DMC.motion_compensation`` (video_model.py:499-500)."""
    mv2 = bilinear_down2(mv) / 2
    mv3 = bilinear_down2(mv2) / 2
    return mv2, mv3


def motion_compensation_warps(x_ref, feat1, feat2, feat3, mv):
    """The warp half of ``DMC.motion_compensation`` (video_model.py:497-504);
    the conv nets at :501 and :505 are outside the hot path, their outputs
    (``feat1..3``) are inputs here."""
    warpframe = flow_warp(x_ref, mv)
    mv2, mv3 = flow_pyramid(mv)
    return (flow_warp(feat1, mv), flow_warp(feat2, mv2), flow_warp(feat3, mv3),
            warpframe)


# --------------------------------------------------------------------------
# piece 2: quantisation  (utils.py:149-152, video_model.py:152-216, 222-224)
# --------------------------------------------------------------------------
class ResBlockRef(torch.nn.Module):
    """``ResBlock(channel)`` with the defaults the context-fusion net uses
    (layers.py:59-81: LeakyReLU(0.01) -> conv -> LeakyReLU -> conv, + x)."""

    def __init__(self, channel, slope=0.01):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(channel, channel, 3, padding=1)
        self.conv2 = torch.nn.Conv2d(channel, channel, 3, padding=1)
        self.slope = slope

    def forward(self, x):
        out = F.leaky_relu(x, self.slope)
        out = self.conv1(out)
        out = F.leaky_relu(out, self.slope)
        out = self.conv2(out)
        return x + out


def _subpel_conv3x3(in_ch, out_ch, r):
    # layers.py:52-56
    return torch.nn.Sequential(torch.nn.Conv2d(in_ch, out_ch * r ** 2, 3, padding=1),
                               torch.nn.PixelShuffle(r))


class MultiScaleContextFusionRef(torch.nn.Module):
    """Restatement of ``MultiScaleContextFusion`` (video_model.py:37-66) with the
    reference's parameter names, so a reference ``state_dict`` loads."""

    def __init__(self, channel_in=64, channel_out=64):
        super().__init__()
        c = channel_out
        self.conv3_up = _subpel_conv3x3(channel_in, c, 2)
        self.res_block3_up = ResBlockRef(c)
        self.conv3_out = torch.nn.Conv2d(c, c, 3, padding=1)
        self.res_block3_out = ResBlockRef(c)
        self.conv2_up = _subpel_conv3x3(c * 2, c, 2)
        self.res_block2_up = ResBlockRef(c)
        self.conv2_out = torch.nn.Conv2d(c * 2, c, 3, padding=1)
        self.res_block2_out = ResBlockRef(c)
        self.conv1_out = torch.nn.Conv2d(c * 2, c, 3, padding=1)
        self.res_block1_out = ResBlockRef(c)

    def forward(self, context1, context2, context3):
        context3_up = self.res_block3_up(self.conv3_up(context3))
        context3_out = self.res_block3_out(self.conv3_out(context3))
        cat32 = torch.cat((context3_up, context2), dim=1)
        context2_up = self.res_block2_up(self.conv2_up(cat32))
        context2_out = self.res_block2_out(self.conv2_out(cat32))
        context1_out = self.res_block1_out(self.conv1_out(torch.cat((context2_up, context1), dim=1)))
        return context1 + context1_out, context2 + context2_out, context3 + context3_out


def motion_compensation(x_ref, feat1, feat2, feat3, mv, fusion_net):
    """``DMC.motion_compensation`` (video_model.py:497-506) given the three
    reference features: warps, then the context-fusion net."""
    c1, c2, c3, wf = motion_compensation_warps(x_ref, feat1, feat2, feat3, mv)
    c1, c2, c3 = fusion_net(c1, c2, c3)
    return c1, c2, c3, wf


def quantize_ste(x):
    """``quantize_ste`` (utils.py:149-152): round-half-even forward, identity
    backward."""
    return (torch.round(x) - x).detach() + x


def checkerboard_masks(h, w, device):
    """``get_mask`` (video_model.py:152-159): mask_0[h,w] = 1 iff (h+w) even."""
    cell = torch.tensor(((1, 0), (0, 1)), dtype=torch.float32, device=device)
    m0 = cell.repeat(h // 2, w // 2)[None, None]
    return m0, torch.ones_like(m0) - m0


def process_with_mask(y, means, scales, mask):
    """``process_with_mask`` (video_model.py:161-167)."""
    means_hat = means * mask
    scales_hat = scales * mask
    y_quant = quantize_ste((y - means_hat) * mask)
    return y_quant, y_quant + means_hat, means_hat, scales_hat


def dual_prior_stage_a(y, means, scales):
    """First half of ``forward_dual_prior`` (video_model.py:176-189): returns
    the spatial-prior input ``cat(y_hat_00, y_hat_11, means, scales)`` and the
    stage-A intermediates."""
    m0, m1 = checkerboard_masks(y.size(2), y.size(3), y.device)
    y0, y1 = y.chunk(2, 1)
    mu0, mu1 = means.chunk(2, 1)
    s0, s1 = scales.chunk(2, 1)
    a00 = process_with_mask(y0, mu0, s0, m0)
    a11 = process_with_mask(y1, mu1, s1, m1)
    params = torch.cat((a00[1], a11[1], means, scales), dim=1)
    return params, a00, a11


def dual_prior(y, means, scales, spatial_prior, mode="trainval"):
    """``forward_dual_prior`` (video_model.py:169-216 == :341-388).
    ``spatial_prior`` is the 3-conv module (kept on cuDNN) or any callable
    mapping the ``[N,3C,h,w]`` stage-A tensor to ``[N,2C,h,w]``."""
    m0, m1 = checkerboard_masks(y.size(2), y.size(3), y.device)
    y0, y1 = y.chunk(2, 1)
    params, a00, a11 = dual_prior_stage_a(y, means, scales)
    mu0p, s0p, mu1p, s1p = spatial_prior(params).chunk(4, 1)
    b01 = process_with_mask(y0, mu0p, s0p, m1)
    b10 = process_with_mask(y1, mu1p, s1p, m0)
    y_hat = torch.cat((a00[1] + b01[1], a11[1] + b10[1]), dim=1)
    means_hat = torch.cat((a00[2] + b01[2], a11[2] + b10[2]), dim=1)
    scales_hat = torch.cat((a00[3] + b01[3], a11[3] + b10[3]), dim=1)
    if mode == "compress":
        return (y_hat, a00[0] + a11[0], b01[0] + b10[0],
                a00[3] + a11[3], b01[3] + b10[3])
    return y_hat, means_hat, scales_hat


def quantize_hyper(z, medians):
    """z quantisation (video_model.py:222-224 / :394-396)."""
    return quantize_ste(z - medians) + medians


def context_model_forward(y, z, means, scales, spatial_prior, eb, gc):
    """Non-conv part of ``MotionContextModel.forward`` / ``FrameContextModel
    .forward`` (video_model.py:218-233 / :390-406) with the conv outputs
    (``z``, prior-fusion ``means``/``scales``) supplied as inputs."""
    _, z_lik = eb(z)
    z_hat = quantize_hyper(z, eb._get_medians())
    y_hat, means_hat, scales_hat = dual_prior(y, means, scales, spatial_prior)
    _, y_lik = gc(y, scales_hat, means_hat)
    return y_hat, z_hat, {"y": y_lik, "z": z_lik}


def context_model_compress(y, z, prior_fusion, spatial_prior, eb, gc):
    """Non-conv part of ``MotionContextModel.compress`` / ``FrameContextModel
    .compress`` (video_model.py:236-253 / :408-427).  ``prior_fusion(z_hat)``
    stands for the conv stack between ``z_hat`` and ``(means, scales)``
    (hyper decoder, temporal prior, ``y_prior_fusion`` + ``chunk``)."""
    z_strings = eb.compress(z)
    z_hat = eb.decompress(z_strings, z.size()[-2:])
    means, scales = prior_fusion(z_hat)
    y_hat, q_w0, q_w1, s_w0, s_w1 = dual_prior(y, means, scales, spatial_prior, mode="compress")
    indexes_0 = gc.build_indexes(s_w0)
    indexes_1 = gc.build_indexes(s_w1)
    y_strings_0 = gc.compress(q_w0, indexes_0)
    y_strings_1 = gc.compress(q_w1, indexes_1)
    return y_hat, {"strings": [y_strings_0, y_strings_1, z_strings], "shape": z.size()[-2:]}


def context_model_decompress(strings, shape, prior_fusion, spatial_prior, eb, gc):
    """Non-conv part of ``MotionContextModel.decompress`` / ``FrameContextModel
    .decompress`` (video_model.py:255-291 / :429-466)."""
    assert isinstance(strings, list) and len(strings) == 3
    z_hat = eb.decompress(strings[2], shape)
    means, scales = prior_fusion(z_hat)
    m0, m1 = checkerboard_masks(means.size(2), means.size(3), means.device)
    mu0, mu1 = means.chunk(2, 1)
    s0, s1 = scales.chunk(2, 1)
    q_r0 = gc.decompress(strings[0], gc.build_indexes(s0 * m0 + s1 * m1))
    y00 = (q_r0 + mu0) * m0
    y11 = (q_r0 + mu1) * m1
    mu0p, s0p, mu1p, s1p = spatial_prior(torch.cat((y00, y11, means, scales), dim=1)).chunk(4, 1)
    q_r1 = gc.decompress(strings[1], gc.build_indexes(s0p * m1 + s1p * m0))
    y01 = (q_r1 + mu0p) * m1
    y10 = (q_r1 + mu1p) * m0
    return torch.cat((y00 + y01, y11 + y10), dim=1)


# --------------------------------------------------------------------------
# piece 4: rate  (dmc/train.py:74-93)
# --------------------------------------------------------------------------
def collect_likelihoods_list(likelihoods_list, num_pixels):
    """``collect_likelihoods_list`` (train.py:74-93)."""
    info = defaultdict(int)
    total = 0
    for i, frame in enumerate(likelihoods_list):
        frame_bpp = 0
        for label, fields in frame.items():
            label_bpp = 0
            for field, p in fields.items():
                bpp = torch.log(p).sum(dim=(1, 2, 3)) / (-math.log(2) * num_pixels)
                total = total + bpp
                frame_bpp = frame_bpp + bpp
                label_bpp = label_bpp + bpp
                info[f"bpp_loss.{label}"] += bpp.sum()
                info[f"bpp_loss.{label}.{i}.{field}"] = bpp.sum()
            info[f"bpp_loss.{label}.{i}"] = label_bpp.sum()
        info[f"bpp_loss.{i}"] = frame_bpp.sum()
    return total, info


def frame_bits(likelihoods):
    """Bits of one P-frame = -sum log2 p over its likelihood tensors, per
    sample (``[B]``); SURVEY.md A.6."""
    bits = 0
    for fields in likelihoods.values():
        for p in fields.values():
            bits = bits + torch.log(p).sum(dim=(1, 2, 3)) / (-math.log(2))
    return bits


# --------------------------------------------------------------------------
# the timed region of one P-frame, reference ops only
# --------------------------------------------------------------------------
def pframe_hot_path(inp, eb_modules, gc, num_pixels=None):
    """Reference-op restatement of the benchmark's timed region: the four
    warps + flow pyramid of ``DMC.motion_compensation`` (video_model.py:497-504)
    and, for both context models, the non-conv part of ``forward``
    (video_model.py:218-233 / :390-406) with the spatial-prior conv replaced by
    its precomputed output ``inp['<label>.prior']``; then the rate
    (train.py:74-93).  ``inp`` uses the keys of
    ``deepvideocodec_b200.pipeline.synthetic_pframe_inputs``."""
    c1, c2, c3, wf = motion_compensation_warps(
        inp["x_ref"], inp["feat1"], inp["feat2"], inp["feat3"], inp["mv"])
    out = {"context1": c1, "context2": c2, "context3": c3, "warpframe": wf}
    liks = {}
    for label in ("motion", "frame"):
        prior = inp[f"{label}.prior"]
        y_hat, z_hat, lik = context_model_forward(
            inp[f"{label}.y"], inp[f"{label}.z"], inp[f"{label}.means"], inp[f"{label}.scales"],
            lambda params, _p=prior: _p, eb_modules[label], gc)
        out[f"{label}.y_hat"], out[f"{label}.z_hat"] = y_hat, z_hat
        out[f"{label}.y_lik"], out[f"{label}.z_lik"] = lik["y"], lik["z"]
        liks[label] = lik
    h, w = inp["x_ref"].shape[-2:]
    npx = num_pixels if num_pixels is not None else h * w
    out["bpp_total"], out["bpp_info"] = collect_likelihoods_list([liks], npx)
    out["bits"] = frame_bits(liks)
    return out
