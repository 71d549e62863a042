"""Drop-ins for the warp half of the reference's ``dmc/models/layers.py``.

Same names, argument meaning and return structure as the reference:

* ``flow_warp(im, flow)``            -- layers.py:196-198 (``torch_warp`` :175-193)
* ``bilineardownsacling(x)``         -- layers.py:201-206
* ``motion_compensation_warps(...)`` -- the non-conv part of
  ``DMC.motion_compensation`` (video_model.py:497-504) in one launch

Each call is one kernel of ``libdvc_b200.so`` (no grid tensor, no flow
normalisation passes, no global cache -- the reference's ``backward_grid``
cache and its CPU/``cuda:7`` aliasing are deliberately not reproduced).
Both NCHW-contiguous and ``torch.channels_last`` tensors are accepted; the
output takes the memory format of ``im``.  channels_last with C % 4 == 0 is
the fast path.
"""
import ctypes
import weakref

import torch

from . import _native as nat

__all__ = ["flow_warp", "torch_warp", "bilineardownsacling", "flow_pyramid",
           "motion_compensation_warps", "warp_conv3x3", "pack_conv3x3_weight"]


def _check_warp_args(im, flow):
    nat.require_cuda_f32(im, "flow_warp(im)")
    nat.require_cuda_f32(flow, "flow_warp(flow)")
    if flow.device != im.device:
        raise nat.DvcError("flow_warp: im and flow are on different devices")
    n, c, h, w = im.shape
    if flow.shape != (n, 2, h, w):
        # the reference builds its base grid from the flow's size and samples
        # `im`; it only ever calls with equal sizes
        raise nat.DvcError(f"flow_warp: flow must be [N,2,H,W] matching im {tuple(im.shape)}, "
                           f"got {tuple(flow.shape)}")
    if h < 2 or w < 2:
        raise nat.DvcError("flow_warp: H and W must be >= 2")


def _warp_fwd(im, flow, flags=0):
    out = torch.empty_like(im)            # preserves NCHW / channels_last
    n, c, h, w = im.shape
    with nat.device_of(im):
        rc = nat.lib().dvc_flow_warp_fwd(im.data_ptr(), flow.data_ptr(), out.data_ptr(),
                                         n, c, h, w, nat.st4(im), nat.st4(flow), nat.st4(out),
                                         flags, nat.stream_of(im))
    nat.check(rc, "dvc_flow_warp_fwd")
    return out


class _FlowWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im, flow, flags):
        ctx.save_for_backward(im, flow)
        ctx.flags = flags
        return _warp_fwd(im, flow, flags)

    @staticmethod
    def backward(ctx, grad_out):
        im, flow = ctx.saved_tensors
        need_im, need_flow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_out = nat.require_cuda_f32(grad_out, "flow_warp(grad_out)")
        n, c, h, w = im.shape
        grad_im = torch.zeros_like(im) if need_im else None
        grad_flow = torch.empty_like(flow) if need_flow else None
        with nat.device_of(im):
            rc = nat.lib().dvc_flow_warp_bwd(
                grad_out.data_ptr(), im.data_ptr(), flow.data_ptr(), nat.ptr(grad_im),
                nat.ptr(grad_flow), n, c, h, w, nat.st4(grad_out), nat.st4(im), nat.st4(flow),
                nat.opt_st4(grad_im), nat.opt_st4(grad_flow), ctx.flags, nat.stream_of(im))
        nat.check(rc, "dvc_flow_warp_bwd")
        return grad_im, grad_flow, None


def flow_warp(im, flow, *, ieee_div=False):
    """Backward-warp ``im[N,C,H,W]`` by ``flow[N,2,H,W]`` (pixels; channel 0 = x).

    Bilinear, border padding, ``align_corners=True`` -- reference
    ``flow_warp`` (layers.py:196).  ``ieee_div=True`` reproduces PyTorch-CPU's
    true division of the flow instead of PyTorch-CUDA's reciprocal multiply.
    """
    _check_warp_args(im, flow)
    flags = nat.DVC_WARP_IEEE_DIV if ieee_div else 0
    if torch.is_grad_enabled() and (im.requires_grad or flow.requires_grad):
        return _FlowWarp.apply(im, flow, flags)
    return _warp_fwd(im, flow, flags)


torch_warp = flow_warp      # layers.py:175 (the reference exposes both names)


# ---------------------------------------------------------------------------
def _down2_fwd(x, post_scale):
    n, c, h, w = x.shape
    fmt = torch.channels_last if (x.is_contiguous(memory_format=torch.channels_last)
                                  and not x.is_contiguous()) else torch.contiguous_format
    y = torch.empty((n, c, h // 2, w // 2), dtype=x.dtype, device=x.device, memory_format=fmt)
    with nat.device_of(x):
        rc = nat.lib().dvc_bilinear_down2_fwd(x.data_ptr(), y.data_ptr(), n, c, h, w,
                                              nat.st4(x), nat.st4(y), float(post_scale),
                                              nat.stream_of(x))
    nat.check(rc, "dvc_bilinear_down2_fwd")
    return y


class _Down2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, post_scale):
        ctx.shape = tuple(x.shape)
        ctx.post = post_scale
        ctx.like = x.new_empty(0)
        ctx.channels_last = x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        return _down2_fwd(x, post_scale)

    @staticmethod
    def backward(ctx, grad_y):
        grad_y = nat.require_cuda_f32(grad_y, "bilineardownsacling(grad)")
        n, c, h, w = ctx.shape
        fmt = torch.channels_last if ctx.channels_last else torch.contiguous_format
        # odd sizes scatter with atomics into a zero-filled gradient
        grad_x = torch.empty(ctx.shape, dtype=grad_y.dtype, device=grad_y.device, memory_format=fmt)
        if h % 2 or w % 2:
            grad_x.zero_()
        with nat.device_of(grad_y):
            rc = nat.lib().dvc_bilinear_down2_bwd(grad_y.data_ptr(), grad_x.data_ptr(), n, c, h, w,
                                                  nat.st4(grad_y), nat.st4(grad_x), float(ctx.post),
                                                  nat.stream_of(grad_y))
        nat.check(rc, "dvc_bilinear_down2_bwd")
        return grad_x, None


def bilineardownsacling(inputfeature, *, post_scale=1.0):
    """``F.interpolate(x, (H//2, W//2), 'bilinear', align_corners=False)`` --
    reference ``bilineardownsacling`` (layers.py:201-206; the spelling is the
    reference's).  ``post_scale`` folds the ``/ 2`` of video_model.py:499-500."""
    x = nat.require_cuda_f32(inputfeature, "bilineardownsacling(x)")
    if x.size(2) < 2 or x.size(3) < 2:
        raise nat.DvcError("bilineardownsacling: H and W must be >= 2")
    if torch.is_grad_enabled() and x.requires_grad:
        return _Down2.apply(x, float(post_scale))
    return _down2_fwd(x, float(post_scale))


def flow_pyramid(mv):
    """``mv2 = bilineardownsacling(mv) / 2; mv3 = bilineardownsacling(mv2) / 2``
    (video_model.py:499-500).  One launch when H, W are multiples of 4 and no
    gradient is needed; two single-level launches otherwise."""
    mv = nat.require_cuda_f32(mv, "flow_pyramid(mv)")
    n, c, h, w = mv.shape
    fused = (c == 2 and h % 4 == 0 and w % 4 == 0
             and not (torch.is_grad_enabled() and mv.requires_grad))
    if not fused:
        mv2 = bilineardownsacling(mv, post_scale=0.5)
        return mv2, bilineardownsacling(mv2, post_scale=0.5)
    mv2 = torch.empty((n, 2, h // 2, w // 2), dtype=mv.dtype, device=mv.device)
    mv3 = torch.empty((n, 2, h // 4, w // 4), dtype=mv.dtype, device=mv.device)
    with nat.device_of(mv):
        rc = nat.lib().dvc_flow_pyramid_fwd(mv.data_ptr(), mv2.data_ptr(), mv3.data_ptr(), n, h, w,
                                            nat.st4(mv), nat.st4(mv2), nat.st4(mv3),
                                            nat.stream_of(mv))
    nat.check(rc, "dvc_flow_pyramid_fwd")
    return mv2, mv3


def _task(im, flow, out, level=0):
    t = nat.WarpTask()
    t.im, t.flow, t.out = im.data_ptr(), flow.data_ptr(), out.data_ptr()
    t.N, t.C, t.H, t.W = im.shape
    t.im_st = nat.st4(im)
    t.flow_st = nat.st4(flow)
    t.out_st = nat.st4(out)
    t.flow_downscale = level
    return t


def warp_multi(pairs, *, ieee_div=False):
    """Warp up to 4 independent ``(im, flow[, level])`` problems in ONE launch
    (forward only).  ``level`` k in {0,1,2}: ``flow`` is the motion field at 2^k
    times the resolution of ``im`` and is reduced on the fly with the
    reference's pyramid arithmetic (``bilineardownsacling(.) / 2`` k times,
    video_model.py:499-500).  Returns the list of warped tensors."""
    if not 1 <= len(pairs) <= 4:
        raise nat.DvcError("warp_multi: 1..4 (im, flow) pairs")
    outs = []
    tasks = (nat.WarpTask * len(pairs))()
    for i, item in enumerate(pairs):
        im, flow = item[0], item[1]
        level = item[2] if len(item) > 2 else 0
        if level not in (0, 1, 2):
            raise nat.DvcError("warp_multi: level must be 0, 1 or 2")
        nat.require_cuda_f32(im, "warp_multi(im)")
        nat.require_cuda_f32(flow, "warp_multi(flow)")
        n, c, h, w = im.shape
        if flow.shape != (n, 2, h << level, w << level):
            raise nat.DvcError(f"warp_multi: flow must be {(n, 2, h << level, w << level)} for im "
                               f"{tuple(im.shape)} at level {level}, got {tuple(flow.shape)}")
        if h < 2 or w < 2:
            raise nat.DvcError("warp_multi: H and W must be >= 2")
        if im.device != pairs[0][0].device or flow.device != im.device:
            raise nat.DvcError("warp_multi: all tensors must be on one device")
        out = torch.empty_like(im)
        outs.append(out)
        tasks[i] = _task(im, flow, out, level)
    im0 = pairs[0][0]
    with nat.device_of(im0):
        rc = nat.lib().dvc_warp_multi_fwd(ctypes.cast(tasks, ctypes.c_void_p), len(pairs),
                                          nat.DVC_WARP_IEEE_DIV if ieee_div else 0,
                                          nat.stream_of(im0))
    nat.check(rc, "dvc_warp_multi_fwd")
    return outs


def motion_compensation_warps(x_ref, feat1, feat2, feat3, mv):
    """The warps of ``DMC.motion_compensation`` (video_model.py:497-504):

        warpframe = flow_warp(x_ref, mv);  mv2, mv3 = pyramid(mv)
        context_k = flow_warp(feat_k, mv_k)

    Returns ``(context1, context2, context3, warpframe)``.  In inference this
    is ONE launch: ``mv2``/``mv3`` are intermediates of the reference function
    and are evaluated inside the warp kernel (bit-identical to materialising
    them).  When a gradient is required it composes the per-op autograd
    functions instead."""
    need_grad = torch.is_grad_enabled() and any(
        t.requires_grad for t in (x_ref, feat1, feat2, feat3, mv))
    fusable = (mv.size(2) % 4 == 0 and mv.size(3) % 4 == 0
               and feat2.shape[-2:] == (mv.size(2) // 2, mv.size(3) // 2)
               and feat3.shape[-2:] == (mv.size(2) // 4, mv.size(3) // 4))
    if need_grad or not fusable:
        mv2, mv3 = flow_pyramid(mv)
        return (flow_warp(feat1, mv), flow_warp(feat2, mv2), flow_warp(feat3, mv3),
                flow_warp(x_ref, mv))
    # big task first so the small ones fill the tail of the grid
    c1, c2, c3, wf = warp_multi([(feat1, mv, 0), (feat2, mv, 1), (feat3, mv, 2), (x_ref, mv, 0)])
    return c1, c2, c3, wf


# ---------------------------------------------------------------------------
# warp fused into the 3x3 conv that consumes it (tcgen05)
# ---------------------------------------------------------------------------
_packed_weights = {}   # id(weight) -> (weakref to the weight object, {ce: (key, packed)})


def pack_conv3x3_weight(weight, ce=0):
    """Re-lay a ``[64, ce + cf, 3, 3]`` conv weight into the UMMA operand layout.

    ``ce`` is the number of leading input channels that belong to ``extra``.
    Cached per weight OBJECT (weakly: a freed tensor whose storage address is
    reused cannot alias), re-packed when its version counter, storage or
    strides change.
    """
    nat.require_cuda_f32(weight, "pack_conv3x3_weight(weight)")
    co, ci, kh, kw = weight.shape
    if (kh, kw) != (3, 3):
        raise nat.DvcError(f"pack_conv3x3_weight: 3x3 kernels only, got {kh}x{kw}")
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), tuple(weight.stride()))
    entry = _packed_weights.get(id(weight))
    if entry is not None and entry[0]() is not weight:     # id reused by another object
        entry = None
    if entry is not None and ce in entry[1] and entry[1][ce][0] == key:
        return entry[1][ce][1]
    n = nat.lib().dvc_conv3x3_packed_weight_floats(co, ci)
    if n <= 0 or ce % 16 or not 0 <= ce < ci:
        raise nat.DvcError(f"pack_conv3x3_weight: need Co == 64 and channel counts that are "
                           f"multiples of 16, got Co={co}, Ci={ci}, Ce={ce}")
    packed = torch.empty(n, dtype=torch.float32, device=weight.device)
    w = weight.detach()
    with nat.device_of(w):
        rc = nat.lib().dvc_conv3x3_pack_weights(w.data_ptr(), nat.st4(w), co, ce, ci - ce,
                                                packed.data_ptr(), nat.stream_of(w))
    nat.check(rc, "dvc_conv3x3_pack_weights")
    if entry is None:
        wid = id(weight)
        entry = (weakref.ref(weight, lambda _r, wid=wid: _packed_weights.pop(wid, None)), {})
        _packed_weights[wid] = entry
    entry[1][ce] = (key, packed)
    return packed


def _nhwc_dense(t):
    return t if t.is_contiguous(memory_format=torch.channels_last) else \
        t.contiguous(memory_format=torch.channels_last)


def warp_conv3x3(feat, flow, weight, bias=None, extra=None, *, flow_downscale=0,
                 want_warp=True, packed=None, _debug=0):
    """``ctx = flow_warp(feat, flow)``; ``conv2d(cat((extra, ctx), 1), weight, bias, padding=1)``.

    One tcgen05 implicit-GEMM kernel (TF32 operands, fp32 accumulation): the
    warped tile goes from the gather straight into shared memory as the GEMM
    operand and is written to HBM once, never re-read.  Replaces
    ``flow_warp`` (layers.py:196) + ``conv{1,2,3}_out`` of
    ``MultiScaleContextFusion`` (video_model.py:55-61).  Inference only.

    Returns ``(ctx, conv)``, both channels_last; ``ctx`` is bit-identical to
    :func:`flow_warp` (``None`` when ``want_warp`` is false).
    """
    nat.require_cuda_f32(feat, "warp_conv3x3(feat)")
    nat.require_cuda_f32(flow, "warp_conv3x3(flow)")
    nat.require_cuda_f32(weight, "warp_conv3x3(weight)")
    n, cf, h, w = feat.shape
    ce = 0
    if extra is not None:
        nat.require_cuda_f32(extra, "warp_conv3x3(extra)")
        if extra.shape[0] != n or extra.shape[2:] != feat.shape[2:]:
            raise nat.DvcError(f"warp_conv3x3: extra {tuple(extra.shape)} does not match feat "
                               f"{tuple(feat.shape)}")
        ce = extra.shape[1]
        extra = _nhwc_dense(extra)
    if flow.shape != (n, 2, h << flow_downscale, w << flow_downscale):
        raise nat.DvcError(f"warp_conv3x3: flow must be [N,2,{h << flow_downscale},"
                           f"{w << flow_downscale}], got {tuple(flow.shape)}")
    co = weight.shape[0]
    if weight.shape[1] != ce + cf:
        raise nat.DvcError(f"warp_conv3x3: weight expects {weight.shape[1]} input channels, "
                           f"got {ce} + {cf}")
    if bias is not None and (not bias.is_cuda or bias.dtype != torch.float32 or
                             bias.numel() != co or not bias.is_contiguous()):
        raise nat.DvcError("warp_conv3x3: bias must be a contiguous CUDA fp32 [Co] tensor")
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad
                                       for t in (feat, flow, extra, weight, bias)):
        raise nat.DvcError("warp_conv3x3 is inference only (no backward); call it under "
                           "torch.no_grad() or use flow_warp + conv2d for training")
    feat = _nhwc_dense(feat)
    if packed is None:
        packed = pack_conv3x3_weight(weight, ce)
    ctx = torch.empty_like(feat) if want_warp else None
    conv = torch.empty((n, co, h, w), dtype=torch.float32, device=feat.device,
                       memory_format=torch.channels_last)
    with nat.device_of(feat):
        rc = nat.lib().dvc_warp_conv3x3_fwd(
            feat.data_ptr(), flow.data_ptr(), nat.ptr(extra), packed.data_ptr(),
            nat.ptr(None if bias is None else bias.detach()), nat.ptr(ctx), conv.data_ptr(),
            n, cf, ce, co, h, w, nat.st4(flow), flow_downscale, (_debug & 0xff) << 8,
            nat.stream_of(feat))
    nat.check(rc, "dvc_warp_conv3x3_fwd")
    return ctx, conv
