"""Fused replacements for the non-conv part of the reference's two context
models (``MotionContextModel`` / ``FrameContextModel``,
``dmc/models/video_model.py:128-233`` and ``:294-406``).

The reference runs ``get_mask`` (host tensor + H2D copy per forward, :152-159),
four ``process_with_mask`` calls (:161-167), three ``cat``s, the Gaussian
conditional (~20 ops) and the z quantisation as ~70 eager launches around the
3-conv spatial prior.  Here it is two kernels around that conv:

    stage A : y, means, scales            -> params = cat(y_hat_00, y_hat_11, means, scales)
    (cuDNN) : self.y_spatial_prior(params)
    stage B : y, means, scales, prior_out -> y_hat, likelihood (+ sum ln p)
                                             [+ means_hat, scales_hat | compress planes]

The functions below are written as *methods* (first argument ``self`` is the
reference's context-model instance) so ``deepvideocodec_b200.patch`` can bind
them onto the stock classes; signatures and return structures are the
reference's.
"""
import math

import torch

from . import _native as nat
from .entropy_models import _launch_noise_like, eb_forward

__all__ = ["dual_prior_stage_a", "dual_prior_stage_b_gc", "forward_dual_prior",
           "motion_context_forward", "frame_context_forward",
           "motion_context_compress", "frame_context_compress",
           "motion_context_decompress", "frame_context_decompress"]


def _check_latents(y, means, scales, who):
    for t, nm in ((y, "y"), (means, "means"), (scales, "scales")):
        nat.require_cuda_f32(t, f"{who}({nm})")
    if means.shape != y.shape or scales.shape != y.shape:
        raise nat.DvcError(f"{who}: y, means, scales must have one shape")
    n, c, h, w = y.shape
    if c % 2 or h % 2 or w % 2:
        # get_mask (video_model.py:155) repeats a 2x2 cell: odd sizes break the
        # reference as well
        raise nat.DvcError(f"{who}: C, H, W must be even, got {tuple(y.shape)}")


# ---------------------------------------------------------------------------
# stage A
# ---------------------------------------------------------------------------
def _stage_a_fwd(y, means, scales):
    n, c, h, w = y.shape
    # keyed on the prior means (not on y): the decoder has no y, and both sides must hand
    # y_spatial_prior the same memory format (coder.decode_stage_a)
    cl = means.is_contiguous(memory_format=torch.channels_last) and not means.is_contiguous()
    params = torch.empty((n, 3 * c, h, w), dtype=y.dtype, device=y.device,
                         memory_format=torch.channels_last if cl else torch.contiguous_format)
    with nat.device_of(y):
        rc = nat.lib().dvc_dual_prior_stage_a_fwd(
            y.data_ptr(), means.data_ptr(), scales.data_ptr(), params.data_ptr(), n, c, h, w,
            nat.st4(y), nat.st4(means), nat.st4(scales), nat.st4(params), nat.stream_of(y))
    nat.check(rc, "dvc_dual_prior_stage_a_fwd")
    return params


class _StageAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, means, scales):
        ctx.cshape = tuple(y.shape)
        return _stage_a_fwd(y, means, scales)

    @staticmethod
    def backward(ctx, g_params):
        from .autograd_kernels import stage_a_bwd
        return stage_a_bwd(g_params, ctx.cshape)


def dual_prior_stage_a(y, means, scales):
    """``params = cat(y_hat_0_0, y_hat_1_1, means, scales)`` (video_model.py:176-189)."""
    _check_latents(y, means, scales, "dual_prior_stage_a")
    if torch.is_grad_enabled() and (y.requires_grad or means.requires_grad or scales.requires_grad):
        return _StageAFn.apply(y, means, scales)
    return _stage_a_fwd(y, means, scales)


# ---------------------------------------------------------------------------
# stage B + Gaussian conditional
# ---------------------------------------------------------------------------
def _stage_b_fwd(y, means, scales, prior, noise, scale_bound, lik_bound, want_params, compress):
    n, c, h, w = y.shape
    y_hat = torch.empty_like(y)
    lik = torch.empty_like(y)
    means_hat = torch.empty_like(y) if want_params else None
    scales_hat = torch.empty_like(y) if want_params else None
    planes = [None] * 4
    if compress:
        # the two checkerboard passes are stacked ([2, N, C/2, H, W]) so that one
        # encoder launch codes both (context._compress_tail)
        stacked = torch.empty((2, 2, n, c // 2, h, w), dtype=y.dtype, device=y.device)
        planes = [stacked[0, 0], stacked[0, 1], stacked[1, 0], stacked[1, 1]]
    logsum = torch.empty(n, dtype=torch.float64, device=y.device)
    ws = nat.rate_workspace(y.device, n)
    with nat.device_of(y):
        rc = nat.lib().dvc_dual_prior_stage_b_gc_fwd(
            y.data_ptr(), means.data_ptr(), scales.data_ptr(), prior.data_ptr(), nat.ptr(noise),
            y_hat.data_ptr(), nat.ptr(means_hat), nat.ptr(scales_hat), lik.data_ptr(),
            nat.ptr(planes[0]), nat.ptr(planes[1]), nat.ptr(planes[2]), nat.ptr(planes[3]),
            logsum.data_ptr(), ws.data_ptr(), n, c, h, w,
            nat.st4(y), nat.st4(means), nat.st4(scales), nat.st4(prior), nat.opt_st4(noise),
            nat.st4(y_hat), nat.opt_st4(planes[0]), scale_bound, lik_bound, nat.stream_of(y))
    nat.check(rc, "dvc_dual_prior_stage_b_gc_fwd")
    return y_hat, means_hat, scales_hat, lik, logsum, planes


class _StageBGcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, means, scales, prior, noise, scale_bound, lik_bound, want_params):
        y_hat, means_hat, scales_hat, lik, logsum, _ = _stage_b_fwd(
            y, means, scales, prior, noise, scale_bound, lik_bound, want_params, False)
        ctx.save_for_backward(y, means, scales, prior, noise)
        ctx.bounds = (scale_bound, lik_bound)
        if not want_params:
            means_hat = y.new_empty(0)
            scales_hat = y.new_empty(0)
        return y_hat, means_hat, scales_hat, lik, logsum

    @staticmethod
    def backward(ctx, g_yhat, g_mh, g_sh, g_lik, g_logsum):
        from .autograd_kernels import stage_b_gc_bwd
        y, means, scales, prior, noise = ctx.saved_tensors
        gy, gm, gs, gp = stage_b_gc_bwd(y, means, scales, prior, noise, g_yhat, g_mh, g_sh,
                                        g_lik, g_logsum, ctx.bounds[0], ctx.bounds[1])
        return gy, gm, gs, gp, None, None, None, None


def _gc_bounds(gc):
    """(scale_bound, likelihood_bound) of a GaussianConditional -- ours or the
    real CompressAI module (reads the buffers once and caches the floats)."""
    sb_mod = gc.lower_bound_scale
    lb_mod = gc.likelihood_lower_bound if getattr(gc, "use_likelihood_bound", True) else None
    if hasattr(sb_mod, "value") and (lb_mod is None or hasattr(lb_mod, "value")):
        # our own holders keep a host float that load_state_dict refreshes: nothing to cache
        return sb_mod.value(), (lb_mod.value() if lb_mod is not None else float("-inf"))
    # foreign (real CompressAI) modules: one device read, cached until a bound buffer changes
    bufs = [sb_mod.bound] + ([lb_mod.bound] if lb_mod is not None else [])
    key = tuple((b.data_ptr(), b._version) for b in bufs)
    cached = getattr(gc, "_dvc_bounds", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    sb = float(sb_mod.bound.detach().cpu().reshape(-1)[0])
    lb = float(lb_mod.bound.detach().cpu().reshape(-1)[0]) if lb_mod is not None else float("-inf")
    gc._dvc_bounds = (key, (sb, lb))
    return sb, lb


def dual_prior_stage_b_gc(y, means, scales, prior, gc, training, want_params=False,
                          compress=False):
    """Stage B + merge + Gaussian conditional.  Returns
    ``(y_hat, means_hat, scales_hat, likelihood, planes)``; ``means_hat`` /
    ``scales_hat`` are ``None`` unless ``want_params``; ``planes`` is
    ``(q_w0, q_w1, s_w0, s_w1)`` when ``compress`` else ``None``.  The
    likelihood carries ``._dvc_logsum``."""
    _check_latents(y, means, scales, "dual_prior_stage_b_gc")
    nat.require_cuda_f32(prior, "dual_prior_stage_b_gc(prior)")
    n, c, h, w = y.shape
    if prior.shape != (n, 2 * c, h, w):
        raise nat.DvcError(f"dual_prior_stage_b_gc: spatial prior output must be "
                           f"{(n, 2 * c, h, w)}, got {tuple(prior.shape)}")
    sb, lb = _gc_bounds(gc)
    noise = _launch_noise_like(y) if training else None
    needs_grad = torch.is_grad_enabled() and any(
        t.requires_grad for t in (y, means, scales, prior))
    if needs_grad and not compress:
        y_hat, means_hat, scales_hat, lik, logsum = _StageBGcFn.apply(
            y, means, scales, prior, noise, sb, lb, want_params)
        planes = None
        if not want_params:
            means_hat = scales_hat = None
    else:
        y_hat, means_hat, scales_hat, lik, logsum, planes = _stage_b_fwd(
            y, means, scales, prior, noise, sb, lb, want_params, compress)
        planes = tuple(planes) if compress else None
    lik._dvc_logsum = logsum
    return y_hat, means_hat, scales_hat, lik, planes


# ---------------------------------------------------------------------------
# method drop-ins
# ---------------------------------------------------------------------------
def forward_dual_prior(self, y, means, scales, mode="trainval"):
    """Drop-in for ``forward_dual_prior`` (video_model.py:169-216 == :341-388):
    same arguments, same returns.  The likelihood computed by stage B is kept
    on ``self`` so a following ``gaussian_conditional`` call could reuse it;
    the fused ``*_context_forward`` below skip this method entirely."""
    params = dual_prior_stage_a(y, means, scales)
    prior = self.y_spatial_prior(params)
    compress = mode == "compress"
    y_hat, means_hat, scales_hat, _, planes = dual_prior_stage_b_gc(
        y, means, scales, prior, self.gaussian_conditional, training=False,
        want_params=not compress, compress=compress)
    if compress:
        return (y_hat,) + planes
    return y_hat, means_hat, scales_hat


def _context_tail(self, y, means_hat, scales_hat, z_likelihoods):
    params = dual_prior_stage_a(y, means_hat, scales_hat)
    prior = self.y_spatial_prior(params)
    y_hat, _, _, y_likelihoods, _ = dual_prior_stage_b_gc(
        y, means_hat, scales_hat, prior, self.gaussian_conditional,
        training=self.gaussian_conditional.training)
    return y_hat, {"y": y_likelihoods, "z": z_likelihoods}


def motion_context_forward(self, y, y_ref):
    """Drop-in for ``MotionContextModel.forward`` (video_model.py:218-233)."""
    z = self.hyper_encoder(y)
    _, z_hat, z_likelihoods = eb_forward(self.entropy_bottleneck, z, want_outputs=False,
                                         want_zhat=True)
    params = self.hyper_decoder(z_hat)
    if y_ref is None:
        y_ref = torch.zeros_like(y)
    means_hat, scales_hat = self.y_prior_fusion(torch.cat((params, y_ref), dim=1)).chunk(2, 1)
    return _context_tail(self, y, means_hat, scales_hat, z_likelihoods)


def frame_context_forward(self, y, y_ref, context):
    """Drop-in for ``FrameContextModel.forward`` (video_model.py:390-406)."""
    z = self.hyper_encoder(y)
    _, z_hat, z_likelihoods = eb_forward(self.entropy_bottleneck, z, want_outputs=False,
                                         want_zhat=True)
    params = self.hyper_decoder(z_hat)
    if y_ref is None:
        y_ref = torch.zeros_like(y)
    temporal_params = self.temporal_prior_encoder(context)
    means_hat, scales_hat = self.y_prior_fusion(
        torch.cat((temporal_params, params, y_ref), dim=1)).chunk(2, 1)
    return _context_tail(self, y, means_hat, scales_hat, z_likelihoods)


# ---------------------------------------------------------------------------
# real bit streams: drop-ins for compress / decompress
# of both context models (video_model.py:236-291, :408-466)
# ---------------------------------------------------------------------------
def _coder_tables(module):
    from . import coder
    return coder.Tables(module._quantized_cdf, module._cdf_length, module._offset)


def _medians(eb):
    return eb._get_medians().detach().reshape(1, -1, 1, 1)


def _compress_tail(self, y, z, z_hat, means_hat, scales_hat, z_pending):
    """stage A -> spatial prior -> stage B (compress planes) -> two encoder
    launches with the table look-up fused -> ONE device->host transfer for the
    three strings.  The reference does 2 build_indexes (63 passes each), 3
    compress calls (device->host copy + ``.tolist()`` + CPU coder each) and a
    decompress of z just to obtain z_hat (:238-251)."""
    from . import coder
    gc = self.gaussian_conditional
    params = dual_prior_stage_a(y, means_hat, scales_hat)
    prior = self.y_spatial_prior(params)
    y_hat, _, _, lik, planes = dual_prior_stage_b_gc(
        y, means_hat, scales_hat, prior, gc, training=False, compress=True)
    q_w0, q_w1, s_w0, s_w1 = planes
    # estimated payload of one (sample, checkerboard pass): the likelihood kernel's fused
    # sum(ln p) -- one 8-byte read -- sizes the sub-streams so that the container overhead
    # stays at ~1 % of the bytes written (coder.auto_stream_symbols)
    est_bytes = float(-lik._dvc_logsum.sum().item()) / (math.log(2.0) * 8.0 * 2.0 * y.size(0))
    tables = _coder_tables(gc)
    sb, _ = _gc_bounds(gc)
    n, ch, h, w = q_w0.shape
    # both passes in ONE launch: the stage-B kernel wrote them as one
    # [2, N, C/2, H, W] buffer, coded here as a batch of 2N samples
    q_both, s_both = stacked_planes(planes)
    pending = coder.rans_encode_async(tables, x=q_both, scales=s_both,
                                      scale_table=gc.scale_table, scale_bound=sb,
                                      est_bytes=est_bytes)
    y_strings, z_strings = coder.collect([pending, z_pending])
    return y_hat, {"strings": [y_strings[:n], y_strings[n:], z_strings],
                   "shape": z.size()[-2:]}


def stacked_planes(planes):
    """``(q_w0, q_w1, s_w0, s_w1)`` of stage B -> ``(q [2N,C/2,H,W], s [2N,C/2,H,W])``
    without a copy (they are views of one buffer, see ``_stage_b_fwd``)."""
    q_w0, q_w1, s_w0, s_w1 = planes
    n, ch, h, w = q_w0.shape
    base = q_w0._base
    if base is not None and base.dim() == 6 and s_w1._base is base:
        return base[0].view(2 * n, ch, h, w), base[1].view(2 * n, ch, h, w)
    return torch.cat((q_w0, q_w1), 0), torch.cat((s_w0, s_w1), 0)


def _compress_head(self, y):
    from . import coder
    eb = self.entropy_bottleneck
    z = self.hyper_encoder(y)
    # side stream: the hyper-decoder / prior convolutions run beside the z coder
    z_pending = coder.rans_encode_async(_coder_tables(eb), x=z, means=_medians(eb).expand_as(z),
                                        overlap=True)
    # z_hat = decompress(compress(z)) = round(z - median) + median: the likelihood
    # kernel's z_hat output (bit-identical, tests/test_gpu_coder.py)
    _, z_hat, _ = eb_forward(eb, z, training=False, want_outputs=False, want_zhat=True)
    return z, z_hat, z_pending


def motion_context_compress(self, y, y_ref):
    """Drop-in for ``MotionContextModel.compress`` (video_model.py:236-253)."""
    z, z_hat, z_pending = _compress_head(self, y)
    params = self.hyper_decoder(z_hat)
    if y_ref is None:
        y_ref = torch.zeros_like(y)
    means_hat, scales_hat = self.y_prior_fusion(torch.cat((params, y_ref), dim=1)).chunk(2, 1)
    return _compress_tail(self, y, z, z_hat, means_hat, scales_hat, z_pending)


def frame_context_compress(self, y, y_ref, context):
    """Drop-in for ``FrameContextModel.compress`` (video_model.py:408-427)."""
    z, z_hat, z_pending = _compress_head(self, y)
    params = self.hyper_decoder(z_hat)
    if y_ref is None:
        y_ref = torch.zeros_like(y)
    temporal_params = self.temporal_prior_encoder(context)
    means_hat, scales_hat = self.y_prior_fusion(
        torch.cat((temporal_params, params, y_ref), dim=1)).chunk(2, 1)
    return _compress_tail(self, y, z, z_hat, means_hat, scales_hat, z_pending)


def _decompress_head(self, strings, shape, statuses):
    from . import coder
    assert isinstance(strings, list) and len(strings) == 3
    eb = self.entropy_bottleneck
    tables = _coder_tables(eb)
    out_shape = (len(strings[2]), tables.cdf.size(0), int(shape[0]), int(shape[1]))
    return coder.rans_decode(strings[2], tables, out_shape, means=_medians(eb),
                             device=tables.cdf.device, statuses=statuses)


def _decompress_tail(self, strings, means_hat, scales_hat, statuses):
    """Two decoding passes around the spatial prior (video_model.py:259-289):
    the checkerboard scale planes are read in place by the decoder (no masks, no
    build_indexes tensor), symbols stay int32 on the device, and two
    element-wise kernels replace the ~20 mask multiplies / adds / cats."""
    from . import coder
    gc = self.gaussian_conditional
    tables = _coder_tables(gc)
    sb, _ = _gc_bounds(gc)
    n, c, h, w = means_hat.shape
    _check_latents(means_hat, means_hat, scales_hat, "decompress")
    half = c // 2
    q0 = coder.rans_decode(strings[0], tables, (n, half, h, w), scales=scales_hat[:, :half],
                           scale_table=gc.scale_table, scale_bound=sb, want_symbols=True,
                           cb=(0, half * scales_hat.stride(1)), device=means_hat.device,
                           statuses=statuses)
    params = coder.decode_stage_a(q0, means_hat, scales_hat)
    prior = self.y_spatial_prior(params)
    q1 = coder.rans_decode(strings[1], tables, (n, half, h, w), scales=prior[:, half:c],
                           scale_table=gc.scale_table, scale_bound=sb, want_symbols=True,
                           cb=(1, c * prior.stride(1)), device=means_hat.device,
                           statuses=statuses)
    y_hat = coder.decode_stage_b(q0, q1, means_hat, prior)
    # the three decoder launches and everything between them are queued: ONE host read
    coder.check_decode_status(statuses)
    return y_hat


def motion_context_decompress(self, strings, shape, y_ref):
    """Drop-in for ``MotionContextModel.decompress`` (video_model.py:255-291)."""
    statuses = []
    z_hat = _decompress_head(self, strings, shape, statuses)
    n, _, h, w = z_hat.shape
    params = self.hyper_decoder(z_hat)
    if y_ref is None:
        y_ref = torch.zeros([n, params.size(1) // 2, h * 4, w * 4], device=z_hat.device)
    means_hat, scales_hat = self.y_prior_fusion(torch.cat((params, y_ref), dim=1)).chunk(2, 1)
    return _decompress_tail(self, strings, means_hat, scales_hat, statuses)


def frame_context_decompress(self, strings, shape, y_ref, context):
    """Drop-in for ``FrameContextModel.decompress`` (video_model.py:429-466)."""
    statuses = []
    z_hat = _decompress_head(self, strings, shape, statuses)
    n, _, h, w = z_hat.shape
    params = self.hyper_decoder(z_hat)
    if y_ref is None:
        y_ref = torch.zeros([n, params.size(1) // 2, h * 4, w * 4], device=z_hat.device)
    temporal_params = self.temporal_prior_encoder(context)
    means_hat, scales_hat = self.y_prior_fusion(
        torch.cat((temporal_params, params, y_ref), dim=1)).chunk(2, 1)
    return _decompress_tail(self, strings, means_hat, scales_hat, statuses)
