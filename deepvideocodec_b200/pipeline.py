"""One P-frame of the hot path on resident tensors, with preallocated outputs.

This is the region BASELINE.json's metric is quoted on (SURVEY.md 8d "timed
region"): per P-frame

    flow pyramid (mv -> mv2, mv3; evaluated inside the warp) video_model.py:499-500
    warp x_ref, feat1, feat2, feat3 (ONE launch)            video_model.py:498, 502-504
    for both context models (motion, frame):
        entropy bottleneck likelihood + z_hat               video_model.py:220-224 / 392-396
        dual prior stage A -> spatial-prior conv input      video_model.py:176-189 / 348-361
        dual prior stage B + Gaussian conditional + ln-sum  video_model.py:192-207, 232 / 364-379, 405
    rate finalise (bits, bpp)                               train.py:74-93

The convolutions between those steps are outside the path; their outputs are
inputs here.  ``PFramePath`` binds one set of input tensors, allocates every
output once, pre-builds the C-ABI argument lists and then ``launch()`` only
enqueues kernels: no allocation, no host sync, CUDA-graph capturable.
"""
import ctypes
import math

import torch

from . import _native as nat
from .entropy_models import _packed_cached

__all__ = ["PFramePath", "synthetic_pframe_inputs", "pframe_algorithmic_bytes", "DPB_KEYS",
           "frame_keys", "SpyNetWarps", "spynet_algorithmic_bytes"]

# What the codec keeps on the device between frames: the decoded picture buffer
# (``dpb``, video_model.py:529-534, 544-549).  ``x_ref`` is the previous
# reconstruction, ``feat1..3`` are the feature pyramid extracted from
# ``dpb["feature_ref"]`` (video_model.py:490-504).  Everything else a P-frame
# of the path consumes depends on the NEW frame (motion field, latents, priors).
DPB_KEYS = ("x_ref", "feat1", "feat2", "feat3")


def frame_keys(inputs):
    """Input keys that change with every new frame (everything but the dpb)."""
    return tuple(k for k in inputs if k not in DPB_KEYS)


def spynet_algorithmic_bytes(h, w, n=1):
    """The four 3-channel warps of ``ME_Spynet.forward`` (layers.py:261) at
    H/8, H/4, H/2, H: SURVEY.md 8d's optional second figure (88.78 MB at 1080p)."""
    return sum(4 * n * (h >> k) * (w >> k) * (2 * 3 + 2) for k in range(4))


def pframe_algorithmic_bytes(h, w, n=1, c_feat=64, c_mv=64, c_y=96, c_z=64):
    """Algorithmic bytes of one P-frame: fp32, every operand
    read once and every result written once."""
    def warp(c, hh, ww):
        return 4 * n * hh * ww * (2 * c + 2)
    pyr = 4 * n * 2 * (h * w + h * w // 4) + 4 * n * 2 * (h * w // 4 + h * w // 16)
    lh, lw = h // 16, w // 16
    zh, zw = h // 64, w // 64
    d = {
        "warp_x_ref": warp(3, h, w),
        "warp_ctx1": warp(c_feat, h, w),
        "warp_ctx2": warp(c_feat, h // 2, w // 2),
        "warp_ctx3": warp(c_feat, h // 4, w // 4),
        "flow_pyramid": pyr,
        "gc_dual_prior_motion": 40 * n * c_mv * lh * lw,
        "gc_dual_prior_frame": 40 * n * c_y * lh * lw,
        "eb": 2 * 12 * n * c_z * zh * zw,
    }
    d["total"] = sum(d.values())
    d["warp_multi"] = d["warp_x_ref"] + d["warp_ctx1"] + d["warp_ctx2"] + d["warp_ctx3"]
    return d


def synthetic_pframe_inputs(h, w, device, seed, n=1, regime="smooth", c_feat=64, c_mv=64,
                            c_y=96, c_z=64, layout="channels_last"):
    """Synthetic tensors of SURVEY.md 8d config 2 (shapes of a random-init DMC
    at H x W; values from the controlled distributions, not the degenerate
    random-init activations).  ``layout``: memory format of the three feature
    scales -- ``channels_last`` (the fast path) or ``nchw`` (what the stock
    reference allocates); values are identical in both."""
    g = torch.Generator(device=device).manual_seed(seed)

    def randn(*s):
        return torch.randn(*s, device=device, generator=g)

    if layout not in ("channels_last", "nchw"):
        raise ValueError(layout)
    cl = torch.channels_last if layout == "channels_last" else torch.contiguous_format
    inp = {"x_ref": torch.rand(n, 3, h, w, device=device, generator=g)}
    inp["feat1"] = randn(n, c_feat, h, w).contiguous(memory_format=cl)
    inp["feat2"] = randn(n, c_feat, h // 2, w // 2).contiguous(memory_format=cl)
    inp["feat3"] = randn(n, c_feat, h // 4, w // 4).contiguous(memory_format=cl)
    if regime == "smooth":       # N(0,1) field, 31x31 box low-pass, sigma = 4 px
        f = randn(n, 2, h, w)
        f = torch.nn.functional.avg_pool2d(f, 31, stride=1, padding=15, count_include_pad=False)
        inp["mv"] = (f / f.std() * 4.0).contiguous()
    elif regime == "adversarial":  # i.i.d. N(0, 16^2) px
        inp["mv"] = randn(n, 2, h, w) * 16.0
    else:
        raise ValueError(regime)
    lh, lw, zh, zw = h // 16, w // 16, h // 64, w // 64

    def latents(c):
        mu = randn(n, c, lh, lw) * 3
        sg = torch.exp(torch.empty(n, c, lh, lw, device=device).uniform_(
            math.log(0.05), math.log(32), generator=g))
        return mu, sg

    for name, c in (("motion", c_mv), ("frame", c_y)):
        mu, sg = latents(c)
        inp[f"{name}.means"], inp[f"{name}.scales"] = mu, sg
        inp[f"{name}.y"] = mu + sg * randn(n, c, lh, lw)
        pm0, ps0 = latents(c // 2)
        pm1, ps1 = latents(c // 2)
        inp[f"{name}.prior"] = torch.cat((pm0, ps0, pm1, ps1), dim=1)   # chunk(4,1) layout
        inp[f"{name}.z"] = randn(n, c_z, zh, zw) * 10
    return inp


class PFramePath:
    """Binds one input set; ``launch()`` enqueues the whole P-frame hot path."""

    LABELS = ("motion", "frame")

    def __init__(self, inputs, eb_modules, gc_bounds=(0.11, 1e-9), num_pixels=None,
                 materialize_pyramid=False, outputs=None):
        """``outputs``: optional preallocated tensors for ``warpframe`` /
        ``context1..3`` (same shape and memory format as the matching input).  A
        GOP runner passes the *next* frame's dpb buffers here, so the warped
        frame and contexts of frame t are written straight into the reference
        slots of frame t + 1 (no copy; ``deepvideocodec_b200.gop``)."""
        self.inp = inputs
        outputs = outputs or {}
        # mv2 / mv3 are only intermediates of DMC.motion_compensation
        # (video_model.py:499-500); by default they are derived inside the warp
        # kernel and never written to HBM
        self.materialize_pyramid = materialize_pyramid
        x_ref = inputs["x_ref"]
        self.device = x_ref.device
        n, _, h, w = x_ref.shape
        self.n, self.h, self.w = n, h, w
        self.num_pixels = float(num_pixels if num_pixels is not None else h * w)
        L = nat.lib()
        self._lib = L
        dev = self.device
        o = {}
        o["mv2"] = torch.empty((n, 2, h // 2, w // 2), device=dev)
        o["mv3"] = torch.empty((n, 2, h // 4, w // 4), device=dev)
        def out_like(name, like):
            t = outputs.get(name)
            if t is None:
                return torch.empty_like(like)
            if t.shape != like.shape or t.stride() != like.stride() or t.dtype != like.dtype \
                    or t.device != like.device:
                raise nat.DvcError(f"PFramePath: outputs[{name!r}] must match its input's "
                                   "shape, strides, dtype and device")
            return t

        o["warpframe"] = out_like("warpframe", x_ref)
        for k in (1, 2, 3):
            o[f"context{k}"] = out_like(f"context{k}", inputs[f"feat{k}"])
        # [4 likelihood tensors][N] ln-sums: motion.y, motion.z, frame.y, frame.z (dict order
        # of the reference's likelihood dicts, video_model.py:233, 577-579)
        o["logsums"] = torch.zeros((4, n), dtype=torch.float64, device=dev)
        o["bpp"] = torch.empty((4, n), dtype=torch.float32, device=dev)
        o["bpp_total"] = torch.empty(n, dtype=torch.float32, device=dev)
        o["bits"] = torch.empty(n, dtype=torch.float64, device=dev)
        self._ws = [torch.zeros(L.dvc_rate_workspace_bytes(n), dtype=torch.uint8, device=dev)
                    for _ in range(4)]
        self._eb_params = {}
        for li, name in enumerate(self.LABELS):
            y = inputs[f"{name}.y"]
            z = inputs[f"{name}.z"]
            c = y.size(1)
            o[f"{name}.params"] = torch.empty((n, 3 * c, y.size(2), y.size(3)), device=dev)
            o[f"{name}.y_hat"] = torch.empty_like(y)
            o[f"{name}.y_lik"] = torch.empty_like(y)
            o[f"{name}.z_hat"] = torch.empty_like(z)
            o[f"{name}.z_lik"] = torch.empty_like(z)
            self._eb_params[name] = _packed_cached(eb_modules[name])
        self.out = o
        self._sb, self._lb = float(gc_bounds[0]), float(gc_bounds[1])
        self._build_calls()

    # -- argument lists are built once: launch() is a handful of foreign calls --
    def _build_calls(self):
        i, o, L = self.inp, self.out, self._lib
        n, h, w = self.n, self.h, self.w
        st, P = nat.st4, (lambda t: t.data_ptr())
        self._keep = []

        def keep(x):
            self._keep.append(x)
            return x

        calls = []
        mv = i["mv"]
        if self.materialize_pyramid:
            calls.append((L.dvc_flow_pyramid_fwd, "dvc_flow_pyramid_fwd",
                          (P(mv), P(o["mv2"]), P(o["mv3"]), n, h, w, keep(st(mv)),
                           keep(st(o["mv2"])), keep(st(o["mv3"])))))
        tasks = (nat.WarpTask * 4)()
        # largest task first; the pyramid levels are derived inside the kernel
        for k, (im, level, out) in enumerate((
                (i["feat1"], 0, o["context1"]), (i["feat2"], 1, o["context2"]),
                (i["feat3"], 2, o["context3"]), (i["x_ref"], 0, o["warpframe"]))):
            t = tasks[k]
            t.im, t.flow, t.out = P(im), P(mv), P(out)
            t.N, t.C, t.H, t.W = im.shape
            t.im_st, t.flow_st, t.out_st = st(im), st(mv), st(out)
            t.flow_downscale = level
        keep(tasks)
        self._warp_call = (L.dvc_warp_multi_fwd, "dvc_warp_multi_fwd",
                           (ctypes.cast(tasks, ctypes.c_void_p), 4, 0))
        ent = []
        for li, name in enumerate(self.LABELS):
            y, mu, sg = i[f"{name}.y"], i[f"{name}.means"], i[f"{name}.scales"]
            prior, z = i[f"{name}.prior"], i[f"{name}.z"]
            c, lh, lw = y.size(1), y.size(2), y.size(3)
            mats, bias, fact, med = self._eb_params[name]
            zl = o[f"{name}.z_lik"]
            ls_y = o["logsums"][2 * li]
            ls_z = o["logsums"][2 * li + 1]
            ent.append((L.dvc_eb_likelihood_fwd, "dvc_eb_likelihood_fwd",
                        (P(z), None, P(mats), P(bias), P(fact), P(med), None,
                         P(o[f"{name}.z_hat"]), P(zl), P(ls_z), P(self._ws[2 * li + 1]),
                         n, z.size(1), z.size(2), z.size(3), keep(st(z)), None, keep(st(zl)),
                         self._lb)))
            pr = o[f"{name}.params"]
            ent.append((L.dvc_dual_prior_stage_a_fwd, "dvc_dual_prior_stage_a_fwd",
                        (P(y), P(mu), P(sg), P(pr), n, c, lh, lw, keep(st(y)), keep(st(mu)),
                         keep(st(sg)), keep(st(pr)))))
            yh, yl = o[f"{name}.y_hat"], o[f"{name}.y_lik"]
            ent.append((L.dvc_dual_prior_stage_b_gc_fwd, "dvc_dual_prior_stage_b_gc_fwd",
                        (P(y), P(mu), P(sg), P(prior), None, P(yh), None, None, P(yl),
                         None, None, None, None, P(ls_y), P(self._ws[2 * li]), n, c, lh, lw,
                         keep(st(y)), keep(st(mu)), keep(st(sg)), keep(st(prior)), None,
                         keep(st(yh)), None, self._sb, self._lb)))
        self._fin_head = (P(o["logsums"]), 4, n, self.num_pixels, P(o["bpp"]), P(o["bpp_total"]))
        self._fin_bits = P(o["bits"])
        self._pre_calls = calls
        self._ent_calls = ent

    def _finalize(self, stream, bits_ptr):
        rc = self._lib.dvc_rate_finalize(*self._fin_head, bits_ptr or self._fin_bits, stream)
        if rc:
            nat.check(rc, "dvc_rate_finalize")

    def launch(self, warp_events=None, concurrent=True, bits_ptr=None):
        """Enqueue the P-frame on the current stream of ``self.device``.

        ``bits_ptr``: device address of ``N`` fp64 words that receive this frame's
        bits instead of ``out["bits"]`` (a GOP runner points it at row t of a
        per-unit table, so no frame needs a host read).

        The motion-compensation branch (one launch) and the entropy branch (six
        small launches + the rate finalise) are independent given the conv
        outputs, so with ``concurrent=True`` the entropy branch goes to a
        high-priority side stream, forked from and joined back into the current
        stream with events: its latency-bound kernels then run inside the
        bandwidth-bound warp instead of after it.  ``warp_events=(start, end)``
        records two CUDA events around the dominant kernel (the multi-scale
        warp) for the roofline measurement."""
        main = torch.cuda.current_stream(self.device)
        s = main.cuda_stream
        side = None
        if concurrent:
            side = self._side_stream()
            self._fork.record(main)
            side.wait_event(self._fork)
            s2 = side.cuda_stream
            for fn, name, args in self._ent_calls:
                rc = fn(*args, s2)
                if rc:
                    nat.check(rc, name)
            self._finalize(s2, bits_ptr)
            self._join.record(side)
        for fn, name, args in self._pre_calls:
            rc = fn(*args, s)
            if rc:
                nat.check(rc, name)
        if warp_events is not None:
            warp_events[0].record(main)
        fn, name, args = self._warp_call
        rc = fn(*args, s)
        if rc:
            nat.check(rc, name)
        if warp_events is not None:
            warp_events[1].record(main)
        if concurrent:
            main.wait_event(self._join)
        else:
            for fn, name, args in self._ent_calls:
                rc = fn(*args, s)
                if rc:
                    nat.check(rc, name)
            self._finalize(s, bits_ptr)
        return self.out

    _side = {}

    def _side_stream(self):
        key = self.device.index
        st = PFramePath._side.get(key)
        if st is None:
            import os
            st = torch.cuda.Stream(self.device, priority=int(os.environ.get("DVC_SIDE_PRIORITY", "-1")))
            PFramePath._side[key] = st
        if not hasattr(self, "_fork"):
            self._fork = torch.cuda.Event()
            self._join = torch.cuda.Event()
        return st

    @property
    def n_launches(self):
        """Kernel launches of one ``launch()``.  NCHW (non channels_last) feature
        scales take the staged planar path: one launch for the staged tiles of
        all scales and one for their complement, beside the multi-scale launch
        that keeps the 3-channel frame."""
        planar = any(self.inp[f"feat{k}"].is_contiguous() and self.inp[f"feat{k}"].size(1) >= 8
                     for k in (1, 2, 3))
        return len(self._pre_calls) + 1 + (2 if planar else 0) + len(self._ent_calls) + 1


class SpyNetWarps:
    """The four 3-channel warps of ``ME_Spynet.forward`` (layers.py:242-264, the
    ``flow_warp(im2_list[level], flow_up)`` at :261) as ONE launch: SURVEY.md 8d's
    optional second figure.  Inputs are synthetic pyramids of the reference frame
    and flows of the matching sizes (in the codec each level's flow depends on
    the previous level's conv output, so the four warps cannot be one launch
    there; this measures their cost, not a drop-in)."""

    def __init__(self, h, w, device, seed, n=1):
        g = torch.Generator(device=device).manual_seed(seed)
        self.ims, self.flows, self.outs = [], [], []
        tasks = (nat.WarpTask * 4)()
        self._keep = [tasks]
        for k in range(4):
            hh, ww = h >> (3 - k), w >> (3 - k)
            im = torch.rand(n, 3, hh, ww, device=device, generator=g)
            f = torch.randn(n, 2, hh, ww, device=device, generator=g)
            f = torch.nn.functional.avg_pool2d(f, 15, stride=1, padding=7, count_include_pad=False)
            flow = (f / f.std() * 2.0).contiguous()
            out = torch.empty_like(im)
            t = tasks[k]
            t.im, t.flow, t.out = im.data_ptr(), flow.data_ptr(), out.data_ptr()
            t.N, t.C, t.H, t.W = im.shape
            t.im_st, t.flow_st, t.out_st = nat.st4(im), nat.st4(flow), nat.st4(out)
            t.flow_downscale = 0
            self.ims.append(im), self.flows.append(flow), self.outs.append(out)
        self._args = (ctypes.cast(tasks, ctypes.c_void_p), 4, 0)
        self.device = device
        self.bytes = spynet_algorithmic_bytes(h, w, n)

    def launch(self):
        rc = nat.lib().dvc_warp_multi_fwd(*self._args,
                                          torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            nat.check(rc, "dvc_warp_multi_fwd")
