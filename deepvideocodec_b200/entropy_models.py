"""Drop-ins for the CompressAI entropy models the reference constructs:

* ``GaussianConditional(None)``  -- video_model.py:150, :322; called :232, :405
* ``EntropyBottleneck(channels)`` -- base_model.py:63; called video_model.py:220, :392;
  ``_get_medians()`` :222, :394; ``loss()`` base_model.py:76

Same constructor arguments, parameter/buffer names (SURVEY.md Appendix B --
``DMC.load_state_dict`` validates them, video_model.py:626-656), call signature
``module(x, ...) -> (outputs, likelihoods)`` and training/eval semantics.  The
forward of each module is ONE kernel of ``libdvc_b200.so`` (the eager original
is ~20 resp. ~81 launches); the returned likelihood tensor also carries the
fused per-sample ``sum(ln p)`` as ``likelihoods._dvc_logsum`` which
``deepvideocodec_b200.rate.collect_likelihoods_list`` consumes so the rate
never re-reads the likelihoods from HBM.

CUDA fp32 tensors only.  There is no CPU path: a CPU tensor raises.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _native as nat

__all__ = ["EntropyModel", "EntropyBottleneck", "GaussianConditional", "LowerBound"]


# ---------------------------------------------------------------------------
# LowerBound (CompressAI ops.bound_ops) -- kept for state_dict compatibility
# (``*.likelihood_lower_bound.bound``, ``*.lower_bound_scale.bound``).  The fused
# kernels apply the bounds themselves; this module is only a holder for them.
# ---------------------------------------------------------------------------
class LowerBound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))
        self._bound_f = float(np.float32(bound))

    def value(self) -> float:
        return self._bound_f

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._bound_f = float(self.bound.detach().cpu().reshape(-1)[0])

    def forward(self, x):
        raise nat.DvcError(
            "LowerBound is fused into the likelihood kernels of deepvideocodec_b200; "
            "call GaussianConditional / EntropyBottleneck instead")


def _launch_noise_like(x):
    # CompressAI: torch.empty_like(inputs).uniform_(-0.5, 0.5) from torch's generator
    return torch.empty_like(x).uniform_(-0.5, 0.5)


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder=None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder = entropy_coder
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def _lik_bound(self) -> float:
        # a bound of -inf disables the clamp inside the kernels
        return self.likelihood_lower_bound.value() if self.use_likelihood_bound else -math.inf

    def forward(self, *args):
        raise NotImplementedError()

    def quantize(self, inputs, mode, means=None):
        from .utils import _round_fwd
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return inputs + _launch_noise_like(inputs)
        shifted = inputs if means is None else inputs - means
        q = _round_fwd(shifted.contiguous())
        if mode == "dequantize":
            return q if means is None else q + means
        return q.int()

    # ---- real entropy coding: GPU rANS, csrc/dvc_coder.cu
    @staticmethod
    def dequantize(inputs, means=None, dtype=torch.float):
        outputs = inputs.type_as(means) if means is not None else inputs.type(dtype)
        return outputs if means is None else outputs + means

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """Rows of ``pmf`` (+ their tail mass) -> int32 ``[n, max_length + 2]``
        quantised CDFs (host, setup time: ``dvc_pmf_to_quantized_cdf``)."""
        from . import coder
        lengths = [int(v) for v in pmf_length.reshape(-1).tolist()]
        rows = pmf.detach().to("cpu", torch.float32).numpy()
        tails = tail_mass.detach().to("cpu", torch.float32).numpy().reshape(len(lengths), -1)
        cdf = np.zeros((len(lengths), int(max_length) + 2), dtype=np.int32)
        for i, n in enumerate(lengths):
            q = coder.pmf_to_quantized_cdf(np.concatenate((rows[i, :n], tails[i, :1])),
                                           self.entropy_coder_precision)
            cdf[i, :q.size] = q
        return torch.from_numpy(cdf).to(pmf.device)

    def _tables(self):
        from . import coder
        return coder.Tables(self._quantized_cdf, self._cdf_length, self._offset)

    def compress(self, inputs, indexes, means=None):
        """``EntropyModel.compress``: one ``bytes`` per batch sample.  Symbols
        (``round(inputs - means)``), table look-ups and the coder run on the GPU;
        only the finished streams cross to the host."""
        from . import coder
        if inputs.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        tables = self._tables()
        shape = inputs.shape
        x, idx, mu = _as4d(inputs), _as4d(indexes), None
        if means is not None:
            mu = _as4d(means.expand(shape))
        return coder.rans_encode(tables, x=x, means=mu, indexes=idx)

    def decompress(self, strings, indexes, dtype=torch.float, means=None):
        from . import coder
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if len(strings) != indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        tables = self._tables()
        shape = indexes.shape
        if means is not None:
            if means.shape[:2] != shape[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.shape != shape and any(means.size(i) != 1 for i in range(2, len(shape))):
                raise ValueError("Invalid means parameters")
            means = _as4d(means.expand(shape))
        idx = _as4d(indexes)
        out = coder.rans_decode(strings, tables, idx.shape, indexes=idx, means=means,
                                device=tables.cdf.device)
        out = out.reshape(shape)
        return out if dtype in (torch.float, None) or means is not None else out.type(dtype)


def _as4d(t):
    """View a ``[N, C, *spatial]`` tensor as ``[N, C, H, W]`` (the coder's
    NCHW enumeration equals the flattened order for any number of spatial dims)."""
    if t.dim() == 4:
        return t
    if t.dim() < 2:
        raise ValueError("expected a tensor with at least 2 dimensions")
    return t.reshape(t.size(0), t.size(1), 1, -1)


# ---------------------------------------------------------------------------
# Gaussian conditional
# ---------------------------------------------------------------------------
def _gc_fwd(inputs, scales, means, noise, scale_bound, lik_bound, want_outputs=True):
    n, c, h, w = inputs.shape
    outputs = torch.empty_like(inputs) if want_outputs else None
    lik = torch.empty_like(inputs)
    logsum = torch.empty(n, dtype=torch.float64, device=inputs.device)
    ws = nat.rate_workspace(inputs.device, n)
    with nat.device_of(inputs):
        rc = nat.lib().dvc_gc_likelihood_fwd(
            inputs.data_ptr(), scales.data_ptr(), nat.ptr(means), nat.ptr(noise),
            nat.ptr(outputs), lik.data_ptr(), logsum.data_ptr(), ws.data_ptr(), n, c, h, w,
            nat.st4(inputs), nat.st4(scales), nat.opt_st4(means), nat.opt_st4(noise),
            nat.st4(lik), scale_bound, lik_bound, nat.stream_of(inputs))
    nat.check(rc, "dvc_gc_likelihood_fwd")
    return outputs, lik, logsum


class _GaussianConditionalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, scales, means, noise, scale_bound, lik_bound):
        outputs, lik, logsum = _gc_fwd(inputs, scales, means, noise, scale_bound, lik_bound)
        ctx.save_for_backward(inputs, scales, means, noise)
        ctx.bounds = (scale_bound, lik_bound)
        ctx.mark_non_differentiable(outputs) if noise is None else None
        return outputs, lik, logsum

    @staticmethod
    def backward(ctx, g_out, g_lik, g_logsum):
        inputs, scales, means, noise = ctx.saved_tensors
        return _gc_bwd(ctx, inputs, scales, means, noise, g_out, g_lik, g_logsum)


def _gc_bwd(ctx, inputs, scales, means, noise, g_out, g_lik, g_logsum):
    from .autograd_kernels import gc_likelihood_bwd
    gi, gs, gm = gc_likelihood_bwd(inputs, scales, means, noise, g_out, g_lik, g_logsum,
                                   ctx.bounds[0], ctx.bounds[1], ctx.needs_input_grad[:3])
    return gi, gs, gm, None, None, None


class GaussianConditional(EntropyModel):
    """``GaussianConditional(scale_table, scale_bound=0.11, tail_mass=1e-9)``."""

    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if scale_table and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer(
            "scale_table",
            torch.Tensor(tuple(float(s) for s in scale_table)) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_cumulative(inputs):
        return 0.5 * torch.erfc(-(2 ** -0.5) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        import scipy.stats
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force=False):
        """CDF tables of the scale table (``DMC.update``, video_model.py:669-677).
        Setup time, 64 rows: plain torch for the pmf, the library for the CDFs."""
        if self._offset.numel() > 0 and not force:
            return False
        self.scale_table = self._prepare_scale_table(scale_table).to(self.scale_table.device)
        self.update()
        return True

    def update(self):
        reach = -self._standardized_quantile(self.tail_mass / 2)
        centre = torch.ceil(self.scale_table * reach).int()
        length = 2 * centre + 1
        widest = int(length.max().item())
        k = torch.arange(widest, device=centre.device).int()
        dist = torch.abs(k - centre[:, None]).float()
        sigma = self.scale_table.unsqueeze(1).float()
        upper = self._standardized_cumulative((0.5 - dist) / sigma)
        lower = self._standardized_cumulative((-0.5 - dist) / sigma)
        self._quantized_cdf = self._pmf_to_cdf(upper - lower, 2 * lower[:, :1], length, widest)
        self._offset = -centre
        self._cdf_length = length + 2

    def build_indexes(self, scales):
        """int32 table index per element -- one launch (``dvc_symbols_indexes_fwd``)
        instead of the original's ``len(scale_table) - 1`` compare passes."""
        from . import coder
        return coder.build_indexes(scales, self.scale_table, self.lower_bound_scale.value())

    def forward(self, inputs, scales, means=None, training=None):
        if training is None:
            training = self.training
        inputs = nat.require_cuda_f32(inputs, "GaussianConditional(inputs)")
        scales = nat.require_cuda_f32(scales, "GaussianConditional(scales)")
        if scales.shape != inputs.shape:
            raise nat.DvcError("GaussianConditional: scales must have the shape of inputs")
        if means is not None:
            means = nat.require_cuda_f32(means, "GaussianConditional(means)")
            if means.shape != inputs.shape:
                raise nat.DvcError("GaussianConditional: means must have the shape of inputs")
        noise = _launch_noise_like(inputs) if training else None
        sb, lb = self.lower_bound_scale.value(), self._lik_bound()
        needs_grad = torch.is_grad_enabled() and any(
            t is not None and t.requires_grad for t in (inputs, scales, means))
        if needs_grad:
            outputs, lik, logsum = _GaussianConditionalFn.apply(inputs, scales, means, noise, sb, lb)
        else:
            outputs, lik, logsum = _gc_fwd(inputs, scales, means, noise, sb, lb)
        lik._dvc_logsum = logsum
        return outputs, lik


# ---------------------------------------------------------------------------
# factorised entropy bottleneck
# ---------------------------------------------------------------------------
def pack_eb_params(eb):
    """Concatenate the per-layer parameters into the kernel's per-channel
    layout: matrices [C,33], biases [C,13], factors [C,12], medians [C].
    Differentiable (plain ``torch.cat``), so gradients reach the module's own
    parameters.  Works on any module with CompressAI's attribute names."""
    c = eb._matrix0.size(0)
    mats = torch.cat([getattr(eb, f"_matrix{k}").reshape(c, -1) for k in range(5)], dim=1)
    bias = torch.cat([getattr(eb, f"_bias{k}").reshape(c, -1) for k in range(5)], dim=1)
    fact = torch.cat([getattr(eb, f"_factor{k}").reshape(c, -1) for k in range(4)], dim=1)
    med = eb.quantiles[:, 0, 1]
    if mats.size(1) != 33 or bias.size(1) != 13 or fact.size(1) != 12:
        raise nat.DvcError("EntropyBottleneck kernel supports filters=(3,3,3,3) only")
    return mats.contiguous(), bias.contiguous(), fact.contiguous(), med.contiguous()


def _packed_cached(eb):
    """In no-grad mode the packing is cached until a parameter changes."""
    names = [f"_matrix{k}" for k in range(5)] + [f"_bias{k}" for k in range(5)] + \
            [f"_factor{k}" for k in range(4)] + ["quantiles"]
    key = tuple((getattr(eb, n).data_ptr(), getattr(eb, n)._version) for n in names)
    cache = getattr(eb, "_dvc_packed", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, tuple(t.detach() for t in pack_eb_params(eb)))
        eb._dvc_packed = cache
    return cache[1]


def _eb_fwd(z, noise, mats, bias, fact, med, lik_bound, want_outputs, want_zhat):
    n, c, h, w = z.shape
    outputs = torch.empty_like(z) if want_outputs else None
    z_hat = torch.empty_like(z) if want_zhat else None
    lik = torch.empty_like(z)
    logsum = torch.empty(n, dtype=torch.float64, device=z.device)
    ws = nat.rate_workspace(z.device, n)
    with nat.device_of(z):
        rc = nat.lib().dvc_eb_likelihood_fwd(
            z.data_ptr(), nat.ptr(noise), mats.data_ptr(), bias.data_ptr(), fact.data_ptr(),
            med.data_ptr(), nat.ptr(outputs), nat.ptr(z_hat), lik.data_ptr(), logsum.data_ptr(),
            ws.data_ptr(), n, c, h, w, nat.st4(z), nat.opt_st4(noise), nat.st4(lik),
            lik_bound, nat.stream_of(z))
    nat.check(rc, "dvc_eb_likelihood_fwd")
    return outputs, z_hat, lik, logsum


class _EntropyBottleneckFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, noise, mats, bias, fact, med, lik_bound, want_zhat):
        outputs, z_hat, lik, logsum = _eb_fwd(z, noise, mats, bias, fact, med, lik_bound,
                                              True, want_zhat)
        ctx.save_for_backward(z, noise, mats, bias, fact, med)
        ctx.lik_bound = lik_bound
        if z_hat is None:
            z_hat = z.new_empty(0)
        return outputs, z_hat, lik, logsum

    @staticmethod
    def backward(ctx, g_out, g_zhat, g_lik, g_logsum):
        from .autograd_kernels import eb_likelihood_bwd
        z, noise, mats, bias, fact, med = ctx.saved_tensors
        gz, gm, gb, gf, gmed = eb_likelihood_bwd(z, noise, mats, bias, fact, med, g_out, g_zhat,
                                                 g_lik, g_logsum, ctx.lik_bound,
                                                 ctx.needs_input_grad)
        return gz, None, gm, gb, gf, gmed, None, None


def eb_forward(eb, z, training=None, want_outputs=True, want_zhat=False):
    """Kernel entry shared by the module and the fused context models.
    Returns ``(outputs, z_hat, likelihood)``; works for any module exposing
    CompressAI's EntropyBottleneck attribute names."""
    if training is None:
        training = eb.training
    z = nat.require_cuda_f32(z, "EntropyBottleneck(x)")
    if z.size(1) != eb._matrix0.size(0):
        raise nat.DvcError(f"EntropyBottleneck: expected {eb._matrix0.size(0)} channels, "
                           f"got {z.size(1)}")
    bound = getattr(eb, "likelihood_lower_bound", None)
    if not getattr(eb, "use_likelihood_bound", True) or bound is None:
        lb = -math.inf
    elif hasattr(bound, "value"):
        lb = bound.value()
    else:                                   # real CompressAI LowerBound module
        key = (bound.bound.data_ptr(), bound.bound._version)
        cached = getattr(eb, "_dvc_lik_bound", None)
        if cached is None or cached[0] != key:
            cached = (key, float(bound.bound.detach().cpu().reshape(-1)[0]))
            eb._dvc_lik_bound = cached
        lb = cached[1]
    noise = None
    if training:
        # CompressAI draws the noise on its permuted [C, 1, N*H*W] view; drawing it
        # in that layout and viewing it back as [N,C,H,W] (the kernel takes
        # strides) consumes torch's generator exactly like the reference does
        n_, c_, h_, w_ = z.shape
        noise = torch.empty((c_, n_, h_, w_), dtype=z.dtype, device=z.device).uniform_(
            -0.5, 0.5).permute(1, 0, 2, 3)
    params = [getattr(eb, f"_matrix{k}") for k in range(5)] + \
             [getattr(eb, f"_bias{k}") for k in range(5)] + \
             [getattr(eb, f"_factor{k}") for k in range(4)] + [eb.quantiles]
    needs_grad = torch.is_grad_enabled() and (
        z.requires_grad or any(p.requires_grad for p in params))
    if needs_grad:
        mats, bias, fact, med = pack_eb_params(eb)
        outputs, z_hat, lik, logsum = _EntropyBottleneckFn.apply(
            z, noise, mats, bias, fact, med, lb, want_zhat)
        if not want_zhat:
            z_hat = None
    else:
        mats, bias, fact, med = _packed_cached(eb)
        outputs, z_hat, lik, logsum = _eb_fwd(z, noise, mats, bias, fact, med, lb,
                                              want_outputs, want_zhat)
    lik._dvc_logsum = logsum
    return outputs, z_hat, lik


class EntropyBottleneck(EntropyModel):
    """``EntropyBottleneck(channels, tail_mass=1e-9, init_scale=10, filters=(3,3,3,3))``."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3),
                 **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise NotImplementedError("deepvideocodec_b200 EntropyBottleneck: filters=(3,3,3,3) "
                                      "only (the sole configuration the reference constructs)")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        dims = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for k in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / dims[k + 1]))
            matrix = torch.Tensor(self.channels, dims[k + 1], dims[k])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{k:d}", nn.Parameter(matrix))
            bias = torch.Tensor(self.channels, dims[k + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{k:d}", nn.Parameter(bias))
            if k < len(self.filters):
                factor = torch.Tensor(self.channels, dims[k + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{k:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(self.channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, inputs, stop_gradient=True):
        """Plain torch, used only on parameter-sized tensors (``loss``: the
        ``[C,1,3]`` quantiles; ``update``: ``[C,1,<=~60]`` table samples)."""
        import torch.nn.functional as F
        pick = (lambda p: p.detach()) if stop_gradient else (lambda p: p)
        logits = inputs
        for k in range(5):
            logits = torch.matmul(F.softplus(pick(getattr(self, f"_matrix{k}"))), logits)
            logits = logits + pick(getattr(self, f"_bias{k}"))
            if k < 4:
                logits = logits + torch.tanh(pick(getattr(self, f"_factor{k}"))) * torch.tanh(logits)
        return logits

    def update(self, force=False):
        """Per-channel CDF tables from the learned quantiles (``DMC.update``,
        video_model.py:669-677).  Setup time."""
        if self._offset.numel() > 0 and not force:
            return False
        med = self.quantiles[:, 0, 1]
        below = torch.clamp(torch.ceil(med - self.quantiles[:, 0, 0]).int(), min=0)
        above = torch.clamp(torch.ceil(self.quantiles[:, 0, 2] - med).int(), min=0)
        self._offset = -below
        first = med - below
        length = above + below + 1
        widest = int(length.max().item())
        samples = torch.arange(widest, device=first.device)[None, :] + first[:, None, None]
        lower = self._logits_cumulative(samples - 0.5)
        upper = self._logits_cumulative(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail, length, widest)
        self._cdf_length = length + 2
        return True

    def _build_indexes(self, size):
        shape = [1] * len(size)
        shape[1] = -1
        return torch.arange(size[1]).view(*shape).int().repeat(size[0], 1, *size[2:])

    def compress(self, x):
        """``EntropyBottleneck.compress(x)`` (video_model.py:238, :411): the table
        index is the channel and the mean the channel median -- neither is
        materialised; the kernel derives both."""
        from . import coder
        med = self._get_medians().detach().reshape(1, -1, 1, 1)
        x4 = _as4d(x)
        return coder.rans_encode(self._tables(), x=x4, means=med.expand(x4.shape))

    def decompress(self, strings, size):
        from . import coder
        tables = self._tables()
        shape = (len(strings), tables.cdf.size(0), *size)
        med = self._get_medians().detach().reshape(1, -1, 1, 1)
        n, c = shape[:2]
        flat = (n, c, 1, int(np.prod(shape[2:]))) if len(shape) != 4 else shape
        out = coder.rans_decode(strings, tables, flat, means=med.to(tables.cdf.device),
                                device=tables.cdf.device)
        return out.reshape(shape)

    def loss(self):
        """Auxiliary quantile loss (base_model.py:71-78, train.py:336).  It touches
        only the [C,1,3] quantiles and the (detached) per-channel parameters --
        3*C scalars, no data tensor -- so it is plain torch, not a kernel."""
        return torch.abs(self._logits_cumulative(self.quantiles) - self.target).sum()

    def forward(self, x, training=None):
        outputs, _, lik = eb_forward(self, x, training=training, want_outputs=True)
        return outputs, lik
