"""Oracle restatement of ``compressai.entropy_models`` (forward + autograd).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The arithmetic follows
the published CompressAI (>= 1.2) ``entropy_models/entropy_models.py``; the
package is absent from ``/root/reference`` so parity here is *unpinned* by the
reference.  Specification: SURVEY.md Appendix A.4 / A.5 / B.  Reference call
sites: ``dmc/models/video_model.py:150,220,222,232,322,392,394,405``,
``dmc/models/base_model.py:63,76``.

The op sequence is kept op-for-op (one torch op per arithmetic step, same
constants, same association) because fp32 cancellation in ``upper - lower``
makes 1e-5-relative likelihood parity order sensitive.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import LowerBound


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder=None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder = entropy_coder
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        # CDF tables for the real entropy coder (names validated by the
        # reference at video_model.py:629-654 / utils.py:112-115)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def forward(self, *args):
        raise NotImplementedError()

    def quantize(self, inputs, mode, means=None):
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            half = float(0.5)
            noise = torch.empty_like(inputs).uniform_(-half, half)
            return inputs + noise
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        return outputs.int()

    # entropy-coding surface: SURVEY.md section 8 rows f1/f2 (not on the hot path)
    def compress(self, *a, **k):
        raise NotImplementedError("oracle shim: real entropy coding is out of scope")

    def decompress(self, *a, **k):
        raise NotImplementedError("oracle shim: real entropy coding is out of scope")


class EntropyBottleneck(EntropyModel):
    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10,
                 filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
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

    def update(self, force=False):
        raise NotImplementedError("oracle shim: CDF tables are out of scope")

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def _logits_cumulative(self, inputs, stop_gradient):
        logits = inputs
        for k in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{k:d}")
            if stop_gradient:
                matrix = matrix.detach()
            logits = torch.matmul(F.softplus(matrix), logits)
            bias = getattr(self, f"_bias{k:d}")
            if stop_gradient:
                bias = bias.detach()
            logits = logits + bias
            if k < len(self.filters):
                factor = getattr(self, f"_factor{k:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def _likelihood(self, inputs):
        half = float(0.5)
        lower = self._logits_cumulative(inputs - half, stop_gradient=False)
        upper = self._logits_cumulative(inputs + half, stop_gradient=False)
        sign = -torch.sign(lower + upper)
        sign = sign.detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def forward(self, x, training=None):
        if training is None:
            training = self.training
        perm = list(range(x.dim()))
        perm[0], perm[1] = perm[1], perm[0]
        inv_perm = perm  # swapping two axes is its own inverse
        x = x.permute(*perm).contiguous()
        shape = x.size()
        values = x.reshape(x.size(0), 1, -1)
        outputs = self.quantize(values, "noise" if training else "dequantize",
                                self._get_medians())
        likelihood = self._likelihood(outputs)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        outputs = outputs.reshape(shape).permute(*inv_perm).contiguous()
        likelihood = likelihood.reshape(shape).permute(*inv_perm).contiguous()
        return outputs, likelihood


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if scale_table and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table)
                            or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer(
            "scale_table",
            torch.Tensor(tuple(float(s) for s in scale_table)) if scale_table else torch.Tensor())
        self.register_buffer(
            "scale_bound",
            torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    def update_scale_table(self, scale_table, force=False):
        raise NotImplementedError("oracle shim: CDF tables are out of scope")

    def build_indexes(self, scales):
        raise NotImplementedError("oracle shim: CDF tables are out of scope")

    def _standardized_cumulative(self, inputs):
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    def _likelihood(self, inputs, scales, means=None):
        half = float(0.5)
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((half - values) / scales)
        lower = self._standardized_cumulative((-half - values) / scales)
        return upper - lower

    def forward(self, inputs, scales, means=None, training=None):
        if training is None:
            training = self.training
        outputs = self.quantize(inputs, "noise" if training else "dequantize", means)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood
