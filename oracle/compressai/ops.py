"""``LowerBound`` (CompressAI ``ops/bound_ops.py``), restated.

forward  : max(x, bound)
backward : gradient passes where (x >= bound) or (grad_output < 0); the bound
           itself receives no gradient.
"""
import torch
import torch.nn as nn


class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through = (x >= bound) | (grad_output < 0)
        return pass_through.to(grad_output.dtype) * grad_output, None


class LowerBound(nn.Module):
    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)
