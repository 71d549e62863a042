"""Drop-in for the reference's ``dmc/models/utils.py:quantize_ste`` (:149-152)."""
import torch

from . import _native as nat

__all__ = ["quantize_ste", "quantize_around"]


def _as4(x):
    """View an arbitrary-rank dense tensor as [1,1,1,numel] / keep 4-D as is."""
    if x.dim() == 4:
        return x, None
    if not x.is_contiguous():
        raise nat.DvcError("quantize_ste: non-4-D inputs must be contiguous")
    return x.reshape(1, 1, 1, -1), x.shape


def _round_fwd(x, offset=None):
    if not x.is_cuda:
        raise nat.DvcError(f"quantize_ste: CUDA tensors only (got {x.device}); no CPU fallback")
    if x.dtype != torch.float32:
        raise nat.DvcError(f"quantize_ste: fp32 only (got {x.dtype})")
    x4, shape = _as4(x)
    q = torch.empty_like(x4)
    n, c, h, w = x4.shape
    off_ptr, off_st = None, 0
    if offset is not None:
        if offset.numel() != c or not offset.is_cuda or offset.dtype != torch.float32:
            raise nat.DvcError("quantize_around: offset must be a CUDA fp32 tensor with C elements")
        flat = offset.reshape(-1)
        off_ptr, off_st = flat.data_ptr(), flat.stride(0)
    with nat.device_of(x4):
        rc = nat.lib().dvc_quantize_fwd(x4.data_ptr(), off_ptr, q.data_ptr(), n, c, h, w,
                                        nat.st4(x4), off_st, nat.st4(q), nat.stream_of(x4))
    nat.check(rc, "dvc_quantize_fwd")
    return q if shape is None else q.reshape(shape)


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _round_fwd(x)

    @staticmethod
    def backward(ctx, g):
        return g          # straight-through: d round(x)/dx := 1


class _RoundAroundSTE(torch.autograd.Function):
    """round(x - m_c) + m_c; dx = g (STE), dm = 0 (the -m and +m cancel)."""

    @staticmethod
    def forward(ctx, x, offset):
        return _round_fwd(x, offset)

    @staticmethod
    def backward(ctx, g):
        return g, None


def quantize_ste(x):
    """``(round(x) - x).detach() + x``: round half to even, identity gradient."""
    if torch.is_grad_enabled() and x.requires_grad:
        return _RoundSTE.apply(x)
    return _round_fwd(x)


def quantize_around(x, offset):
    """``quantize_ste(x - offset) + offset`` with a per-channel ``offset`` (the
    hyper-latent form of video_model.py:222-224), one kernel."""
    if torch.is_grad_enabled() and x.requires_grad:
        return _RoundAroundSTE.apply(x, offset)
    return _round_fwd(x, offset)
