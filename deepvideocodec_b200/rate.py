"""Drop-in for the rate reduction of the reference's ``dmc/train.py``:

* ``collect_likelihoods_list(likelihoods_list, num_pixels)`` -- train.py:74-93
  (same return: ``(bpp_loss[B], dict)`` with the keys ``bpp_loss.{label}``,
  ``bpp_loss.{label}.{i}.{field}``, ``bpp_loss.{label}.{i}``, ``bpp_loss.{i}``)
* ``frame_bits(likelihoods)`` -- bits of one P-frame

Likelihood tensors produced by this package carry their fused per-sample
``sum(ln p)`` (``._dvc_logsum``, fp64), so no likelihood is re-read from HBM;
a foreign likelihood tensor gets one ``dvc_log_sum_fwd`` launch.
"""
import math
from collections import defaultdict

import torch

from . import _native as nat

__all__ = ["log_sum", "collect_likelihoods_list", "frame_bits", "rate_finalize"]


def _log_sum_fwd(lik):
    n, c, h, w = lik.shape
    logsum = torch.empty(n, dtype=torch.float64, device=lik.device)
    ws = nat.rate_workspace(lik.device, n)
    with nat.device_of(lik):
        rc = nat.lib().dvc_log_sum_fwd(lik.data_ptr(), logsum.data_ptr(), ws.data_ptr(),
                                       n, c, h, w, nat.st4(lik), nat.stream_of(lik))
    nat.check(rc, "dvc_log_sum_fwd")
    return logsum


class _LogSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lik):
        ctx.save_for_backward(lik)
        return _log_sum_fwd(lik)

    @staticmethod
    def backward(ctx, g):
        (lik,) = ctx.saved_tensors
        # d/dp sum(ln p) = 1/p : one elementwise op on a tensor we did not fuse
        return g.to(lik.dtype).view(-1, 1, 1, 1) / lik


def log_sum(lik):
    """``torch.log(lik).sum(dim=(1,2,3))`` in fp64 (``[N]``).  Uses the value
    fused into the producing kernel when there is one."""
    fused = getattr(lik, "_dvc_logsum", None)
    if fused is not None:
        return fused
    nat.require_cuda_f32(lik, "log_sum(likelihood)")
    if torch.is_grad_enabled() and lik.requires_grad:
        return _LogSumFn.apply(lik)
    return _log_sum_fwd(lik)


def rate_finalize(logsums, num_pixels):
    """``logsums`` fp64 ``[K,N]`` -> ``(bpp[K,N] fp32, bpp_total[N] fp32, bits[N] fp64)``
    in one launch (no autograd)."""
    if logsums.dtype != torch.float64 or not logsums.is_cuda or logsums.dim() != 2:
        raise nat.DvcError("rate_finalize: logsums must be a CUDA fp64 [K,N] tensor")
    logsums = logsums.contiguous()
    k, n = logsums.shape
    bpp = torch.empty((k, n), dtype=torch.float32, device=logsums.device)
    total = torch.empty(n, dtype=torch.float32, device=logsums.device)
    bits = torch.empty(n, dtype=torch.float64, device=logsums.device)
    with nat.device_of(logsums):
        rc = nat.lib().dvc_rate_finalize(logsums.data_ptr(), k, n, float(num_pixels),
                                         bpp.data_ptr(), total.data_ptr(), bits.data_ptr(),
                                         nat.stream_of(logsums))
    nat.check(rc, "dvc_rate_finalize")
    return bpp, total, bits


def collect_likelihoods_list(likelihoods_list, num_pixels: int):
    """Same contract as the reference (train.py:74-93)."""
    entries = []          # (frame, label, field, logsum[N])
    for i, frame_likelihoods in enumerate(likelihoods_list):
        for label, likelihoods in frame_likelihoods.items():
            for field, v in likelihoods.items():
                entries.append((i, label, field, log_sum(v)))
    bpp_info_dict = defaultdict(int)
    if not entries:
        return 0, bpp_info_dict
    differentiable = torch.is_grad_enabled() and any(e[3].requires_grad for e in entries)
    stacked = torch.stack([e[3] for e in entries])                      # [K,N] fp64
    if differentiable:
        bpp_all = (stacked / (-math.log(2) * num_pixels)).to(torch.float32)
        bpp_loss = bpp_all.sum(dim=0)
    else:
        bpp_all, bpp_loss, _ = rate_finalize(stacked, num_pixels)
    per_key = bpp_all.sum(dim=1)                                        # bpp.sum() per tensor
    frame_sum = defaultdict(int)
    label_sum = defaultdict(int)
    for k, (i, label, field, _) in enumerate(entries):
        s = per_key[k]
        bpp_info_dict[f"bpp_loss.{label}"] = bpp_info_dict[f"bpp_loss.{label}"] + s
        bpp_info_dict[f"bpp_loss.{label}.{i}.{field}"] = s
        label_sum[(label, i)] = label_sum[(label, i)] + s
        frame_sum[i] = frame_sum[i] + s
    # insertion order of the reference: fields, then the label total, then the frame total
    ordered = defaultdict(int)
    seen_labels = set()
    for i, frame_likelihoods in enumerate(likelihoods_list):
        for label, likelihoods in frame_likelihoods.items():
            if label not in seen_labels:
                ordered[f"bpp_loss.{label}"] = bpp_info_dict[f"bpp_loss.{label}"]
                seen_labels.add(label)
            for field in likelihoods:
                key = f"bpp_loss.{label}.{i}.{field}"
                ordered[key] = bpp_info_dict[key]
            ordered[f"bpp_loss.{label}.{i}"] = label_sum[(label, i)]
        ordered[f"bpp_loss.{i}"] = frame_sum[i]
    return bpp_loss, ordered


def frame_bits(likelihoods):
    """Bits of one P-frame per sample: ``-sum log2 p`` over its likelihood
    tensors -> fp64 ``[N]``."""
    sums = [log_sum(v) for fields in likelihoods.values() for v in fields.values()]
    stacked = torch.stack(sums)
    if torch.is_grad_enabled() and stacked.requires_grad:
        return -stacked.sum(dim=0) / math.log(2)
    return rate_finalize(stacked, 1.0)[2]
