"""Python side of the backward kernels (``dvc_*_bwd`` in include/dvc_b200.h),
called from the ``torch.autograd.Function``s in ``entropy_models.py`` and
``context.py``.  Each function is one launch; gradients that are not required
are not computed (NULL output pointers)."""
import torch

from . import _native as nat


def _c(t):
    """Incoming gradients may arrive as expanded / non-dense views."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.stride(-1) != 0 and all(s != 0 for s in t.stride()) else t.contiguous()


def _logsum_grad(g):
    if g is None:
        return None
    return g.to(torch.float64).contiguous()


def gc_likelihood_bwd(inputs, scales, means, noise, g_out, g_lik, g_logsum, scale_bound,
                      lik_bound, needs):
    """Backward of ``dvc_gc_likelihood_fwd``.  Returns (g_inputs, g_scales, g_means)."""
    n, c, h, w = inputs.shape
    g_lik, g_out, g_logsum = _c(g_lik), _c(g_out), _logsum_grad(g_logsum)
    if noise is None:
        g_out = None                 # eval-mode outputs are not differentiable
    if g_lik is None and g_logsum is None and g_out is None:
        return None, None, None
    gi = torch.empty_like(inputs) if needs[0] else None
    gs = torch.empty_like(inputs) if needs[1] else None
    gm = torch.empty_like(inputs) if (needs[2] and means is not None) else None
    ref = gi if gi is not None else (gs if gs is not None else gm)
    if ref is None:
        return None, None, None
    with nat.device_of(inputs):
        rc = nat.lib().dvc_gc_likelihood_bwd(
            nat.ptr(g_lik), nat.ptr(g_logsum), nat.ptr(g_out), inputs.data_ptr(),
            scales.data_ptr(), nat.ptr(means), nat.ptr(noise), nat.ptr(gi), nat.ptr(gs),
            nat.ptr(gm), n, c, h, w, nat.st4(inputs), nat.st4(scales), nat.opt_st4(means),
            nat.opt_st4(noise), nat.opt_st4(g_lik), nat.opt_st4(g_out), nat.st4(ref),
            scale_bound, lik_bound, nat.stream_of(inputs))
    nat.check(rc, "dvc_gc_likelihood_bwd")
    return gi, gs, gm


def stage_a_bwd(g_params, cshape):
    """Backward of ``dvc_dual_prior_stage_a_fwd``.  Returns (g_y, g_means, g_scales)."""
    n, c, h, w = cshape
    g_params = _c(g_params)
    cl = g_params.is_contiguous(memory_format=torch.channels_last) and not g_params.is_contiguous()
    fmt = torch.channels_last if cl else torch.contiguous_format
    outs = [torch.empty(cshape, dtype=g_params.dtype, device=g_params.device, memory_format=fmt)
            for _ in range(3)]
    with nat.device_of(g_params):
        rc = nat.lib().dvc_dual_prior_stage_a_bwd(
            g_params.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
            n, c, h, w, nat.st4(g_params), nat.st4(outs[0]), nat.stream_of(g_params))
    nat.check(rc, "dvc_dual_prior_stage_a_bwd")
    return tuple(outs)


def stage_b_gc_bwd(y, means, scales, prior, noise, g_yhat, g_mh, g_sh, g_lik, g_logsum,
                   scale_bound, lik_bound):
    """Backward of ``dvc_dual_prior_stage_b_gc_fwd``.
    Returns (g_y, g_means, g_scales, g_prior)."""
    n, c, h, w = y.shape

    def usable(g):
        return None if (g is None or g.numel() == 0) else _c(g).contiguous()

    g_yhat, g_mh, g_sh, g_lik = usable(g_yhat), usable(g_mh), usable(g_sh), usable(g_lik)
    g_logsum = _logsum_grad(g_logsum)
    gin = next((g for g in (g_yhat, g_mh, g_sh, g_lik) if g is not None), None)
    if gin is None and g_logsum is None:
        return None, None, None, None
    gy = torch.empty(y.shape, dtype=y.dtype, device=y.device)
    gm = torch.empty_like(gy)
    gs = torch.empty_like(gy)
    gp = torch.empty(prior.shape, dtype=y.dtype, device=y.device)
    with nat.device_of(y):
        rc = nat.lib().dvc_dual_prior_stage_b_gc_bwd(
            nat.ptr(g_yhat), nat.ptr(g_mh), nat.ptr(g_sh), nat.ptr(g_lik), nat.ptr(g_logsum),
            y.data_ptr(), means.data_ptr(), scales.data_ptr(), prior.data_ptr(), nat.ptr(noise),
            gy.data_ptr(), gm.data_ptr(), gs.data_ptr(), gp.data_ptr(), n, c, h, w,
            nat.st4(y), nat.st4(means), nat.st4(scales), nat.st4(prior), nat.opt_st4(noise),
            nat.opt_st4(gin), nat.st4(gy), nat.st4(gp), scale_bound, lik_bound,
            nat.stream_of(y))
    nat.check(rc, "dvc_dual_prior_stage_b_gc_bwd")
    return gy, gm, gs, gp


def eb_likelihood_bwd(z, noise, mats, bias, fact, med, g_out, g_zhat, g_lik, g_logsum,
                      lik_bound, needs):
    """Backward of ``dvc_eb_likelihood_fwd``.
    Returns (g_z, g_matrices, g_biases, g_factors, g_medians)."""
    n, c, h, w = z.shape

    def usable(g):
        return None if (g is None or g.numel() == 0) else _c(g).contiguous()

    g_out, g_zhat, g_lik = usable(g_out), usable(g_zhat), usable(g_lik)
    if noise is None:
        pass            # eval: d outputs / d z = 0, d outputs / d median = 1 (kernel handles it)
    g_logsum = _logsum_grad(g_logsum)
    gin = next((g for g in (g_out, g_zhat, g_lik) if g is not None), None)
    if gin is None and g_logsum is None:
        return None, None, None, None, None
    gz = torch.empty(z.shape, dtype=z.dtype, device=z.device)
    gmat = torch.empty_like(mats)
    gbias = torch.empty_like(bias)
    gfact = torch.empty_like(fact)
    gmed = torch.empty_like(med)
    with nat.device_of(z):
        rc = nat.lib().dvc_eb_likelihood_bwd(
            nat.ptr(g_out), nat.ptr(g_zhat), nat.ptr(g_lik), nat.ptr(g_logsum), z.data_ptr(),
            nat.ptr(noise), mats.data_ptr(), bias.data_ptr(), fact.data_ptr(), med.data_ptr(),
            gz.data_ptr(), gmat.data_ptr(), gbias.data_ptr(), gfact.data_ptr(), gmed.data_ptr(),
            n, c, h, w, nat.st4(z), nat.opt_st4(noise), nat.opt_st4(gin), nat.st4(gz),
            lik_bound, nat.stream_of(z))
    nat.check(rc, "dvc_eb_likelihood_bwd")
    return gz, gmat, gbias, gfact, gmed
