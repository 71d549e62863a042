"""GPU gradient parity: analytic backward kernels vs the reference autograd
graph (oracle ops executed by PyTorch-CUDA eager) on identical inputs.
Tolerance: 1e-4 relative to the gradient's scale."""
import math

import pytest
import torch

from test_gpu_entropy import _latents, _oracle_entropy_models

pytestmark = pytest.mark.gpu

GRAD_RTOL = 1e-4


def _close(a, b, what, rtol=GRAD_RTOL):
    assert a.shape == b.shape, what
    scale = b.abs().max().item()
    err = (a - b).abs().max().item()
    assert err <= rtol * max(scale, 1e-12), f"{what}: max err {err} vs scale {scale}"


def _smooth_flow(n, h, w, sigma, dev, g):
    f = torch.randn(n, 2, h, w, device=dev, generator=g)
    f = torch.nn.functional.avg_pool2d(f, 9, stride=1, padding=4, count_include_pad=False)
    return (f / f.std() * sigma).contiguous()


@pytest.mark.parametrize("shape", [(2, 3, 24, 40), (1, 64, 32, 48), (1, 8, 17, 23)])
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_flow_warp_backward(cuda_dev, shape, layout):
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    n, c, h, w = shape
    g = torch.Generator(device=cuda_dev).manual_seed(51)
    im = torch.randn(n, c, h, w, device=cuda_dev, generator=g)
    flow = _smooth_flow(n, h, w, 3.0, cuda_dev, g)
    flow[0, :, 0, 0] = 100.0           # clamped: zero gradient through the border clip
    flow[0, :, 1, 1] = -100.0
    if layout == "nhwc":
        im = im.contiguous(memory_format=torch.channels_last)
    gout = torch.randn(n, c, h, w, device=cuda_dev, generator=g)
    res = []
    for fn in (dmc_ref.flow_warp, dvc.flow_warp):
        a = im.clone().requires_grad_(True)
        b = flow.clone().requires_grad_(True)
        out = fn(a, b)
        out.backward(gout)
        res.append((out.detach(), a.grad, b.grad))
    assert torch.equal(res[0][0], res[1][0])
    _close(res[1][1], res[0][1], "grad_im")
    _close(res[1][2], res[0][2], "grad_flow")
    assert res[1][2][0, :, 0, 0].abs().max().item() == 0.0


@pytest.mark.parametrize("shape", [(1, 2, 32, 48), (2, 6, 19, 27)])
def test_bilinear_down_backward(cuda_dev, shape):
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    g = torch.Generator(device=cuda_dev).manual_seed(52)
    x = torch.randn(*shape, device=cuda_dev, generator=g)
    a = x.clone().requires_grad_(True)
    b = x.clone().requires_grad_(True)
    ya = dmc_ref.bilinear_down2(a) / 2
    yb = dvc.bilineardownsacling(b, post_scale=0.5)
    gy = torch.randn_like(ya)
    ya.backward(gy)
    yb.backward(gy)
    _close(b.grad, a.grad, "grad_x", 1e-6)


def test_motion_compensation_backward(cuda_dev):
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    g = torch.Generator(device=cuda_dev).manual_seed(53)
    h, w = 32, 48
    base = [torch.rand(2, 3, h, w, device=cuda_dev, generator=g),
            torch.randn(2, 64, h, w, device=cuda_dev, generator=g),
            torch.randn(2, 64, h // 2, w // 2, device=cuda_dev, generator=g),
            torch.randn(2, 64, h // 4, w // 4, device=cuda_dev, generator=g),
            _smooth_flow(2, h, w, 2.0, cuda_dev, g)]
    grads = []
    for fn in (dmc_ref.motion_compensation_warps, dvc.motion_compensation_warps):
        leaves = [t.clone().requires_grad_(True) for t in base]
        outs = fn(*leaves)
        loss = sum((o * o).sum() for o in outs)
        loss.backward()
        grads.append([t.grad for t in leaves])
    for a, b, nm in zip(grads[1], grads[0], ("x_ref", "f1", "f2", "f3", "mv")):
        _close(a, b, nm)


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("with_means", [True, False])
def test_gaussian_conditional_backward(cuda_dev, training, with_means):
    import deepvideocodec_b200 as dvc
    oem = _oracle_entropy_models()
    g = torch.Generator(device=cuda_dev).manual_seed(54)
    y, mu, sg = _latents(2, 16, 8, 12, cuda_dev, g)
    sg[0, 0, 0, :4] = torch.tensor([0.0, -1.0, 0.05, 0.2], device=cuda_dev)   # LowerBound(scale) cases
    mods = (oem.GaussianConditional(None).to(cuda_dev), dvc.GaussianConditional(None).to(cuda_dev))
    res = []
    for mod in mods:
        mod.train(training)
        a, b, c = (t.clone().requires_grad_(True) for t in (y, sg, mu))
        torch.manual_seed(9)
        out, lik = mod(a, b, c if with_means else None)
        # rate-like loss (drives most likelihoods towards the 1e-9 floor rule) + a direct term
        loss = -torch.log2(lik).sum() + (lik * lik).sum()
        if training:
            loss = loss + (out * 0.01).sum()
        loss.backward()
        res.append((a.grad, b.grad, c.grad if with_means else None))
    _close(res[1][1], res[0][1], "grad_scales")
    if training:
        _close(res[1][0], res[0][0], "grad_inputs")
        if with_means:
            _close(res[1][2], res[0][2], "grad_means")
    else:
        assert res[1][0] is None or res[1][0].abs().max().item() == 0.0
        assert res[0][0].abs().max().item() == 0.0


def test_gaussian_conditional_fused_logsum_gradient(cuda_dev):
    """The rate taken from the fused per-sample ln-sum must back-propagate like
    log(lik).sum() of the reference (train.py:83)."""
    import deepvideocodec_b200 as dvc
    oem = _oracle_entropy_models()
    from oracle import dmc_ref
    g = torch.Generator(device=cuda_dev).manual_seed(55)
    y, mu, sg = _latents(2, 16, 8, 12, cuda_dev, g)
    ref_mod = oem.GaussianConditional(None).to(cuda_dev).train()
    mod = dvc.GaussianConditional(None).to(cuda_dev).train()
    res = []
    for m, collect in ((ref_mod, dmc_ref.collect_likelihoods_list), (mod, dvc.collect_likelihoods_list)):
        a, b, c = (t.clone().requires_grad_(True) for t in (y, sg, mu))
        torch.manual_seed(3)
        _, lik = m(a, b, c)
        bpp, info = collect([{"motion": {"y": lik}}], 8 * 12 * 256)
        bpp.mean().backward()
        res.append((bpp.detach(), a.grad, b.grad, c.grad))
    _close(res[1][0], res[0][0], "bpp")
    for k, nm in ((1, "grad_inputs"), (2, "grad_scales"), (3, "grad_means")):
        _close(res[1][k], res[0][k], nm)


@pytest.mark.parametrize("training", [True, False])
def test_entropy_bottleneck_backward(cuda_dev, training):
    import deepvideocodec_b200 as dvc
    oem = _oracle_entropy_models()
    torch.manual_seed(56)
    ref_mod = oem.EntropyBottleneck(8).to(cuda_dev)
    with torch.no_grad():
        for name, p in ref_mod.named_parameters():
            if name.startswith("_factor"):
                p.uniform_(-0.8, 0.8)
            elif name.startswith("_matrix"):
                p.add_(torch.randn_like(p) * 0.3)
            elif name == "quantiles":
                p[:, 0, 1] = torch.randn(8, device=cuda_dev)
    mod = dvc.EntropyBottleneck(8).to(cuda_dev)
    mod.load_state_dict(ref_mod.state_dict())
    z = torch.randn(3, 8, 5, 7, device=cuda_dev) * 4
    res = []
    for m in (ref_mod, mod):
        m.train(training)
        m.zero_grad()
        a = z.clone().requires_grad_(True)
        torch.manual_seed(4)
        out, lik = m(a)
        loss = -torch.log2(lik).sum() + (lik * lik).sum()
        if training:
            loss = loss + (out * 0.01).sum()
        loss.backward()
        res.append((a.grad, {n: p.grad for n, p in m.named_parameters()}, lik.detach()))
    assert ((res[1][2] - res[0][2]).abs() / res[0][2]).max().item() <= 5e-5
    if training:
        _close(res[1][0], res[0][0], "grad_z")
    for name, gref in res[0][1].items():
        if gref is None:
            continue
        gmine = res[1][1][name]
        assert gmine is not None, name
        _close(gmine, gref, name, 2e-4)


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("fmt", ["nchw", "nhwc"])
def test_context_model_tail_backward(cuda_dev, training, fmt, monkeypatch):
    """dual prior (stage A -> conv -> stage B) + Gaussian conditional + rate,
    against reference ops + CompressAI restatement, gradients w.r.t. the
    latent, both priors and the conv weights."""
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200 import context as ctxmod
    from oracle import dmc_ref
    oem = _oracle_entropy_models()
    n, c, h, w = 2, 16, 8, 12
    g = torch.Generator(device=cuda_dev).manual_seed(57)
    mf = torch.channels_last if fmt == "nhwc" else torch.contiguous_format
    y, mu, sg = _latents(n, c, h, w, cuda_dev, g, mf)
    torch.manual_seed(58)
    conv = torch.nn.Conv2d(3 * c, 2 * c, 3, padding=1).to(cuda_dev)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)      # restored after the test
    gc_ref = oem.GaussianConditional(None).to(cuda_dev).train(training)
    gc = dvc.GaussianConditional(None).to(cuda_dev).train(training)

    class Holder:
        pass

    res = []
    for impl in ("ref", "dvc"):
        conv.zero_grad()
        a, b, cc = (t.clone().requires_grad_(True) for t in (y, mu, sg))
        torch.manual_seed(5)
        if impl == "ref":
            y_hat, mh, sh = dmc_ref.dual_prior(a, b, cc, conv)
            _, lik = gc_ref(a, sh, mh)
        else:
            hold = Holder()
            hold.y_spatial_prior = conv
            hold.gaussian_conditional = gc
            y_hat, liks = ctxmod._context_tail(hold, a, b, cc, None)
            lik = liks["y"]
        loss = (y_hat * torch.linspace(-1, 1, y_hat.numel(), device=cuda_dev).view_as(y_hat)).sum() \
            - torch.log2(lik).sum() / 64.0
        loss.backward()
        res.append((y_hat.detach(), lik.detach(), a.grad, b.grad, cc.grad, conv.weight.grad.clone()))
    assert torch.equal(res[1][0], res[0][0])
    assert ((res[1][1] - res[0][1]).abs() / res[0][1]).max().item() <= 1e-5
    for k, nm in ((2, "grad_y"), (3, "grad_means"), (4, "grad_scales"), (5, "grad_conv_weight")):
        _close(res[1][k], res[0][k], nm, 2e-4)


class _MiniMotionContext(torch.nn.Module):
    """Stand-in with the attribute names of the reference's MotionContextModel
    (video_model.py:128-150) and small conv nets, so the *bound method* drop-in
    `motion_context_forward` can be exercised end to end on the GPU box (the
    reference itself cannot travel there)."""

    def __init__(self, c, cz, eb_cls, gc_cls):
        super().__init__()
        nn = torch.nn
        self.hyper_encoder = nn.Conv2d(c, cz, 3, stride=4, padding=1)
        self.hyper_decoder = nn.ConvTranspose2d(cz, c, 4, stride=4)
        self.y_prior_fusion = nn.Conv2d(2 * c, 2 * c, 3, padding=1)
        self.y_spatial_prior = nn.Conv2d(3 * c, 2 * c, 3, padding=1)
        self.entropy_bottleneck = eb_cls(cz)
        self.gaussian_conditional = gc_cls(None)


@pytest.mark.parametrize("training", [False, True])
def test_motion_context_model_drop_in(cuda_dev, training, monkeypatch):
    """`MotionContextModel.forward` (video_model.py:218-233) restated with oracle
    ops vs the fused drop-in bound onto the same module: outputs, likelihoods,
    bpp and every parameter gradient."""
    import deepvideocodec_b200 as dvc
    from oracle import dmc_ref
    oem = _oracle_entropy_models()
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)      # restored after the test
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    c, cz = 16, 8
    torch.manual_seed(60)
    ref = _MiniMotionContext(c, cz, oem.EntropyBottleneck, oem.GaussianConditional).to(cuda_dev)
    mine = _MiniMotionContext(c, cz, dvc.EntropyBottleneck, dvc.GaussianConditional).to(cuda_dev)
    mine.load_state_dict(ref.state_dict())
    ref.train(training)
    mine.train(training)
    g = torch.Generator(device=cuda_dev).manual_seed(61)
    y = torch.randn(2, c, 16, 16, device=cuda_dev, generator=g) * 3
    y_ref = torch.randn(2, c, 16, 16, device=cuda_dev, generator=g)

    def reference_forward(m, y, y_ref):          # video_model.py:218-233, op for op
        z = m.hyper_encoder(y)
        _, z_lik = m.entropy_bottleneck(z)
        z_hat = dmc_ref.quantize_hyper(z, m.entropy_bottleneck._get_medians())
        params = m.hyper_decoder(z_hat)
        means, scales = m.y_prior_fusion(torch.cat((params, y_ref), dim=1)).chunk(2, 1)
        y_hat, mh, sh = dmc_ref.dual_prior(y, means, scales, m.y_spatial_prior)
        _, y_lik = m.gaussian_conditional(y, sh, mh)
        return y_hat, {"y": y_lik, "z": z_lik}

    out = []
    for m, fwd, collect in ((ref, reference_forward, dmc_ref.collect_likelihoods_list),
                            (mine, dvc.motion_context_forward, dvc.collect_likelihoods_list)):
        m.zero_grad()
        a = y.clone().requires_grad_(True)
        torch.manual_seed(7)
        y_hat, liks = fwd(m, a, y_ref)
        bpp, info = collect([{"motion": liks}], 256 * 256)
        loss = bpp.mean() + 1e-3 * (y_hat * y_hat).mean()
        loss.backward()
        out.append((y_hat.detach(), liks["y"].detach(), liks["z"].detach(), bpp.detach(), a.grad,
                    {n: p.grad for n, p in m.named_parameters()}))
    assert torch.equal(out[1][0], out[0][0])
    for k in (1, 2):
        assert ((out[1][k] - out[0][k]).abs() / out[0][k]).max().item() <= 5e-5
    _close(out[1][3], out[0][3], "bpp")
    _close(out[1][4], out[0][4], "grad_y", 5e-4)
    for name, gref in out[0][5].items():
        if gref is None:
            assert out[1][5][name] is None or out[1][5][name].abs().max().item() == 0.0, name
            continue
        assert out[1][5][name] is not None, name
        _close(out[1][5][name], gref, name, 1e-3)
