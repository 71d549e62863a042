"""CPU: the CompressAI restatement against independent fp64 closed forms and
invariants (this is what compensates for 'parity unpinned')."""
import math

import numpy as np
import pytest
import torch
from scipy import special

from test_oracle_golden import _oem


def test_gaussian_conditional_closed_form():
    oem = _oem()
    g = torch.Generator().manual_seed(1)
    y = torch.randn(1, 4, 16, 16, generator=g) * 5
    mu = torch.randn(1, 4, 16, 16, generator=g) * 3
    sg = torch.exp(torch.empty(1, 4, 16, 16).uniform_(math.log(0.05), math.log(2.0), generator=g))
    with torch.no_grad():
        out, lik = oem.GaussianConditional(None).eval()(y, sg, mu)
    assert torch.equal(out, torch.round(y - mu) + mu)
    v = np.abs((out - mu).double().numpy())
    s = np.maximum(sg.double().numpy(), np.float32(0.11).astype(np.float64))
    p = special.ndtr((0.5 - v) / s) - special.ndtr((-0.5 - v) / s)
    p = np.maximum(p, 1e-9)
    rel = np.abs(lik.double().numpy() - p) / p
    assert rel.max() < 2e-5          # fp32 conditioning for scales < 2


def test_gaussian_conditional_sums_to_one_and_is_symmetric():
    oem = _oem()
    gc = oem.GaussianConditional(None).eval()
    k = torch.arange(-60, 61, dtype=torch.float32).view(1, 1, 1, -1)
    for s in (0.05, 0.11, 0.5, 3.0, 9.0):
        for m in (0.0, 0.3, -1.7):
            with torch.no_grad():
                _, lik = gc(k + m, torch.full_like(k, s), torch.full_like(k, m))
            assert abs(lik.double().sum().item() - 1.0) < 1e-5, (s, m)
            if m == 0.0:       # (k + m) - m is not exactly k in fp32 otherwise
                assert torch.equal(lik, lik.flip(-1))


def test_bounds_are_exact():
    oem = _oem()
    gc = oem.GaussianConditional(None).eval()
    y = torch.tensor([0.0, 1000.0]).view(1, 1, 1, 2)
    with torch.no_grad():
        _, lik = gc(y, torch.tensor([1e-6, 1.0]).view(1, 1, 1, 2), torch.zeros(1, 1, 1, 2))
    assert lik[0, 0, 0, 1].item() == np.float32(1e-9)          # likelihood floor
    v = special.ndtr(0.5 / np.float64(np.float32(0.11))) - special.ndtr(-0.5 / np.float64(np.float32(0.11)))
    assert abs(lik[0, 0, 0, 0].item() - v) < 1e-6               # scale floor 0.11


def test_lower_bound_gradient_rule():
    oem = _oem()
    from oracle_compressai.ops import LowerBound
    lb = LowerBound(1.0)
    x = torch.tensor([0.5, 0.5, 2.0, 2.0], requires_grad=True)
    y = lb(x)
    y.backward(torch.tensor([1.0, -1.0, 1.0, -1.0]))
    assert torch.equal(y.detach(), torch.tensor([1.0, 1.0, 2.0, 2.0]))
    assert torch.equal(x.grad, torch.tensor([0.0, -1.0, 1.0, -1.0]))


def test_entropy_bottleneck_closed_form_and_normalisation():
    oem = _oem()
    torch.manual_seed(2)
    eb = oem.EntropyBottleneck(3).eval()
    with torch.no_grad():
        for name, p in eb.named_parameters():
            if name.startswith("_factor"):
                p.uniform_(-0.5, 0.5)
    sd = {k: v.double().numpy() for k, v in eb.state_dict().items()}

    def logits(t, c):
        l = np.array([[t]])
        for k in range(5):
            m = np.log1p(np.exp(sd[f"_matrix{k}"][c]))
            l = m @ l + sd[f"_bias{k}"][c]
            if k < 4:
                l = l + np.tanh(sd[f"_factor{k}"][c]) * np.tanh(l)
        return l[0, 0]

    ks = torch.arange(-400, 401, dtype=torch.float32)
    z = ks.view(1, 1, 1, -1).repeat(1, 3, 1, 1)
    with torch.no_grad():
        _, lik = eb(z)
    for c in range(3):
        assert abs(lik[0, c].double().sum().item() - 1.0) < 1e-4     # a proper pmf over the integers
        for k in (-3.0, 0.0, 2.0, 17.0):
            lo, up = logits(k - 0.5, c), logits(k + 0.5, c)
            sgn = -np.sign(lo + up)
            p = abs(1 / (1 + np.exp(-sgn * up)) - 1 / (1 + np.exp(-sgn * lo)))
            got = lik[0, c, 0, int(k) + 400].item()
            assert abs(got - max(p, 1e-9)) / max(p, 1e-9) < 1e-4, (c, k)


def test_state_dict_names_match_survey_appendix_b():
    oem = _oem()
    eb = set(oem.EntropyBottleneck(4).state_dict())
    assert eb == {*(f"_matrix{k}" for k in range(5)), *(f"_bias{k}" for k in range(5)),
                  *(f"_factor{k}" for k in range(4)), "quantiles", "_offset", "_quantized_cdf",
                  "_cdf_length", "target", "likelihood_lower_bound.bound"}
    gc = set(oem.GaussianConditional(None).state_dict())
    assert gc == {"_offset", "_quantized_cdf", "_cdf_length", "scale_table", "scale_bound",
                  "likelihood_lower_bound.bound", "lower_bound_scale.bound"}
