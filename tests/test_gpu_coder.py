"""GPU parity of the entropy-coder path through the C
ABI (``dvc_symbols_indexes_fwd``, ``dvc_rans_encode``, ``dvc_rans_decode``):

* bit streams are compared BYTE FOR BYTE with the plain-C oracle
  (``oracle/c/rans_ref.c``): the raw mode (``stream_symbols = 0``) against one
  stock stream, the ``DVC1`` container sub-stream by sub-stream against stock
  streams of the corresponding slices, the lane-interleaved ``DVC3`` / ``DVS3``
  containers against the sequential restatement of their schedule
  (``dvcref_ilv_encode``), which in turn decodes the GPU's bytes;
* decode(encode(x)) == round(x - means) + means, also at BASELINE.json's full
  1080p latent sizes, where the coded size is additionally checked against the
  estimated rate of the likelihood kernels (a size-independent property);
* edge cases: one symbol, ragged last sub-stream, escapes (symbols outside the
  table, both signs), bad indexes, malformed containers, strided inputs.
"""
import struct

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scale_table():
    return np.exp(np.linspace(np.log(0.11), np.log(256), 64)).tolist()


@pytest.fixture(scope="module")
def gc_pair(cuda_dev):
    """(oracle module on CPU, product module on the GPU) with identical tables."""
    import deepvideocodec_b200 as dvc
    from oracle.compressai import entropy_models as oem
    o = oem.GaussianConditional(None)
    o.update_scale_table(_scale_table())
    p = dvc.GaussianConditional(None)
    p.update_scale_table(_scale_table())
    for name in ("_quantized_cdf", "_cdf_length", "_offset", "scale_table"):
        assert torch.equal(getattr(o, name), getattr(p, name)), name
    return o, p.to(cuda_dev).eval()


@pytest.fixture(scope="module")
def eb_pair(cuda_dev):
    import deepvideocodec_b200 as dvc
    from oracle.compressai import entropy_models as oem
    torch.manual_seed(11)
    o = oem.EntropyBottleneck(64)
    torch.manual_seed(11)
    p = dvc.EntropyBottleneck(64)
    with torch.no_grad():
        for m in (o, p):
            m.quantiles[:, 0, 0] = -torch.linspace(3.2, 40, 64)
            m.quantiles[:, 0, 2] = torch.linspace(1.5, 25, 64)
            m.quantiles[:, 0, 1] = torch.linspace(-2, 2, 64)
    o.update()
    p.update()
    for name in ("_quantized_cdf", "_cdf_length", "_offset"):
        assert torch.equal(getattr(o, name), getattr(p, name)), name
    return o, p.to(cuda_dev).eval()


def _latents(shape, seed, dev, escapes=True):
    g = torch.Generator().manual_seed(seed)
    scales = torch.exp(torch.empty(shape).uniform_(np.log(0.05), np.log(64), generator=g))
    means = torch.randn(shape, generator=g) * 3
    y = means + scales * torch.randn(shape, generator=g)
    if escapes:
        flat = y.view(-1)
        k = min(flat.numel(), 8)
        flat[:k] += torch.tensor([4000.0, -4000.0, 70000.0, -70000.0, 1e6, -1e6, 300.0, -300.0])[:k]
    return y.to(dev), means.to(dev), scales.to(dev)


def _oracle_tables(o):
    return (o._quantized_cdf.numpy(), o._cdf_length.numpy(), o._offset.numpy())


def _split_container(s, L):
    magic, n_sym, S, ns = struct.unpack_from("<4I", s, 0)
    assert magic == 0x31435644 and n_sym == L and ns == (L + S - 1) // S
    words = struct.unpack_from(f"<{ns}I", s, 16)
    pos = 16 + 4 * ns
    subs = []
    for wds in words:
        subs.append(s[pos:pos + 4 * wds])
        pos += 4 * wds
    assert pos == len(s)
    return S, subs


# ---------------------------------------------------------------------------
# f1
# ---------------------------------------------------------------------------
def test_build_indexes_and_symbols_match_oracle(cuda_dev, gc_pair):
    o, p = gc_pair
    y, means, scales = _latents((2, 6, 10, 14), 1, cuda_dev)
    scales.view(-1)[:64] = p.scale_table                    # exactly on a table entry
    scales.view(-1)[64:127] = torch.nextafter(p.scale_table[:-1], p.scale_table[1:])
    scales.view(-1)[200] = float("nan")
    got = p.build_indexes(scales)
    want = o.build_indexes(scales.cpu())
    assert got.dtype == torch.int32 and torch.equal(got.cpu(), want)
    # channels_last input: same values
    got_cl = p.build_indexes(scales.contiguous(memory_format=torch.channels_last))
    assert torch.equal(got_cl, got)
    # non-4D input keeps its shape
    assert torch.equal(p.build_indexes(scales.reshape(2, -1)).reshape(scales.shape), got)
    sym = p.quantize(y, "symbols", means)
    assert torch.equal(sym.cpu(), o.quantize(y.cpu(), "symbols", means.cpu()))
    from deepvideocodec_b200 import coder
    assert torch.equal(coder.quantize_symbols(y, means), sym)
    assert torch.equal(coder.quantize_symbols(y), torch.round(y).int())


# ---------------------------------------------------------------------------
# f2: byte parity
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (1, 2, 2, 2), (2, 6, 10, 14), (1, 32, 68, 120)])
def test_raw_mode_is_byte_identical_to_stock_stream(cuda_dev, gc_pair, shape):
    from deepvideocodec_b200 import coder
    from oracle import rans
    o, p = gc_pair
    y, means, scales = _latents(shape, 2, cuda_dev)
    idx = p.build_indexes(scales)
    strings = coder.rans_encode(p._tables(), x=y, means=means, indexes=idx, stream_symbols=0)
    sym = torch.round(y - means).int().cpu().numpy()
    cdf, size, off = _oracle_tables(o)
    for n in range(shape[0]):
        want = rans.encode_with_indexes(sym[n], idx[n].cpu().numpy(), cdf, size, off)
        assert strings[n] == want
    # ... and the GPU decodes the oracle's stock stream
    stock = [rans.encode_with_indexes(sym[n], idx[n].cpu().numpy(), cdf, size, off)
             for n in range(shape[0])]
    back = coder.rans_decode(stock, p._tables(), shape, indexes=idx, means=means)
    assert torch.equal(back, torch.round(y - means) + means)
    # the oracle decodes the GPU's stream
    for n in range(shape[0]):
        dec = rans.decode_with_indexes(strings[n], idx[n].cpu().numpy(), cdf, size, off)
        assert np.array_equal(dec, sym[n])


@pytest.mark.parametrize("shape,S", [((1, 2, 2, 2), 3), ((2, 6, 10, 14), 256), ((2, 6, 10, 14), 100),
                                      ((1, 3, 5, 7), 1000), ((1, 32, 68, 120), 1024)])
def test_container_substreams_are_stock_streams(cuda_dev, gc_pair, shape, S):
    from deepvideocodec_b200 import coder
    from oracle import rans
    o, p = gc_pair
    y, means, scales = _latents(shape, 3, cuda_dev)
    L = shape[1] * shape[2] * shape[3]
    # indexes derived from the scales inside the coder (fused build_indexes)
    strings = coder.rans_encode(p._tables(), x=y, means=means, scales=scales,
                                scale_table=p.scale_table, scale_bound=0.11, stream_symbols=S,
                                lanes=1)
    idx = o.build_indexes(scales.cpu()).numpy().reshape(shape[0], -1)
    sym = torch.round(y - means).int().cpu().numpy().reshape(shape[0], -1)
    cdf, size, off = _oracle_tables(o)
    for n in range(shape[0]):
        got_S, subs = _split_container(strings[n], L)
        assert got_S == S and coder.stream_symbols_of(strings[n], L) == S
        for j, sub in enumerate(subs):
            sl = slice(j * S, min(L, (j + 1) * S))
            assert sub == rans.encode_with_indexes(sym[n, sl], idx[n, sl], cdf, size, off), (n, j)
    back = coder.rans_decode(strings, p._tables(), shape, scales=scales,
                             scale_table=p.scale_table, scale_bound=0.11, means=means)
    assert torch.equal(back, torch.round(y - means) + means)
    syms = coder.rans_decode(strings, p._tables(), shape, indexes=torch.from_numpy(idx).to(
        cuda_dev).reshape(shape), want_symbols=True)
    assert torch.equal(syms.cpu().reshape(shape[0], -1), torch.from_numpy(sym))


def test_module_compress_decompress_match_oracle_modules(cuda_dev, gc_pair, eb_pair, monkeypatch):
    from deepvideocodec_b200 import coder
    o, p = gc_pair
    y, means, scales = _latents((2, 8, 12, 16), 4, cuda_dev)
    idx = p.build_indexes(scales)
    idx_cpu = o.build_indexes(scales.cpu())
    want = o.compress(y.cpu(), idx_cpu, means.cpu())
    monkeypatch.setattr(coder, "DEFAULT_STREAM_SYMBOLS", 0)
    assert p.compress(y, idx, means) == want
    # the reference's own call shape: compress(y_quant, indexes) without means
    q = torch.round(y - means)
    assert p.compress(q, idx) == o.compress(q.cpu(), idx_cpu)
    assert torch.equal(p.decompress(want, idx, means=means).cpu(),
                       o.decompress(want, idx_cpu, means=means.cpu()))
    assert torch.equal(p.decompress(o.compress(q.cpu(), idx_cpu), idx).cpu(), q.cpu())
    monkeypatch.setattr(coder, "DEFAULT_STREAM_SYMBOLS", 512)
    monkeypatch.setattr(coder, "DEFAULT_LANES", 1)
    s = p.compress(y, idx, means)
    assert s != want and coder.container_of(s[0], 8 * 12 * 16) == (512, 1, False)
    assert torch.equal(p.decompress(s, idx, means=means), torch.round(y - means) + means)
    # the default layout: 32 lane-interleaved coders per sub-stream, implied zeros on
    monkeypatch.setattr(coder, "DEFAULT_LANES", 32)
    s = p.compress(y, idx, means)
    assert coder.container_of(s[0], 8 * 12 * 16)[:2] == (coder.LANES_STREAM_SYMBOLS, 32)
    assert torch.equal(p.decompress(s, idx, means=means), torch.round(y - means) + means)

    eo, ep = eb_pair
    z = (torch.randn(3, 64, 17, 30, generator=torch.Generator().manual_seed(5)) * 8).to(cuda_dev)
    z[0, 0, 0, :2] += torch.tensor([500.0, -500.0], device=cuda_dev)
    monkeypatch.setattr(coder, "DEFAULT_STREAM_SYMBOLS", 0)
    zs = ep.compress(z)
    assert zs == eo.compress(z.cpu())
    zb = ep.decompress(zs, z.shape[-2:])
    assert torch.equal(zb.cpu(), eo.decompress(zs, z.shape[-2:]))
    med = ep._get_medians().detach().reshape(1, -1, 1, 1)
    assert torch.equal(zb, torch.round(z - med) + med)
    # z_hat of the coder == z_hat of the likelihood kernel (video_model.py:222-224 vs :239)
    from deepvideocodec_b200.entropy_models import eb_forward
    _, z_hat, _ = eb_forward(ep, z, want_outputs=False, want_zhat=True)
    assert torch.equal(zb, z_hat)
    monkeypatch.setattr(coder, "DEFAULT_STREAM_SYMBOLS", 1024)
    assert torch.equal(ep.decompress(ep.compress(z), z.shape[-2:]), zb)


def test_strided_inputs_and_broadcast_means(cuda_dev, gc_pair):
    from deepvideocodec_b200 import coder
    o, p = gc_pair
    y, means, scales = _latents((2, 6, 10, 14), 6, cuda_dev, escapes=False)
    idx = p.build_indexes(scales)
    ref = coder.rans_encode(p._tables(), x=y, means=means, indexes=idx, stream_symbols=128)
    cl = torch.channels_last
    got = coder.rans_encode(p._tables(), x=y.contiguous(memory_format=cl),
                            means=means.contiguous(memory_format=cl),
                            scales=scales.contiguous(memory_format=cl), scale_table=p.scale_table,
                            stream_symbols=128)
    assert got == ref
    # per-channel broadcast mean (the entropy bottleneck's medians)
    med = torch.linspace(-1, 1, 6, device=cuda_dev).reshape(1, 6, 1, 1)
    a = coder.rans_encode(p._tables(), x=y, means=med.expand_as(y), indexes=idx, stream_symbols=128)
    b = coder.rans_encode(p._tables(), x=y, means=med.expand_as(y).contiguous(), indexes=idx,
                          stream_symbols=128)
    assert a == b
    out = coder.rans_decode(a, p._tables(), y.shape, indexes=idx, means=med)
    assert torch.equal(out, torch.round(y - med) + med)


def test_errors(cuda_dev, gc_pair):
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200 import coder
    o, p = gc_pair
    y, means, scales = _latents((1, 4, 6, 8), 7, cuda_dev)
    idx = p.build_indexes(scales)
    bad = idx.clone()
    bad[0, 1, 2, 3] = 64
    with pytest.raises(ValueError, match="outside the CDF tables"):
        coder.rans_encode(p._tables(), x=y, means=means, indexes=bad)
    # the status word is reset: the next call succeeds
    s = coder.rans_encode(p._tables(), x=y, means=means, indexes=idx, stream_symbols=64)
    with pytest.raises(ValueError, match="outside the CDF tables"):
        coder.rans_decode(s, p._tables(), y.shape, indexes=bad)
    assert torch.equal(coder.rans_decode(s, p._tables(), y.shape, indexes=idx, means=means),
                       torch.round(y - means) + means)
    # malformed containers: a sub-stream length pointing past the end; wrong symbol count
    w = bytearray(s[0])
    struct.pack_into("<I", w, 16, 1 << 20)
    with pytest.raises(ValueError, match="malformed"):
        coder.rans_decode([bytes(w)], p._tables(), y.shape, indexes=idx)
    with pytest.raises(ValueError):
        coder.rans_decode([s[0][:6]], p._tables(), y.shape, indexes=idx)
    with pytest.raises(ValueError):
        coder.rans_decode(s + s, p._tables(), y.shape, indexes=idx)
    with pytest.raises(ValueError, match="same size"):
        p.compress(y, idx[:, :2])
    with pytest.raises(dvc.DvcError):
        p.compress(y.cpu(), idx.cpu())
    # a truncated (but well-formed-looking) raw stream decodes zeros past its end, never faults
    raw = coder.rans_encode(p._tables(), x=y, means=means, indexes=idx, stream_symbols=0)
    coder.rans_decode([raw[0][:8]], p._tables(), y.shape, indexes=idx)
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------
# full 1080p latent sizes: round trip + coded size against the estimated rate
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("c", [64, 96])
def test_full_size_round_trip_and_rate(cuda_dev, gc_pair, c):
    from deepvideocodec_b200 import coder
    o, p = gc_pair
    shape = (1, c, 68, 120)
    y, means, scales = _latents(shape, 8 + c, cuda_dev, escapes=False)
    idx = p.build_indexes(scales)
    _, lik = p(y, scales, means)
    est_bits = float(-torch.log2(lik.double()).sum())
    for S in (256, 1024, 4096):
        s = coder.rans_encode(p._tables(), x=y, means=means, indexes=idx, stream_symbols=S,
                              lanes=1)
        back = coder.rans_decode(s, p._tables(), shape, indexes=idx, means=means)
        assert torch.equal(back, torch.round(y - means) + means)
        n_streams = -(-c * 68 * 120 // S)
        real_bits = 8 * len(s[0])
        overhead = 32 * (4 + n_streams) + 48 * n_streams      # header + ~one state flush each
        # 16-bit tables on a 64-entry scale grid: within a few percent of the estimate
        assert abs(real_bits - overhead - est_bits) < 0.04 * est_bits + 20 * n_streams, \
            (S, real_bits, est_bits)
    # lane-interleaved: <= 65 flush words + the length word per sub-stream of 32 chains
    for S, skip in ((8192, False), (65536, True)):
        s = coder.rans_encode(p._tables(), x=y, means=means, indexes=idx, stream_symbols=S,
                              lanes=32, skip=skip)
        back = coder.rans_decode(s, p._tables(), shape, indexes=idx, means=means)
        assert torch.equal(back, torch.round(y - means) + means)
        n_streams = -(-c * 68 * 120 // S)
        real_bits = 8 * len(s[0])
        assert coder.container_of(s[0], c * 68 * 120) == (S, 32, skip)
        assert abs(real_bits - est_bits) < 0.04 * est_bits + 32 * (4 + 66 * n_streams), \
            (S, real_bits, est_bits)


def test_reference_compress_call_sequence(cuda_dev, gc_pair):
    """The sequence video_model.py:245-251 / :268-289 runs around the spatial
    prior: compress planes of forward_dual_prior(mode='compress') -> two
    strings -> decompress -> the same y_hat."""
    from deepvideocodec_b200 import context
    from oracle import dmc_ref
    o, p = gc_pair

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(21)
            self.y_spatial_prior = torch.nn.Conv2d(3 * 8, 2 * 8, 3, padding=1)
            self.gaussian_conditional = p

    m = Holder().to(cuda_dev).eval()
    y, means, scales = _latents((1, 8, 12, 16), 22, cuda_dev, escapes=False)
    with torch.no_grad():
        y_hat, q0, q1, s0, s1 = context.forward_dual_prior(m, y, means, scales, mode="compress")
        i0, i1 = p.build_indexes(s0), p.build_indexes(s1)
        st0, st1 = p.compress(q0, i0), p.compress(q1, i1)
        # decoder side, video_model.py:259-289
        mask_0, mask_1 = dmc_ref.checkerboard_masks(12, 16, cuda_dev)
        m0, m1 = means.chunk(2, 1)
        sc0, sc1 = scales.chunk(2, 1)
        r0 = p.decompress(st0, p.build_indexes(sc0 * mask_0 + sc1 * mask_1))
        assert torch.equal(r0, q0)
        y00, y11 = (r0 + m0) * mask_0, (r0 + m1) * mask_1
        pm0, ps0, pm1, ps1 = m.y_spatial_prior(torch.cat((y00, y11, means, scales), 1)).chunk(4, 1)
        r1 = p.decompress(st1, p.build_indexes(ps0 * mask_1 + ps1 * mask_0))
        assert torch.equal(r1, q1)
        y01, y10 = (r1 + pm0) * mask_1, (r1 + pm1) * mask_0
        assert torch.equal(torch.cat((y00 + y01, y11 + y10), 1), y_hat)


# ---------------------------------------------------------------------------
# context-model level: fused compress / decompress drop-ins vs the oracle's
# restatement of the reference's methods, executed by CUDA eager on this device
# ---------------------------------------------------------------------------
class _Ctx(torch.nn.Module):
    """Stand-in with the attribute names of the reference's context models
    (video_model.py:128-150 / :294-322) and small conv stacks."""

    def __init__(self, c, cz, frame, gc, eb):
        super().__init__()
        conv = torch.nn.Conv2d
        self.hyper_encoder = torch.nn.Sequential(conv(c, cz, 3, 2, 1), torch.nn.LeakyReLU(0.1),
                                                 conv(cz, cz, 3, 2, 1))
        self.hyper_decoder = torch.nn.Sequential(
            torch.nn.ConvTranspose2d(cz, c, 4, 2, 1), torch.nn.LeakyReLU(0.1),
            torch.nn.ConvTranspose2d(c, c, 4, 2, 1))
        extra = c if frame else 0
        self.y_prior_fusion = conv(2 * c + extra, 2 * c, 3, 1, 1)
        self.y_spatial_prior = conv(3 * c, 2 * c, 3, 1, 1)
        if frame:
            self.temporal_prior_encoder = conv(8, c, 3, 4, 1)
        self.gaussian_conditional = gc
        self.entropy_bottleneck = eb


def test_payload_driven_substreams(cuda_dev, gc_pair):
    """at the 0.03 - 0.15 bits/symbol of a low-rate P-frame a fixed 4 096-symbol
    partition costs +15 ... 80 % bytes.  The default partition is sized from the estimated
    payload: container overhead <= 1 % of the bytes (+ the 8-stream floor, 96 B), at every rate;
    and the bytes still decode to the symbols."""
    import math
    from deepvideocodec_b200 import coder
    _, p = gc_pair
    tables = p._tables()
    g = torch.Generator().manual_seed(77)
    shape = (1, 48, 68, 120)                                  # one checkerboard pass, frame model
    L = shape[1] * shape[2] * shape[3]
    for lo, hi, label in ((0.16, 0.2, "low"), (0.5, 1.0, "mid"), (4.0, 32.0, "high")):
        scales = (torch.empty(shape).uniform_(lo, hi, generator=g)).to(cuda_dev)
        x = torch.round(torch.randn(shape, generator=g).to(cuda_dev) * scales)
        kw = dict(x=x, scales=scales, scale_table=p.scale_table, scale_bound=0.11)
        raw = coder.rans_encode(tables, stream_symbols=0, **kw)[0]         # one stock stream
        fixed = coder.rans_encode(tables, stream_symbols=4096, lanes=1, **kw)[0]
        est = len(raw)                                        # what the likelihood sum predicts
        auto = coder.rans_encode(tables, est_bytes=est, lanes=1, **kw)[0]
        S = coder.stream_symbols_of(auto, L)
        n_streams = (L + S - 1) // S
        assert S == coder.auto_stream_symbols(L, est) and S % 256 == 0
        overhead = len(auto) - len(raw)
        assert overhead <= 0.0125 * len(raw) + coder.MIN_STREAMS * 16 + 32, (label, overhead, len(raw))
        if label == "low":
            assert len(raw) * 8 / L < 0.2                     # the regime the advice is about
            assert len(fixed) - len(raw) > 5 * overhead       # what the fixed partition cost
            assert n_streams == coder.MIN_STREAMS
        if label == "high":
            assert n_streams > (L + 4095) // 4096             # more parallel than the old default
        out = coder.rans_decode([auto], tables, shape, scales=scales, scale_table=p.scale_table,
                                scale_bound=0.11, device=cuda_dev)
        assert torch.equal(out, x)
    assert coder.auto_stream_symbols(L, None) == coder.DEFAULT_STREAM_SYMBOLS
    assert math.isfinite(est)


@pytest.mark.parametrize("frame", [False, True])
@pytest.mark.parametrize("S", [0, 256, None])
def test_context_model_compress_decompress(cuda_dev, gc_pair, frame, S, monkeypatch):
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200 import coder
    from oracle import dmc_ref
    from oracle.compressai import entropy_models as oem
    monkeypatch.setattr(coder, "PINNED_STREAM_SYMBOLS", S)
    o_gc, p_gc = gc_pair
    torch.manual_seed(31)
    o_eb = oem.EntropyBottleneck(6)
    torch.manual_seed(31)
    p_eb = dvc.EntropyBottleneck(6)
    o_eb.update()
    p_eb.update()
    torch.manual_seed(32)
    m = _Ctx(12, 6, frame, p_gc, p_eb).to(cuda_dev).eval()
    o_eb = o_eb.to(cuda_dev)
    o_gc_dev = oem.GaussianConditional(None)
    o_gc_dev.update_scale_table(_scale_table())
    o_gc_dev = o_gc_dev.to(cuda_dev)
    g = torch.Generator().manual_seed(33)
    y = (torch.randn(2, 12, 16, 24, generator=g) * 5).to(cuda_dev)
    y_ref = torch.randn(2, 12, 16, 24, generator=g).to(cuda_dev)
    context = torch.randn(2, 8, 64, 96, generator=g).to(cuda_dev)

    def prior_fusion(z_hat):
        params = m.hyper_decoder(z_hat)
        parts = (m.temporal_prior_encoder(context), params, y_ref) if frame else (params, y_ref)
        return m.y_prior_fusion(torch.cat(parts, 1)).chunk(2, 1)

    with torch.no_grad():
        if frame:
            y_hat, out = dvc.frame_context_compress(m, y, y_ref, context)
            dec = dvc.frame_context_decompress(m, out["strings"], out["shape"], y_ref, context)
        else:
            y_hat, out = dvc.motion_context_compress(m, y, y_ref)
            dec = dvc.motion_context_decompress(m, out["strings"], out["shape"], y_ref)
        z = m.hyper_encoder(y)
        y_hat_o, out_o = dmc_ref.context_model_compress(
            y, z, prior_fusion, m.y_spatial_prior, o_eb, o_gc_dev)
        dec_o = dmc_ref.context_model_decompress(
            out_o["strings"], out_o["shape"], prior_fusion, m.y_spatial_prior, o_eb, o_gc_dev)
    assert torch.equal(y_hat, y_hat_o)
    assert torch.equal(dec, y_hat) and torch.equal(dec_o, y_hat_o)
    assert tuple(out["shape"]) == tuple(out_o["shape"])
    assert [len(s) for s in out["strings"]] == [2, 2, 2]
    if S == 0:
        assert out["strings"] == out_o["strings"]            # byte-identical to the stock streams
    else:
        # the fused decoder also decodes the oracle's stock streams
        with torch.no_grad():
            fn = dvc.frame_context_decompress if frame else dvc.motion_context_decompress
            extra = (context,) if frame else ()
            assert torch.equal(fn(m, out_o["strings"], out_o["shape"], y_ref, *extra), y_hat)


def test_golden_reference_bitstreams(cuda_dev, golden_dir):
    """Vectors written by the reference's own ``MotionContextModel.compress``
    (tests/golden/make_golden_coder.py): same tensors in -> same bytes out."""
    import os
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200 import coder
    z = np.load(os.path.join(golden_dir, "coder.npz"))
    t = lambda k: torch.from_numpy(z[k]).to(cuda_dev)      # noqa: E731
    gc = dvc.GaussianConditional(None)
    gc.update_scale_table(z["scale_table"].tolist())
    assert np.array_equal(gc._quantized_cdf.numpy(), z["gc.cdf"])
    assert np.array_equal(gc._cdf_length.numpy(), z["gc.len"])
    assert np.array_equal(gc._offset.numpy(), z["gc.off"])
    gc = gc.to(cuda_dev)
    eb = dvc.EntropyBottleneck(8)
    eb.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ebp.")},
                       strict=False)
    eb.update(force=True)
    assert np.array_equal(eb._quantized_cdf.numpy(), z["eb.cdf"])
    assert np.array_equal(eb._offset.numpy(), z["eb.off"])
    eb = eb.to(cuda_dev)
    strs = lambda name: [z[f"str.{name}.{n}"].tobytes() for n in range(2)]   # noqa: E731
    assert torch.equal(gc.build_indexes(t("s0")).cpu(), torch.from_numpy(z["i0"]))
    assert torch.equal(gc.build_indexes(t("s1")).cpu(), torch.from_numpy(z["i1"]))
    raw = dict(stream_symbols=0)
    assert coder.rans_encode(gc._tables(), x=t("q0"), scales=t("s0"), scale_table=gc.scale_table,
                             **raw) == strs("y0")
    assert coder.rans_encode(gc._tables(), x=t("q1"), indexes=t("i1"), **raw) == strs("y1")
    med = eb._get_medians().detach().reshape(1, -1, 1, 1)
    assert coder.rans_encode(eb._tables(), x=t("z"), means=med.expand_as(t("z")), **raw) == strs("z")
    assert torch.equal(eb.decompress(strs("z"), tuple(z["shape"])).cpu(), torch.from_numpy(z["z_hat"]))
    assert torch.equal(gc.decompress(strs("y0"), t("i0")).cpu(), torch.from_numpy(z["q0"]))
    # decoder glue against the reference's tensors
    means, scales, prior = t("means"), t("scales"), t("prior")
    n, c, h, w = means.shape
    q0 = coder.rans_decode(strs("y0"), gc._tables(), (n, c // 2, h, w), scales=scales[:, :c // 2],
                           scale_table=gc.scale_table, want_symbols=True,
                           cb=(0, (c // 2) * scales.stride(1)))
    assert torch.equal(q0.float().cpu(), torch.from_numpy(z["q0"]))
    q1 = coder.rans_decode(strs("y1"), gc._tables(), (n, c // 2, h, w), scales=prior[:, c // 2:c],
                           scale_table=gc.scale_table, want_symbols=True,
                           cb=(1, c * prior.stride(1)))
    assert torch.equal(q1.float().cpu(), torch.from_numpy(z["q1"]))
    y_hat = coder.decode_stage_b(q0, q1, means, prior)
    assert torch.equal(y_hat.cpu(), torch.from_numpy(z["y_hat"]))


# ---------------------------------------------------------------------------
# the lane-interleaved layouts ('DVC3', 'DVS3'): 32 stock rans64 coders per sub-stream
# ---------------------------------------------------------------------------
def _floor_heavy_latents(shape, seed, dev, floor_frac, exception_frac):
    """Scales mostly at the 0.11 floor (marked table rows), symbols drawn from the model,
    plus a few symbols a marked row does not expect."""
    g = torch.Generator().manual_seed(seed)
    scales = torch.exp(torch.empty(shape).uniform_(np.log(0.05), np.log(40), generator=g))
    scales[torch.rand(shape, generator=g) < floor_frac] = 0.05
    x = torch.round(torch.randn(shape, generator=g) * scales.clamp_min(0.11))
    wrong = (torch.rand(shape, generator=g) < exception_frac) & (scales <= 0.11)
    x[wrong] = torch.randint(-3, 4, shape, generator=g).float()[wrong]
    return x.to(dev), scales.to(dev)


@pytest.mark.parametrize("shape,S,floor,skip", [
    ((1, 1, 1, 1), 1024, 0.0, True), ((2, 6, 10, 14), 256, 0.5, True),
    ((2, 6, 10, 14), 100, 0.5, False), ((1, 3, 5, 7), 4096, 0.9, True),
    ((1, 4, 33, 31), 1024, 0.97, True), ((2, 5, 41, 50), 3000, 0.97, True),
    ((1, 32, 68, 120), 32768, 0.9, True), ((1, 32, 68, 120), 5000, 0.0, False)])
def test_lane_interleaved_container_matches_oracle(cuda_dev, gc_pair, shape, S, floor, skip,
                                                   monkeypatch):
    """Byte for byte against the sequential CPU restatement of the layout
    (oracle/c/rans_ref.c::dvcref_ilv_encode over the stock per-symbol arithmetic); the oracle
    decodes the GPU's bytes, the GPU decodes its own, with the table look-up fused (scales)
    and with explicit indexes."""
    from deepvideocodec_b200 import coder
    from oracle import rans
    o, p = gc_pair
    tables = p._tables()
    cdf, size, off = _oracle_tables(o)
    marks = rans.skip_rows_of(cdf, size, off)
    assert torch.equal(tables.skip_rows().cpu(), torch.from_numpy(marks))
    assert marks[0] == 1 and marks[-1] == 0 and 1 <= marks.sum() <= 4
    x, scales = _floor_heavy_latents(shape, S + shape[1], cuda_dev, floor, 0.01)
    L = shape[1] * shape[2] * shape[3]
    flat = x.view(shape[0], -1)
    for pos, val in ((0, 5.0), (31, -2.0), (32, 70000.0), (1023, -4000.0), (1024, 1.0), (L - 1, -1e6)):
        if pos < L:
            flat[0, pos] = val                                # escapes at group / chunk / stream edges
    kw = dict(x=x, scales=scales, scale_table=p.scale_table, scale_bound=0.11)
    got = coder.rans_encode(tables, stream_symbols=S, lanes=32, skip=skip, **kw)
    S = -(-S // 1024) * 1024                                 # sub-streams are whole 1 024-position chunks
    idx = o.build_indexes(scales.cpu()).numpy().reshape(shape[0], -1)
    sym = x.int().cpu().numpy().reshape(shape[0], -1)
    for n in range(shape[0]):
        want = rans.encode_container(sym[n], idx[n], cdf, size, off, S, 32, marks if skip else None)
        assert got[n] == want, (n, len(got[n]), len(want))
        assert coder.container_of(got[n], L) == (S, 32, skip)
        assert np.array_equal(
            rans.decode_container(got[n], idx[n], cdf, size, off, marks if skip else None), sym[n])
    out = coder.rans_decode(got, tables, shape, scales=scales, scale_table=p.scale_table,
                            scale_bound=0.11, device=cuda_dev)
    assert torch.equal(out, x)
    syms = coder.rans_decode(got, tables, shape, indexes=torch.from_numpy(idx).to(cuda_dev).reshape(
        shape), want_symbols=True)
    assert torch.equal(syms, x.int())
    # the packed tables the decoder stages in shared memory: their definition, and the same
    # symbols without them (search in global memory)
    blob, entries = tables.cdf_pack()
    raw = blob.cpu().numpy().tobytes()
    n_rows = cdf.shape[0]
    assert entries == int(size.sum()) + 4 * n_rows
    lut = np.frombuffer(raw, dtype=np.uint16, count=n_rows * coder.LUT_STRIDE).reshape(n_rows, -1)
    start_off = n_rows * coder.LUT_STRIDE * 2
    starts = np.frombuffer(raw, dtype=np.uint32, count=n_rows, offset=start_off)
    packed = np.frombuffer(raw, dtype=np.uint16, count=entries, offset=start_off + 4 * n_rows)
    assert np.array_equal(starts, np.cumsum(size + 4) - (size + 4))
    rng = np.random.default_rng(S)
    for r in (0, 1, 17, 40, 63):
        row = cdf[r, :size[r]]
        assert np.array_equal(packed[starts[r]:starts[r] + size[r]], ((row - 1) & 0xFFFF).astype(np.uint16))
        assert (packed[starts[r] + size[r]:starts[r] + size[r] + 4] == 0xFFFF).all()
        # the look-up never starts past the symbol: row[lo] <= cum, and lo is tight for the
        # first cum of the key (the symbol of that cum starts at lo)
        cmin, _ = coder._lut_ranges()
        for cum in list(range(0, 40)) + list(range(65496, 65536)) + rng.integers(0, 65536, 300).tolist():
            key = coder.lut_key(int(cum))
            lo = int(lut[r, key])
            assert lo <= size[r] - 2 and row[lo] <= cum, (r, cum, lo)
            assert row[lo] <= cmin[key] < row[lo + 1], (r, cum, lo)
            assert cum < row[int(lut[r, key + 1]) + 1], (r, cum)     # entry k + 1 bounds the bracket
        assert (np.diff(lut[r, :coder.LUT_KEYS + 1].astype(np.int64)) >= 0).all()
        assert lut[r, coder.LUT_KEYS] == size[r] - 2
    monkeypatch.setattr(coder.Tables, "cdf_pack", lambda self: (None, 0))
    assert torch.equal(coder.rans_decode(got, tables, shape, scales=scales,
                                         scale_table=p.scale_table, scale_bound=0.11,
                                         device=cuda_dev), x)


def test_lane_interleaved_entropy_bottleneck_and_errors(cuda_dev, gc_pair, eb_pair):
    """Channel-indexed tables (EntropyBottleneck) with medians; malformed sub-streams are flagged
    (a lane-interleaved sub-stream must end on the encoder's initial states)."""
    from deepvideocodec_b200 import coder
    from oracle import rans
    eo, ep = eb_pair
    z = (torch.randn(2, 64, 17, 30, generator=torch.Generator().manual_seed(15)) * 6).to(cuda_dev)
    z[1, 3, 0, :2] += torch.tensor([900.0, -900.0], device=cuda_dev)
    med = ep._get_medians().detach().reshape(1, -1, 1, 1)
    s = coder.rans_encode(ep._tables(), x=z, means=med.expand_as(z), stream_symbols=8192, lanes=32)
    cdf, size, off = _oracle_tables(eo)
    marks = rans.skip_rows_of(cdf, size, off)
    sym = torch.round(z - med).int().cpu().numpy().reshape(2, -1)
    idx = np.repeat(np.arange(64, dtype=np.int32), 17 * 30)
    for n in range(2):
        assert s[n] == rans.encode_container(sym[n], idx, cdf, size, off, 8192, 32, marks)
    back = coder.rans_decode(s, ep._tables(), z.shape, means=med, device=cuda_dev)
    assert torch.equal(back, torch.round(z - med) + med)
    # corrupt payload word / truncated container / wrong layout for the magic
    w = bytearray(s[0])
    w[-5] ^= 0x40
    with pytest.raises(ValueError, match="malformed"):
        coder.rans_decode([bytes(w), s[1]], ep._tables(), z.shape, means=med, device=cuda_dev)
    with pytest.raises(ValueError, match="malformed"):
        coder.rans_decode([s[0][:-8], s[1]], ep._tables(), z.shape, means=med, device=cuda_dev)
    # the status word is per launch: the next call succeeds
    assert torch.equal(coder.rans_decode(s, ep._tables(), z.shape, means=med, device=cuda_dev), back)
    o, p = gc_pair
    y, means, scales = _latents((1, 4, 6, 8), 7, cuda_dev)
    with pytest.raises(Exception, match="lanes"):
        coder.rans_encode(p._tables(), x=y, means=means, scales=scales, scale_table=p.scale_table,
                          lanes=7)


def test_lane_interleaved_payload_policy(cuda_dev, gc_pair):
    """32 chains cost ~200 bytes: the sub-stream count follows the payload (1 % overhead
    target), one sub-stream at the low end -- where implied zeros keep the chain short and cost
    no more bytes than the stock stream codes them with."""
    from deepvideocodec_b200 import coder
    _, p = gc_pair
    tables = p._tables()
    g = torch.Generator().manual_seed(78)
    shape = (1, 48, 68, 120)
    L = shape[1] * shape[2] * shape[3]
    for lo, hi, label in ((0.05, 0.11, "floor"), (0.16, 0.2, "low"), (4.0, 32.0, "high")):
        scales = (torch.empty(shape).uniform_(lo, hi, generator=g)).to(cuda_dev)
        x = torch.round(torch.randn(shape, generator=g).to(cuda_dev) * scales.clamp_min(0.11))
        kw = dict(x=x, scales=scales, scale_table=p.scale_table, scale_bound=0.11)
        raw = coder.rans_encode(tables, stream_symbols=0, **kw)[0]
        est = len(raw)
        auto = coder.rans_encode(tables, est_bytes=est, **kw)[0]
        S, lanes, skip = coder.container_of(auto, L)
        assert (lanes, skip) == (32, True) and S == coder.auto_stream_symbols(L, est, 32)
        assert S % 1024 == 0
        n_streams = (L + S - 1) // S
        overhead = len(auto) - len(raw)
        assert overhead <= 0.0125 * len(raw) + 280, (label, overhead, len(raw))
        if label != "high":
            assert n_streams == 1
        else:
            assert n_streams >= 8
        out = coder.rans_decode([auto], tables, shape, scales=scales, scale_table=p.scale_table,
                                scale_bound=0.11, device=cuda_dev)
        assert torch.equal(out, x)


def test_implied_zeros_are_adaptive(cuda_dev, gc_pair):
    """Default (skip=None): the encoder counts, on the device, the groups whose flag would be
    set and only uses the marks when those flags fit the budget -- data the tables describe
    keeps 'DVS3', data they do not (marked rows full of non-zero symbols) falls back to a plain
    'DVC3' that is no larger than coding without marks; forcing the marks there costs bytes."""
    from deepvideocodec_b200 import coder
    _, p = gc_pair
    tables = p._tables()
    shape = (1, 16, 68, 120)
    L = shape[1] * shape[2] * shape[3]
    for exc, want_skip in ((0.0, True), (0.05, False)):
        x, scales = _floor_heavy_latents(shape, 91, cuda_dev, 0.9, exc)
        kw = dict(x=x, scales=scales, scale_table=p.scale_table, scale_bound=0.11, stream_symbols=32768)
        raw = coder.rans_encode(tables, **dict(kw, stream_symbols=0))[0]
        auto = coder.rans_encode(tables, est_bytes=len(raw), **kw)[0]
        plain = coder.rans_encode(tables, skip=False, **kw)[0]
        forced = coder.rans_encode(tables, skip=True, **kw)[0]
        assert coder.container_of(auto, L) == (32768, 32, want_skip)
        assert coder.container_of(plain, L) == (32768, 32, False)
        assert coder.container_of(forced, L) == (32768, 32, True)
        if want_skip:
            assert auto == forced and len(auto) <= len(plain) + 64
        else:
            assert auto == plain and len(forced) > len(plain) + 200
        for s in (auto, plain, forced):
            out = coder.rans_decode([s], tables, shape, scales=scales, scale_table=p.scale_table,
                                    scale_bound=0.11, device=cuda_dev)
            assert torch.equal(out, x)


def test_lane_interleaved_decoder_survives_corrupt_streams(cuda_dev, gc_pair):
    """Random byte flips anywhere in a valid 'DVS3' / 'DVC3' container: the decoder either
    flags the stream (ValueError) or returns symbols -- it never faults, hangs or reads out of
    bounds (every word read is bounds-checked, every loop is bounded by the table or by 64
    bypass nibbles) -- and the next, clean, decode is unaffected."""
    from deepvideocodec_b200 import coder
    _, p = gc_pair
    tables = p._tables()
    shape = (1, 6, 40, 52)
    rng = np.random.default_rng(5)
    for skip, floor in ((True, 0.9), (False, 0.2)):
        x, scales = _floor_heavy_latents(shape, 17, cuda_dev, floor, 0.002)
        x.view(-1)[:4] = torch.tensor([900.0, -70000.0, 3.0, -5.0], device=cuda_dev)   # bypass-coded
        kw = dict(scales=scales, scale_table=p.scale_table, scale_bound=0.11)
        good = coder.rans_encode(tables, x=x, stream_symbols=4096, lanes=32, skip=skip, **kw)[0]
        flagged = 0
        for trial in range(48):
            bad = bytearray(good)
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(16, len(bad)))           # keep the magic / L / S / n words
                bad[pos] ^= int(rng.integers(1, 256))
            try:
                out = coder.rans_decode([bytes(bad)], tables, shape, device=cuda_dev, **kw)
                assert out.shape == x.shape
            except ValueError:
                flagged += 1
        torch.cuda.synchronize()
        assert flagged >= 40                                    # the end-state check catches nearly all
        assert torch.equal(coder.rans_decode([good], tables, shape, device=cuda_dev, **kw), x)


def test_decode_calls_do_not_need_a_host_sync_between_them(cuda_dev, gc_pair):
    """`rans_decode` stages its input through a ring of pinned buffers and one asynchronous
    copy: 24 calls queued back to back (three times the ring), different payloads and sizes, status
    words read once at the end -- every result is right, i.e. no staging slot is overwritten
    before its copy has run."""
    from deepvideocodec_b200 import coder
    _, p = gc_pair
    tables = p._tables()
    cases = []
    for k in range(24):
        shape = (1, 4 + (k % 3) * 2, 24 + 2 * (k % 5), 40)
        x, scales = _floor_heavy_latents(shape, 100 + k, cuda_dev, 0.5 if k % 2 else 0.0, 0.002)
        kw = dict(scales=scales, scale_table=p.scale_table, scale_bound=0.11)
        cases.append((shape, x, kw, coder.rans_encode(tables, x=x, stream_symbols=2048 << (k % 3),
                                                      lanes=32 if k % 4 else 1, **kw)))
    torch.cuda.synchronize()
    sts, outs = [], []
    for shape, x, kw, strings in cases:
        outs.append(coder.rans_decode(strings, tables, shape, device=cuda_dev, statuses=sts, **kw))
    coder.check_decode_status(sts)
    for (shape, x, kw, strings), out in zip(cases, outs):
        assert torch.equal(out, x)


def test_lane_interleaved_wide_rows(cuda_dev, gc_pair):
    """Scales 30 ... 256: table rows of up to ~3 000 symbols, where a look-up key covers more
    table positions than the decoder probes at once (its bisection runs) and symbols beyond the
    tables are bypass-coded.  Bytes equal the oracle's, the oracle decodes them, the GPU decodes
    them with and without the packed tables."""
    from deepvideocodec_b200 import coder
    from oracle import rans
    o, p = gc_pair
    tables = p._tables()
    shape = (2, 8, 36, 52)
    g = torch.Generator().manual_seed(77)
    scales = torch.exp(torch.empty(shape).uniform_(np.log(30.0), np.log(256.0), generator=g))
    x = torch.round(torch.randn(shape, generator=g) * scales)
    x.view(-1)[:3] = torch.tensor([5000.0, -4200.0, 0.0])            # beyond the widest row
    x, scales = x.to(cuda_dev), scales.to(cuda_dev)
    kw = dict(scales=scales, scale_table=p.scale_table, scale_bound=0.11)
    got = coder.rans_encode(tables, x=x, stream_symbols=4096, lanes=32, skip=False, **kw)
    idx = o.build_indexes(scales.cpu()).numpy().reshape(2, -1)
    assert idx.min() >= 45 and idx.max() == 63
    sym = x.int().cpu().numpy().reshape(2, -1)
    cdf, size, off = _oracle_tables(o)
    for n in range(2):
        assert got[n] == rans.encode_container(sym[n], idx[n], cdf, size, off, 4096, 32, None)
        assert np.array_equal(rans.decode_container(got[n], idx[n], cdf, size, off, None), sym[n])
    assert torch.equal(coder.rans_decode(got, tables, shape, device=cuda_dev, **kw), x)
    import unittest.mock as mock
    with mock.patch.object(coder.Tables, "cdf_pack", lambda self: (None, 0)):
        assert torch.equal(coder.rans_decode(got, tables, shape, device=cuda_dev, **kw), x)
