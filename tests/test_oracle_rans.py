"""CPU: the plain-C oracle of the entropy-coding arithmetic (oracle/c/rans_ref.c,
SURVEY.md 8f rows f1/f2) against

* an independent arbitrary-precision restatement of rans64 written here from
  the published recurrences (Python big integers, no 64-bit wrap-around
  anywhere), and hand-derived known-answer streams for tiny inputs;
* the invariants of a quantised CDF;
* the oracle's torch ``build_indexes`` loop;
* encode -> decode round trips through the oracle's CompressAI-shaped modules.

CompressAI itself is absent from /root/reference and not installed: PARITY
UNPINNED for these bit streams (no reference-held vector exists); the
known-answer vectors below are derived by hand from the published algorithm.
Also checks the product's host-side table builder (``dvc_pmf_to_quantized_cdf``
and the modules' ``update()``) against the oracle -- CPU only, no launch.
"""
import struct

import numpy as np
import pytest
import torch

from oracle import rans
from oracle.compressai import entropy_models as oem

L31 = 1 << 31


# ---------------------------------------------------------------------------
# independent big-integer rans64 (Rans64EncPut / EncPutBits / EncFlush and the
# matching decoder), written from the recurrences, used as the known-answer
# generator for small cases
# ---------------------------------------------------------------------------
def _ops_of(sym, ci, cdfs, sizes, offsets):
    cdf = cdfs[ci]
    max_value = int(sizes[ci]) - 2
    value = int(sym) - int(offsets[ci])
    raw = 0
    if value < 0:
        raw, value = -2 * value - 1, max_value
    elif value >= max_value:
        raw, value = 2 * (value - max_value), max_value
    ops = [("sym", int(cdf[value]), int(cdf[value + 1] - cdf[value]))]
    if value == max_value:
        nb = 0
        while (raw >> (4 * nb)) != 0:
            nb += 1
        v = nb
        while v >= 15:
            ops.append(("bits", 15, 0))
            v -= 15
        ops.append(("bits", v, 0))
        for j in range(nb):
            ops.append(("bits", (raw >> (4 * j)) & 15, 0))
    return ops


def big_encode(symbols, indexes, cdfs, sizes, offsets):
    ops = []
    for s, ci in zip(symbols, indexes):
        ops += _ops_of(s, ci, cdfs, sizes, offsets)
    x, words = L31, []
    for kind, a, freq in reversed(ops):
        if kind == "sym":
            if x >= ((L31 >> 16) << 32) * freq:
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // freq) << 16) + (x % freq) + a
        else:
            if x >= ((L31 >> 16) << 32) * (1 << 12):
                words.append(x & 0xFFFFFFFF)
                x >>= 32
            x = (x << 4) | a
    words.append(x >> 32)
    words.append(x & 0xFFFFFFFF)
    words.reverse()
    return struct.pack(f"<{len(words)}I", *words)


def _tables(n_tab=5, width=9, seed=0):
    """Random strictly increasing 16-bit CDF rows of ragged length."""
    g = np.random.default_rng(seed)
    cdfs = np.zeros((n_tab, width + 2), dtype=np.int32)
    sizes = np.zeros(n_tab, dtype=np.int32)
    offsets = np.zeros(n_tab, dtype=np.int32)
    for i in range(n_tab):
        n = int(g.integers(2, width + 1))              # pmf entries incl. the escape entry
        pmf = g.random(n).astype(np.float32) ** 3 + 1e-4
        pmf /= pmf.sum()
        row = rans.pmf_to_quantized_cdf(pmf)
        cdfs[i, :row.size] = row
        sizes[i] = row.size
        offsets[i] = -int(g.integers(0, n))
    return cdfs, sizes, offsets


def test_known_answer_single_symbol():
    # cdf [0, 2^15, 2^16], symbol 0: x = 2^31 -> (2^31 / 2^15) << 16 = 2^32;
    # flush writes the low word then the high word
    cdfs = np.array([[0, 32768, 65536]], dtype=np.int32)
    s = rans.encode_with_indexes([0], [0], cdfs, [3], [0])
    assert s == bytes([0, 0, 0, 0, 1, 0, 0, 0])
    assert rans.decode_with_indexes(s, np.zeros(1, np.int32), cdfs, [3], [0]).tolist() == [0]


def test_known_answer_two_symbols_and_escape():
    # table: values {0,1} + escape entry; cdf = [0, 16384, 49152, 65536], offset 0
    cdfs = np.array([[0, 16384, 49152, 65536]], dtype=np.int32)
    # symbols [1, 0], coded last first:
    #   put 0: start 0, freq 2^14: x = (2^31 >> 14) << 16          = 2^33
    #   put 1: start 2^14, freq 2^15: x = (2^33 >> 15) << 16 + 2^14 = 2^34 + 2^14
    s = rans.encode_with_indexes([1, 0], [0, 0], cdfs, [4], [0])
    x = (1 << 34) + (1 << 14)
    assert s == struct.pack("<2I", x & 0xFFFFFFFF, x >> 32)
    # symbol 5 -> escape entry (start 49152, freq 16384), raw = 2*(5-2) = 6,
    # one payload nibble: ops = sym, bits(1), bits(6); coded in reverse:
    #   bits 6: x = 2^31 << 4 | 6 ; bits 1: x = x << 4 | 1 ;
    #   sym: x = (x // 2^14) << 16 + x % 2^14 + 49152
    x = ((L31 << 4 | 6) << 4) | 1
    x = ((x >> 14) << 16) + (x & 16383) + 49152
    assert x < (1 << 63)
    s = rans.encode_with_indexes([5], [0], cdfs, [4], [0])
    assert s == struct.pack("<2I", x & 0xFFFFFFFF, x >> 32)
    assert rans.decode_with_indexes(s, np.zeros(1, np.int32), cdfs, [4], [0]).tolist() == [5]
    # a negative out-of-table symbol: raw = -2*(-3) - 1 = 5
    s = rans.encode_with_indexes([-3], [0], cdfs, [4], [0])
    assert rans.decode_with_indexes(s, np.zeros(1, np.int32), cdfs, [4], [0]).tolist() == [-3]


@pytest.mark.parametrize("n", [1, 2, 7, 64, 1000])
def test_c_oracle_matches_big_integer_rans(n):
    cdfs, sizes, offsets = _tables(seed=n)
    g = np.random.default_rng(100 + n)
    idx = g.integers(0, cdfs.shape[0], n).astype(np.int32)
    sym = np.array([g.integers(-4, sizes[i] + 3) + offsets[i] for i in idx], dtype=np.int32)
    sym[::11] += 40000                                   # long escapes (many nibbles)
    sym[5::13] -= 70000
    got = rans.encode_with_indexes(sym, idx, cdfs, sizes, offsets)
    want = big_encode(sym.tolist(), idx.tolist(), cdfs, sizes, offsets)
    assert got == want
    back = rans.decode_with_indexes(got, idx, cdfs, sizes, offsets)
    assert np.array_equal(back, sym)


def test_largest_escapes():
    # the escape payload is held in 32 bits (upstream): |symbol| <= 2^30 is the
    # domain; the nibble count is then <= 8, always a single count nibble
    cdfs, sizes, offsets = _tables(seed=3)
    sym = np.array([2**30, -2**30, 0, 2**27, -2**27 - 1, 2**28 - 1], dtype=np.int32)
    idx = np.zeros(sym.size, dtype=np.int32)
    s = rans.encode_with_indexes(sym, idx, cdfs, sizes, offsets)
    assert s == big_encode(sym.tolist(), idx.tolist(), cdfs, sizes, offsets)
    assert np.array_equal(rans.decode_with_indexes(s, idx, cdfs, sizes, offsets), sym)


def test_pmf_to_quantized_cdf_invariants_and_errors():
    g = np.random.default_rng(1)
    for n in (1, 2, 3, 17, 300, 2000):
        pmf = g.random(n).astype(np.float32) ** 8          # many near-zero entries
        pmf /= pmf.sum()
        cdf = rans.pmf_to_quantized_cdf(pmf)
        assert cdf.shape == (n + 1,) and cdf[0] == 0 and cdf[-1] == 65536
        assert np.all(np.diff(cdf) >= 1)
    with pytest.raises(ValueError):
        rans.pmf_to_quantized_cdf(np.array([0.5, -0.1], np.float32))
    with pytest.raises(ValueError):
        rans.pmf_to_quantized_cdf(np.array([0.5, np.nan], np.float32))
    with pytest.raises(ValueError):
        rans.pmf_to_quantized_cdf(np.zeros(4, np.float32))
    # exact small case: [0.5, 0.25, 0.25] -> [0, 32768, 49152, 65536]
    assert rans.pmf_to_quantized_cdf(np.array([0.5, 0.25, 0.25], np.float32)).tolist() == \
        [0, 32768, 49152, 65536]
    # a zero-probability entry steals one count from the narrowest donor
    assert rans.pmf_to_quantized_cdf(np.array([0.75, 0.0, 0.25], np.float32)).tolist() == \
        [0, 49152, 49153, 65536]


def test_product_host_cdf_builder_equals_oracle():
    from deepvideocodec_b200 import coder
    g = np.random.default_rng(2)
    for n in (1, 2, 5, 64, 777, 4000):
        pmf = g.random(n).astype(np.float32) ** 6
        pmf /= pmf.sum()
        assert np.array_equal(coder.pmf_to_quantized_cdf(pmf), rans.pmf_to_quantized_cdf(pmf))
    with pytest.raises(ValueError):
        coder.pmf_to_quantized_cdf(np.array([0.5, -0.1], np.float32))
    with pytest.raises(ValueError):
        coder.pmf_to_quantized_cdf(np.zeros(3, np.float32))


def _scale_table():
    # dmc/models/base_model.py:43-49
    return np.exp(np.linspace(np.log(0.11), np.log(256), 64)).tolist()


def test_build_indexes_c_equals_torch_loop():
    torch.manual_seed(0)
    gc = oem.GaussianConditional(None)
    gc.update_scale_table(_scale_table())
    tab = gc.scale_table
    scales = torch.exp(torch.empty(4000).uniform_(np.log(0.01), np.log(600)))
    scales[:64] = tab                                     # exactly on the table
    scales[64:127] = torch.nextafter(tab[:-1], torch.tensor(1e9))
    scales[127] = float("nan")
    want = gc.build_indexes(scales).numpy()
    got = rans.build_indexes(scales.numpy(), tab.numpy(), 0.11)
    assert np.array_equal(got, want)
    assert got.min() == 0 and got.max() == 63


def test_oracle_modules_round_trip():
    torch.manual_seed(0)
    gc = oem.GaussianConditional(None)
    assert gc.update_scale_table(_scale_table()) is True
    assert gc.update_scale_table(_scale_table()) is False
    assert gc._quantized_cdf.shape[0] == 64 and int(gc._cdf_length.max()) == gc._quantized_cdf.shape[1]
    scales = torch.exp(torch.empty(2, 6, 8, 10).uniform_(np.log(0.05), np.log(40)))
    means = torch.randn(2, 6, 8, 10) * 3
    y = means + scales * torch.randn(2, 6, 8, 10)
    y[0, 0, 0, :4] += torch.tensor([4000.0, -4000.0, 70000.0, -70000.0])   # escapes
    idx = gc.build_indexes(scales)
    strings = gc.compress(y, idx, means)
    assert len(strings) == 2 and all(isinstance(s, bytes) and len(s) % 4 == 0 for s in strings)
    back = gc.decompress(strings, idx, means=means)
    assert torch.equal(back, torch.round(y - means) + means)
    eb = oem.EntropyBottleneck(7)
    assert eb.update() is True and eb.update() is False and eb.update(force=True) is True
    z = torch.randn(3, 7, 5, 4) * 6
    zs = eb.compress(z)
    zb = eb.decompress(zs, z.shape[-2:])
    med = eb._get_medians().detach().reshape(1, -1, 1, 1)
    assert torch.equal(zb, torch.round(z - med) + med)
    # real bits track the estimated rate of the same symbols (16-bit tables)
    _, lik = eb(z, training=False)
    est = float(-torch.log2(lik).sum())
    real = 8 * sum(len(s) for s in zs)
    assert abs(real - est) < 0.03 * est + 64 * len(zs)


def test_product_update_tables_equal_oracle_tables():
    """The product modules' ``update()`` (plain torch + dvc_pmf_to_quantized_cdf,
    host side, setup time) builds the same tables as the oracle's on CPU."""
    import deepvideocodec_b200 as dvc
    torch.manual_seed(3)
    a, b = oem.GaussianConditional(None), dvc.GaussianConditional(None)
    a.update_scale_table(_scale_table())
    assert b.update_scale_table(_scale_table()) is True
    assert b.update_scale_table(_scale_table()) is False
    for name in ("_quantized_cdf", "_cdf_length", "_offset", "scale_table"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    torch.manual_seed(4)
    ea = oem.EntropyBottleneck(64)
    torch.manual_seed(4)
    eb = dvc.EntropyBottleneck(64)
    with torch.no_grad():
        for m in (ea, eb):
            m.quantiles[:, 0, 0] = -torch.linspace(3.2, 40, 64)
            m.quantiles[:, 0, 2] = torch.linspace(1.5, 25, 64)
            m.quantiles[:, 0, 1] = torch.linspace(-2, 2, 64)
    assert ea.update() and eb.update()
    for name in ("_quantized_cdf", "_cdf_length", "_offset"):
        assert torch.equal(getattr(ea, name), getattr(eb, name)), name
    assert eb._quantized_cdf.dtype == torch.int32
    # state_dict carries the tables (DMC.load_state_dict resizes them, video_model.py:626-656)
    sd = eb.state_dict()
    assert {"_quantized_cdf", "_cdf_length", "_offset"} <= set(sd)


def test_product_coder_refuses_cpu_tensors():
    import deepvideocodec_b200 as dvc
    gc = dvc.GaussianConditional(None)
    gc.update_scale_table(_scale_table())
    y = torch.zeros(1, 2, 4, 4)
    with pytest.raises(dvc.DvcError):
        gc.build_indexes(torch.ones(1, 2, 4, 4))
    with pytest.raises(dvc.DvcError):
        gc.compress(y, torch.zeros(1, 2, 4, 4, dtype=torch.int32))
    fresh = dvc.GaussianConditional(None)
    with pytest.raises(ValueError, match="update"):
        fresh.compress(y, torch.zeros(1, 2, 4, 4, dtype=torch.int32))


def test_lane_interleaved_restatement_round_trip_and_known_answer():
    """The repository's own lane-interleaved sub-stream (oracle/c/rans_ref.c::dvcref_ilv_*):
    round trips, and a hand-computed stream for a single symbol: mask word, 32 one-word states,
    lane 0 = the stock state after one Rans64EncPut, the other lanes untouched at 2^31."""
    rng = np.random.default_rng(11)
    rows = [rans.pmf_to_quantized_cdf(np.array(p, dtype=np.float32)) for p in (
        [3e-5, 1 - 7e-5, 3e-5, 1e-5], [0.1, 0.2, 0.39, 0.2, 0.1, 0.01], [1 / 41] * 41)]
    cdf = np.zeros((3, max(len(r) for r in rows)), np.int32)
    for i, r in enumerate(rows):
        cdf[i, :len(r)] = r
    sizes = np.array([len(r) for r in rows], np.int32)
    offs = np.array([-1, -2, -20], np.int32)
    marks = rans.skip_rows_of(cdf, sizes, offs)
    assert marks.tolist() == [1, 0, 0]
    for L in (1, 32, 33, 1023, 1024, 1025, 3000):
        for p_floor in (0.0, 0.6, 0.98):
            idx = np.where(rng.random(L) < p_floor, 0, rng.integers(1, 3, L)).astype(np.int32)
            sym = np.where(idx == 0, (rng.random(L) < 0.01) * rng.integers(-3, 4, L),
                           rng.integers(-30, 31, L)).astype(np.int32)
            sym[rng.random(L) < 0.01] = 123456               # bypass-coded
            for sk in (None, marks):
                b = rans.ilv_encode(sym, idx, cdf, sizes, offs, sk)
                assert np.array_equal(rans.ilv_decode(b, idx, cdf, sizes, offs, sk), sym)
                c = rans.encode_container(sym, idx, cdf, sizes, offs, 700, 32, sk)
                assert np.array_equal(rans.decode_container(c, idx, cdf, sizes, offs, sk), sym)
    # known answer: symbol value 1 of row 1 (table position 3): start = cdf[1][3], freq = next - start
    start, freq = int(cdf[1, 3]), int(cdf[1, 4] - cdf[1, 3])
    x = ((1 << 31) // freq << 16) + (1 << 31) % freq + start
    assert (1 << 32) <= x < (1 << 63)                        # p = 0.2: the state needs its high word
    want = np.array([1, x & 0xFFFFFFFF, x >> 32] + [1 << 31] * 31, dtype=np.uint32).tobytes()
    assert rans.ilv_encode(np.array([1]), np.array([1]), cdf, sizes, offs) == want
    # a truncated stream is rejected by the restatement's end-state check
    b = rans.ilv_encode(np.arange(-20, 20), np.full(40, 2), cdf, sizes, offs)
    with pytest.raises(RuntimeError):
        rans.ilv_decode(b[:-4], np.full(40, 2), cdf, sizes, offs)
