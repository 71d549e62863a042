"""ctypes binding of ``oracle/c/librans_ref.so`` -- TEST INFRASTRUCTURE ONLY.

numpy-facing wrappers of the plain-C restatement of CompressAI's entropy-coding
arithmetic (``oracle/c/rans_ref.c``: ``pmf_to_quantized_cdf``, stock single
stream ``encode_with_indexes`` / ``decode_with_indexes``, ``build_indexes``).
Built by ``__graft_entry__.build()`` (``make -C oracle/c``).
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "librans_ref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
                os.path.join(_DIR, "rans_ref.c")):
            subprocess.run(["make", "-s", "-C", _DIR], check=True)
        h = ctypes.CDLL(_SO)
        h.dvcref_pmf_to_quantized_cdf.restype = c_int
        h.dvcref_pmf_to_quantized_cdf.argtypes = [POINTER(c_float), c_int, c_int, POINTER(c_uint32)]
        h.dvcref_build_indexes.restype = None
        h.dvcref_build_indexes.argtypes = [POINTER(c_float), c_int64, POINTER(c_float), c_int,
                                           c_float, POINTER(c_int32)]
        h.dvcref_rans_encode_with_indexes.restype = c_int64
        h.dvcref_rans_encode_with_indexes.argtypes = [
            POINTER(c_int32), POINTER(c_int32), c_int64, POINTER(c_int32), c_int64,
            POINTER(c_int32), POINTER(c_int32), c_int32, POINTER(c_uint8), c_int64]
        h.dvcref_rans_decode_with_indexes.restype = c_int64
        h.dvcref_rans_decode_with_indexes.argtypes = [
            POINTER(c_uint8), c_int64, POINTER(c_int32), c_int64, POINTER(c_int32), c_int64,
            POINTER(c_int32), POINTER(c_int32), c_int32, POINTER(c_int32)]
        h.dvcref_ilv_encode.restype = c_int64
        h.dvcref_ilv_encode.argtypes = [
            POINTER(c_int32), POINTER(c_int32), c_int64, POINTER(c_int32), c_int64,
            POINTER(c_int32), POINTER(c_int32), c_int32, POINTER(c_uint8), POINTER(c_uint32),
            c_int64]
        h.dvcref_ilv_decode.restype = c_int64
        h.dvcref_ilv_decode.argtypes = [
            POINTER(c_uint32), c_int64, POINTER(c_int32), c_int64, POINTER(c_int32), c_int64,
            POINTER(c_int32), POINTER(c_int32), c_int32, POINTER(c_uint8), POINTER(c_int32)]
        _lib = h
    return _lib


def _p(a, t):
    return a.ctypes.data_as(POINTER(t))


def pmf_to_quantized_cdf(pmf, precision=16):
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    cdf = np.zeros(pmf.size + 1, dtype=np.uint32)
    rc = lib().dvcref_pmf_to_quantized_cdf(_p(pmf, c_float), pmf.size, precision, _p(cdf, c_uint32))
    if rc == -1:
        raise ValueError("Invalid `pmf`, non-finite or negative element found")
    if rc == -2:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability")
    if rc != 0:
        raise RuntimeError(f"pmf_to_quantized_cdf failed ({rc})")
    return cdf.astype(np.int32)


def build_indexes(scales, table, bound):
    s = np.ascontiguousarray(scales, dtype=np.float32)
    t = np.ascontiguousarray(table, dtype=np.float32)
    out = np.empty(s.shape, dtype=np.int32)
    lib().dvcref_build_indexes(_p(s, c_float), s.size, _p(t, c_float), t.size, float(bound),
                               _p(out, c_int32))
    return out


def _tables(cdfs, cdf_sizes, offsets):
    cdfs = np.ascontiguousarray(cdfs, dtype=np.int32)
    assert cdfs.ndim == 2
    return (cdfs, np.ascontiguousarray(cdf_sizes, dtype=np.int32).reshape(-1),
            np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1))


def encode_with_indexes(symbols, indexes, cdfs, cdf_sizes, offsets):
    """Stock CompressAI single-stream encoding -> ``bytes``."""
    sym = np.ascontiguousarray(symbols, dtype=np.int32).reshape(-1)
    idx = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
    assert sym.size == idx.size
    cdfs, sizes, offs = _tables(cdfs, cdf_sizes, offsets)
    cap = 8 * sym.size + 64
    out = np.empty(cap, dtype=np.uint8)
    n = lib().dvcref_rans_encode_with_indexes(
        _p(sym, c_int32), _p(idx, c_int32), sym.size, _p(cdfs, c_int32), cdfs.shape[1],
        _p(sizes, c_int32), _p(offs, c_int32), cdfs.shape[0], _p(out, c_uint8), cap)
    if n < 0:
        raise RuntimeError(f"encode_with_indexes failed ({n})")
    return out[:n].tobytes()


def decode_with_indexes(encoded, indexes, cdfs, cdf_sizes, offsets):
    """Stock CompressAI single-stream decoding -> int32 array shaped like ``indexes``."""
    idx = np.ascontiguousarray(indexes, dtype=np.int32)
    cdfs, sizes, offs = _tables(cdfs, cdf_sizes, offsets)
    # the decoder may peek one word past the stream when it renormalises last
    enc = np.frombuffer(bytes(encoded) + b"\0" * 8, dtype=np.uint8).copy()
    out = np.empty(idx.size, dtype=np.int32)
    n = lib().dvcref_rans_decode_with_indexes(
        _p(enc, c_uint8), len(encoded), _p(idx.reshape(-1), c_int32), idx.size,
        _p(cdfs, c_int32), cdfs.shape[1], _p(sizes, c_int32), _p(offs, c_int32), cdfs.shape[0],
        _p(out, c_int32))
    if n < 0:
        raise RuntimeError(f"decode_with_indexes failed ({n})")
    return out.reshape(idx.shape)


# ---------------------------------------------------------------------------
# The GPU containers (csrc/dvc_coder.cu, include/dvc_b200.h), restated on the CPU.  These are
# the repository's OWN layouts around the stock per-symbol arithmetic:
#   'DVC1'  L symbols cut into sub-streams of S; each one a stock stream of its slice
#   'DVC3'  each sub-stream coded by 32 interleaved stock rans64 coders (rans_ref.c)
#   'DVS3'  'DVC3' + implied zeros: symbols of marked table rows are coded as one flag per
#           group of 32 positions, and individually only in flagged groups
# ---------------------------------------------------------------------------
MAGIC_DVC1 = 0x31435644
MAGIC_DVC3 = 0x33435644
MAGIC_DVS3 = 0x33535644


def skip_rows_of(cdfs, cdf_sizes, offsets, min_freq=(1 << 16) - 8):
    """One byte per table row: 1 where value 0 (table position -offset) holds at least
    ``min_freq`` of the 2^16 probability mass."""
    cdfs, sizes, offs = _tables(cdfs, cdf_sizes, offsets)
    out = np.zeros(cdfs.shape[0], dtype=np.uint8)
    for r in range(cdfs.shape[0]):
        p = -int(offs[r])
        if 0 <= p < int(sizes[r]) - 2 and int(cdfs[r, p + 1]) - int(cdfs[r, p]) >= min_freq:
            out[r] = 1
    return out


def ilv_encode(symbols, indexes, cdfs, cdf_sizes, offsets, skip_rows=None):
    """One lane-interleaved sub-stream -> ``bytes``."""
    sym = np.ascontiguousarray(symbols, dtype=np.int32).reshape(-1)
    idx = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
    cdfs, sizes, offs = _tables(cdfs, cdf_sizes, offsets)
    marks = None if skip_rows is None else np.ascontiguousarray(skip_rows, dtype=np.uint8)
    cap = 2 * sym.size + 160
    out = np.empty(cap, dtype=np.uint32)
    n = lib().dvcref_ilv_encode(
        _p(sym, c_int32), _p(idx, c_int32), sym.size, _p(cdfs, c_int32), cdfs.shape[1],
        _p(sizes, c_int32), _p(offs, c_int32), cdfs.shape[0],
        None if marks is None else _p(marks, c_uint8), _p(out, c_uint32), cap)
    if n < 0:
        raise RuntimeError(f"ilv_encode failed ({n})")
    return out[:n].tobytes()


def ilv_decode(encoded, indexes, cdfs, cdf_sizes, offsets, skip_rows=None):
    idx = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
    cdfs, sizes, offs = _tables(cdfs, cdf_sizes, offsets)
    marks = None if skip_rows is None else np.ascontiguousarray(skip_rows, dtype=np.uint8)
    w = np.frombuffer(bytes(encoded), dtype=np.uint32).copy()
    out = np.zeros(idx.size, dtype=np.int32)
    n = lib().dvcref_ilv_decode(
        _p(w, c_uint32), w.size, _p(idx, c_int32), idx.size, _p(cdfs, c_int32), cdfs.shape[1],
        _p(sizes, c_int32), _p(offs, c_int32), cdfs.shape[0],
        None if marks is None else _p(marks, c_uint8), _p(out, c_int32))
    if n < 0:
        raise RuntimeError(f"ilv_decode failed ({n})")
    return out


def encode_container(symbols, indexes, cdfs, cdf_sizes, offsets, stream_symbols, lanes=1,
                     skip_rows=None):
    sym = np.ascontiguousarray(symbols, dtype=np.int32).reshape(-1)
    idx = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
    L, S = sym.size, int(stream_symbols)
    ns = (L + S - 1) // S
    subs = []
    for j in range(ns):
        s, i = sym[j * S:(j + 1) * S], idx[j * S:(j + 1) * S]
        if lanes == 1:
            subs.append(encode_with_indexes(s, i, cdfs, cdf_sizes, offsets))
        else:
            subs.append(ilv_encode(s, i, cdfs, cdf_sizes, offsets, skip_rows))
    magic = MAGIC_DVC1 if lanes == 1 else (MAGIC_DVC3 if skip_rows is None else MAGIC_DVS3)
    header = np.array([magic, L, S, ns] + [len(b) // 4 for b in subs], dtype=np.uint32).tobytes()
    return header + b"".join(subs)


def decode_container(encoded, indexes, cdfs, cdf_sizes, offsets, skip_rows=None):
    idx = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
    w = np.frombuffer(encoded, dtype=np.uint32)
    magic, L, S, ns = (int(v) for v in w[:4])
    assert L == idx.size and ns == (L + S - 1) // S
    assert magic in (MAGIC_DVC1, MAGIC_DVC3, MAGIC_DVS3)
    assert (magic == MAGIC_DVS3) == (skip_rows is not None)
    lens = w[4:4 + ns].astype(np.int64)
    pos = 4 + ns
    out = np.zeros(L, dtype=np.int32)
    for j in range(ns):
        body = w[pos:pos + int(lens[j])].tobytes()
        pos += int(lens[j])
        i = idx[j * S:(j + 1) * S]
        if magic == MAGIC_DVC1:
            out[j * S:(j + 1) * S] = decode_with_indexes(body, i, cdfs, cdf_sizes, offsets)
        else:
            out[j * S:(j + 1) * S] = ilv_decode(body, i, cdfs, cdf_sizes, offsets, skip_rows)
    assert pos == w.size
    return out.reshape(np.asarray(indexes).shape)
