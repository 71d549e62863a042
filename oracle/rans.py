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
