"""Entropy-coder inputs and range-ANS bit streams on the GPU (SURVEY.md 8f rows
f1 / f2) -- Python face of ``csrc/dvc_coder.cu``.

These are the pieces of CompressAI the reference reaches only on its
real-bitstream path (``dmc/test.py:187-188`` -> ``DMC.encode_inter`` /
``decode_inter``): ``build_indexes``, ``quantize(.., "symbols")``,
``compress`` and ``decompress`` of both entropy models
(``dmc/models/video_model.py:238-283, 411-458``) and the CDF-table
construction behind ``DMC.update`` (``:669-677``).

A bit stream returned here is a ``bytes`` object per batch sample, like
CompressAI's.  With ``stream_symbols > 0`` it is one of this repository's
containers around the stock rans64 arithmetic -- ``DVC3`` / ``DVS3`` (default:
every sub-stream coded by the 32 lanes of one warp, optionally with implied zeros
for near-deterministic table rows) or ``DVC1`` (one stock stream per sub-stream),
see ``include/dvc_b200.h`` -- with ``stream_symbols == 0`` it is one raw stock
stream, byte-compatible with CompressAI's ``RansEncoder.encode_with_indexes``.
The decoder tells them apart from the first word.
"""
import os
import functools
import struct
import threading

import numpy as np
import torch

from . import _native as nat

__all__ = ["DEFAULT_STREAM_SYMBOLS", "PINNED_STREAM_SYMBOLS", "MAGIC", "MAGIC_LANES", "MAGIC_LANES_SKIP",
           "DEFAULT_LANES", "DEFAULT_SKIP", "SKIP_MIN_FREQ", "PendingStreams", "auto_stream_symbols",
           "build_indexes", "check_decode_status", "collect", "container_of",
           "decode_stage_a", "decode_stage_b", "pmf_to_quantized_cdf", "quantize_symbols",
           "rans_decode", "rans_encode", "rans_encode_async", "stream_symbols_of"]

MAGIC = 0x31435644                     # "DVC1": one stock rans64 stream per sub-stream (lanes = 1)
MAGIC_LANES = 0x33435644               # "DVC3": 32 lane-interleaved rans64 coders per sub-stream
MAGIC_LANES_SKIP = 0x33535644          # "DVS3": "DVC3" + implied zeros (group flags)
SKIP_MIN_FREQ = (1 << 16) - 8          # a table row is marked when value 0 holds this much mass
# Layout of the container (stream_symbols > 0): 32 = lane-interleaved (default), 1 = 'DVC1'
DEFAULT_LANES = int(os.environ.get("DVC_RANS_LANES", "32"))
DEFAULT_SKIP = os.environ.get("DVC_RANS_SKIP", "1") != "0"
_ENV_S = os.environ.get("DVC_RANS_STREAM_SYMBOLS")
# Sub-stream length when nothing is known about the payload (module-level compress()):
DEFAULT_STREAM_SYMBOLS = int(_ENV_S) if _ENV_S is not None else 4096
# A pinned length overrides the payload-driven policy (DVC_RANS_STREAM_SYMBOLS; 0 = raw stock stream)
PINNED_STREAM_SYMBOLS = int(_ENV_S) if _ENV_S is not None else None
MIN_STREAMS = int(os.environ.get("DVC_RANS_MIN_STREAMS", "8"))
MIN_STREAM_SYMBOLS = 256
STREAM_OVERHEAD_BYTES = 12             # per sub-stream: 4 B length word + 8 B rans64 flush
LANES_OVERHEAD_BYTES = 200             # lanes = 32: length + mask words + 32 states of 1-2 words
LANES_STREAM_SYMBOLS = 65536           # lanes = 32, payload unknown: 2 048 rounds per sub-stream
LANES_CHUNK = 1024                     # positions per chunk of the lane-interleaved kernels
SKIP_FLAG_BUDGET = 0.005               # adaptive implied zeros: set flags may cost this share of the payload
SKIP_FLAG_BITS = 12.0                  #   at ~12 bits each
OVERHEAD_TARGET = float(os.environ.get("DVC_RANS_OVERHEAD", "0.01"))


def auto_stream_symbols(n_symbols, est_bytes=None, lanes=1):
    """Sub-stream length used when the caller does not choose one.

    Range coding is serial inside a chain, so more chains are faster, but every
    chain costs its flush bytes.  The policy is driven by the PAYLOAD: with an
    estimate of the coded size (the fused ``sum ln p`` of the likelihood kernel,
    free in the fused context models) the number of sub-streams is the largest
    that keeps the container overhead at ``DVC_RANS_OVERHEAD`` (1 %) of it.

    ``lanes = 32`` (``DVC3`` / ``DVS3``, the default layout): a sub-stream is 32
    chains for ~200 bytes and a whole number of 1024-position chunks; at least
    one sub-stream (one warp: ``n_symbols / 32`` rounds, far fewer with implied
    zeros); without an estimate ``LANES_STREAM_SYMBOLS``.
    ``lanes = 1`` (``DVC1``): 12 bytes per chain; at least ``DVC_RANS_MIN_STREAMS``
    (8; 96 bytes) so that a near-empty tensor does not decode as one serial chain;
    without an estimate ``DEFAULT_STREAM_SYMBOLS``.  A fixed 4 096 symbols was
    +0.3 % bytes at 7 bits/symbol but +15 ... 80 % at the 0.03 - 0.15 bits/symbol
    of a low-rate P-frame.
    ``DVC_RANS_STREAM_SYMBOLS`` pins a length (0 = one raw stock stream,
    byte-compatible with CompressAI, for interop).  The length is recorded in the
    container header, so decoders need no matching policy."""
    if PINNED_STREAM_SYMBOLS is not None:
        return PINNED_STREAM_SYMBOLS
    if lanes == 32 and DEFAULT_STREAM_SYMBOLS > 0:
        # 32 chains per sub-stream for ~200 bytes: the sub-stream count follows the payload
        # (>= 1: one warp, n_symbols / 32 rounds -- and far fewer with implied zeros)
        if est_bytes is None:
            return LANES_STREAM_SYMBOLS
        n = max(1, int(OVERHEAD_TARGET * float(est_bytes) // LANES_OVERHEAD_BYTES))
        s = (n_symbols + n - 1) // n
        return max(LANES_CHUNK, (s + LANES_CHUNK - 1) // LANES_CHUNK * LANES_CHUNK)
    if est_bytes is None or DEFAULT_STREAM_SYMBOLS <= 0:
        return DEFAULT_STREAM_SYMBOLS
    n = int(OVERHEAD_TARGET * float(est_bytes) // STREAM_OVERHEAD_BYTES)
    n = max(n, MIN_STREAMS, 1)
    n = min(n, max(1, (n_symbols + MIN_STREAM_SYMBOLS - 1) // MIN_STREAM_SYMBOLS))
    s = (n_symbols + n - 1) // n
    return max(MIN_STREAM_SYMBOLS, (s + 255) // 256 * 256)     # whole staging chunks


_side_streams = {}


def _side_stream(device):
    s = _side_streams.get(device.index)
    if s is None:
        s = torch.cuda.Stream(device=device)
        _side_streams[device.index] = s
    return s


class _PinnedRing:
    """A few pinned host buffers per device and thread, reused round robin: slot k is handed
    out again once the copy that last read it has completed (an event; normally long past).
    torch's own pinned allocator is not used per call because a block freed while its copy is
    still queued cannot be reused yet, and a fresh ``cudaHostAlloc`` costs more than the
    decode it would feed."""

    def __init__(self, slots=8):
        self.bufs = [None] * slots
        self.events = [None] * slots
        self.k = 0

    def take(self, nbytes):
        k = self.k
        self.k = (k + 1) % len(self.bufs)
        if self.events[k] is not None:
            self.events[k].synchronize()
            self.events[k] = None
        b = self.bufs[k]
        if b is None or b.numel() < nbytes:
            size = 1 << max(16, int(nbytes - 1).bit_length())
            b = torch.empty(size, dtype=torch.uint8, pin_memory=True)
            self.bufs[k] = b
        return k, b

    def release(self, k, stream):
        ev = torch.cuda.Event()
        ev.record(stream)
        self.events[k] = ev


_staging_local = threading.local()


def _staging(device):
    rings = getattr(_staging_local, "rings", None)
    if rings is None:
        rings = _staging_local.rings = {}
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    r = rings.get(idx)
    if r is None:
        r = rings[idx] = _PinnedRing()
    return r


def _dev_i32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.int32:
        raise nat.DvcError(f"{name}: expected a CUDA int32 tensor (run update() and move the "
                           f"module to the GPU); there is no CPU coder in deepvideocodec_b200")
    return t.contiguous()


class Tables:
    """Device views of a module's ``_quantized_cdf`` / ``_cdf_length`` / ``_offset``."""

    def __init__(self, quantized_cdf, cdf_length, offset):
        if quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if quantized_cdf.dim() != 2:
            raise ValueError(f"Invalid CDF size {quantized_cdf.size()}")
        if offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if offset.dim() != 1:
            raise ValueError(f"Invalid offsets size {offset.size()}")
        if cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if cdf_length.dim() != 1:
            raise ValueError(f"Invalid offsets size {cdf_length.size()}")
        self.cdf = _dev_i32(quantized_cdf, "_quantized_cdf")
        self.size = _dev_i32(cdf_length, "_cdf_length")
        self.offset = _dev_i32(offset, "_offset")
        if self.size.numel() != self.cdf.size(0) or self.offset.numel() != self.cdf.size(0):
            raise ValueError("CDF tables disagree on the number of distributions")

    def skip_rows(self):
        """uint8 ``[n_cdf]`` on the device: 1 where value 0 (table position
        ``-offset``) holds at least ``SKIP_MIN_FREQ`` / 65536 of the row's mass.
        Encoder and decoder derive it from the tables; symbols of such rows are
        coded as one flag per group of 32 positions (``DVS3``, include/dvc_b200.h).
        Cached on the CDF buffer, keyed on the buffers' versions."""
        key = (self.cdf._version, self.size.data_ptr(), self.size._version,
               self.offset.data_ptr(), self.offset._version)
        hit = getattr(self.cdf, "_dvc_skip_rows", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        n = self.cdf.size(0)
        pos = (-self.offset).long()
        ok = (pos >= 0) & (pos < (self.size.long() - 2))
        p = pos.clamp(0, self.cdf.size(1) - 2)
        rows = torch.arange(n, device=self.cdf.device)
        freq = self.cdf[rows, p + 1] - self.cdf[rows, p]
        marks = (ok & (freq >= SKIP_MIN_FREQ)).to(torch.uint8).contiguous()
        try:
            self.cdf._dvc_skip_rows = (key, marks)
        except AttributeError:
            pass
        return marks

    def cdf_pack(self):
        """``(blob, entries)``: the tables re-packed for the decoder's shared memory
        (``dvc_rans_decode(cdf_pack)``, layout in include/dvc_b200.h) -- a look-up
        from the 16-bit ``cum`` to the first table position it can fall on
        (``_lut_ranges``), the row offsets, and the rows back to back as uint16
        ``(value - 1) mod 2^16`` with 4 pad entries each -- or ``(None, 0)`` for tables the kernel does not stage
        (more than 256 rows, more than 124 KB).  Never changes a result.  Cached on
        the CDF buffer."""
        key = (self.cdf._version, self.size.data_ptr(), self.size._version)
        hit = getattr(self.cdf, "_dvc_cdf_pack", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        n, width = self.cdf.shape
        out = (None, 0)
        sizes = self.size.long()
        total = int(sizes.sum().item()) + PACK_PAD * n
        start_off = n * LUT_STRIDE * 2
        tbl_off = start_off + 4 * n
        nbytes = (tbl_off + 2 * total + 15) // 16 * 16
        if n <= 256 and nbytes <= PACK_MAX_BYTES and int(sizes.min().item()) >= 2 and width < 65536:
            dev = self.cdf.device
            cols = torch.arange(width, device=dev)
            inside = cols[None, :] < sizes[:, None]
            rows = torch.where(inside, self.cdf, torch.full_like(self.cdf, 1 << 17))   # sorted rows
            cmin = torch.from_numpy(_lut_ranges()[0]).to(dev).expand(n, LUT_KEYS).contiguous()
            lo = (torch.searchsorted(rows, cmin, right=True) - 1).clamp_(min=0)
            lo = torch.cat((lo, (sizes - 2)[:, None], torch.zeros_like(sizes)[:, None]), 1)   # sentinel, pad
            lut = torch.where(lo > 32767, lo - 65536, lo).to(torch.int16)     # u16 bit patterns
            starts = torch.cumsum(sizes + PACK_PAD, 0) - (sizes + PACK_PAD)
            blob = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            blob[:start_off] = lut.contiguous().view(torch.uint8).reshape(-1)
            blob[start_off:tbl_off] = starts.to(torch.int32).contiguous().view(torch.uint8).reshape(-1)
            # (value - 1) mod 2^16, PACK_PAD entries 0xffff after every row
            padded = torch.cat((rows, torch.full((n, PACK_PAD), 1 << 17, dtype=rows.dtype, device=dev)), 1)
            keep = torch.arange(width + PACK_PAD, device=dev)[None, :] < (sizes + PACK_PAD)[:, None]
            vals = torch.where(padded > 65536, torch.full_like(padded, 65536), padded)
            flat = ((vals[keep] - 1) & 0xFFFF).to(torch.int32)
            flat = torch.where(flat > 32767, flat - 65536, flat).to(torch.int16)
            blob[tbl_off:tbl_off + 2 * total] = flat.contiguous().view(torch.uint8).reshape(-1)
            out = (blob, total)
        try:
            self.cdf._dvc_cdf_pack = (key, out)
        except AttributeError:
            pass
        return out


LUT_KEYS = 416                  # keys of the decoder's look-up (csrc/dvc_coder.cu::lut_key)
LUT_STRIDE = LUT_KEYS + 2        # + sentinel (size - 2) + pad: u16 entries per row
PACK_PAD = 4                    # entries 0xffff after every packed row
PACK_MAX_BYTES = 124 * 1024


def lut_key(cum):
    """Key of a 16-bit ``cum`` in the decoder's look-up, monotone in ``cum``: less than 2048
    counts from either end, eight keys per octave of the distance ``d`` (the exponent and the
    three leading mantissa bits of ``float(d | 1)``): 0..87 from the lower end, 415..328 from the
    upper one; ``80 + (cum >> 8)`` (88..327) in between."""
    up = cum >> 15
    d = 65535 - cum if up else cum
    if d >= 2048:
        return 80 + (cum >> 8)
    v = d | 1
    e = v.bit_length() - 1
    t = 8 * e + (((v << 3) >> e) & 7)
    return LUT_KEYS - 1 - t if up else t


@functools.lru_cache(maxsize=1)
def _lut_ranges():
    """``(cmin, cmax)`` int32 ``[LUT_KEYS]``: the ``cum`` values of every key (a contiguous
    range each).  A key no ``cum`` maps to (the logarithmic keys have holes at small distances)
    takes the first ``cum`` of the next key that is used, so that entry ``k + 1`` of a row's
    look-up always bounds the bracket of entry ``k`` from above."""
    keys = np.fromiter((lut_key(c) for c in range(65536)), dtype=np.int64, count=65536)
    assert (np.diff(keys) >= 0).all() and keys[0] == 0 and keys[-1] == LUT_KEYS - 1
    cmin = np.full(LUT_KEYS, -1, dtype=np.int32)
    cmax = np.full(LUT_KEYS, -1, dtype=np.int32)
    edges = np.flatnonzero(np.diff(keys)) + 1
    for a, b in zip(np.r_[0, edges], np.r_[edges, 65536]):
        cmin[keys[a]], cmax[keys[a]] = a, b - 1
    for k in range(LUT_KEYS - 2, -1, -1):
        if cmin[k] < 0:
            cmin[k], cmax[k] = cmin[k + 1], cmin[k + 1] - 1          # an empty range
    return cmin, cmax


def pmf_to_quantized_cdf(pmf, precision=16):
    """CompressAI ``_CXX.pmf_to_quantized_cdf`` (host, setup time)."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32).reshape(-1))
    cdf = np.zeros(p.size + 1, dtype=np.int32)
    rc = nat.lib().dvc_pmf_to_quantized_cdf(p.ctypes.data, p.size, int(precision), cdf.ctypes.data)
    if rc != 0:
        msg = nat.lib().dvc_last_error_string().decode()
        raise ValueError(msg)
    return cdf


def _table_f32(scale_table, device):
    if scale_table is None or scale_table.numel() == 0:
        raise ValueError("Uninitialized scale table. Run update_scale_table() first")
    t = scale_table.detach().to(device=device, dtype=torch.float32).contiguous()
    if t.numel() > 256:
        raise nat.DvcError("scale tables of more than 256 entries are not supported")
    return t


def build_indexes(scales, scale_table, scale_bound):
    """``GaussianConditional.build_indexes``: int32 tensor shaped like ``scales``.
    One launch instead of the reference's 63 compare-and-subtract passes."""
    squeeze = None
    if scales.dim() != 4:
        squeeze = scales.shape
        scales = scales.reshape(1, 1, 1, -1)
    scales = nat.require_cuda_f32(scales, "build_indexes(scales)")
    tab = _table_f32(scale_table, scales.device)
    n, c, h, w = scales.shape
    out = torch.empty((n, c, h, w), dtype=torch.int32, device=scales.device)
    with nat.device_of(scales):
        rc = nat.lib().dvc_symbols_indexes_fwd(
            None, None, scales.data_ptr(), tab.data_ptr(), tab.numel(), None, out.data_ptr(),
            n, c, h, w, None, None, nat.st4(scales), float(scale_bound), nat.stream_of(scales))
    nat.check(rc, "dvc_symbols_indexes_fwd")
    return out if squeeze is None else out.reshape(squeeze)


def quantize_symbols(x, means=None):
    """``EntropyModel.quantize(x, "symbols", means)``: ``int32(round(x - means))``."""
    x = nat.require_cuda_f32(x, "quantize(inputs)")
    if means is not None:
        means = means.expand_as(x)
    n, c, h, w = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.int32, device=x.device)
    with nat.device_of(x):
        rc = nat.lib().dvc_symbols_indexes_fwd(
            x.data_ptr(), nat.ptr(means), None, None, 0, out.data_ptr(), None, n, c, h, w,
            nat.st4(x), nat.opt_st4(means), None, 0.0, nat.stream_of(x))
    nat.check(rc, "dvc_symbols_indexes_fwd")
    return out


def _index_args(indexes, scales, scale_table, shape, device):
    """(indexes_ptr, scales_ptr, table_ptr, T, scales_st, keepalive)."""
    keep = []
    if indexes is not None:
        if tuple(indexes.shape) != tuple(shape):
            raise ValueError("`inputs` and `indexes` should have the same size.")
        idx = indexes
        if not idx.is_cuda:
            raise nat.DvcError("indexes must be a CUDA tensor; there is no CPU coder")
        idx = idx.to(torch.int32).contiguous()
        keep.append(idx)
        return idx.data_ptr(), None, None, 0, None, keep
    if scales is not None:
        scales = nat.require_cuda_f32(scales, "scales")
        if tuple(scales.shape) != tuple(shape):
            raise ValueError("`scales` must have the shape of the coded tensor")
        tab = _table_f32(scale_table, device)
        keep += [scales, tab]
        return None, scales.data_ptr(), tab.data_ptr(), tab.numel(), nat.st4(scales), keep
    return None, None, None, 0, None, keep        # channel index (entropy bottleneck)


def _status_word(device):
    """One status word per launch: a failed or abandoned encode cannot leave a
    flag behind that an unrelated later decode would trip over."""
    return torch.zeros(1, dtype=torch.int32, device=device)


class PendingStreams:
    """Bit streams of one ``rans_encode_async`` call, still on the device."""

    def __init__(self, out, out_bytes, status, cap, keep, done=None):
        self.out, self.out_bytes, self.status, self.cap, self._keep = out, out_bytes, status, cap, keep
        self.done = done       # CUDA event when the encoder ran on the side stream

    def result(self):
        return collect([self])[0]


def rans_encode_async(tables, x=None, means=None, symbols=None, indexes=None, scales=None,
                      scale_table=None, scale_bound=0.11, stream_symbols=None, overlap=False,
                      est_bytes=None, lanes=None, skip=None):
    """Launch the encoder for one tensor ``[N,C,H,W]``; no host synchronisation.

    Symbols: ``symbols`` (int32) or ``round(x - means)``.  Table indexes:
    ``indexes`` (int), or derived from ``scales`` like ``build_indexes``, or --
    neither -- the channel number.  ``collect`` turns pending results into
    ``bytes`` with a single device->host round trip for any number of them.

    ``est_bytes``: estimated coded size of ONE sample (``-sum log2 p / 8``), used by
    ``auto_stream_symbols`` to bound the container overhead.

    ``lanes`` (default ``DVC_RANS_LANES`` = 32): 32 = every sub-stream is coded by
    32 lane-interleaved rans64 coders (``DVC3``), 1 = one stock stream per
    sub-stream (``DVC1``).  ``skip`` (lanes = 32 only): symbols of (almost)
    deterministic table rows are coded as one flag per group of 32 (``DVS3``),
    which shortens the serial chain by up to 32x on them -- most of a low-rate
    P-frame.  ``True`` forces it, ``False`` turns it off, the default (``None``,
    with ``DVC_RANS_SKIP`` on) lets the encoder decide on the device: it counts
    the flags that would be set and falls back to ``DVC3`` when they would cost
    more than 0.5 % of the payload.  Both are ignored for a raw stock stream
    (``stream_symbols == 0``); all layouts are lossless.

    ``overlap=True`` runs the encoder on a per-device side stream (ordered after
    everything already queued on the current stream): the coder keeps a handful
    of warps busy for a fraction of a millisecond, and whatever the caller
    launches next (the hyper-decoder convolutions) proceeds beside it instead of
    behind it.  ``collect`` joins the streams."""
    src = x if x is not None else symbols
    if src is None or (x is not None and symbols is not None):
        raise ValueError("give exactly one of x / symbols")
    if src.dim() != 4:
        raise ValueError("expected a 4-D [N,C,H,W] tensor")
    n, c, h, w = src.shape
    dev = src.device
    if x is not None:
        x = nat.require_cuda_f32(x, "compress(inputs)")
        if means is not None:
            means = nat.require_cuda_f32(means, "compress(means)").expand_as(x)
    else:
        if not symbols.is_cuda:
            raise nat.DvcError("symbols must be a CUDA tensor; there is no CPU coder")
        symbols = symbols.to(torch.int32).contiguous()
    ip, sp, tp, T, sst, keep = _index_args(indexes, scales, scale_table, src.shape, dev)
    L = c * h * w
    lanes = DEFAULT_LANES if lanes is None else int(lanes)
    if lanes not in (1, 32):
        raise ValueError("lanes must be 1 or 32")
    if stream_symbols is None:
        stream_symbols = auto_stream_symbols(L, est_bytes, lanes)   # est_bytes: per sample
    if int(stream_symbols) <= 0:
        lanes = 1                                   # raw stock stream
    if lanes == 32:                                 # sub-streams are whole chunks of 1 024 positions
        stream_symbols = -(-int(stream_symbols) // LANES_CHUNK) * LANES_CHUNK
    # skip: True = always use the marks, None (default) = let the encoder count the groups a
    # flag would be set in and use the marks only when those flags stay within SKIP_FLAG_BUDGET
    # of the estimated payload (data the tables do not describe would otherwise grow), False = off
    adaptive = skip is None
    skip = (DEFAULT_SKIP if skip is None else bool(skip)) and lanes == 32
    marks = tables.skip_rows() if skip else None
    max_flagged = -1
    if skip and adaptive:
        if est_bytes is not None:
            max_flagged = int(SKIP_FLAG_BUDGET * 8.0 * float(est_bytes) * n / SKIP_FLAG_BITS)
        else:
            max_flagged = n * L // 8192
        max_flagged = max(max_flagged, n * L // 65536)      # a handful is never worth a fallback
    if marks is not None:
        keep = keep + [marks]
    lib = nat.lib()
    cap = lib.dvc_rans_max_bytes(L, int(stream_symbols), lanes)
    scratch_bytes = lib.dvc_rans_scratch_bytes(n, L, int(stream_symbols), lanes)
    if cap < 0 or scratch_bytes < 0:
        raise nat.DvcError(f"rans_encode: unsupported partition (L={L}, S={stream_symbols})")
    out = torch.empty((n, cap), dtype=torch.uint8, device=dev)
    out_bytes = torch.empty(n, dtype=torch.int64, device=dev)
    scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=dev)
    status = _status_word(dev)
    done = None
    with nat.device_of(src):
        stream = nat.stream_of(src)
        if overlap:
            side = _side_stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            stream = side.cuda_stream
        rc = lib.dvc_rans_encode(
            nat.ptr(x), nat.ptr(means), nat.ptr(symbols), ip, sp, tp, T, float(scale_bound),
            tables.cdf.data_ptr(), tables.size.data_ptr(), tables.offset.data_ptr(),
            tables.cdf.size(0), tables.cdf.size(1), out.data_ptr(), cap, out_bytes.data_ptr(),
            scratch.data_ptr(), status.data_ptr(), n, c, h, w, nat.opt_st4(x),
            nat.opt_st4(means), sst, int(stream_symbols), lanes, nat.ptr(marks),
            int(max_flagged), stream)
        if overlap:
            done = torch.cuda.Event()
            done.record(side)
            # these blocks belong to the current stream's allocator pool: tell it the side
            # stream uses them, so dropping the pending object before collect() (an exception
            # between the head and the tail of a compress call) cannot hand them out early
            for t in [out, out_bytes, scratch, status, x, means, symbols] + list(keep):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(side)
    nat.check(rc, "dvc_rans_encode")
    return PendingStreams(out, out_bytes, status, cap, keep + [x, means, symbols, scratch], done)


def collect(pendings):
    """``[PendingStreams] -> [[bytes per sample]]``: one D2H read of all sizes,
    one D2H copy of all payload bytes."""
    if not pendings:
        return []
    for p in pendings:
        if p.done is not None:
            torch.cuda.current_stream(p.out.device).wait_event(p.done)
    sizes = torch.cat([p.out_bytes for p in pendings] +
                      [p.status.to(torch.int64) for p in pendings]).cpu().tolist()
    if any(sizes[-len(pendings):]):
        raise ValueError("compress: an index lies outside the CDF tables")
    pieces, k = [], 0
    for p in pendings:
        for i in range(p.out.size(0)):
            if sizes[k] < 0:
                raise nat.DvcError(f"rans_encode: output capacity too small "
                                   f"({-sizes[k]} > {p.cap})")
            pieces.append(p.out[i, :sizes[k]])
            k += 1
    blob = (torch.cat(pieces) if len(pieces) > 1 else pieces[0]).cpu().numpy().tobytes()
    strings, k, pos = [], 0, 0
    for p in pendings:
        row = []
        for _ in range(p.out.size(0)):
            row.append(blob[pos:pos + sizes[k]])
            pos += sizes[k]
            k += 1
        strings.append(row)
    return strings


def rans_encode(tables, **kwargs):
    """Encode one tensor ``[N,C,H,W]`` -> ``list`` of ``N`` ``bytes``
    (``rans_encode_async`` + ``collect``)."""
    return rans_encode_async(tables, **kwargs).result()


def stream_symbols_of(string, n_symbols):
    """Sub-stream length a bit stream was written with (0 = raw stock stream)."""
    return container_of(string, n_symbols)[0]


def container_of(string, n_symbols):
    """``(stream_symbols, lanes, skip)`` of a bit stream: ``(0, 1, False)`` for a raw
    stock stream, else the sub-stream length and the layout its magic names."""
    if len(string) >= 16 and len(string) % 4 == 0:
        magic, L, S, ns = struct.unpack_from("<4I", string, 0)
        if magic in (MAGIC, MAGIC_LANES, MAGIC_LANES_SKIP) and L == n_symbols and S > 0 and \
                ns == (L + S - 1) // S and len(string) >= 4 * (4 + ns):
            return S, (1 if magic == MAGIC else 32), magic == MAGIC_LANES_SKIP
    return 0, 1, False


def check_decode_status(statuses):
    """One device->host read for the status words of any number of decoder
    launches (``rans_decode(..., statuses=[...])``); raises like ``rans_decode``."""
    if not statuses:
        return
    st = 0
    for v in torch.cat(statuses).cpu().tolist():
        st |= int(v)
    if st != 0:
        # bit 0: an index outside the tables (wins: the stream then cannot decode either),
        # bit 1: malformed container
        raise ValueError("decompress: " + ("an index lies outside the CDF tables" if st & 1
                                           else "malformed bit-stream container"))


def rans_decode(strings, tables, shape, indexes=None, scales=None, scale_table=None,
                scale_bound=0.11, means=None, device=None, want_symbols=False, cb=None,
                statuses=None):
    """Decode ``len(strings)`` bit streams into a ``[N,C,H,W]`` tensor
    ``float(symbol) + means`` (``EntropyModel.dequantize``), or int32 symbols.
    ``cb = (parity, alt_elements)``: checkerboard pairing of ``scales`` (see
    ``dvc_rans_decode`` in include/dvc_b200.h).  ``statuses``: a list that
    receives this launch's status word instead of a host synchronisation here;
    the caller hands the list to ``check_decode_status`` once everything that
    depends on the decoded values is queued (the context models do: one read per
    ``decompress`` instead of three stalls)."""
    if not isinstance(strings, (tuple, list)):
        raise ValueError("Invalid `strings` parameter type.")
    n, c, h, w = (int(v) for v in shape)
    if len(strings) != n:
        raise ValueError("Invalid strings or indexes parameters")
    dev = tables.cdf.device if device is None else device
    L = c * h * w
    S = {container_of(s, L) for s in strings}
    if len(S) != 1:
        raise ValueError("bit streams of one batch were written with different layouts")
    S, lanes, skip = S.pop()
    marks = tables.skip_rows() if skip else None
    pack, pack_entries = tables.cdf_pack() if lanes == 32 else (None, 0)
    for s in strings:
        if len(s) < 8 or len(s) % 4:
            raise ValueError("truncated bit stream")
    # ONE pinned staging buffer per call -- the n sizes (int64), then the n streams at a common
    # 16-byte-aligned stride -- and ONE asynchronous copy: no host synchronisation here, so
    # the host side of the next launch (and whatever the caller queues) overlaps this one's
    # serial chain.
    stride = (max(len(s) for s in strings) + 15) // 16 * 16
    head = (8 * n + 15) // 16 * 16
    slot, stage = _staging(dev).take(head + n * stride)
    host = stage.numpy()
    host[:8 * n].view(np.int64)[:] = [len(s) for s in strings]
    for i, s in enumerate(strings):
        at = head + i * stride
        host[at:at + len(s)] = np.frombuffer(s, dtype=np.uint8)
        host[at + len(s):at + stride] = 0
    dbuf = torch.empty(head + n * stride, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        dbuf.copy_(stage[:head + n * stride], non_blocking=True)
        _staging(dev).release(slot, torch.cuda.current_stream(dev))
    in_bytes = dbuf[:8 * n].view(torch.int64)
    buf = dbuf[head:].view(n, stride)
    ip, sp, tp, T, sst, keep = _index_args(indexes, scales, scale_table, (n, c, h, w), dev)
    if means is not None:
        means = means.to(device=dev, dtype=torch.float32).expand(n, c, h, w)
    out_f = out_s = None
    if want_symbols:
        out_s = torch.empty((n, c, h, w), dtype=torch.int32, device=dev)
    else:
        out_f = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
    status = _status_word(dev)
    scratch = None
    if lanes == 32:
        nbytes = nat.lib().dvc_rans_decode_scratch_bytes(n, L, int(S), lanes)
        if nbytes < 0:
            raise nat.DvcError(f"rans_decode: unsupported partition (L={L}, S={S})")
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with nat.device_of(buf):
        rc = nat.lib().dvc_rans_decode(
            buf.data_ptr(), stride, in_bytes.data_ptr(), ip, sp, tp, T, float(scale_bound),
            tables.cdf.data_ptr(), tables.size.data_ptr(), tables.offset.data_ptr(),
            tables.cdf.size(0), tables.cdf.size(1), nat.ptr(means), nat.ptr(out_f),
            nat.ptr(out_s), status.data_ptr(), n, c, h, w, sst, nat.opt_st4(means),
            nat.opt_st4(out_f), int(S), -1 if cb is None else int(cb[0]),
            0 if cb is None else int(cb[1]), lanes, nat.ptr(marks), nat.ptr(pack),
            int(pack_entries), nat.ptr(scratch), nat.stream_of(buf))
    nat.check(rc, "dvc_rans_decode")
    if statuses is not None:
        statuses.append(status)
    else:
        check_decode_status([status])
    return out_s if want_symbols else out_f


def decode_stage_a(q0, means, scales):
    """``params = cat((q0 + means_0) * mask_0, (q0 + means_1) * mask_1, means,
    scales)`` (video_model.py:274-277 == :449-452): the spatial-prior conv input
    on the decoder side.  ``q0``: int32 symbols ``[N, C/2, H, W]``."""
    means = nat.require_cuda_f32(means, "decode_stage_a(means)")
    scales = nat.require_cuda_f32(scales, "decode_stage_a(scales)")
    n, c, h, w = means.shape
    if q0.dtype != torch.int32 or tuple(q0.shape) != (n, c // 2, h, w) or not q0.is_contiguous():
        raise nat.DvcError("decode_stage_a: q0 must be contiguous int32 [N, C/2, H, W]")
    # same memory-format rule as the encoder's stage A (context._stage_a_fwd), keyed on the
    # prior means both sides compute identically: y_spatial_prior must see the same layout in
    # the encoder and the decoder, or cuDNN may pick different algorithms and a 1-ulp
    # difference in its output can move a scale across a table threshold (decoder desync)
    cl = means.is_contiguous(memory_format=torch.channels_last) and not means.is_contiguous()
    params = torch.empty((n, 3 * c, h, w), dtype=torch.float32, device=means.device,
                         memory_format=torch.channels_last if cl else torch.contiguous_format)
    with nat.device_of(means):
        rc = nat.lib().dvc_dual_prior_decode_stage_a(
            q0.data_ptr(), means.data_ptr(), scales.data_ptr(), params.data_ptr(), n, c, h, w,
            nat.st4(means), nat.st4(scales), nat.st4(params), nat.stream_of(means))
    nat.check(rc, "dvc_dual_prior_decode_stage_a")
    return params


def decode_stage_b(q0, q1, means, prior):
    """``y_hat`` of the decoder (video_model.py:284-289 == :459-464)."""
    means = nat.require_cuda_f32(means, "decode_stage_b(means)")
    prior = nat.require_cuda_f32(prior, "decode_stage_b(prior)")
    n, c, h, w = means.shape
    for q in (q0, q1):
        if q.dtype != torch.int32 or tuple(q.shape) != (n, c // 2, h, w) or not q.is_contiguous():
            raise nat.DvcError("decode_stage_b: q0/q1 must be contiguous int32 [N, C/2, H, W]")
    if tuple(prior.shape) != (n, 2 * c, h, w):
        raise nat.DvcError("decode_stage_b: prior must be [N, 2C, H, W]")
    y_hat = torch.empty((n, c, h, w), dtype=torch.float32, device=means.device)
    with nat.device_of(means):
        rc = nat.lib().dvc_dual_prior_decode_stage_b(
            q0.data_ptr(), q1.data_ptr(), means.data_ptr(), prior.data_ptr(), y_hat.data_ptr(),
            n, c, h, w, nat.st4(means), nat.st4(prior), nat.st4(y_hat), nat.stream_of(means))
    nat.check(rc, "dvc_dual_prior_decode_stage_b")
    return y_hat
