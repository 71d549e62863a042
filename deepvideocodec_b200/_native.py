"""ctypes binding of ``libdvc_b200.so`` (the C ABI in ``include/dvc_b200.h``).

There is NO fallback: if the library is missing, or a tensor is not a CUDA fp32
tensor, the call raises.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os
import subprocess
import sys
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_void_p

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_PKG_DIR)
LIB_PATH = os.environ.get("DVC_B200_LIB") or os.path.join(_PKG_DIR, "libdvc_b200.so")   # override: A/B kernel builds
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
INCLUDE_DIR = os.path.join(_REPO_DIR, "include")
SOURCES = ("dvc_api.cu", "dvc_warp.cu", "dvc_warp_bwd.cu", "dvc_entropy.cu", "dvc_entropy_bwd.cu",
           "dvc_coder.cu", "dvc_warp_conv.cu")

DVC_WARP_IEEE_DIV = 1
DVC_RATE_MAX_BLOCKS = 1024

_I64x4 = c_int64 * 4
_P4 = POINTER(c_int64)


class DvcError(RuntimeError):
    pass


_NVCC_FLAGS = ("-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared")


def _dep_files():
    srcs = [os.path.join(CSRC_DIR, s) for s in SOURCES if os.path.exists(os.path.join(CSRC_DIR, s))]
    hdrs = sorted(os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith(".cuh"))
    return srcs, hdrs + [os.path.join(INCLUDE_DIR, "dvc_b200.h")]


def source_hash(extra=()):
    """sha256[:16] of every CUDA source, header and the compiler flags: what
    ``dvc_build_info()`` of a library built from this tree must report."""
    import hashlib
    h = hashlib.sha256()
    srcs, hdrs = _dep_files()
    for path in srcs + hdrs:
        h.update(os.path.basename(path).encode() + b"\0")
        h.update(open(path, "rb").read())
    h.update(" ".join(_NVCC_FLAGS + tuple(e for e in extra if e not in ("-Xptxas", "-v"))).encode())
    return h.hexdigest()[:16]


def nvcc_command(out_path=LIB_PATH, extra=()):
    srcs, _ = _dep_files()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return [nvcc, *_NVCC_FLAGS, f"-I{INCLUDE_DIR}", f"-I{CSRC_DIR}",
            f'-DDVC_SRC_HASH="{source_hash(extra)}"', *extra, "-o", out_path, *srcs]


def built_hash(path=None):
    """The source hash recorded inside a built library (None if absent / unreadable)."""
    import re
    path = path or LIB_PATH
    if not os.path.exists(path):
        return None
    # read the string out of the file instead of dlopen()ing it: a library that is about to
    # be rebuilt must not already be mapped into this process
    m = re.search(rb"src=([0-9a-f]{16}) arch=sm_100a", open(path, "rb").read())
    return m.group(1).decode() if m else None


def build_library(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into ``libdvc_b200.so`` (in-tree)."""
    # up to date = the hash compiled into the binary equals the hash of the sources on disk
    # (mtimes do not survive a checkout or the copy to the GPU box)
    if not force and built_hash() == source_hash():
        return LIB_PATH
    cmd = nvcc_command(extra=("-Xptxas", "-v") if verbose else ())
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise DvcError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    global _lib
    _lib = None
    return LIB_PATH


_lib = None

_SIGNATURES = {
    "dvc_version": (c_int, []),
    "dvc_last_error_string": (c_char_p, []),
    "dvc_build_info": (c_char_p, []),
    "dvc_device_info": (c_int, [POINTER(c_int)] * 3),
    "dvc_rate_workspace_bytes": (c_int64, [c_int64]),
    "dvc_flow_warp_fwd": (c_int, [c_void_p] * 3 + [c_int64] * 4 + [_P4] * 3 + [c_int, c_void_p]),
    "dvc_flow_warp_bwd": (c_int, [c_void_p] * 5 + [c_int64] * 4 + [_P4] * 5 + [c_int, c_void_p]),
    "dvc_bilinear_down2_fwd": (c_int, [c_void_p] * 2 + [c_int64] * 4 + [_P4] * 2 + [c_float, c_void_p]),
    "dvc_bilinear_down2_bwd": (c_int, [c_void_p] * 2 + [c_int64] * 4 + [_P4] * 2 + [c_float, c_void_p]),
    "dvc_flow_pyramid_fwd": (c_int, [c_void_p] * 3 + [c_int64] * 3 + [_P4] * 3 + [c_void_p]),
    "dvc_warp_multi_fwd": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "dvc_conv3x3_packed_weight_floats": (c_int64, [c_int64] * 2),
    "dvc_conv3x3_pack_weights": (c_int, [c_void_p, _P4, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "dvc_warp_conv3x3_fwd": (c_int, [c_void_p] * 7 + [c_int64] * 6 + [_P4, c_int64, c_int, c_void_p]),
    "dvc_quantize_fwd": (c_int, [c_void_p] * 3 + [c_int64] * 4 + [_P4, c_int64, _P4, c_void_p]),
    "dvc_dual_prior_stage_a_fwd": (c_int, [c_void_p] * 4 + [c_int64] * 4 + [_P4] * 4 + [c_void_p]),
    "dvc_dual_prior_stage_b_gc_fwd": (
        c_int, [c_void_p] * 15 + [c_int64] * 4 + [_P4] * 7 + [c_float, c_float, c_void_p]),
    "dvc_gc_likelihood_fwd": (
        c_int, [c_void_p] * 8 + [c_int64] * 4 + [_P4] * 5 + [c_float, c_float, c_void_p]),
    "dvc_gc_likelihood_bwd": (
        c_int, [c_void_p] * 10 + [c_int64] * 4 + [_P4] * 7 + [c_float, c_float, c_void_p]),
    "dvc_dual_prior_stage_a_bwd": (c_int, [c_void_p] * 4 + [c_int64] * 4 + [_P4] * 2 + [c_void_p]),
    "dvc_dual_prior_stage_b_gc_bwd": (
        c_int, [c_void_p] * 14 + [c_int64] * 4 + [_P4] * 8 + [c_float, c_float, c_void_p]),
    "dvc_eb_likelihood_bwd": (
        c_int, [c_void_p] * 15 + [c_int64] * 4 + [_P4] * 4 + [c_float, c_void_p]),
    "dvc_eb_likelihood_fwd": (
        c_int, [c_void_p] * 11 + [c_int64] * 4 + [_P4] * 3 + [c_float, c_void_p]),
    "dvc_rate_finalize": (c_int, [c_void_p, c_int, c_int64, c_double] + [c_void_p] * 4),
    "dvc_log_sum_fwd": (c_int, [c_void_p] * 3 + [c_int64] * 4 + [_P4, c_void_p]),
    "dvc_symbols_indexes_fwd": (
        c_int, [c_void_p] * 4 + [c_int64] + [c_void_p] * 2 + [c_int64] * 4 + [_P4] * 3 +
        [c_float, c_void_p]),
    "dvc_pmf_to_quantized_cdf": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
    "dvc_rans_scratch_bytes": (c_int64, [c_int64] * 3 + [c_int]),
    "dvc_rans_max_bytes": (c_int64, [c_int64] * 2 + [c_int]),
    "dvc_rans_decode_scratch_bytes": (c_int64, [c_int64] * 3 + [c_int]),
    "dvc_rans_encode": (
        c_int, [c_void_p] * 6 + [c_int64, c_float] + [c_void_p] * 3 + [c_int64] * 2 +
        [c_void_p, c_int64, c_void_p, c_void_p, c_void_p] + [c_int64] * 4 + [_P4] * 3 +
        [c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "dvc_rans_decode": (
        c_int, [c_void_p, c_int64, c_void_p] + [c_void_p] * 3 + [c_int64, c_float] +
        [c_void_p] * 3 + [c_int64] * 2 + [c_void_p] * 4 + [c_int64] * 4 + [_P4] * 3 +
        [c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "dvc_dual_prior_decode_stage_a": (c_int, [c_void_p] * 4 + [c_int64] * 4 + [_P4] * 3 + [c_void_p]),
    "dvc_dual_prior_decode_stage_b": (c_int, [c_void_p] * 5 + [c_int64] * 4 + [_P4] * 3 + [c_void_p]),
}


class WarpTask(ctypes.Structure):
    """Mirror of ``dvc_warp_task`` (include/dvc_b200.h)."""
    _fields_ = [("im", c_void_p), ("flow", c_void_p), ("out", c_void_p),
                ("N", c_int64), ("C", c_int64), ("H", c_int64), ("W", c_int64),
                ("im_st", _I64x4), ("flow_st", _I64x4), ("out_st", _I64x4),
                ("flow_downscale", c_int64)]


def lib():
    """The loaded library; raises ``DvcError`` when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DvcError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  deepvideocodec_b200 has no CPU or eager fallback.")
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError:
            continue     # reported by exported_symbols()/tests; calling it raises below
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def declared_symbols():
    """Every function name declared in include/dvc_b200.h."""
    import re
    text = open(os.path.join(INCLUDE_DIR, "dvc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dvc_[a-z0-9_]+)\s*\(", text)))


def check(rc, what):
    if rc != 0:
        msg = lib().dvc_last_error_string()
        raise DvcError(f"{what} failed ({rc}): {msg.decode() if msg else 'unknown error'}")


# ---------------------------------------------------------------------------
# tensor plumbing
# ---------------------------------------------------------------------------
def require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise DvcError(f"{name}: deepvideocodec_b200 runs on CUDA tensors only (got {t.device}); "
                       "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise DvcError(f"{name}: fp32 only (got {t.dtype})")
    if t.dim() != 4:
        raise DvcError(f"{name}: expected a 4-D [N,C,H,W] tensor, got shape {tuple(t.shape)}")
    return t


def st4(t):
    return _I64x4(*t.stride())


def ptr(t):
    return None if t is None else t.data_ptr()


def opt_st4(t):
    return None if t is None else _I64x4(*t.stride())


def stream_of(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class device_of:
    """Make ``t.device`` current for the duration of a launch (no-op if it is)."""

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


_workspaces = {}


def rate_workspace(device, n, slot=0):
    """Per-(device, stream, slot) rate workspace (partials + zeroed tickets)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, slot, int(n))
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = lib().dvc_rate_workspace_bytes(int(n))
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws
