"""CPU: the C-ABI library builds, loads and exports every symbol that
include/dvc_b200.h declares; argument validation happens before any launch."""
import ctypes
import os
import re

import pytest


def test_header_symbols_are_exported():
    import deepvideocodec_b200 as dvc
    declared = dvc.declared_symbols()
    assert len(declared) >= 20
    lib = dvc.lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # and the Python binding table covers exactly the header
    from deepvideocodec_b200 import _native
    assert set(_native._SIGNATURES) == set(declared)


def test_header_cites_reference_interfaces():
    import deepvideocodec_b200._native as nat
    text = open(os.path.join(nat.INCLUDE_DIR, "dvc_b200.h")).read()
    for cite in ("layers.py:175-198", "layers.py:201-206", "video_model.py:497-504",
                 "utils.py:149-152", "video_model.py:176-189", "train.py:74-93",
                 "video_model.py:55-61", "video_model.py:502-504"):
        assert cite in text, cite
    assert "torch" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S).lower()


def test_binary_matches_committed_sources():
    """The shipped .so is git-ignored: it must say which sources it was built from."""
    import deepvideocodec_b200 as dvc
    info = dvc.lib().dvc_build_info().decode()
    assert info.startswith("src=") and "sm_100a" in info
    assert dvc.built_hash() == dvc.source_hash(), (info, dvc.source_hash())


def test_version_and_sizes():
    import deepvideocodec_b200 as dvc
    lib = dvc.lib()
    assert lib.dvc_version() == 100
    assert lib.dvc_rate_workspace_bytes(1) == 1024 * 8 + 16
    assert lib.dvc_rate_workspace_bytes(8) == 8 * 1024 * 8 + 32
    assert lib.dvc_last_error_string() is not None


def test_invalid_arguments_fail_without_launching():
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200 import _native as nat
    lib = dvc.lib()
    st = nat._I64x4(1, 1, 1, 1)
    rc = lib.dvc_flow_warp_fwd(None, None, None, 1, 3, 8, 8, st, st, st, 0, None)
    assert rc == -1
    assert b"null pointer" in lib.dvc_last_error_string()
    rc = lib.dvc_dual_prior_stage_a_fwd(1, 1, 1, 1, 1, 3, 8, 8, st, st, st, st, None)
    assert rc == -1 and b"even" in lib.dvc_last_error_string()
    rc = lib.dvc_flow_pyramid_fwd(1, 1, 1, 1, 6, 8, st, st, st, None)
    assert rc == -1 and b"multiples of 4" in lib.dvc_last_error_string()
    rc = lib.dvc_rate_finalize(1, 4, 1, 0.0, None, None, None, None)
    assert rc == -1
    # fused warp + 3x3 conv (row f3): every shape rule is checked before the launch
    a = 0x1000                      # 16-byte aligned, never dereferenced
    call = lambda **k: lib.dvc_warp_conv3x3_fwd(   # noqa: E731
        k.get("feat", a), a, k.get("extra", a), a, None, a, a, 1, k.get("cf", 64), k.get("ce", 64),
        k.get("co", 64), 16, 128, st, k.get("level", 0), 0, None)
    assert call(co=32) == -1 and b"Co must be 64" in lib.dvc_last_error_string()
    assert call(cf=24) == -1 and b"multiple of 16" in lib.dvc_last_error_string()
    assert call(ce=8) == -1 and b"multiple of 16" in lib.dvc_last_error_string()
    assert call(extra=None) == -1 and b"mismatch" in lib.dvc_last_error_string()
    assert call(feat=a + 4) == -1 and b"aligned" in lib.dvc_last_error_string()
    assert call(level=3) == -1 and b"flow_downscale" in lib.dvc_last_error_string()
    assert lib.dvc_conv3x3_packed_weight_floats(64, 128) == 128 * 9 * 64
    assert lib.dvc_conv3x3_packed_weight_floats(32, 128) == 0
    assert lib.dvc_conv3x3_pack_weights(a, st, 64, 64, 40, a, None) == -1


def test_library_is_sm100a_only():
    import subprocess
    from deepvideocodec_b200 import _native as nat
    out = subprocess.run(["cuobjdump", "-lelf", nat.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_no_oracle_import_in_product():
    """The product package must never route through oracle/ (or any CPU path)."""
    import deepvideocodec_b200._native as nat
    pkg = os.path.dirname(nat.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
