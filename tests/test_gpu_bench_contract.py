"""GPU: the JSON line `bench.py` prints obeys the measurement contract (keys,
units, internal consistency) and the GOP mode sums what it should."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout[-500:]          # exactly ONE JSON line
    return json.loads(lines[0])


def test_bench_line_contract(cuda_dev):
    d = _run("--steps", "8", "--warmup", "3", "--e2e-steps", "8")
    assert d["metric"].startswith("1080p P-frames/sec") and d["unit"] == "P-frames/s"
    assert d["n_gpus"] == 1 and d["steps"] == 8 and d["warmup"] == 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "configs[1]" in d["config"]["workload"] and "model" not in d["config"]
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) <= 1e-6 * d["value"]
    assert d["gpu_launches"] == d["launches_per_step"] * d["steps"] and d["launches_per_step"] == 8
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] == "warp_multi_kernel"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-9
    assert r["algorithmic_bytes_per_launch"] == 1492561920
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    assert r["kernel_ms"] <= d["ms_per_step"] * 1.05          # a kernel cannot outlast its step
    assert 0.3 < r["frac"] < 1.1 and r["traffic"] and "static" in r["traffic_source"]
    assert r["whole_step"]["algorithmic_bytes"] == 1571681280 - 26112000
    assert set(r["regimes"]) == {"smooth", "adversarial"} and set(r["layouts"]) == {"channels_last", "nchw"}
    assert r["layouts"]["nchw"]["value"] < r["layouts"]["channels_last"]["value"]
    assert d["with_spynet"]["spynet_warp_bytes"] == 88780800
    assert d["gpu_eager_baseline"]["value"] < d["value"] and d["gpu_eager_baseline"]["kind"] == "port"
    e = d["e2e"]
    assert e["unit"] == "P-frames/s" and e["h2d_bytes_per_step"] > 40e6 and e["d2h_bytes_per_step"] == 24
    assert e["value"] < d["value"]                            # host copies are inside the e2e region
    assert 0.5 < e["frac_of_copy_ceiling"] < 1.2
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == "P-frames/s" and c["sample"]
    assert c["bits_rel_err_gpu_vs_cpu"] <= 1e-4                # same bits on both arms
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


def test_gop_mode_line(cuda_dev):
    d = _run("--gop", "--sequences", "1")
    assert d["mode"] == "gop" and d["n_gpus"] == 1
    assert d["config"]["units"] == 3 and d["config"]["p_frames"] == 93 and d["sum_frames"] == 93
    assert d["sum_pixels"] == 93 * 1088 * 1920 and d["sum_bits"] > 0
    assert abs(d["bpp"] - d["sum_bits"] / d["sum_pixels"]) <= 1e-12 * d["bpp"]
