"""CPU: host-side logic of the drop-in layer -- module surface, error
behaviour without a GPU, patching of the reference, sharding and the scalar
all-reduce (gloo, world_size 2), bench.py's reference arm."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_entropy_module_surface_matches_compressai_names():
    import deepvideocodec_b200 as dvc
    from test_oracle_golden import _oem
    oem = _oem()
    for ours, theirs in ((dvc.EntropyBottleneck(6), oem.EntropyBottleneck(6)),
                         (dvc.GaussianConditional(None), oem.GaussianConditional(None))):
        a, b = ours.state_dict(), theirs.state_dict()
        assert list(a) == list(b)
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        ours.load_state_dict(theirs.state_dict(), strict=True)
    eb = dvc.EntropyBottleneck(6)
    assert eb._get_medians().shape == (6, 1, 3 - 2)
    assert [n for n, _ in eb.named_parameters() if n.endswith("quantiles")] == ["quantiles"]
    assert torch.isfinite(eb.loss())
    with pytest.raises(NotImplementedError):
        dvc.EntropyBottleneck(4, filters=(3, 3))
    # the entropy-coding surface exists and fails loudly
    # before update() / on CPU tensors instead of falling back
    for name in ("update", "compress", "decompress", "_build_indexes", "quantize", "dequantize"):
        assert callable(getattr(eb, name))
    gc = dvc.GaussianConditional(None)
    for name in ("update_scale_table", "update", "build_indexes", "compress", "decompress"):
        assert callable(getattr(gc, name))
    with pytest.raises(ValueError, match="update"):
        eb.compress(torch.zeros(1, 6, 2, 2))
    assert eb.update() is True and eb._quantized_cdf.dtype == torch.int32
    with pytest.raises(dvc.DvcError):
        eb.compress(torch.zeros(1, 6, 2, 2))


def test_aux_loss_matches_oracle():
    import deepvideocodec_b200 as dvc
    from test_oracle_golden import _oem
    oem = _oem()
    torch.manual_seed(3)
    theirs = oem.EntropyBottleneck(5)
    ours = dvc.EntropyBottleneck(5)
    ours.load_state_dict(theirs.state_dict())
    la, lb = ours.loss(), theirs.loss()
    assert torch.equal(la, lb)
    la.backward()
    lb.backward()
    assert torch.equal(ours.quantiles.grad, theirs.quantiles.grad)
    assert ours._matrix0.grad is None          # parameters are detached in the aux loss


def test_cpu_tensors_raise_no_fallback():
    import deepvideocodec_b200 as dvc
    x = torch.zeros(1, 4, 8, 8)
    f = torch.zeros(1, 2, 8, 8)
    calls = [
        lambda: dvc.flow_warp(x, f),
        lambda: dvc.bilineardownsacling(x),
        lambda: dvc.flow_pyramid(f),
        lambda: dvc.quantize_ste(x),
        lambda: dvc.GaussianConditional(None)(x, x, x),
        lambda: dvc.EntropyBottleneck(4)(x),
        lambda: dvc.dual_prior_stage_a(x, x, x),
        lambda: dvc.log_sum(x),
        lambda: dvc.collect_likelihoods_list([{"motion": {"y": x}}], 64),
    ]
    for call in calls:
        with pytest.raises(dvc.DvcError):
            call()


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from deepvideocodec_b200 import _native as nat
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(nat.DvcError, match="no CPU or eager fallback"):
        nat.lib()


def test_algorithmic_bytes_match_survey():
    from deepvideocodec_b200.pipeline import pframe_algorithmic_bytes
    b = pframe_algorithmic_bytes(1088, 1920)
    assert b["total"] == 1571681280                      # 1 571.68 MB, SURVEY.md 8d
    assert b["warp_ctx1"] == 1086259200 and b["warp_x_ref"] == 66846720
    assert b["flow_pyramid"] == 26112000
    assert pframe_algorithmic_bytes(256, 256)["total"] == 49307648   # config 1: 49.31 MB


def test_units_and_sharding():
    from deepvideocodec_b200.dist import make_units, shard_units
    units = make_units([96, 96, 50, 7])
    assert [u.p_frames for u in units] == [31, 31, 31, 31, 31, 31, 31, 17, 6]
    for world in (1, 2, 4, 8):
        shards = [shard_units(units, r, world) for r in range(world)]
        flat = sorted((u.sequence, u.start) for s in shards for u in s)
        assert flat == sorted((u.sequence, u.start) for u in units)          # a partition
        loads = [sum(u.p_frames for u in s) for s in shards]
        assert max(loads) - min(loads) <= 31
    with pytest.raises(ValueError):
        shard_units(units, 2, 2)


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from deepvideocodec_b200.dist import RateStats, make_units, reduce_stats, shard_units
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = make_units([96, 96, 40])
    mine = shard_units(units, rank, world)
    frames = sum(u.p_frames for u in mine)
    # deterministic fake per-frame bits so the reduced total is checkable
    bits = float(sum(1000.0 * u.sequence + u.start for u in mine for _ in range(u.p_frames)))
    red = reduce_stats(RateStats(bits=bits, sq_err=0.5 * frames, frames=frames,
                                 pixels=frames * 1088.0 * 1920.0))
    with open(os.path.join(out_dir, f"r{rank}.json"), "w") as f:
        json.dump({"frames": red.frames, "bits": red.bits, "bpp": red.bpp, "local": frames}, f)
    dist.destroy_process_group()


def test_scalar_all_reduce_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    from deepvideocodec_b200.dist import make_units
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [json.load(open(tmp_path / f"r{r}.json")) for r in range(2)]
    units = make_units([96, 96, 40])
    total_frames = sum(u.p_frames for u in units)
    total_bits = float(sum(1000.0 * u.sequence + u.start for u in units for _ in range(u.p_frames)))
    for r in res:
        assert r["frames"] == total_frames and r["bits"] == total_bits    # sums identical on both ranks
    assert res[0]["local"] + res[1]["local"] == total_frames


@pytest.mark.skipif(not os.path.isdir("/root/reference/dmc"), reason="reference not present")
def test_patch_rebinds_reference_names():
    code = r'''
import sys, torch
sys.path.insert(0, %r)
import deepvideocodec_b200 as dvc
assert dvc.install_compressai_shim() is True
sys.path.insert(0, "/root/reference/dmc")
import models
vm = sys.modules["models.video_model"]; ly = sys.modules["models.layers"]
torch.manual_seed(0)
net = models.DMC()
assert len(net.state_dict()) == 438
assert type(net.motion_context_model.gaussian_conditional).__module__ == "deepvideocodec_b200.entropy_models"
assert type(net.frame_context_model.entropy_bottleneck).__module__ == "deepvideocodec_b200.entropy_models"
net.load_state_dict(net.state_dict())          # DMC.load_state_dict validates the buffer names
stock = vm.flow_warp
dvc.patch(models)
assert vm.flow_warp is dvc.flow_warp and ly.flow_warp is dvc.flow_warp
assert vm.quantize_ste is dvc.quantize_ste and vm.bilineardownsacling is dvc.bilineardownsacling
assert vm.MotionContextModel.forward is dvc.motion_context_forward
assert vm.FrameContextModel.forward_dual_prior is dvc.forward_dual_prior
try:
    net([torch.rand(1, 3, 64, 64), torch.rand(1, 3, 64, 64)])
    raise SystemExit("expected DvcError on CPU tensors")
except dvc.DvcError:
    pass
dvc.unpatch()
assert vm.flow_warp is stock
# opt-in: every context warp fused into the conv that consumes it (tcgen05, inference)
stock_mc = vm.DMC.motion_compensation
dvc.patch(models, fuse_warp_conv=True)
assert vm.DMC.motion_compensation is dvc.motion_compensation_fused
net.eval()
with torch.no_grad():
    try:
        net([torch.rand(1, 3, 64, 64), torch.rand(1, 3, 64, 64)])
        raise SystemExit("expected DvcError on CPU tensors")
    except dvc.DvcError:
        pass
dvc.unpatch()
assert vm.DMC.motion_compensation is stock_mc
print("ok")
''' % ROOT
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr


def test_bench_reference_arm_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "P-frames/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["frame"] == [1088, 1920]


def test_substream_policy_follows_the_payload():
    """coder.auto_stream_symbols (pure host logic): <= 1 % container overhead, a floor of
    MIN_STREAMS sub-streams, whole 256-symbol chunks, pinned lengths win."""
    from deepvideocodec_b200 import coder
    L = 48 * 68 * 120
    assert coder.auto_stream_symbols(L, None) == coder.DEFAULT_STREAM_SYMBOLS
    for est in (50.0, 2_000.0, 50_000.0, 400_000.0):
        S = coder.auto_stream_symbols(L, est)
        n = (L + S - 1) // S
        assert S % 256 == 0 and S >= coder.MIN_STREAM_SYMBOLS
        assert n >= min(coder.MIN_STREAMS, (L + 255) // 256) - 1
        if n > coder.MIN_STREAMS:                      # above the floor the 1 % target binds
            assert n * coder.STREAM_OVERHEAD_BYTES <= coder.OVERHEAD_TARGET * est + 12
    # more payload -> more sub-streams (faster), never fewer
    sizes = [coder.auto_stream_symbols(L, e) for e in (1e3, 1e4, 1e5, 1e6)]
    assert sizes == sorted(sizes, reverse=True) and sizes[-1] < 4096 < sizes[0]
    saved = coder.PINNED_STREAM_SYMBOLS
    try:
        coder.PINNED_STREAM_SYMBOLS = 0
        assert coder.auto_stream_symbols(L, 1e6) == 0     # interop: one raw stock stream
    finally:
        coder.PINNED_STREAM_SYMBOLS = saved


def test_decoder_lookup_keys_are_monotone_and_bound_their_brackets():
    """Host half of the decoder's packed look-up (coder.lut_key / _lut_ranges, mirrored by
    csrc/dvc_coder.cu::lut_key): every 16-bit cum has a key below LUT_KEYS, keys never decrease
    with cum (so entry k + 1 bounds the bracket of entry k), every key -- used or not -- has a
    first cum, and the formula equals the float-exponent form the device evaluates."""
    import numpy as np
    from deepvideocodec_b200 import coder
    keys = np.array([coder.lut_key(c) for c in range(65536)])
    assert keys.min() == 0 and keys.max() == coder.LUT_KEYS - 1
    assert (np.diff(keys) >= 0).all()
    cmin, cmax = coder._lut_ranges()
    assert (np.diff(cmin) >= 0).all() and cmin[0] == 0 and (cmin >= 0).all()
    for k in np.unique(keys):
        assert (keys[cmin[k]:cmax[k] + 1] == k).all() and (keys == k).sum() == cmax[k] - cmin[k] + 1
    c = np.arange(65536, dtype=np.uint32)
    up = c >> 15
    d = np.where(up == 1, 65535 - c, c)
    t = ((d | 1).astype(np.float32).view(np.uint32) >> 20) - (127 << 3)
    dev = np.where(d >= 2048, 80 + (c >> 8), np.where(up == 1, coder.LUT_KEYS - 1 - t, t))
    assert np.array_equal(dev, keys)
    # widest central key 256 counts, logarithmic keys at most an eighth of their distance
    width = cmax - cmin + 1
    assert width.max() == 256


def test_reference_loads_twice_stock_and_patched():
    """oracle/load_reference.py: the unmodified reference as two independent packages in one
    process (stock over the eager restatement, patched over this package's modules), identical
    state-dict layouts, `patch` rebinding only the patched copy -- the setup of
    tests/test_gpu_dropin.py, checked here without a GPU."""
    from oracle.load_reference import load_stock_and_patched, reference_available
    if not reference_available():
        pytest.skip("reference not present")
    import deepvideocodec_b200 as dvc
    stock_pkg, patched_pkg = load_stock_and_patched()
    try:
        assert "compressai" not in sys.modules or not getattr(sys.modules["compressai"], "__file__", "").startswith(ROOT)
        vm_s = sys.modules["dvc_ref_stock.video_model"]
        vm_p = sys.modules["dvc_ref_patched.video_model"]
        assert vm_p.flow_warp is dvc.flow_warp and vm_s.flow_warp is not dvc.flow_warp
        assert vm_s.flow_warp.__module__ == "dvc_ref_stock.layers"
        torch.manual_seed(0)
        a, b = stock_pkg.DMC(), patched_pkg.DMC()
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb) and len(sa) == 438
        assert sum(p.numel() for p in a.parameters()) == 16884403
        b.load_state_dict(sa)
        assert type(b.motion_context_model.entropy_bottleneck).__module__ == \
            "deepvideocodec_b200.entropy_models"
        assert type(a.motion_context_model.entropy_bottleneck).__module__.startswith("compressai")
        with pytest.raises(dvc.DvcError):                             # no CPU fallback
            b([torch.zeros(1, 3, 64, 64)] * 2)
    finally:
        dvc.unpatch()
