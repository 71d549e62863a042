"""GPU: BASELINE.json configs[2] -- GOP-serial units sharded over ranks
(dmc/test.py:162-173 I-frame reset, :190-195 dpb feedback, :275-281 loop over
sequences).  The sharded sum must equal the single-GPU run: frame/pixel counts
exactly, fp64 bit sums to 1e-12 (SURVEY.md 4 "Multi-GPU without a cluster")."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W = 256, 384


def _runner(dev):
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200.gop import GopRunner
    torch.manual_seed(1234)
    ebs = {"motion": dvc.EntropyBottleneck(64).to(dev).eval(),
           "frame": dvc.EntropyBottleneck(64).to(dev).eval()}
    return GopRunner(H, W, dev, ebs, frame_pool=3, max_frames=32)


def test_unit_is_serial_and_deterministic(cuda_dev):
    """Frame t+1 reads what frame t wrote (dpb feedback), a unit's results do not
    depend on what ran before it, and the bits of every frame and the dpb left
    behind by the last one match the oracle run on the same chain."""
    from deepvideocodec_b200.dist import Unit
    from oracle import dmc_ref
    from bench import oracle_modules
    r = _runner(cuda_dev)
    u = Unit(3, 32, 40)                                 # 7 P-frames
    n = r.launch_unit(u)
    a = r.bits[:n, 0].cpu().clone()
    last = {k: v.clone() for k, v in r.dpb[n & 1].items()}     # frame n-1 wrote dpb (n-1)^1
    r.launch_unit(Unit(0, 0, 5))                        # something else in between
    n2 = r.launch_unit(u)
    b = r.bits[:n2, 0].cpu().clone()
    assert n == n2 == 7 and torch.equal(a, b)
    for k, v in last.items():
        assert torch.equal(v, r.dpb[n & 1][k]), k
    # oracle on the same chain: warped outputs of frame t are the dpb of frame t+1
    o_ebs, o_gc = oracle_modules(cuda_dev)
    r._reset_dpb(u)
    dpb = {k: v.clone() for k, v in r.dpb[0].items()}
    with torch.no_grad():
        for t in range(n):
            inp = dict(dpb)
            inp.update(r.frames[t % r.frame_pool])
            ref = dmc_ref.pframe_hot_path(inp, o_ebs, o_gc)
            assert abs(float(ref["bits"][0]) - float(a[t])) <= 1e-4 * abs(float(a[t])), t
            dpb = {"x_ref": ref["warpframe"], "feat1": ref["context1"], "feat2": ref["context2"],
                   "feat3": ref["context3"]}
    # seven warps deep, the chain still agrees with eager (each warp is bit-identical)
    for k, v in dpb.items():
        assert (v - last[k]).abs().max().item() <= 1e-5, k
    r._reset_dpb(u)
    assert not torch.equal(last["x_ref"], r.dpb[0]["x_ref"])      # the dpb really moved


def test_sharded_sum_equals_single_run(cuda_dev):
    """world = 1 vs the union of world = 2 / 4 shards executed one after the other
    on this GPU: identical integer counts, bit sums equal to 1e-12."""
    from deepvideocodec_b200.dist import make_units, shard_units
    r = _runner(cuda_dev)
    units = make_units([40, 33, 12])
    single = r.run_units(units)
    for world in (2, 4):
        parts = [r.run_units(shard_units(units, k, world)) for k in range(world)]
        assert sum(p.frames for p in parts) == single.frames == sum(u.p_frames for u in units)
        assert sum(p.pixels for p in parts) == single.pixels
        tot = sum(p.bits for p in parts)
        assert abs(tot - single.bits) <= 1e-12 * abs(single.bits)


_WORKER = r"""
import json, os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import deepvideocodec_b200 as dvc
from deepvideocodec_b200.dist import make_units, shard_units, reduce_stats
from deepvideocodec_b200.gop import GopRunner
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1234)
ebs = {{"motion": dvc.EntropyBottleneck(64).to(dev).eval(),
        "frame": dvc.EntropyBottleneck(64).to(dev).eval()}}
r = GopRunner({h}, {w}, dev, ebs, frame_pool=3, max_frames=32)
units = make_units([40, 33, 12])
mine = r.run_units(shard_units(units, rank, world))
total = reduce_stats(mine, device=dev)                      # NCCL all-reduce
if rank == 0:
    single = r.run_units(units)
    print(json.dumps({{"total": [total.bits, total.frames, total.pixels],
                       "single": [single.bits, single.frames, single.pixels]}}))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_gpu_nccl_reduce_equals_single_gpu(cuda_dev, tmp_path):
    """Two ranks, two GPUs, NCCL: sum over ranks == rank 0 running everything."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2); the same sum is checked on one "
                    "GPU by test_sharded_sum_equals_single_run and over gloo on CPU")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, h=H, w=W))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["total"][1] == d["single"][1] and d["total"][2] == d["single"][2]
    assert abs(d["total"][0] - d["single"][0]) <= 1e-12 * abs(d["single"][0])
