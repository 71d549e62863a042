#!/usr/bin/env python
"""Concurrent pinned host->device (and device->host) copy bandwidth per GPU
at N = 1/2/4/8 ranks -- the ceiling the end-to-end (`e2e`) bench line lives
under (e2e scaling 0.52 / 0.42 at N = 4 / 8).

    python tools/h2d_probe.py                                    # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
        --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py

Every rank pins its own buffers, all ranks start together (barrier) and copy
for the same number of iterations; per-rank GB/s come from CUDA events on the
copy stream, the aggregate is total bytes / max-over-ranks time.  Variants:

  default   pinned buffer allocated wherever the process happens to run
  bound     process bound to the CPUs NVML reports as local to its GPU
            (``nvmlDeviceGetCpuAffinity``) *before* the pinned allocation, so the
            pages are first-touched on the GPU's NUMA node
  streams2  the same bytes as two halves on two copy streams

Rank 0 prints one JSON object (also written to gpurun_out/h2d_probe_N<world>.json).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES_MB = (44, 256)


def topology(local):
    info = {"cpu_count": os.cpu_count()}
    try:
        info["affinity_default"] = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        pass
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        info["numa_nodes"] = len(nodes)
    except Exception:  # noqa: BLE001
        info["numa_nodes"] = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        info["gpu_local_cpus"] = [min(cpus), max(cpus), len(cpus)] if cpus else None
        info["_cpus"] = cpus
        pci = pynvml.nvmlDeviceGetPciInfo(h)
        bus = pci.busId.decode() if isinstance(pci.busId, bytes) else pci.busId
        info["pci_bus"] = bus
        try:
            info["pcie_gen"] = pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h)
            info["pcie_width"] = pynvml.nvmlDeviceGetCurrPcieLinkWidth(h)
        except Exception:  # noqa: BLE001
            pass
        p = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node"
        if os.path.exists(p):
            info["gpu_numa_node"] = int(open(p).read().strip())
    except Exception as e:  # noqa: BLE001
        info["nvml_error"] = repr(e)[:120]
    return info


def copy_rate(dev, host_bufs, dev_bufs, streams, iters, world, direction="h2d"):
    """GB/s of this rank while every rank copies concurrently."""
    def go():
        for s, hb, db in zip(streams, host_bufs, dev_bufs):
            with torch.cuda.stream(s):
                if direction == "h2d":
                    db.copy_(hb, non_blocking=True)
                else:
                    hb.copy_(db, non_blocking=True)
    for _ in range(2):
        go()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    for s, a in zip(streams, e0):
        a.record(s)
    for _ in range(iters):
        go()
    for s, b in zip(streams, e1):
        b.record(s)
    torch.cuda.synchronize()
    ms = max(a.elapsed_time(b) for a, b in zip(e0, e1))
    nbytes = sum(h.numel() * h.element_size() for h in host_bufs) * iters
    return nbytes / (ms * 1e-3) / 1e9, ms


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    topo = topology(local)
    cpus = topo.pop("_cpus", None)
    rows = []
    for variant in ("default", "bound", "streams2"):
        if variant == "bound":
            if not cpus or (topo.get("numa_nodes") or 1) <= 1:
                topo["bound_skipped"] = "one NUMA node visible: nothing to bind to"
                continue
            try:
                os.sched_setaffinity(0, cpus)
            except Exception as e:  # noqa: BLE001
                topo["bind_error"] = repr(e)[:120]
                continue
        for mb in SIZES_MB:
            n = mb * (1 << 20)
            parts = 2 if variant == "streams2" else 1
            host = [torch.empty(n // parts, dtype=torch.uint8).pin_memory() for _ in range(parts)]
            for hb in host:
                hb.fill_(1)                      # first touch on the (possibly bound) CPU
            devb = [torch.empty(n // parts, dtype=torch.uint8, device=dev) for _ in range(parts)]
            streams = [torch.cuda.Stream(dev) for _ in range(parts)]
            iters = max(4, min(64, (4 << 30) // n))
            for direction in ("h2d", "d2h"):
                if direction == "d2h" and (variant != "default" or mb != 256):
                    continue
                gbs, ms = copy_rate(dev, host, devb, streams, iters, world, direction)
                t = torch.tensor([gbs, ms], dtype=torch.float64, device=dev)
                if world > 1:
                    allv = [torch.zeros_like(t) for _ in range(world)]
                    dist.all_gather(allv, t)
                else:
                    allv = [t]
                per = [float(v[0]) for v in allv]
                worst_ms = max(float(v[1]) for v in allv)
                rows.append({"variant": variant, "direction": direction, "MB": mb, "iters": iters,
                             "per_rank_GBps": [round(v, 2) for v in per],
                             "aggregate_GBps": round(world * n * iters / (worst_ms * 1e-3) / 1e9, 2)})
            del host, devb
    topos = [None] * world
    if world > 1:
        dist.all_gather_object(topos, topo)
        dist.barrier()
        dist.destroy_process_group()
    else:
        topos = [topo]
    if rank == 0:
        out = {"world": world, "topology": topos, "rows": rows}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"h2d_probe_N{world}.json"), "w") as fh:
            json.dump(out, fh, indent=1)
        print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
