#!/usr/bin/env python
"""bench.py -- 1080p P-frames/sec of the DMC hot path (warp + quantise +
likelihood + rate) on B200, with the HBM roofline and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = the hot path of ONE P-frame at 1920x1088 (BASELINE.json configs[1],
SURVEY.md 8d config 2) per rank: flow pyramid, the four motion-compensation
warps, and for both context models the entropy bottleneck, the checkerboard
dual prior, the Gaussian conditional and the rate.  Convolutions are outside
the path; their outputs are synthetic inputs.  Ranks process independent
sequences (no data-path collective): weak scaling; ``value`` = frames all ranks
processed / max-over-ranks device time.

Printed keys are documented in the task contract; see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1088, 1920          # 1080p padded to x64 (dmc/test.py:75-88)
N_SETS = 4                 # rotating input sets, each ~0.78 GB >> 126 MB L2
METRIC = "1080p P-frames/sec (warp+quant+likelihood+rate)"
UNIT = "P-frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--regime", default="smooth", choices=["smooth", "adversarial"])
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(regime, n_gpus):
    return {
        "workload": "DMC P-frame coding hot path at 1920x1088 (padded 1080p), batch 1 per GPU "
                    "(BASELINE.json configs[1])",
        "frame": [H, W], "latents": [H // 16, W // 16], "hyper_latents": [H // 64, W // 64],
        "channels": {"feature": 64, "y_motion": 64, "y_frame": 96, "z": 64},
        "flow_regime": regime,
        "layout": "features channels_last (NHWC), frame/flow/latents NCHW",
        "l2_policy": f"inputs larger than L2: {N_SETS} rotating input sets of ~0.78 GB each",
        "parallelism": f"{n_gpus} x independent sequences, no data-path collective",
    }


# ---------------------------------------------------------------------------
# CPU path (oracle port of the reference's PyTorch ops) -- cpu_baseline leg and
# the --impl reference arm.  The only place bench.py executes oracle/.
# ---------------------------------------------------------------------------
def oracle_modules(device="cpu"):
    import importlib.util
    import torch
    name = "oracle_compressai"
    if name + ".entropy_models" not in sys.modules:
        pkg_dir = os.path.join(ROOT, "oracle", "compressai")
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules[name] = pkg
        spec.loader.exec_module(pkg)
    import importlib
    oem = importlib.import_module(name + ".entropy_models")
    torch.manual_seed(1234)
    ebs = {"motion": oem.EntropyBottleneck(64).to(device).eval(),
           "frame": oem.EntropyBottleneck(64).to(device).eval()}
    gc = oem.GaussianConditional(None).to(device).eval()
    return ebs, gc


def time_cpu_path(inputs_cpu, ebs, gc, frames, threads):
    import torch
    from oracle import dmc_ref
    torch.set_num_threads(threads)
    with torch.no_grad():
        out = dmc_ref.pframe_hot_path(inputs_cpu, ebs, gc)       # warm-up
        t0 = time.perf_counter()
        for _ in range(frames):
            out = dmc_ref.pframe_hot_path(inputs_cpu, ebs, gc)
        dt = time.perf_counter() - t0
    return frames / dt, dt / frames, out


def run_reference_arm(args):
    """The reference's own CPU implementation of the path.  The reference is
    Python on PyTorch and cannot travel to the GPU box, so this times the
    oracle port (identical ATen CPU kernels: grid_sampler_2d, upsample_bilinear2d,
    elementwise, erfc, bmm) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from deepvideocodec_b200.pipeline import synthetic_pframe_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ebs, gc = oracle_modules("cpu")
    inp = synthetic_pframe_inputs(H, W, torch.device("cpu"), 1234, regime=args.regime)
    frames = max(1, min(args.steps, 24))            # bounded sample (~1.7 s per frame)
    warm = max(1, min(args.warmup, 2))
    with torch.no_grad():
        from oracle import dmc_ref
        for _ in range(warm):
            dmc_ref.pframe_hot_path(inp, ebs, gc)
    fps, spf, _ = time_cpu_path(inp, ebs, gc, frames, cores)
    sample = (f"{frames} P-frames of the same 1920x1088 workload on host cores, "
              f"torch {torch.__version__} CPU kernels, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": frames, "warmup": warm, "ms_per_step": spf * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.regime, args.gpus),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# clocks during the timed region (NVML)
# ---------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index, period=0.02):
        self.samples, self.reasons, self.ok = [], set(), False
        self.max_mhz = None
        self.period = period
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                try:
                    index = int(ids[index])
                except (ValueError, IndexError):
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def sample(self):
        if not self.ok:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:  # noqa: BLE001
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:  # noqa: BLE001
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(self.period)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self._stop.set()
        self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [],
                    "note": "NVML unavailable" if not self.ok else "no sample"}
        s = sorted(self.samples)
        reasons = sorted(r for r in self.reasons if r != "gpu_idle")
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(s)}


# ---------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import deepvideocodec_b200 as dvc
    from deepvideocodec_b200.pipeline import (PFramePath, pframe_algorithmic_bytes,
                                              synthetic_pframe_inputs)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    dvc.lib()   # fail loudly if the extension is missing
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, Wm = args.steps, max(args.warmup, 3)
    torch.manual_seed(1234)
    ebs = {"motion": dvc.EntropyBottleneck(64).to(dev).eval(),
           "frame": dvc.EntropyBottleneck(64).to(dev).eval()}
    paths = []
    with torch.no_grad():
        for s in range(N_SETS):
            inp = synthetic_pframe_inputs(H, W, dev, 1234 + 100 * rank + s, regime=args.regime)
            paths.append(PFramePath(inp, ebs))
    alg = pframe_algorithmic_bytes(H, W)
    launches_per_step = paths[0].n_launches

    # ---- device-resident throughput (`value`) --------------------------------
    for i in range(Wm):
        paths[i % N_SETS].launch()
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    wev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(K)]
    sampler = ClockSampler(local)
    if sampler.ok:
        sampler.start()
    ev0.record()
    for i in range(K):
        paths[i % N_SETS].launch(warp_events=wev[i])
    ev1.record()
    sampler.sample()
    barrier()
    if sampler.ok:
        sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    warp_ms = sum(a.elapsed_time(b) for a, b in wev) / K
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * K / (ms_total * 1e-3)
    bits = [float(p.out["bits"][0].item()) for p in paths]

    # ---- end to end: host buffers in, rate scalars out (`e2e`) ----------------
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(torch, dist, dvc, paths, dev, world, args, K)

    # ---- roofline of the dominant kernel --------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    achieved = alg["warp_multi"] / (warp_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "warp_multi_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg["warp_multi"],
                "kernel_ms": warp_ms,
                "whole_step": {"algorithmic_bytes": alg["total"],
                               "achieved": alg["total"] * K * 1e-9 / (ms_total * 1e-3),
                               "frac": alg["total"] * K * 1e-9 / (ms_total * 1e-3) / peak,
                               "frac_of_nominal_8TBps": alg["total"] * K * 1e-9 / (ms_total * 1e-3) / 8000.0}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = measure_cpu_baseline(torch, paths[0], bits[0])

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.regime, world),
        "clocks": sampler.summary(), "gpu_launches": launches_per_step * K,
        "launches_per_step": launches_per_step, "roofline": roofline,
        "bits_per_frame_set0": bits[0],
    }
    if e2e is not None:
        line["e2e"] = e2e
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line), flush=True)
    return 0


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu
    capture (profiles/*traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "warp_multi_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:  # noqa: BLE001
            return None
    return None


def measure_e2e(torch, dist, dvc, paths, dev, world, args, K):
    """Same metric through the public path object with HOST buffers: every step
    copies that step's inputs from pinned host memory to the device, runs the
    P-frame path and reads the rate scalars back.  Two device input sets are
    double-buffered so the copy of frame i+1 overlaps the kernels of frame i."""
    steps = args.e2e_steps or max(4, min(K, 40))
    host = {k: v.cpu().pin_memory() for k, v in paths[0].inp.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    bits_host = torch.empty(1, dtype=torch.float64).pin_memory()
    bpp_host = torch.empty(4, dtype=torch.float32).pin_memory()
    d2h = bits_host.numel() * 8 + bpp_host.numel() * 4
    copy_stream = torch.cuda.Stream(dev)
    comp = torch.cuda.current_stream(dev)
    slots = paths[:2]
    ready = [torch.cuda.Event() for _ in slots]
    freed = [torch.cuda.Event() for _ in slots]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[slot])
            for k, v in host.items():
                slots[slot].inp[k].copy_(v, non_blocking=True)
            ready[slot].record(copy_stream)

    def step(i):
        slot = i % 2
        comp.wait_event(ready[slot])
        out = slots[slot].launch()
        bits_host.copy_(out["bits"], non_blocking=True)
        bpp_host.copy_(out["bpp"].view(-1), non_blocking=True)
        freed[slot].record(comp)

    for f in freed:
        f.record(comp)
    torch.cuda.synchronize()
    upload(0)
    for i in range(3):                       # warm-up
        upload((i + 1) % 2)
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(comp)                          # the first frame's upload is inside the region
    upload(1)                                # warm-up ended on slot 0 -> frame 0 uses slot 1
    for i in range(steps):
        if i + 1 < steps:
            upload(i % 2)
        step(i + 1)
    e1.record(comp)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"value": world * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": steps,
            "note": "pinned host inputs -> H2D every step (double-buffered, copy overlaps "
                    "kernels) -> PFramePath.launch -> D2H of bits/bpp; PCIe-bound"}


def measure_cpu_baseline(torch, path0, gpu_bits):
    """Oracle port of the reference's PyTorch CPU path on the same input set,
    bounded to a few frames, all host threads and one thread."""
    cores = os.cpu_count() or 1
    ebs, gc = oracle_modules("cpu")
    inp = {k: v.cpu() for k, v in path0.inp.items()}
    frames = 3
    fps, spf, out = time_cpu_path(inp, ebs, gc, frames, cores)
    fps1, _, _ = time_cpu_path(inp, ebs, gc, 1, 1)
    cpu_bits = float(out["bits"][0])
    return {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{frames} P-frames (after 1 warm-up) of input set 0 of this run, oracle port of "
                      f"the reference's PyTorch CPU ops, {cores} threads",
            "value_1_thread": fps1,
            "bits_rel_err_gpu_vs_cpu": abs(gpu_bits - cpu_bits) / abs(cpu_bits)}


if __name__ == "__main__":
    sys.exit(main())
