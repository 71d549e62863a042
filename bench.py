#!/usr/bin/env python
"""bench.py -- 1080p P-frames/sec of the DMC hot path (warp + quantise +
likelihood + rate) on B200, with the HBM roofline and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--gop]

One "step" = the hot path of ONE P-frame at 1920x1088 (BASELINE.json configs[1],
SURVEY.md 8d config 2) per rank: flow pyramid, the four motion-compensation
warps, and for both context models the entropy bottleneck, the checkerboard
dual prior, the Gaussian conditional and the rate.  Convolutions are outside
the path; their outputs are synthetic inputs.  Ranks process independent
sequences (no data-path collective): weak scaling; ``value`` = frames all ranks
processed / max-over-ranks device time.

``--gop`` runs BASELINE.json configs[2] instead (96-frame sequences cut into
GOP units, sharded over the ranks, frames serial inside a unit through the
decoded picture buffer, rate statistics summed over NCCL once) and prints its
own JSON line (``"mode": "gop"``).

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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the secondary figures (adversarial flow, NCHW layout, SpyNet "
                         "warps, GPU eager baseline)")
    ap.add_argument("--gop", action="store_true", help="BASELINE.json configs[2]: GOP-serial units")
    ap.add_argument("--sequences", type=int, default=2, help="--gop: 96-frame sequences per rank")
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
# the --impl reference arm -- and the same eager ops on the GPU as the
# "what stock PyTorch does on this device" baseline.
# The only places bench.py executes oracle/.
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


def time_cpu_path(inputs_cpu, ebs, gc, frames, threads, warm=1):
    import torch
    from oracle import dmc_ref
    torch.set_num_threads(threads)
    with torch.no_grad():
        out = None
        for _ in range(warm):
            out = dmc_ref.pframe_hot_path(inputs_cpu, ebs, gc)
        t0 = time.perf_counter()
        for _ in range(frames):
            out = dmc_ref.pframe_hot_path(inputs_cpu, ebs, gc)
        dt = time.perf_counter() - t0
    return frames / dt, dt / frames, out


REF_BUDGET_S = 150.0       # wall-clock target of a whole --impl reference run


def run_reference_arm(args):
    """The reference's own CPU implementation of the path.  The reference is
    Python on PyTorch and cannot travel to the GPU box, so this times the
    oracle port (identical ATen CPU kernels: grid_sampler_2d, upsample_bilinear2d,
    elementwise, erfc, bmm) with all host threads.

    ``--steps K --warmup W`` are honoured exactly.  A step is one P-frame of the
    same workload; when K + W full frames would not finish within a couple of
    minutes, a step becomes a bounded sample of the frame -- the top H/f rows
    (f = 2, 4; every op on the path is per-pixel work, cost is linear in rows) --
    and the value is scaled by the sampled fraction."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from deepvideocodec_b200.pipeline import synthetic_pframe_inputs
    from oracle import dmc_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ebs, gc = oracle_modules("cpu")
    steps, warm = max(1, args.steps), max(0, args.warmup)
    inp = synthetic_pframe_inputs(H, W, torch.device("cpu"), 1234, regime=args.regime)
    with torch.no_grad():
        dmc_ref.pframe_hot_path(inp, ebs, gc)                     # page-in, thread pool
        t0 = time.perf_counter()
        dmc_ref.pframe_hot_path(inp, ebs, gc)
        t_frame = time.perf_counter() - t0
    rows = H
    for div in (2, 4):
        if (steps + warm) * t_frame * rows / H > REF_BUDGET_S:
            rows = H // div // 64 * 64
    if rows != H:
        inp = synthetic_pframe_inputs(rows, W, torch.device("cpu"), 1234, regime=args.regime)
    frac = rows / H
    fps_s, spf_s, _ = time_cpu_path(inp, ebs, gc, steps, cores, warm=warm)
    fps, spf = fps_s * frac, spf_s / frac
    sample = (f"{steps} steps after {warm} warm-ups; each step = "
              f"{'one full' if rows == H else f'the top {rows} of {H} rows ({frac:.2f}) of a'} "
              f"1920x1088 P-frame of the same workload on host cores, torch {torch.__version__} "
              f"CPU kernels, {cores} threads"
              + ("" if rows == H else "; value scaled by the sampled fraction"))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": spf * 1e3, "higher_is_better": True,
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
class Ctx:
    """torch + process-group plumbing shared by the measurement legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        import deepvideocodec_b200 as dvc
        dvc.lib()   # fail loudly if the extension is missing
        self.dvc = dvc
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        torch.manual_seed(1234)
        self.ebs = {"motion": dvc.EntropyBottleneck(64).to(self.dev).eval(),
                    "frame": dvc.EntropyBottleneck(64).to(self.dev).eval()}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


EVENT_EVERY = 4     # the dominant kernel is bracketed by CUDA events on every 4th step


def time_steps(cx, step, K, Wm, warp_events=False):
    """W warm-ups, then K steps between barriers; CUDA events on the launch
    stream; max over ranks.  ``step(i, events_or_None)`` enqueues step i.

    The dominant kernel's duration is measured live inside the timed region, on
    every ``EVENT_EVERY``-th step: an event pair around every launch costs 3-4 % of
    the step (tools/host_overhead.py: 262 -> 273 us at K = 400) because each record
    is a command of its own between the kernels of two streams."""
    torch = cx.torch
    for i in range(Wm):
        step(i, None)
    cx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wev = {i: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for i in range(0, K, EVENT_EVERY)} if warp_events else {}
    ev0.record()
    for i in range(K):
        step(i, wev.get(i))
    ev1.record()
    cx.barrier()
    ms = cx.max_over_ranks(ev0.elapsed_time(ev1))
    warp_ms = sum(a.elapsed_time(b) for a, b in wev.values()) / len(wev) if wev else None
    return ms, warp_ms


def peak_hbm():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)
    cx = Ctx(args)
    if args.gop:
        return run_gop(cx)
    torch = cx.torch
    from deepvideocodec_b200.pipeline import (PFramePath, pframe_algorithmic_bytes,
                                              synthetic_pframe_inputs)
    rank, world, dev = cx.rank, cx.world, cx.dev
    K, Wm = args.steps, max(args.warmup, 3)
    paths = []
    with torch.no_grad():
        for s in range(N_SETS):
            inp = synthetic_pframe_inputs(H, W, dev, 1234 + 100 * rank + s, regime=args.regime)
            paths.append(PFramePath(inp, cx.ebs))
    alg = pframe_algorithmic_bytes(H, W)
    launches_per_step = paths[0].n_launches
    peak, peak_src = peak_hbm()

    # ---- device-resident throughput (`value`) --------------------------------
    sampler = ClockSampler(cx.local)
    if sampler.ok:
        sampler.start()
    ms_total, warp_ms = time_steps(cx, lambda i, ev: paths[i % N_SETS].launch(warp_events=ev),
                                   K, Wm, warp_events=True)
    sampler.sample()
    if sampler.ok:
        sampler.stop()
    value = world * K / (ms_total * 1e-3)
    bits = [float(p.out["bits"][0].item()) for p in paths]

    def frac_of(bytes_, ms_per_launch):
        a = bytes_ / (ms_per_launch * 1e-3) / 1e9
        return {"achieved": a, "frac": a / peak, "kernel_ms": ms_per_launch}

    step_bytes = alg["total"] - alg["flow_pyramid"]     # mv2 / mv3 never touch HBM in this path
    roofline = {"bound": "hbm", "kernel": "warp_multi_kernel",
                **frac_of(alg["warp_multi"], warp_ms), "peak": peak, "unit": "GB/s",
                "traffic": load_traffic(),
                "traffic_source": "static: one `ncu --set full` capture of this kernel, committed "
                                  "as profiles/r02_warp_multi_traffic.json (not re-measured by this "
                                  "run)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg["warp_multi"],
                "kernel_ms_samples": len(range(0, K, EVENT_EVERY)),
                "kernel_ms_note": f"CUDA-event pairs around the kernel on every {EVENT_EVERY}th step "
                                  "of the timed region"}
    whole = {"algorithmic_bytes": step_bytes,
             "algorithmic_bytes_incl_pyramid": alg["total"],
             "note": "the fused path derives mv2/mv3 inside the warp kernel, so the 26.1 MB of "
                     "pyramid traffic SURVEY.md 8d counts are not moved; `algorithmic_bytes` "
                     "excludes them, `*_incl_pyramid` is SURVEY.md's 1 571.68 MB figure",
             "achieved": step_bytes * K * 1e-9 / (ms_total * 1e-3)}
    whole["frac"] = whole["achieved"] / peak
    whole["frac_of_nominal_8TBps"] = whole["achieved"] / 8000.0
    whole["frac_incl_pyramid"] = alg["total"] * K * 1e-9 / (ms_total * 1e-3) / peak
    roofline["whole_step"] = whole
    roofline["regimes"] = {args.regime: {"value": value / world, "ms_per_step": ms_total / K,
                                         **frac_of(alg["warp_multi"], warp_ms)}}
    roofline["layouts"] = {"channels_last": dict(roofline["regimes"][args.regime])}

    extras = {}
    if not args.no_extras:
        extras = measure_extras(cx, paths, alg, peak, K, Wm, roofline)

    # ---- end to end: host buffers in, rate scalars out (`e2e`) ----------------
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(cx, paths, K)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = measure_cpu_baseline(torch, paths[0], bits[0])

    cx.close()
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
    line.update(extras)
    if e2e is not None:
        line["e2e"] = e2e
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line), flush=True)
    return 0


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu
    capture (profiles/*traffic.json), or None."""
    for name in ("r02_warp_multi_traffic.json", "warp_multi_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                return float(json.load(open(p))["dram_bytes_per_launch"])
            except Exception:  # noqa: BLE001
                continue
    return None


def measure_extras(cx, paths, alg, peak, K, Wm, roofline):
    """Secondary figures, timed in the same run on the same box:
    the adversarial flow regime, the reference's own NCHW feature layout, the
    step with SpyNet's four 3-channel warps added, and stock PyTorch-CUDA eager
    (oracle ops) on the same tensors."""
    torch, dev, rank = cx.torch, cx.dev, cx.rank
    from deepvideocodec_b200.pipeline import PFramePath, SpyNetWarps, synthetic_pframe_inputs
    Kx, Wx = max(8, min(K, 100)), 3
    out = {}

    def run(ps):
        ms, wms = time_steps(cx, lambda i, ev: ps[i % len(ps)].launch(warp_events=ev), Kx, Wx,
                             warp_events=True)
        a = alg["warp_multi"] / (wms * 1e-3) / 1e9
        return {"value": Kx / (ms * 1e-3), "ms_per_step": ms / Kx, "achieved": a,
                "frac": a / peak, "kernel_ms": wms, "steps": Kx}

    # -- the other flow regime: same tensors, another motion field, outputs shared -----------
    other = "adversarial" if cx.args.regime == "smooth" else "smooth"
    with torch.no_grad():
        alt = []
        for s, p in enumerate(paths):
            g = torch.Generator(device=dev).manual_seed(777 + 100 * rank + s)
            if other == "adversarial":
                mv = torch.randn(1, 2, H, W, device=dev, generator=g) * 16.0
            else:
                f = torch.randn(1, 2, H, W, device=dev, generator=g)
                f = torch.nn.functional.avg_pool2d(f, 31, stride=1, padding=15,
                                                   count_include_pad=False)
                mv = (f / f.std() * 4.0).contiguous()
            inp = dict(p.inp)
            inp["mv"] = mv
            alt.append(PFramePath(inp, cx.ebs, outputs={k: p.out[k] for k in (
                "warpframe", "context1", "context2", "context3")}))
    roofline["regimes"][other] = run(alt)
    roofline["regimes"][other]["flow"] = ("i.i.d. N(0, 16^2) px per pixel (no tap shared between "
                                          "neighbours)" if other == "adversarial" else
                                          "31x31 box-filtered, sigma = 4 px")
    del alt

    # -- the layout the unmodified reference allocates: NCHW features -------------------------
    with torch.no_grad():
        nchw = [PFramePath(synthetic_pframe_inputs(H, W, dev, 1234 + 100 * rank + s,
                                                   regime=cx.args.regime, layout="nchw"), cx.ebs)
                for s in range(N_SETS)]
    r = run(nchw)
    r["kernel"] = ("warp_planar_kernel<staged> (all three feature scales, one launch) + its "
                   "complement launch <gather> + warp_multi_kernel for x_ref: `kernel_ms` brackets "
                   "the three")
    r["launches_per_step"] = nchw[0].n_launches
    roofline["layouts"]["nchw"] = r
    del nchw
    torch.cuda.empty_cache()

    # -- with SpyNet's four 3-channel warps (layers.py:261; SURVEY.md 8d second figure) --------
    spy = [SpyNetWarps(H, W, dev, 4321 + 100 * rank + s) for s in range(N_SETS)]

    def step_spy(i, ev):
        spy[i % N_SETS].launch()
        paths[i % N_SETS].launch()
    ms, _ = time_steps(cx, step_spy, Kx, Wx)
    tot = alg["total"] - alg["flow_pyramid"] + spy[0].bytes
    out["with_spynet"] = {"value": Kx / (ms * 1e-3), "ms_per_step": ms / Kx,
                          "algorithmic_bytes": tot, "spynet_warp_bytes": spy[0].bytes,
                          "achieved": tot * Kx * 1e-9 / (ms * 1e-3),
                          "frac": tot * Kx * 1e-9 / (ms * 1e-3) / peak,
                          "note": "per-rank; the four SpyNet warps as one extra warp_multi launch"}
    del spy

    # -- stock PyTorch-CUDA eager on the same tensors ------------------
    if rank == 0:
        try:
            from oracle import dmc_ref
            o_ebs, o_gc = oracle_modules(dev)
            n_it = 5
            with torch.no_grad():
                for i in range(2):
                    dmc_ref.pframe_hot_path(paths[i % N_SETS].inp, o_ebs, o_gc)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(n_it):
                    dmc_ref.pframe_hot_path(paths[i % N_SETS].inp, o_ebs, o_gc)
                e1.record()
                torch.cuda.synchronize()
            ems = e0.elapsed_time(e1) / n_it
            out["gpu_eager_baseline"] = {
                "value": 1e3 / ems, "unit": UNIT, "ms_per_step": ems, "kind": "port",
                "sample": f"{n_it} P-frames of this run's input sets through the oracle port of "
                          "the reference's ops executed by PyTorch-CUDA eager on the same GPU "
                          "(grid_sample, interpolate, ~300 elementwise launches, bmm)"}
        except Exception as e:  # noqa: BLE001
            out["gpu_eager_baseline"] = {"error": repr(e)[:200]}
    cx.barrier()
    return out


def measure_e2e(cx, paths, K):
    """Same metric through the public path object with HOST buffers.

    A codec keeps its decoded picture buffer on the device: ``x_ref`` and the
    feature pyramid of frame t are the outputs of frame t-1
    (video_model.py:544-549), 727 MB at 1080p that never cross PCIe.  What
    arrives from the host every frame is what depends on the NEW frame -- here
    the motion field, both latents with their hyper-prior and spatial-prior
    parameters and the hyper-latents (43.1 MB, stand-ins for the conv nets'
    outputs).  The host packs them into ONE pinned staging buffer per frame; every
    step copies it with one cudaMemcpyAsync into the flat device buffer the
    path's input tensors are views of (uploads run two frames ahead of the
    kernels on a copy stream), runs ``PFramePath.launch`` against one of four
    resident dpb sets (aggregate >> L2) and reads bits/bpp back.

    Also reported: the same copies with no kernels (`copy_only`: the PCIe/host
    ceiling of this byte pattern on this box at this rank count) and the round-1
    definition (all 770 MB re-uploaded per frame, `all_inputs_from_host`)."""
    torch, dist, dev, world = cx.torch, cx.dist, cx.dev, cx.world
    from deepvideocodec_b200.pipeline import PFramePath, frame_keys
    # at least one 96-frame sequence (test.py -f 96, BASELINE.json configs[2]): the two uploads
    # that fill the pipeline are inside the timed region and would weigh 10-20 % on 20 frames
    steps = cx.args.e2e_steps or max(96, min(K, 200))
    n_slots = len(paths)
    fk = frame_keys(paths[0].inp)
    # one flat device buffer per slot; the frame-dependent inputs become views of it
    offs, total = {}, 0
    for k in fk:
        offs[k] = total
        total += (paths[0].inp[k].numel() + 63) // 64 * 64          # 256-byte aligned views
    slots, flats, hosts = [], [], []
    with torch.no_grad():
        for p in paths:
            flat = torch.empty(total, dtype=torch.float32, device=dev)
            inp = dict(p.inp)
            for k in fk:
                v = flat[offs[k]:offs[k] + p.inp[k].numel()].view(p.inp[k].shape)
                v.copy_(p.inp[k])
                inp[k] = v
            slots.append(PFramePath(inp, cx.ebs, outputs={
                k: p.out[k] for k in ("warpframe", "context1", "context2", "context3")}))
            flats.append(flat)
            hosts.append(flat.cpu().pin_memory())
    h2d = total * 4
    bits_host = torch.empty(1, dtype=torch.float64).pin_memory()
    bpp_host = torch.empty(4, dtype=torch.float32).pin_memory()
    d2h = bits_host.numel() * 8 + bpp_host.numel() * 4
    copy_stream = torch.cuda.Stream(dev)
    comp = torch.cuda.current_stream(dev)
    ready = [torch.cuda.Event() for _ in range(n_slots)]
    freed = [torch.cuda.Event() for _ in range(n_slots)]

    def upload(slot):
        copy_stream.wait_event(freed[slot])
        with torch.cuda.stream(copy_stream):
            flats[slot].copy_(hosts[slot], non_blocking=True)        # ONE cudaMemcpyAsync
        ready[slot].record(copy_stream)

    def compute(slot, kernels=True):
        comp.wait_event(ready[slot])
        if kernels:
            out = slots[slot].launch()
            bits_host.copy_(out["bits"], non_blocking=True)
            bpp_host.copy_(out["bpp"].view(-1), non_blocking=True)
        freed[slot].record(comp)

    AHEAD = 2

    def run(n, kernels=True):
        for f in freed:
            f.record(comp)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(comp)                       # the first frames' uploads are inside the region
        copy_stream.wait_event(e0)
        for i in range(min(AHEAD, n)):
            upload(i % n_slots)
        for i in range(n):
            compute(i % n_slots, kernels)
            if i + AHEAD < n:
                upload((i + AHEAD) % n_slots)
        e1.record(comp)
        torch.cuda.synchronize()
        return cx.max_over_ranks(e0.elapsed_time(e1))

    # warm-up: one pass over every slot is not enough for the HOST side -- on this pool the same
    # copies run at 46 GB/s right after a short main region and at 55 GB/s once the host has been
    # busy for a while (uncore / link power states), copy-only ceiling and e2e alike
    run(4)
    run(min(64, steps), kernels=False)
    run(min(32, steps))
    ms = run(steps)
    ms_copy = run(steps, kernels=False)
    e2e = {"value": world * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": steps, "ms_per_step": ms / steps,
           "copy_only": {"ms_per_step": ms_copy / steps,
                         "aggregate_GBps": world * h2d * steps / (ms_copy * 1e-3) / 1e9,
                         "note": "the same uploads with no kernels: the PCIe/host-memory "
                                 "ceiling of this byte pattern at this rank count"},
           "frac_of_copy_ceiling": ms_copy / ms,
           "resident_dpb_bytes": sum(paths[0].inp[k].numel() * 4 for k in paths[0].inp
                                     if k not in fk),
           "note": "dpb (x_ref + 3 feature scales, 727 MB) resident in HBM between frames as in "
                   "video_model.py:544-549; per step H2D = motion field + latents + priors + "
                   "hyper-latents packed in one pinned staging buffer (ONE cudaMemcpyAsync, 2 "
                   "frames ahead) -> PFramePath.launch on one of 4 resident dpb sets -> D2H of "
                   "bits/bpp; timed over max(96, min(--steps, 200)) frames (one 96-frame "
                   "sequence of test.py at least), pipeline fill included"}
    del hosts, flats, slots
    copy_streams = [copy_stream]

    # ---- round-1 definition: every input re-uploaded each frame (PCIe-bound, kept for continuity)
    try:
        full_steps = max(4, min(steps, 12))
        hostf = {k: v.cpu().pin_memory() for k, v in paths[0].inp.items()}
        fbytes = sum(v.numel() * v.element_size() for v in hostf.values())
        cs = copy_streams[0]
        rdy = [torch.cuda.Event() for _ in range(2)]
        fre = [torch.cuda.Event() for _ in range(2)]

        def up(slot):
            cs.wait_event(fre[slot])
            with torch.cuda.stream(cs):
                for k, v in hostf.items():
                    paths[slot].inp[k].copy_(v, non_blocking=True)
            rdy[slot].record(cs)

        def runf(n):
            for f in fre:
                f.record(comp)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(comp)
            cs.wait_event(e0)
            up(0)
            for i in range(n):
                if i + 1 < n:
                    up((i + 1) % 2)
                comp.wait_event(rdy[i % 2])
                out = paths[i % 2].launch()
                bits_host.copy_(out["bits"], non_blocking=True)
                fre[i % 2].record(comp)
            e1.record(comp)
            torch.cuda.synchronize()
            return cx.max_over_ranks(e0.elapsed_time(e1))
        runf(2)
        msf = runf(full_steps)
        e2e["all_inputs_from_host"] = {"value": world * full_steps / (msf * 1e-3),
                                       "h2d_bytes_per_step": fbytes, "steps": full_steps,
                                       "aggregate_GBps": world * fbytes * full_steps / (msf * 1e-3) / 1e9}
    except Exception as e:  # noqa: BLE001
        e2e["all_inputs_from_host"] = {"error": repr(e)[:200]}
    return e2e


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


# ---------------------------------------------------------------------------
# BASELINE.json configs[2]: GOP-serial units sharded over the ranks
# ---------------------------------------------------------------------------
def run_gop(cx):
    """96-frame sequences -> (sequence, GOP) units (I-frame every 32 frames,
    test.py:162) -> ``shard_units(rank, world)`` -> frames serial inside a unit
    (frame t's warped frame / contexts are frame t+1's dpb) -> per-rank
    ``RateStats`` -> ONE NCCL all-reduce (test.py:275-281 summed over sequences)."""
    torch = cx.torch
    from deepvideocodec_b200.dist import RateStats, make_units, reduce_stats, shard_units
    from deepvideocodec_b200.gop import GopRunner
    n_seq = cx.args.sequences * cx.world
    units = make_units([96] * n_seq)
    mine = shard_units(units, cx.rank, cx.world)
    runner = GopRunner(H, W, cx.dev, cx.ebs, regime=cx.args.regime)
    warm = runner.run_units(mine[:1])                            # warm-up: one unit ...
    reduce_stats(warm, device=cx.dev)                            # ... and one collective (lazy NCCL init)
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(cx.local)
    if sampler.ok:
        sampler.start()
    e0.record()
    stats = runner.run_units(mine)
    total = reduce_stats(stats, device=cx.dev)                  # NCCL: 4 fp64 words, once
    e1.record()
    cx.barrier()
    if sampler.ok:
        sampler.stop()
    ms = cx.max_over_ranks(e0.elapsed_time(e1))
    cx.close()
    if cx.rank != 0:
        return 0
    p_frames = sum(u.p_frames for u in units)
    assert int(total.frames) == p_frames, (total.frames, p_frames)
    line = {"mode": "gop", "metric": METRIC, "value": p_frames / (ms * 1e-3), "unit": UNIT,
            "n_gpus": cx.world, "ms_per_step": ms / (p_frames / cx.world), "higher_is_better": True,
            "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "config": {"workload": "DMC 96-frame GOP inference hot path at 1920x1088, independent "
                                   "sequences sharded across ranks (BASELINE.json configs[2])",
                       "sequences": n_seq, "frames_per_sequence": 96, "gop": 32,
                       "units": len(units), "p_frames": p_frames,
                       "units_per_rank": len(mine)},
            "sum_bits": total.bits, "sum_frames": total.frames, "sum_pixels": total.pixels,
            "bpp": total.bpp, "clocks": sampler.summary(),
            "collective": "one all_reduce(SUM) of 4 fp64 words over NCCL inside the timed region"}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
