"""Host-side cost of PFramePath.launch(): wall-clock of the enqueue loop vs device time."""
import os, sys, time, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc
from deepvideocodec_b200.pipeline import PFramePath, synthetic_pframe_inputs
dev = torch.device("cuda:0")
torch.manual_seed(1234)
ebs = {"motion": dvc.EntropyBottleneck(64).to(dev).eval(), "frame": dvc.EntropyBottleneck(64).to(dev).eval()}
with torch.no_grad():
    paths = [PFramePath(synthetic_pframe_inputs(1088, 1920, dev, 1234 + s), ebs) for s in range(4)]
res = {}
for K in (100, 400, 1000):
    for events in (False, True):
        for i in range(10):
            paths[i % 4].launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)] if events else None
        t0 = time.perf_counter()
        e0.record()
        for i in range(K):
            paths[i % 4].launch(warp_events=wev[i] if wev else None)
        e1.record()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        res[f"K={K} events={events}"] = {"enqueue_us_per_step": t_enq / K * 1e6, "device_us_per_step": e0.elapsed_time(e1) / K * 1e3,
                                         "wall_us_per_step": t_all / K * 1e6}
print(json.dumps(res, indent=1))
