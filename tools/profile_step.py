"""A few P-frame steps of the bench workload, for ncu (launch list / full capture).
Usage: python tools/profile_step.py [steps] [regime]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepvideocodec_b200 as dvc  # noqa: E402
from deepvideocodec_b200.pipeline import PFramePath, synthetic_pframe_inputs  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
regime = sys.argv[2] if len(sys.argv) > 2 else "smooth"
dev = torch.device("cuda:0")
torch.manual_seed(1234)
with torch.no_grad():
    ebs = {"motion": dvc.EntropyBottleneck(64).to(dev).eval(),
           "frame": dvc.EntropyBottleneck(64).to(dev).eval()}
    paths = [PFramePath(synthetic_pframe_inputs(1088, 1920, dev, 1234 + s, regime=regime), ebs)
             for s in range(2)]
    torch.cuda.synchronize()
    for i in range(steps):
        paths[i % 2].launch()
    torch.cuda.synchronize()
print("bits", float(paths[0].out["bits"][0]))
