"""Minimal driver for ncu: W warm-up steps + K steps of the device-resident hot path (B = 256).
Prints the number of library launches per step so the launch list can be cut to whole steps.
    python tools/profile_step.py [K] [W] [B]
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
a2m = importlib.import_module("audio-to-motion-generation_b200")
pipeline = importlib.import_module("audio-to-motion-generation_b200.pipeline")
rmm = importlib.import_module("audio-to-motion-generation_b200.real_motion_model")
lib = a2m.load_library()
torch.manual_seed(0)
model = rmm.SelfAttention_G()
with torch.no_grad():
    for n, p in model.state_dict().items():
        if n.endswith("gamma"):
            p.fill_(0.5)
model = model.cuda().eval()
pipe = pipeline.AudioToPosePipeline(model)
wav = [0.1 * torch.randn(B, 68267, device="cuda") for _ in range(3)]
gt = [50 * torch.randn(B, 64, 104, device="cuda") for _ in range(3)]
for i in range(W):
    pipe.step(wav[i % 3], gt[i % 3])
torch.cuda.synchronize()
lib.a2m_launch_count_reset()
for i in range(K):
    pipe.step(wav[i % 3], gt[i % 3])
torch.cuda.synchronize()
print("launches_per_step", lib.a2m_launch_count() // K, "steps", K, flush=True)
print(pipe.finish())
