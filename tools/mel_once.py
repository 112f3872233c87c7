"""One log-mel launch sequence for ncu: python tools/mel_once.py [clips] [fp32|int16] [iters]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mods = importlib.import_module("audio-to-motion-generation_b200").install_dropin()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
kind = sys.argv[2] if len(sys.argv) > 2 else "fp32"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lm = mods["pose_video.audio_repr"].log_mel_spectograms
wav = 0.1 * torch.randn(B, 68267, device="cuda")
if kind == "int16":
    wav = (30000 * wav).round().clamp(-32768, 32767).to(torch.int16)
for _ in range(iters):
    out = lm(wav)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.float().mean()))
