"""Print the parity margins of the CUDA path against the committed goldens / the oracle (GPU box).
    python tools/parity_report.py
Tolerances: log-mel max|a-b| <= 1e-4 * max(1,|b|); pose sum|a-b|/sum|b| <= 1e-2; PCK hits bit-exact."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mel_oracle, model_oracle, eval_oracle, synth, weights          # noqa: E402
from oracle.make_golden import MEL_CASES, MODEL_CASES, model_input, real_pose_input  # noqa: E402

a2m = importlib.import_module("audio-to-motion-generation_b200")
mods = a2m.install_dropin()
G = {n: np.load(os.path.join(ROOT, "tests", "golden", n + "_reference.npz")) for n in ("mel", "eval", "model")}

lm = mods["pose_video.audio_repr"].log_mel_spectograms
for name, kind, n, idx in MEL_CASES:
    ref = G["mel"][name]
    if ref.size == 0:
        continue
    got = lm(torch.from_numpy(synth.wav_clip(idx, n, kind)).cuda()).cpu().numpy()
    err = np.abs(got - ref)
    print("mel   %-16s max|d| %.2e  max d/max(1,|ref|) %.2e  sum|d|/sum|ref| %.2e" % (
        name, err.max(), (err / np.maximum(1.0, np.abs(ref))).max(), err.sum() / np.abs(ref).sum()))

rmm = mods["real_motion_model"]
for name, seed, mode, B, T, F, with_pose in MODEL_CASES:
    m = rmm.SelfAttention_G().cuda().eval()
    m.load_state_dict(weights.make_state_dict(seed, mode))
    x = model_input(seed, B, T, F)
    pose, _ = m(x.cuda())
    ref = torch.from_numpy(G["model"][name + "_pose"])
    rel = ((pose.cpu() - ref).abs().sum() / ref.abs().sum()).item()
    print("model %-16s pose rel-L1 %.3e (bar 1e-2)" % (name, rel))
sd = weights.make_state_dict(0, "stress")
m = rmm.SelfAttention_G().cuda().eval()
m.load_state_dict(sd)
x = model_input(0, 2, 64, 64)
enc = m.audio_encoder(x.cuda()).cpu()
ref = torch.from_numpy(G["model"]["stress_b2_enc"])
print("model encoder          rel-L1 %.3e" % ((enc - ref).abs().sum() / ref.abs().sum()).item())
un = m.unet(ref.cuda()).cpu()
ref_u = torch.from_numpy(G["model"]["stress_b2_unet"])
print("model unet             rel-L1 %.3e" % ((un - ref_u).abs().sum() / ref_u.abs().sum()).item())
