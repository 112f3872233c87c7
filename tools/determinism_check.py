"""Diagnose batch-size / run-to-run differences stage by stage (GPU box)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth, weights
a2m = importlib.import_module("audio-to-motion-generation_b200")
mods = a2m.install_dropin()
pipeline = importlib.import_module("audio-to-motion-generation_b200.pipeline")
lm = mods["pose_video.audio_repr"].log_mel_spectograms
model = mods["real_motion_model"].SelfAttention_G().cuda().eval()
model.load_state_dict(weights.make_state_dict(0, "stress"))
wav = torch.from_numpy(synth.wav_batch(10, 8)).cuda()
for rep in range(3):
    m8 = lm(wav)
    m2 = torch.cat([lm(wav[i:i + 2]) for i in range(0, 8, 2)])
    print("rep", rep, "mel   B8 vs 4xB2 max|d|", (m8 - m2).abs().max().item(), " rerun B8", (m8 - lm(wav)).abs().max().item())
    x = pipeline.adapter(m8).contiguous()
    p8, _ = model(x)
    p2 = torch.cat([model(x[i:i + 2])[0] for i in range(0, 8, 2)])
    p8b, _ = model(x)
    print("rep", rep, "model B8 vs 4xB2 max|d|", (p8 - p2).abs().max().item(), " rerun B8", (p8 - p8b).abs().max().item())
    enc8 = model.audio_encoder(x)
    enc2 = torch.cat([model.audio_encoder(x[i:i + 2]) for i in range(0, 8, 2)])
    print("rep", rep, "enc   B8 vs 4xB2 max|d|", (enc8 - enc2).abs().max().item())
    u8 = model.unet(enc8)
    u2 = torch.cat([model.unet(enc8[i:i + 2]) for i in range(0, 8, 2)])
    print("rep", rep, "unet  B8 vs 4xB2 max|d|", (u8 - u2).abs().max().item())
    for part, J in (("body", 10), ("hand", 42)):
        g = torch.randn(8 * 64, J, 64, device="cuda")
        g8 = model.graph_stack(part, g)
        g2 = torch.cat([model.graph_stack(part, g[i:i + 128]) for i in range(0, 512, 128)])
        print("rep", rep, "gnn", part, "max|d|", (g8 - g2).abs().max().item())
