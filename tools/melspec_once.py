import importlib, os, sys, torch
sys.path.insert(0, "/root/repo")
pa = importlib.import_module("audio-to-motion-generation_b200.pats_audio")
wav = 0.1 * torch.randn(256, 188161, device="cuda")
for _ in range(3):
    y = pa.log_mel_512(wav, 44100)
torch.cuda.synchronize()
print(y.shape)
