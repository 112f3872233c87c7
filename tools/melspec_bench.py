"""CUDA-event timing of the PATS-native front ends (log_mel_400 on the fused kernel, log_mel_512 on melspec_wide).
    python tools/melspec_bench.py [n_clips] [seconds] [iters]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pa = importlib.import_module("audio-to-motion-generation_b200.pats_audio")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.2667
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, sr, fn in (("log_mel_400", 16000, lambda w: pa.log_mel_400(w, 16000)), ("log_mel_512", 44100, lambda w: pa.log_mel_512(w, 44100))):
    n = int(round(sr * secs))
    wav = 0.1 * torch.randn(B, n, device="cuda")
    for _ in range(3):
        y = fn(wav)
    ts = []
    for _ in range(iters):
        flush.zero_()                                  # evict the inputs from L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = fn(wav); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    bytes_alg = wav.numel() * 4 + y.numel() * 4      # every sample read once, every output written once
    out[name] = {"clips": B, "samples": n, "frames": y.shape[1], "ms": ms, "clips_per_s": B / ms * 1e3,
                 "algorithmic_GBps": bytes_alg / ms / 1e6, "l2": "256 MB flush between launches"}
print(json.dumps(out))
