"""Timed runs of the BASELINE configs that are not the bench line (3: eval of 100 k clips, 4: long-form streaming +
sliding windows, 5: mel bandwidth sweep), CUDA events, L2 flushed between iterations.  Prints one JSON object.
    python tools/config_runs.py > gpurun_out/configs.json
"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
a2m = importlib.import_module("audio-to-motion-generation_b200")
mods = a2m.install_dropin()
pipeline = importlib.import_module("audio-to-motion-generation_b200.pipeline")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


out = {"hbm_peak_gbs": PEAK}
lm = mods["pose_video.audio_repr"].log_mel_spectograms
sweep = []
for B in (1, 16, 256, 4096, 16384, 65536):
    wav = 0.1 * torch.randn(B, 68267, device="cuda")
    ms = timeit(lambda: lm(wav), iters=3 if B >= 16384 else 7)
    byt = B * (68267 * 4 + 425 * 64 * 4)
    sweep.append({"clips": B, "ms": ms, "clips_per_s": B / ms * 1e3, "gbs": byt / ms / 1e6, "frac_of_hbm": byt / ms / 1e6 / PEAK})
    del wav
out["config5_mel_sweep"] = sweep

ev = mods["motion_evaluation"]
n = 100000
gt = 50 * torch.randn(n, 64, 104, device="cuda")
pr = gt + 12 * torch.randn(n, 64, 104, device="cuda")
acc = ev.new_metrics()
ms = timeit(lambda: ev.evaluate_poses(pr, gt, accum=acc))
out["config3_eval_100k"] = {"clips": n, "ms": ms, "clips_per_s": n / ms * 1e3, "gbs": n * 64 * 104 * 8 / ms / 1e6,
                            "frac_of_hbm": n * 64 * 104 * 8 / ms / 1e6 / PEAK}
sm = ev.new_smoothness()
ms = timeit(lambda: ev.evaluate_smoothness(pr, accum=sm, from_pose=True))
out["smoothness_jerk_100k"] = {"clips": n, "ms": ms, "clips_per_s": n / ms * 1e3, "gbs": n * 64 * 104 * 4 / ms / 1e6,
                               "frac_of_hbm": n * 64 * 104 * 4 / ms / 1e6 / PEAK}
del gt, pr

# PATS-native front ends: 4.27 s clips, log_mel_400 at 16 kHz and log_mel_512 at 44.1 kHz
pa = importlib.import_module("audio-to-motion-generation_b200.pats_audio")
pats = []
for name, sr, fn in (("log_mel_400", 16000, lambda w: pa.log_mel_400(w, 16000)), ("log_mel_512", 44100, lambda w: pa.log_mel_512(w, 44100))):
    for B in (256, 4096):
        wav = 0.1 * torch.randn(B, int(round(sr * 64 / 15)), device="cuda")
        y = fn(wav)
        ms = timeit(lambda: fn(wav), iters=5)
        byt = wav.numel() * 4 + y.numel() * 4
        pats.append({"front_end": name, "clips": B, "samples": wav.shape[1], "frames": y.shape[1], "ms": ms,
                     "clips_per_s": B / ms * 1e3, "gbs": byt / ms / 1e6, "frac_of_hbm": byt / ms / 1e6 / PEAK})
        del wav, y
out["pats_front_ends"] = pats

torch.manual_seed(0)
model = mods["real_motion_model"].SelfAttention_G().cuda().eval()
pipe = pipeline.AudioToPosePipeline(model, lanes=1)
wav = 0.1 * torch.randn(32, 960000, device="cuda")
ms = timeit(lambda: pipe.generate_long(wav), iters=3, warmup=1)
out["config4_long_form"] = {"streams": 32, "seconds_each": 60, "windows_per_stream": 188, "ms": ms,
                            "windows_per_s": 32 * 188 / ms * 1e3, "audio_seconds_per_s": 32 * 60 / ms * 1e3}
print(json.dumps(out, indent=1))
