"""Per-kernel micro-benchmarks (CUDA events, L2 flushed between iterations): mel, eval.
    python tools/kbench.py [mel] [eval]
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
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    which = sys.argv[1:] or ["mel", "eval"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if "mel" in which:
        lm = mods["pose_video.audio_repr"].log_mel_spectograms
        for B in (1, 16, 256, 4096, 16384):
            wav = 0.1 * torch.randn(B, 68267, device="cuda")
            med, best = timeit(lambda: lm(wav), flush=flush)
            byt = B * (68267 * 4 + 425 * 64 * 4)
            print("mel  B=%6d  %.3f ms (best %.3f)  %.2f Mclips/s  %.0f GB/s  %.1f%% of %.0f" % (
                B, med, best, B / med / 1e3, byt / med / 1e6, 100 * byt / med / 1e6 / PEAK, PEAK))
            pcm = (3000 * torch.randn(B, 68267, device="cuda")).round().clamp(-32768, 32767).to(torch.int16)
            med, best = timeit(lambda: lm(pcm), flush=flush)
            byt = B * (68267 * 2 + 425 * 64 * 4)
            print("mel16 B=%5d  %.3f ms (best %.3f)  %.2f Mclips/s  %.0f GB/s  %.1f%% of %.0f" % (
                B, med, best, B / med / 1e3, byt / med / 1e6, 100 * byt / med / 1e6 / PEAK, PEAK))
            del wav, pcm
    if "eval" in which:
        ev = mods["motion_evaluation"]
        for B in (256, 16384, 100000):
            gt = 50 * torch.randn(B, 64, 104, device="cuda")
            pr = gt + 12 * torch.randn(B, 64, 104, device="cuda")
            acc = ev.new_metrics()
            med, best = timeit(lambda: ev.evaluate_poses(pr, gt, accum=acc), flush=flush)
            byt = B * 64 * 104 * 4 * 2
            print("eval B=%6d  %.3f ms (best %.3f)  %.2f Mclips/s  %.0f GB/s  %.1f%% of %.0f" % (
                B, med, best, B / med / 1e3, byt / med / 1e6, 100 * byt / med / 1e6 / PEAK, PEAK))


if __name__ == "__main__":
    main()
