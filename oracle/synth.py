"""Oracle (test infrastructure): seeded synthetic PATS-shaped inputs.

Per-clip seeds make every clip independent of how the clip range is sharded over ranks.
"""
import numpy as np
import torch

SAMPLE_RATE = 16000            # pose_video/consts.py:14
CLIP_SAMPLES = 68267           # 64 pose frames / 15 fps * 16 kHz -> 425 mel frames
POSE_FRAMES = 64               # pose_video/consts.py:15
POSE_FEATS = 104               # 52 keypoints x (x, y)
ADAPTER_STRIDE = 6             # D2: logmel[:, 0:384:6, :]
ADAPTER_SPAN = POSE_FRAMES * ADAPTER_STRIDE


def wav_clip(index, n=CLIP_SAMPLES, kind="noise"):
    """fp32 waveform for global clip `index`."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1234 + int(index))
    noise = torch.randn(n, generator=g, dtype=torch.float32)
    if kind == "noise":                      # 0.1 * N(0,1)
        return (0.1 * noise).numpy()
    if kind == "tone":                       # 440 Hz tone + 1e-4 noise (dynamic-range stress)
        t = torch.arange(n, dtype=torch.float64) / SAMPLE_RATE
        return (0.5 * torch.sin(2 * np.pi * 440.0 * t).float() + 1e-4 * noise).numpy()
    if kind == "int16":                      # int16-scale samples
        return (3000.0 * noise).round().clamp(-32768, 32767).numpy()
    if kind == "zeros":
        return np.zeros(n, dtype=np.float32)
    raise KeyError(kind)


def wav_batch(start, count, n=CLIP_SAMPLES, kind="noise"):
    return np.stack([wav_clip(start + i, n, kind) for i in range(count)])


def gt_pose_clip(index, t=POSE_FRAMES):
    g = torch.Generator(device="cpu")
    g.manual_seed(4321 + int(index))
    return (50.0 * torch.randn(t, POSE_FEATS, generator=g, dtype=torch.float32)).numpy()


def gt_pose_batch(start, count, t=POSE_FRAMES):
    return np.stack([gt_pose_clip(start + i, t) for i in range(count)])


def noisy_pred_batch(start, count, t=POSE_FRAMES, sigma=12.0):
    """gt + sigma * N(0,1): the eval-only micro-benchmark prediction (exercises hit and miss)."""
    out = []
    for i in range(count):
        g = torch.Generator(device="cpu")
        g.manual_seed(9876 + int(start + i))
        out.append(gt_pose_clip(start + i, t) +
                   (sigma * torch.randn(t, POSE_FEATS, generator=g, dtype=torch.float32)).numpy())
    return np.stack(out).astype(np.float32)


def adapter(logmel):
    """D2: [B,425,64] -> [B,64,64], the reference's strided-slice feed (dataUtils.py:654, ratio 6)."""
    return logmel[:, 0:ADAPTER_SPAN:ADAPTER_STRIDE, :]


def shard_range(n_items, rank, world):
    """Contiguous clip range of `rank`: [r*ceil(n/W), min(n,(r+1)*ceil(n/W)))."""
    per = -(-n_items // world)
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
