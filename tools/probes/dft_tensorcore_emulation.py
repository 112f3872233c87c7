"""Would a tensor-core DFT beat the CUDA-core FFT of the log-mel kernel?  CPU emulation of the numerics + the arithmetic.

The 512-point real DFT of a windowed frame as a GEMM: X[frame, 2 x 257] = frames[frame, 400] . [cos | -sin][400, 514]
(the Hann window folded into the matrix).  Tensor-core input types carry 8 (bf16) or 11 (tf32) significand bits, so an
fp32-accurate product needs the operands split into 2-3 terms and 3-6 MMAs per product (the "3xTF32" / "bf16x3" trick).
This script rounds the operands the way the tensor core would, accumulates in fp32 (numpy matmul on rounded fp32 inputs),
finishes the log-mel in fp64 and compares with the oracle on the clip that decides the matter -- tests/golden tone_8000, a
pure tone + 1e-4 noise whose off-peak bins sit 80 dB below the peak -- against the parity bar 1e-4 (max abs, D8).
    python tools/probes/dft_tensorcore_emulation.py          (CPU only; reads tests/golden, imports oracle/ as the checker)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mel_oracle  # noqa: E402


def round_bits(x, bits):
    """round-to-nearest-even to `bits` significand bits (tf32: 11, bf16: 8), fp32 container"""
    x = np.asarray(x, np.float32)
    if bits >= 24:
        return x
    u = x.view(np.uint32).astype(np.uint64)
    drop = 24 - bits
    u = (u + (1 << (drop - 1)) - 1 + ((u >> drop) & 1)) >> drop << drop
    return u.astype(np.uint32).view(np.float32)


def split(x, bits, terms):
    parts, r = [], np.asarray(x, np.float32)
    for _ in range(terms):
        p = round_bits(r, bits)
        parts.append(p)
        r = (r - p).astype(np.float32)
    return parts


def dft_gemm(frames, basis, bits, terms):
    """sum over the cross terms a_i . b_j with i + j < terms (the usual split-precision product), fp32 accumulate"""
    a, b = split(frames, bits, terms), split(basis, bits, terms)
    acc = np.zeros((frames.shape[0], basis.shape[1]), np.float32)
    n_mma = 0
    for i in range(terms):
        for j in range(terms - i):
            acc += a[i] @ b[j]
            n_mma += 1
    return acc, n_mma


def main():
    kw = mel_oracle.AUDIO_REPR_KW
    rng = np.random.default_rng(1)                    # the same kind of clip as golden tone_8000
    t = np.arange(8000) / 16000.0
    wav = np.sin(2 * np.pi * 440.0 * t) + 1e-4 * rng.standard_normal(8000)
    ref = mel_oracle.log_mel(wav, **kw)
    fr = mel_oracle.frames(wav, 400, 160).astype(np.float32)
    n = np.arange(400)
    k = np.arange(257)
    w = mel_oracle.hann(400)
    ang = 2 * np.pi * np.outer(n, k) / 512.0
    basis = np.concatenate([w[:, None] * np.cos(ang), -w[:, None] * np.sin(ang)], axis=1).astype(np.float32)
    melw = mel_oracle.mel_matrix(num_spectrogram_bins=257, audio_sample_rate=16000, num_mel_bins=64,
                                 lower_edge_hertz=125, upper_edge_hertz=7500)
    print("clip: 440 Hz tone + 1e-4 noise, %d frames; parity bar max|a-b| <= 1e-4 (D8)" % fr.shape[0])
    flop_pass = 2.0 * 425 * 400 * 514 * 256            # one GEMM pass over a B = 256 step (425 frames per clip)
    for name, bits, terms in (("bf16 x1", 8, 1), ("tf32 x1", 11, 1), ("bf16 x2 (3 MMAs)", 8, 2), ("tf32 x2 (3 MMAs)", 11, 2),
                              ("bf16 x3 (6 MMAs)", 8, 3), ("tf32 x3 (6 MMAs)", 11, 3), ("fp32 exact GEMM", 24, 1)):
        x, n_mma = dft_gemm(fr, basis, bits, terms)
        mag = np.sqrt(x[:, :257].astype(np.float64) ** 2 + x[:, 257:].astype(np.float64) ** 2)
        out = np.log(mag @ melw + kw["log_offset"])
        err = np.abs(out - ref).max()
        if bits >= 24:
            print("  %-18s max|err| %.2e  %s   (no tensor-core type: the error floor of an fp32-accumulated dense DFT)"
                  % (name, err, "PASS" if err <= 1e-4 else "FAIL"))
            continue
        rate = 1623e12 if bits == 8 else 811e12        # measured bf16 burst; tf32 runs at half the bf16 rate
        us = n_mma * flop_pass / rate * 1e6
        print("  %-18s max|err| %.2e  %s   %d GEMM passes = %5.1f GFLOP per 256 clips -> >= %5.0f us at the measured peak"
              % (name, err, "PASS" if err <= 1e-4 else "FAIL", n_mma, n_mma * flop_pass / 1e9, us))
    print("shipped CUDA-core FFT (packed f32x2): 83 us per 256 clips, PASS (tests/test_mel_gpu.py)")


if __name__ == "__main__":
    main()
