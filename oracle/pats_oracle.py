"""Oracle (test infrastructure): numpy restatement of the PATS-native audio front ends of the reference,
``/root/reference/pats/data_loading/audio.py`` ``log_mel_400`` :86-120 (and the filterbank / framing pieces
``log_mel_512`` :58-79 shares).

PARITY UNPINNED: the arithmetic lives in librosa (``librosa.core.stft``, ``librosa.feature.melspectrogram``,
``librosa.filters.mel``), a third-party dependency that is absent here and unpinned in the reference (no
requirements file; the code's ``np.float`` dates it to numpy < 1.24 / librosa 0.8-0.9).  This file restates
librosa's published algorithm -- Slaney mel scale (linear to 1 kHz, then log with step ln(6.4)/27), triangular
filters in Hz, ``center=False`` framing of n_fft samples with the win_length window zero-padded to n_fft on both
sides, periodic Hann -- and is anchored only by its own known answers (tests/test_pats_audio_cpu.py).  fp64 throughout.
"""
import numpy as np


def slaney_mel(hz):
    hz = np.asarray(hz, dtype=np.float64)
    lin = 3.0 * hz / 200.0
    log_region = 15.0 + 27.0 * np.log(np.maximum(hz, 1.0) / 1000.0) / np.log(6.4)
    return np.where(hz >= 1000.0, log_region, lin)


def slaney_hz(mel):
    mel = np.asarray(mel, dtype=np.float64)
    return np.where(mel >= 15.0, 1000.0 * 6.4 ** ((mel - 15.0) / 27.0), 200.0 * mel / 3.0)


def filterbank(sr, n_fft, n_mels, fmin, fmax, area_norm):
    """[n_mels, n_fft//2+1]: filter i rises from centre i-1... over centres c[i], c[i+1], c[i+2] (Hz)."""
    bins = np.arange(n_fft // 2 + 1) * (sr / n_fft)
    c = slaney_hz(np.linspace(slaney_mel(fmin), slaney_mel(fmax), n_mels + 2))
    w = np.zeros((n_mels, bins.size))
    for i in range(n_mels):
        rise = (bins - c[i]) / (c[i + 1] - c[i])
        fall = (c[i + 2] - bins) / (c[i + 2] - c[i + 1])
        w[i] = np.clip(np.minimum(rise, fall), 0.0, None)
        if area_norm:
            w[i] *= 2.0 / (c[i + 2] - c[i])
    return w.astype(np.float32).astype(np.float64)          # librosa stores the basis as float32


def stft_mag_uncentred(y, n_fft, hop, win_length):
    """|STFT| with center=False: frames y[t*hop : t*hop + n_fft] times the padded periodic Hann -> [frames, bins]."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = np.arange(win_length)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)
    left = (n_fft - win_length) // 2
    win = np.zeros(n_fft)
    win[left:left + win_length] = hann
    n_frames = 1 + (y.size - n_fft) // hop
    if n_frames < 1:
        raise ValueError("too short for one frame")
    frames = np.stack([y[t * hop:t * hop + n_fft] for t in range(n_frames)])
    return np.abs(np.fft.rfft(frames * win, axis=1))


def log_mel_400(y, eps=1e-6):
    """audio.py:86-120 for 16 kHz input: [frames, 64]."""
    mag = stft_mag_uncentred(y, 512, 160, 400)
    spec = mag @ filterbank(16000, 512, 64, 125.0, 7500.0, area_norm=False).T
    spec = np.where(spec == 0, eps, spec)
    return np.log(spec)


def stft_power_centred(y, n_fft, hop, pad_mode="reflect"):
    """|STFT|^2 with center=True: y padded by n_fft // 2 on both sides ('reflect' without edge repeat, or zeros),
    full-length periodic Hann -> [1 + N // hop, bins]."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    yp = np.pad(y, n_fft // 2, mode="reflect" if pad_mode == "reflect" else "constant")
    n = np.arange(n_fft)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)
    n_frames = 1 + (yp.size - n_fft) // hop
    frames = np.stack([yp[t * hop:t * hop + n_fft] for t in range(n_frames)])
    return np.abs(np.fft.rfft(frames * hann, axis=1)) ** 2


def log_mel_512(y, sr, eps=1e-10, pad_mode="reflect"):
    """audio.py:58-79: librosa.feature.melspectrogram(y, sr, n_fft=2048, hop_length=512) defaults (power 2, 128 bands
    0..sr/2, Slaney area normalisation) -> zeros to eps -> log -> [frames, 128]."""
    spec = stft_power_centred(y, 2048, 512, pad_mode) @ filterbank(sr, 2048, 128, 0.0, sr / 2.0, area_norm=True).T
    spec = np.where(spec == 0, eps, spec)
    return np.log(spec)


# ---------------------------------------------------------------------------------------------------------------
# resampling (log_mel_400's first step, audio.py:87: librosa.core.resample(y, orig_sr=sr, target_sr=16000))
# PARITY UNPINNED like the rest of this file: librosa (and resampy, which implements its 'kaiser_best' mode, the
# default up to librosa 0.9) are absent.  Restated from resampy's published algorithm: band-limited sinc interpolation
# (J. O. Smith) with a Kaiser-windowed sinc table -- 64 zero crossings, 2^9 table entries per crossing, beta
# 14.769656459379492, roll-off 0.9475937167399596 -- linearly interpolated between table entries; the output has
# ceil(n * ratio) samples (librosa's fix_length of resampy's int(n * ratio)), no amplitude scaling (scale=False).
# ---------------------------------------------------------------------------------------------------------------
KAISER_BEST = dict(num_zeros=64, precision=9, beta=14.769656459379492, rolloff=0.9475937167399596)


def sinc_table(num_zeros=64, precision=9, beta=14.769656459379492, rolloff=0.9475937167399596):
    """resampy.filters.sinc_window with a Kaiser taper: (interp_win [num_zeros * 2^precision + 1], 2^precision)."""
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits


def resample(y, orig_sr, target_sr):
    """fp64 restatement of resampy.resample(y, orig_sr, target_sr, filter='kaiser_best') + librosa's fix_length."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    if orig_sr == target_sr:
        return y.copy()
    ratio = float(target_sr) / orig_sr
    interp_win, num_table = sinc_table(**KAISER_BEST)
    if ratio < 1:
        interp_win = interp_win * ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    n_orig, n_out = y.size, int(y.size * ratio)
    out = np.zeros(int(np.ceil(y.size * ratio)))
    scale = min(1.0, ratio)
    index_step = int(scale * num_table)
    nwin = interp_win.size
    for t in range(n_out):
        time_register = t / ratio
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i = np.arange(min(n + 1, (nwin - offset) // index_step))
        idx = offset + i * index_step
        acc = np.dot(interp_win[idx] + eta * interp_delta[idx], y[n - i])
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k = np.arange(min(n_orig - n - 1, (nwin - offset) // index_step))
        idx = offset + k * index_step
        acc += np.dot(interp_win[idx] + eta * interp_delta[idx], y[n + k + 1])
        out[t] = acc
    return out
