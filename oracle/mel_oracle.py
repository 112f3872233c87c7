"""Oracle (test infrastructure): fp64 numpy restatement of the reference log-mel front end.

Follows ``/root/reference/pose_video/mel_features.py``:
  frames            <- frame()                     :21-45
  hann              <- periodic_hann()             :48-68
  stft_mag          <- stft_magnitude()            :71-92
  hz_to_mel         <- hertz_to_mel()              :100-111
  mel_matrix        <- spectrogram_to_mel_matrix() :114-189
  log_mel           <- log_mel_spectrogram()       :192-223
and the fixed parameters of ``pose_video/audio_repr.py:18-26`` (``AUDIO_REPR_KW``).

The arithmetic is the same sequence of fp64 element-wise operations as the reference, so
the results are bit-identical to it (checked in tests/test_oracle_golden.py against vectors
produced by the unmodified reference, see oracle/make_golden.py).
"""
import numpy as np

MEL_BREAK_HZ = 700.0      # mel_features.py:96
MEL_Q = 1127.0            # mel_features.py:97

# audio_repr.py:18-26 (the only in-repo caller of log_mel_spectrogram)
AUDIO_REPR_KW = dict(audio_sample_rate=16000, log_offset=0.01, window_length_secs=0.025,
                     hop_length_secs=0.010, num_mel_bins=64, lower_edge_hertz=125,
                     upper_edge_hertz=7500)


def num_frames(num_samples, window_length, hop_length):
    """mel_features.py:41-42 -- no padding, incomplete tail frame dropped."""
    return 1 + int(np.floor((num_samples - window_length) / hop_length))


def frames(signal, window_length, hop_length):
    """Gather formulation of the reference's as_strided view (mel_features.py:41-45)."""
    signal = np.asarray(signal)
    nf = num_frames(signal.shape[0], window_length, hop_length)
    if nf < 0:
        # the reference hands a negative shape to as_strided, numpy raises ValueError
        raise ValueError("negative dimensions are not allowed")
    idx = hop_length * np.arange(nf)[:, None] + np.arange(window_length)[None, :]
    return signal[idx]


def hann(window_length):
    """Periodic (period-N) Hann, mel_features.py:67-68."""
    n = np.arange(window_length)
    return 0.5 - (0.5 * np.cos(2 * np.pi / window_length * n))


def stft_mag(signal, fft_length, hop_length, window_length):
    """|rfft(frames * hann, n=fft_length)|, mel_features.py:86-92."""
    windowed = frames(signal, window_length, hop_length) * hann(window_length)
    return np.abs(np.fft.rfft(windowed, int(fft_length)))


def hz_to_mel(f_hz):
    """HTK mel, mel_features.py:110-111."""
    return MEL_Q * np.log(1.0 + (f_hz / MEL_BREAK_HZ))


def mel_matrix(num_mel_bins=20, num_spectrogram_bins=129, audio_sample_rate=8000,
               lower_edge_hertz=125.0, upper_edge_hertz=3800.0):
    """[num_spectrogram_bins, num_mel_bins] triangular weights, linear in mel
    (mel_features.py:155-189); all bands evaluated at once instead of the reference's loop."""
    nyquist = audio_sample_rate / 2.
    if lower_edge_hertz < 0.0:
        raise ValueError("lower_edge_hertz %.1f must be >= 0" % lower_edge_hertz)
    if lower_edge_hertz >= upper_edge_hertz:
        raise ValueError("lower_edge_hertz %.1f >= upper_edge_hertz %.1f"
                         % (lower_edge_hertz, upper_edge_hertz))
    if upper_edge_hertz > nyquist:
        raise ValueError("upper_edge_hertz %.1f is greater than Nyquist %.1f"
                         % (upper_edge_hertz, nyquist))
    bins_mel = hz_to_mel(np.linspace(0.0, nyquist, num_spectrogram_bins))[:, None]
    edges = np.linspace(hz_to_mel(lower_edge_hertz), hz_to_mel(upper_edge_hertz),
                        num_mel_bins + 2)
    lo, ce, hi = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    rising = (bins_mel - lo) / (ce - lo)
    falling = (hi - bins_mel) / (hi - ce)
    w = np.maximum(0.0, np.minimum(rising, falling))
    w[0, :] = 0.0                      # DC bin excluded, mel_features.py:188
    return w


def stft_geometry(audio_sample_rate, window_length_secs, hop_length_secs):
    """(window, hop, fft) in samples, mel_features.py:212-214."""
    win = int(round(audio_sample_rate * window_length_secs))
    hop = int(round(audio_sample_rate * hop_length_secs))
    nfft = 2 ** int(np.ceil(np.log(win) / np.log(2.0)))
    return win, hop, nfft


def log_mel(data, audio_sample_rate=8000, log_offset=0.0, window_length_secs=0.025,
            hop_length_secs=0.010, **kwargs):
    """log(|STFT| @ mel_matrix + log_offset), mel_features.py:212-223 -> float64 [frames, mel]."""
    win, hop, nfft = stft_geometry(audio_sample_rate, window_length_secs, hop_length_secs)
    spec = stft_mag(data, nfft, hop, win)
    w = mel_matrix(num_spectrogram_bins=spec.shape[1], audio_sample_rate=audio_sample_rate,
                   **kwargs)
    return np.log(np.dot(spec, w) + log_offset)


def log_mel_audio_repr(wav):
    """audio_repr.log_mel_spectograms(array) with its fixed parameters (audio_repr.py:18-26)."""
    return log_mel(wav, **AUDIO_REPR_KW)


def log_mel_batch(wavs):
    """Convenience for tests/benches: per-clip loop (the reference has no batch dim)."""
    return np.stack([log_mel_audio_repr(w) for w in wavs])
