"""B200 drop-in for ``pose_video/audio_repr.py``: the wrapper that pins the hot-path mel parameters
(16 kHz, log offset 0.01, 25 ms / 10 ms, 64 mel bins, 125-7500 Hz; audio_repr.py:18-26) and the
``repr_map`` / ``get_repr`` registry (:29-36).

D6: arrays / tensors are the well-defined input.  The shipped ``str`` branch hands the
``(wav, sr)`` tuple of ``raw_repr`` to the mel function (:21) and imports a non-existent ``common``
package (:4-5).  Here a path is decoded with librosa when librosa is installed (tuple unpacked),
otherwise a clear ImportError is raised -- the mel computation itself always runs on the GPU.
"""
from . import mel_features

SR = 16000                       # pose_video/consts.py:14
RAW, LOG_MEL_SPECT = 'raw', 'log_mel_spect'


def raw_repr(path, sr=None):
    """(waveform, sample_rate) of an audio file, mono (audio_repr.py:13-15)."""
    try:
        import librosa
    except ImportError as exc:
        raise ImportError("raw_repr(path) needs librosa to decode audio files; "
                          "pass a waveform array or tensor instead") from exc
    return librosa.load(path, sr=sr, mono=True)


def log_mel_spectograms(path, audio_sample_rate=SR, log_offset=0.01, window_length_secs=0.025,
                        hop_length_secs=0.010, num_mel_bins=64, num_min_hz=125, num_max_hz=7500):
    """Waveform (array, tensor, batch of either, or a file path) -> log-mel [frames, num_mel_bins]."""
    waveform = raw_repr(path, audio_sample_rate)[0] if isinstance(path, str) else path
    band = dict(num_mel_bins=num_mel_bins, lower_edge_hertz=num_min_hz, upper_edge_hertz=num_max_hz)
    return mel_features.log_mel_spectrogram(waveform, audio_sample_rate, log_offset, window_length_secs,
                                            hop_length_secs, **band)


repr_map = {RAW: raw_repr, LOG_MEL_SPECT: log_mel_spectograms}


def get_repr(repr_name):
    return repr_map[repr_name]
