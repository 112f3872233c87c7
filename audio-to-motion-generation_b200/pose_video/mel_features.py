"""B200 drop-in for ``pose_video/mel_features.py`` of the reference (same names and argument meaning).

``log_mel_spectrogram`` and ``stft_magnitude`` run on the GPU through liba2m_b200.so (fused framing
-> periodic Hann -> real FFT -> |.| -> mel -> log, csrc/logmel.cu: a tuned kernel for the hot path's
512-point transform, a general one for the other power-of-two lengths the reference's rule produces).  The small constant
tables (``periodic_hann``, ``spectrogram_to_mel_matrix``) are evaluated on the host in fp64 with the
reference's formulas (mel_features.py:67-68, :155-189) and uploaded once per parameter set.

Extensions over the reference: a leading batch dimension ([B, N] -> [B, frames, mel]) and
torch tensors (CPU or CUDA).  numpy in -> numpy float64 out (the reference's dtype; values carry
fp32 precision, tolerance 1e-4 per SURVEY.md D8); torch in -> torch float32 out on the GPU.
int16 PCM (numpy or torch) is consumed as int16 by the kernel -- half the bytes to move; every other
dtype is converted to fp32 first.
Errors mirror the reference: bad band edges raise ValueError (:156-163); fewer samples than one
window minus one hop raises ValueError ("negative dimensions"); anything the CUDA path does not
implement (a window longer than 4096 samples) raises NotImplementedError -- never a silent CPU fallback.
"""
import ctypes
import math

import numpy as np
import torch

from .. import _cabi

_MEL_BREAK_FREQUENCY_HERTZ = 700.0
_MEL_HIGH_FREQUENCY_Q = 1127.0

_plans = {}


def frame(data, window_length, hop_length):
    """[num_samples, ...] -> [num_frames, window_length, ...]; no padding, tail dropped (:21-45)."""
    data = np.asarray(data)
    count = 1 + int(math.floor((data.shape[0] - window_length) / hop_length))
    if count < 0:
        raise ValueError("negative dimensions are not allowed")
    starts = np.arange(count) * hop_length
    return data[starts[:, None] + np.arange(window_length)[None, :]]


def periodic_hann(window_length):
    """One full period of a period-N raised cosine (:48-68)."""
    phase = 2 * np.pi / window_length * np.arange(window_length)
    return 0.5 - (0.5 * np.cos(phase))


def hertz_to_mel(frequencies_hertz):
    """HTK mel scale (:100-111)."""
    return _MEL_HIGH_FREQUENCY_Q * np.log(1.0 + (frequencies_hertz / _MEL_BREAK_FREQUENCY_HERTZ))


def spectrogram_to_mel_matrix(num_mel_bins=20, num_spectrogram_bins=129, audio_sample_rate=8000,
                              lower_edge_hertz=125.0, upper_edge_hertz=3800.0):
    """[num_spectrogram_bins, num_mel_bins] fp64 triangular filterbank, linear in mel, DC row zero
    (:114-189); raises ValueError on mis-ordered / out-of-range edges exactly like the reference."""
    nyquist_hertz = audio_sample_rate / 2.
    if lower_edge_hertz < 0.0:
        raise ValueError("lower_edge_hertz %.1f must be >= 0" % lower_edge_hertz)
    if lower_edge_hertz >= upper_edge_hertz:
        raise ValueError("lower_edge_hertz %.1f >= upper_edge_hertz %.1f" % (lower_edge_hertz, upper_edge_hertz))
    if upper_edge_hertz > nyquist_hertz:
        raise ValueError("upper_edge_hertz %.1f is greater than Nyquist %.1f" % (upper_edge_hertz, nyquist_hertz))
    bin_mel = hertz_to_mel(np.linspace(0.0, nyquist_hertz, num_spectrogram_bins))
    edge_mel = np.linspace(hertz_to_mel(lower_edge_hertz), hertz_to_mel(upper_edge_hertz), num_mel_bins + 2)
    weights = np.empty((num_spectrogram_bins, num_mel_bins))
    for band, (left, centre, right) in enumerate(zip(edge_mel[:-2], edge_mel[1:-1], edge_mel[2:])):
        up = (bin_mel - left) / (centre - left)
        down = (right - bin_mel) / (right - centre)
        weights[:, band] = np.maximum(0.0, np.minimum(up, down))
    weights[0, :] = 0.0
    return weights


class _Plan:
    def __init__(self, handle, window, hop, nfft, n_mel):
        self.handle, self.window, self.hop, self.nfft, self.n_mel = handle, window, hop, nfft, n_mel

    def num_frames(self, n_samples):
        return int(_cabi.lib().a2m_mel_num_frames(self.handle, n_samples))


def _get_plan(device_index, window, hop, nfft, log_offset, mel_kwargs):
    key = (device_index, window, hop, nfft, float(log_offset), tuple(sorted(mel_kwargs.items())))
    plan = _plans.get(key)
    if plan is None:
        if nfft < 64 or nfft > 4096:
            raise NotImplementedError(
                "the B200 log-mel kernels implement fft lengths 64 .. 4096 (windows of 33 .. 4096 samples); got %d" % nfft)
        weights = np.ascontiguousarray(spectrogram_to_mel_matrix(num_spectrogram_bins=nfft // 2 + 1, **mel_kwargs),
                                       dtype=np.float64)
        hann = np.ascontiguousarray(periodic_hann(window), dtype=np.float64)
        out = ctypes.c_void_p()
        _cabi.check(_cabi.lib().a2m_mel_plan_create(
            window, hop, nfft, weights.shape[1], hann.ctypes.data_as(ctypes.c_void_p),
            weights.ctypes.data_as(ctypes.c_void_p), float(log_offset), device_index, ctypes.byref(out)))
        plan = _plans[key] = _Plan(out, window, hop, nfft, weights.shape[1])
    return plan


def _to_device(data):
    """-> (CUDA tensor [B, N] with unit inner stride -- int16 if the input is int16, else fp32 --, had_batch_dim,
    input_was_numpy)."""
    was_numpy = not isinstance(data, torch.Tensor)
    t = torch.as_tensor(np.asarray(data)) if was_numpy else data
    if t.dim() not in (1, 2):
        raise ValueError("expected a waveform [num_samples] or a batch [B, num_samples], got shape %s"
                         % (tuple(t.shape),))
    batched = t.dim() == 2
    dtype = torch.int16 if t.dtype == torch.int16 else torch.float32
    if not t.is_cuda:
        t = t.to(device=torch.device("cuda", torch.cuda.current_device()), dtype=dtype, non_blocking=True)
    elif t.dtype != dtype:
        t = t.to(dtype)
    if not batched:
        t = t.unsqueeze(0)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t, batched, was_numpy


def _geometry(audio_sample_rate, window_length_secs, hop_length_secs):
    window = int(round(audio_sample_rate * window_length_secs))
    hop = int(round(audio_sample_rate * hop_length_secs))
    nfft = 2 ** int(np.ceil(np.log(window) / np.log(2.0)))
    return window, hop, nfft


def _run(entry, plan, wav, width):
    frames = plan.num_frames(wav.shape[1])
    if frames < 0:
        raise ValueError("negative dimensions are not allowed")
    out = torch.empty((wav.shape[0], frames, width), dtype=torch.float32, device=wav.device)
    with torch.cuda.device(wav.device):
        _cabi.check(entry(plan.handle, _cabi.ptr(wav), wav.shape[0], wav.shape[1], wav.stride(0),
                          _cabi.ptr(out), _cabi.stream_ptr(wav.device)))
    return out


def _finish(out, batched, was_numpy):
    if not batched:
        out = out[0]
    return out.cpu().numpy().astype(np.float64) if was_numpy else out


def stft_magnitude(signal, fft_length, hop_length=None, window_length=None):
    """|rfft(frames * periodic_hann, fft_length)| -> [frames, fft_length/2+1] (:71-92), on the GPU."""
    _cabi.require_cuda("stft_magnitude")
    wav, batched, was_numpy = _to_device(signal)
    if wav.dtype != torch.float32:
        wav = wav.to(torch.float32)
    plan = _get_plan(wav.device.index, int(window_length), int(hop_length), int(fft_length), 0.0,
                     dict(num_mel_bins=20, audio_sample_rate=8000))          # mel table unused by this entry
    out = _run(_cabi.lib().a2m_stft_magnitude_f32, plan, wav, plan.nfft // 2 + 1)
    return _finish(out, batched, was_numpy)


def log_mel_spectrogram(data, audio_sample_rate=8000, log_offset=0.0, window_length_secs=0.025,
                        hop_length_secs=0.010, **kwargs):
    """log(|STFT| . mel_matrix + log_offset) -> [frames, num_mel_bins] (:192-223), one fused launch."""
    _cabi.require_cuda("log_mel_spectrogram")
    window, hop, nfft = _geometry(audio_sample_rate, window_length_secs, hop_length_secs)
    mel_kwargs = dict(kwargs, audio_sample_rate=audio_sample_rate)
    mel_kwargs.pop("num_spectrogram_bins", None)
    wav, batched, was_numpy = _to_device(data)
    plan = _get_plan(wav.device.index, window, hop, nfft, log_offset, mel_kwargs)
    entry = _cabi.lib().a2m_logmel_i16 if wav.dtype == torch.int16 else _cabi.lib().a2m_logmel_f32
    out = _run(entry, plan, wav, plan.n_mel)
    return _finish(out, batched, was_numpy)
