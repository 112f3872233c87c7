"""B200 drop-in for the PATS-native audio front ends of the reference, ``pats/data_loading/audio.py`` (class ``Audio``,
methods ``log_mel_400`` :86-120 and ``log_mel_512`` :58-79) -- SURVEY.md section 8(f) rank 2.

``log_mel_400`` has the hot path's geometry (16 kHz, 400-sample periodic Hann window centred in a 512-point frame, hop
160, |STFT|, 64 mel bands 125-7500 Hz) with librosa's conventions instead of VGGish's: frames are 512 samples long
(``center=False``: ``1 + (N - 512) // 160`` frames, the window occupies samples 56..455 of each), the filterbank is
librosa's Slaney-scale triangles in Hz with ``norm=None``, and exact zeros are replaced by ``eps`` before the log.
It runs on the same fused CUDA kernel (csrc/logmel.cu) through ``a2m_mel_plan_create_ex`` with a zero-padded
window table and ``A2M_LOG_FLOOR_ZEROS``; the window and the filterbank are evaluated on the host in fp64 from
librosa's published formulas (librosa itself is not a dependency of this package).

``log_mel_512`` (the representation the shipped training configuration reads, ``audio/log_mel_512``) is librosa's
default mel spectrogram: 2048-point centred frames at hop 512, periodic Hann, POWER spectrum, 128 Slaney bands 0..sr/2
with area normalisation, zeros floored to ``eps``, log.  It runs on its own kernel (csrc/melspec_wide.cu, radix-4
Stockham FFT in shared memory).  librosa's padding default changed from 'reflect' (< 0.10, the reference's era) to
'constant' (>= 0.10): ``pad_mode`` selects, default 'reflect'.

Not implemented on the GPU (raises NotImplementedError, never a CPU fallback): resampling (``sr`` must already be
16 kHz for ``log_mel_400``).
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from .pose_video import mel_features

_plans = {}


def hz_to_mel(frequencies, htk=False):
    """librosa.hz_to_mel: Slaney (linear below 1 kHz, log above) or HTK."""
    f = np.asanyarray(frequencies, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz(mels, htk=False):
    """librosa.mel_to_hz."""
    m = np.asanyarray(mels, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm="slaney"):
    """librosa.filters.mel -> [n_mels, 1 + n_fft // 2] float32: triangles in Hz between mel-spaced centres;
    norm='slaney' scales each to unit area, None leaves the peaks at 1."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    fftfreqs = np.fft.rfftfreq(n_fft, 1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin, htk), hz_to_mel(fmax, htk), n_mels + 2), htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == "slaney":
        weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None].astype(np.float32)
    elif norm is not None:
        raise NotImplementedError("mel_filterbank: norm=%r" % (norm,))
    return weights


def centred_window(win_length, n_fft):
    """librosa.stft's window: periodic Hann of win_length (scipy get_window(..., fftbins=True)) zero-padded on both
    sides to n_fft (util.pad_center)."""
    w = mel_features.periodic_hann(win_length)
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad))


def _plan_400(device_index, eps):
    key = (device_index, "log_mel_400", float(eps))
    plan = _plans.get(key)
    if plan is None:
        window = np.ascontiguousarray(centred_window(400, 512), dtype=np.float64)
        weights = np.ascontiguousarray(mel_filterbank(16000, 512, n_mels=64, fmin=125.0, fmax=7500.0, norm=None).T,
                                       dtype=np.float64)                    # [257, 64]
        out = ctypes.c_void_p()
        _cabi.check(_cabi.lib().a2m_mel_plan_create_ex(
            512, 160, 512, 64, window.ctypes.data_as(ctypes.c_void_p), weights.ctypes.data_as(ctypes.c_void_p),
            float(eps), 1, device_index, ctypes.byref(out)))
        plan = _plans[key] = mel_features._Plan(out, 512, 160, 512, 64)
    return plan


_resamplers = {}


def resample(y, orig_sr, target_sr):
    """librosa.core.resample(y, orig_sr=..., target_sr=...) on the GPU (band-limited sinc interpolation, resampy's
    'kaiser_best' table; csrc/resample.cu): waveform [N] or batch [B, N] -> [ceil(N * target_sr / orig_sr)] fp32.
    numpy in -> numpy out, torch in -> torch (CUDA) out."""
    _cabi.require_cuda("resample")
    wav, batched, was_numpy = mel_features._to_device(y)
    if wav.dtype != torch.float32:
        wav = wav.to(torch.float32)
    if float(orig_sr) == float(target_sr):
        out = wav.clone()
    else:
        key = (wav.device.index, float(orig_sr), float(target_sr))
        plan = _resamplers.get(key)
        if plan is None:
            handle = ctypes.c_void_p()
            _cabi.check(_cabi.lib().a2m_resample_plan_create(float(orig_sr), float(target_sr), wav.device.index,
                                                             ctypes.byref(handle)))
            plan = _resamplers[key] = handle
        n_out = int(_cabi.lib().a2m_resample_out_length(plan, wav.shape[1]))
        out = torch.empty((wav.shape[0], n_out), dtype=torch.float32, device=wav.device)
        with torch.cuda.device(wav.device):
            _cabi.check(_cabi.lib().a2m_resample_f32(plan, _cabi.ptr(wav), wav.shape[0], wav.shape[1], wav.stride(0),
                                                     _cabi.ptr(out), _cabi.stream_ptr(wav.device)))
    if not batched:
        out = out[0]
    return out.cpu().numpy() if was_numpy else out


def log_mel_400(y, sr=16000, eps=1e-6):
    """pats/data_loading/audio.py:86-120 on the GPU: waveform [N] (or batch [B, N]) at 16 kHz ->
    log-mel [frames, 64] (``np.log(spec).transpose(1, 0)``), frames = 1 + (N - 512) // 160."""
    _cabi.require_cuda("log_mel_400")
    if not isinstance(y, torch.Tensor):
        y = np.asarray(y)
        if y.ndim == 2 and y.shape[0] != 1 and y.shape[1] == 1:
            y = y.reshape(-1)                       # the reference flattens with y.reshape((-1))
    wav, batched, was_numpy = mel_features._to_device(y)
    if float(sr) != 16000.0:                        # audio.py:87: resample to 16 kHz first
        wav = resample(wav if batched else wav[0], sr, 16000)
        if not batched:
            wav = wav.unsqueeze(0)
    if wav.shape[1] < 512:
        raise ValueError("log_mel_400: %d samples are fewer than one 512-sample frame" % wav.shape[1])
    plan = _plan_400(wav.device.index, eps)
    entry = _cabi.lib().a2m_logmel_i16 if wav.dtype == torch.int16 else _cabi.lib().a2m_logmel_f32
    out = mel_features._run(entry, plan, wav, 64)
    if not batched:
        out = out[0]
    return out.cpu().numpy() if was_numpy else out


_PAD_MODES = {None: 0, "none": 0, "reflect": 1, "constant": 2, "zeros": 2}


class _WidePlan:
    def __init__(self, handle, n_mel):
        self.handle, self.n_mel = handle, n_mel


def _plan_512(device_index, sr, eps, pad_mode):
    key = (device_index, "log_mel_512", float(sr), float(eps), pad_mode)
    plan = _plans.get(key)
    if plan is None:
        window = np.ascontiguousarray(centred_window(2048, 2048), dtype=np.float64)
        weights = np.ascontiguousarray(mel_filterbank(sr, 2048, n_mels=128).T, dtype=np.float64)      # [1025, 128]
        out = ctypes.c_void_p()
        _cabi.check(_cabi.lib().a2m_melspec_plan_create(
            2048, 512, 128, 2, _PAD_MODES[pad_mode], window.ctypes.data_as(ctypes.c_void_p),
            weights.ctypes.data_as(ctypes.c_void_p), float(eps), 1, device_index, ctypes.byref(out)))
        plan = _plans[key] = _WidePlan(out, 128)
    return plan


def log_mel_512(y, sr, eps=1e-10, pad_mode="reflect"):
    """pats/data_loading/audio.py:58-79 on the GPU: waveform [N] (or batch [B, N]) at ``sr`` Hz ->
    log-mel [1 + N // 512, 128]."""
    _cabi.require_cuda("log_mel_512")
    if pad_mode not in ("reflect", "constant", "zeros"):
        raise ValueError("log_mel_512: pad_mode must be 'reflect' or 'constant', got %r" % (pad_mode,))
    wav, batched, was_numpy = mel_features._to_device(y)
    if wav.dtype != torch.float32:
        wav = wav.to(torch.float32)                 # the 2048-point kernel reads fp32
    if pad_mode == "reflect" and wav.shape[1] <= 1024:
        raise ValueError("log_mel_512: reflect padding needs more than 1024 samples, got %d" % wav.shape[1])
    plan = _plan_512(wav.device.index, float(sr), eps, pad_mode)
    frames = int(_cabi.lib().a2m_melspec_num_frames(plan.handle, wav.shape[1]))
    out = torch.empty((wav.shape[0], frames, 128), dtype=torch.float32, device=wav.device)
    with torch.cuda.device(wav.device):
        _cabi.check(_cabi.lib().a2m_melspec_f32(plan.handle, _cabi.ptr(wav), wav.shape[0], wav.shape[1], wav.stride(0),
                                                _cabi.ptr(out), _cabi.stream_ptr(wav.device)))
    if not batched:
        out = out[0]
    return out.cpu().numpy() if was_numpy else out


class Audio:
    """The feature-extraction surface of the reference's ``Audio`` modality (the HDF5 / CSV bookkeeping of its
    constructor is out of scope): ``Audio().log_mel_400(y, sr)``, ``fs_map``."""

    def __init__(self, path2data=None, path2outdata=None, speaker="oliver", preprocess_methods=("log_mel_512",)):
        self.path2data, self.path2outdata, self.speaker = path2data, path2outdata, speaker
        self.preprocess_methods = list(preprocess_methods)

    def log_mel_400(self, y, sr, eps=1e-6):
        return log_mel_400(y, sr, eps)

    def log_mel_512(self, y, sr, eps=1e-10, pad_mode="reflect"):
        return log_mel_512(y, sr, eps, pad_mode)

    @property
    def fs_map(self):
        """Feature rates of the stored representations (pats/data_loading/audio.py:175-180)."""
        return {"log_mel_512": int(45.6 * 1000 / 512), "log_mel_400": int(16.52 * 1000 / 160), "silence": 15}

    def fs(self, modality):
        return self.fs_map[modality.split("/")[-1]]
