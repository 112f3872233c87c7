"""GPU parity: the fused log-mel kernel (through the Python drop-in -> C ABI) against the oracle and
the reference goldens.  Tolerance D8: max|a-b| <= 1e-4*max(1,|b|) and sum|a-b|/sum|b| <= 1e-4."""
import numpy as np
import pytest
import torch

from oracle import mel_oracle, synth
from oracle.make_golden import MEL_CASES, MEL_NFFT_CASES

pytestmark = pytest.mark.gpu


def assert_d8(got, ref):
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == ref.shape
    if ref.size == 0:
        return
    err = np.abs(got - ref)
    assert np.all(err <= 1e-4 * np.maximum(1.0, np.abs(ref))), float((err / np.maximum(1.0, np.abs(ref))).max())
    assert err.sum() / np.abs(ref).sum() <= 1e-4


@pytest.fixture(scope="module")
def mods(pkg):
    return pkg.install_dropin()


@pytest.mark.parametrize("name,kind,n,idx", MEL_CASES)
def test_logmel_vs_reference_golden(mods, golden, name, kind, n, idx):
    got = mods["pose_video.audio_repr"].log_mel_spectograms(synth.wav_clip(idx, n, kind))
    assert isinstance(got, np.ndarray) and got.dtype == np.float64
    assert_d8(got, golden["mel"][name])


def test_logmel_batch_matches_oracle(mods):
    wav = synth.wav_batch(100, 5)                                  # full-length clips
    ref = mel_oracle.log_mel_batch(wav)
    got = mods["pose_video.audio_repr"].log_mel_spectograms(torch.from_numpy(wav).cuda())
    assert got.is_cuda and got.dtype == torch.float32 and got.shape == (5, 425, 64)
    assert_d8(got.cpu().numpy(), ref)
    single = mods["pose_video.audio_repr"].log_mel_spectograms(wav[3])
    assert_d8(single, ref[3])


def test_logmel_misaligned_rows_and_strides(mods):
    big = torch.from_numpy(synth.wav_batch(7, 3, n=5001)).cuda()  # odd row length: rows not 16 B aligned
    ref = mel_oracle.log_mel_batch(big.cpu().numpy())
    assert_d8(mods["pose_video.audio_repr"].log_mel_spectograms(big).cpu().numpy(), ref)
    view = big[:, 3:4003]                                          # strided rows, offset start
    ref_v = mel_oracle.log_mel_batch(view.cpu().numpy())
    assert_d8(mods["pose_video.audio_repr"].log_mel_spectograms(view).cpu().numpy(), ref_v)


def test_logmel_other_parameters(mods):
    lm = mods["pose_video.mel_features"].log_mel_spectrogram
    wav = synth.wav_clip(21, 6000)
    kw = dict(audio_sample_rate=16000, log_offset=1e-3, window_length_secs=0.030, hop_length_secs=0.0125,
              num_mel_bins=40, lower_edge_hertz=60.0, upper_edge_hertz=7000.0)     # window 480, hop 200
    assert_d8(lm(wav, **kw), mel_oracle.log_mel(wav, **kw))
    kw = dict(audio_sample_rate=22050, log_offset=0.1, window_length_secs=0.020, hop_length_secs=0.005,
              num_mel_bins=128, lower_edge_hertz=20.0, upper_edge_hertz=10000.0)   # window 441, hop 110
    assert_d8(lm(wav, **kw), mel_oracle.log_mel(wav, **kw))


def test_logmel_edge_cases(mods):
    lm = mods["pose_video.audio_repr"].log_mel_spectograms
    z = lm(np.zeros(1000, np.float32))
    assert z.shape == (4, 64) and np.allclose(z, np.log(0.01), atol=1e-6)
    assert lm(synth.wav_clip(6, 399)).shape == (0, 64)
    assert lm(np.zeros((0, 1000), np.float32)).shape == (0, 4, 64)
    with pytest.raises(ValueError):
        lm(np.zeros(100, np.float32))
    with pytest.raises(ValueError):
        mods["pose_video.mel_features"].log_mel_spectrogram(np.zeros(1000), audio_sample_rate=16000,
                                                           upper_edge_hertz=9000.0)
    with pytest.raises(NotImplementedError):                                      # a 25 ms window at 200 kHz: 5000 samples
        mods["pose_video.mel_features"].log_mel_spectrogram(np.zeros(20000), audio_sample_rate=200000,
                                                           upper_edge_hertz=3800.0)
    int16 = (synth.wav_clip(3, 4000, "int16")).astype(np.int16)
    assert_d8(lm(int16), mel_oracle.log_mel_audio_repr(int16))


def test_logmel_default_parameters_golden(mods, golden):
    """The function's own defaults (8 kHz -> window 200, hop 80, fft length 256, 20 bands, 125-3800 Hz): the reference
    golden `default_params_4000`."""
    got = mods["pose_video.mel_features"].log_mel_spectrogram(synth.wav_clip(10, 4000), log_offset=1e-3)
    assert got.shape == (48, 20) and got.dtype == np.float64
    assert_d8(got, golden["mel"]["default_params_4000"])


@pytest.mark.parametrize("name,idx,n,kind,kw", MEL_NFFT_CASES)
def test_logmel_other_fft_lengths_golden(mods, golden, name, idx, n, kind, kw):
    """fft lengths 128 / 256 / 1024 / 2048 against the unmodified reference (mel_features.py:212-214)."""
    wav = synth.wav_clip(idx, n, kind)
    if kind == "int16":
        wav = wav.astype(np.int16)
    ref = golden["melnfft"][name]
    got = mods["pose_video.mel_features"].log_mel_spectrogram(wav, **kw)
    assert_d8(got, ref)
    both = mods["pose_video.mel_features"].log_mel_spectrogram(torch.from_numpy(np.stack([wav, wav])).cuda(), **kw)
    assert both.shape == (2,) + ref.shape and torch.equal(both[0], both[1])
    assert_d8(both[1].cpu().numpy(), ref)


def test_logmel_int16_pcm_path(mods, golden):
    """int16 PCM is consumed as int16 by the kernel (a2m_logmel_i16): golden `int16_8000`, odd / even row alignment,
    a strided view, and equality with the fp32 path on the same (exactly representable) samples to fp32 rounding."""
    lm = mods["pose_video.audio_repr"].log_mel_spectograms
    pcm = synth.wav_clip(3, 8000, "int16").astype(np.int16)
    assert_d8(lm(pcm), golden["mel"]["int16_8000"])
    batch = np.stack([synth.wav_clip(40 + i, 5001, "int16").astype(np.int16) for i in range(5)])      # odd rows: 2-byte aligned only
    ref = mel_oracle.log_mel_batch(batch)
    t = torch.from_numpy(batch).cuda()
    got = lm(t)
    assert got.dtype == torch.float32 and got.shape == ref.shape
    assert_d8(got.cpu().numpy(), ref)
    assert_d8(lm(t[:, 3:4504]).cpu().numpy(), mel_oracle.log_mel_batch(batch[:, 3:4504]))
    as_f32 = lm(t.float())
    assert torch.allclose(got, as_f32, rtol=0, atol=2e-5)
    full = torch.from_numpy(np.stack([synth.wav_clip(50 + i, synth.CLIP_SAMPLES, "int16").astype(np.int16) for i in range(3)])).cuda()
    assert_d8(lm(full).cpu().numpy(), mel_oracle.log_mel_batch(full.cpu().numpy()))


def test_stft_magnitude_1024(mods, golden):
    got = mods["pose_video.mel_features"].stft_magnitude(synth.wav_clip(35, 5000), fft_length=1024, hop_length=300,
                                                        window_length=700)
    ref = golden["melnfft"]["stft_mag_1024"]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_stft_magnitude(mods, golden):
    got = mods["pose_video.mel_features"].stft_magnitude(synth.wav_clip(9, 2000), fft_length=512, hop_length=160,
                                                        window_length=400)
    ref = golden["mel"]["stft_mag_2000"]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_logmel_full_size_properties(mods):
    """BASELINE config-2 size (256 full clips): every clip equals its single-clip result bit for bit
    (batch invariance), and scaling the input by 2 shifts exp(logmel)-offset by exactly 2x (linearity
    of |STFT|.mel in fp32 for a power-of-two scale)."""
    lm = mods["pose_video.audio_repr"].log_mel_spectograms
    wav = torch.from_numpy(synth.wav_batch(1000, 8)).cuda().repeat(32, 1)          # [256, 68267]
    out = lm(wav)
    assert out.shape == (256, 425, 64)
    assert torch.equal(out[:8], out[248:])
    one = lm(wav[5])
    assert torch.equal(one, out[5])
    out2 = lm(2.0 * wav[:8])
    a = torch.exp(out[:8].double()) - 0.01
    b = torch.exp(out2.double()) - 0.01
    assert torch.allclose(b, 2 * a, rtol=2e-5, atol=1e-7)
