"""GPU parity: the fused log-mel kernel (through the Python drop-in -> C ABI) against the oracle and
the reference goldens.  Tolerance D8: max|a-b| <= 1e-4*max(1,|b|) and sum|a-b|/sum|b| <= 1e-4."""
import numpy as np
import pytest
import torch

from oracle import mel_oracle, synth
from oracle.make_golden import MEL_CASES

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
    with pytest.raises(NotImplementedError):
        mods["pose_video.mel_features"].log_mel_spectrogram(np.zeros(1000))       # 8 kHz default -> nfft 256
    int16 = (synth.wav_clip(3, 4000, "int16")).astype(np.int16)
    assert_d8(lm(int16), mel_oracle.log_mel_audio_repr(int16))


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
