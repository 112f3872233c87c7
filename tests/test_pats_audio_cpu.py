"""CPU: the PATS-native front end (pats/data_loading/audio.py log_mel_400) -- host tables of the drop-in against the
oracle's independent restatement of librosa's published formulas, and the known answers of those formulas.  librosa
is absent and unpinned in the reference: parity is UNPINNED at that boundary (oracle/pats_oracle.py header)."""
import importlib

import numpy as np
import pytest

from oracle import mel_oracle, pats_oracle, synth

PKG = "audio-to-motion-generation_b200"


@pytest.fixture(scope="module")
def pa():
    return importlib.import_module(PKG + ".pats_audio")


def test_slaney_scale_known_answers(pa):
    # librosa documentation values: hz_to_mel(60) = 0.9, hz_to_mel([110, 220, 440]) = [1.65, 3.3, 6.6]; 1 kHz = mel 15
    np.testing.assert_allclose(pa.hz_to_mel([60, 110, 220, 440, 1000]), [0.9, 1.65, 3.3, 6.6, 15.0], rtol=1e-12)
    np.testing.assert_allclose(pa.mel_to_hz([1, 2, 3]), [200 / 3, 400 / 3, 200.0], rtol=1e-12)
    np.testing.assert_allclose(pa.mel_to_hz(pa.hz_to_mel([125.0, 999.0, 1000.0, 7500.0])), [125.0, 999.0, 1000.0, 7500.0], rtol=1e-12)
    np.testing.assert_allclose(pa.hz_to_mel(6400.0), 15.0 + 27.0, rtol=1e-12)              # one 6.4x step = 27 mels
    np.testing.assert_allclose(pa.hz_to_mel([300.0, 4000.0], htk=True), 2595.0 * np.log10(1 + np.array([300.0, 4000.0]) / 700.0))
    np.testing.assert_allclose(pats_oracle.slaney_mel([60, 1000, 6400]), pa.hz_to_mel([60, 1000, 6400]), rtol=1e-12)


def test_filterbank_matches_oracle_and_shape_rules(pa):
    w = pa.mel_filterbank(16000, 512, n_mels=64, fmin=125.0, fmax=7500.0, norm=None)
    assert w.shape == (64, 257) and w.dtype == np.float32
    np.testing.assert_allclose(w, pats_oracle.filterbank(16000, 512, 64, 125.0, 7500.0, False), rtol=0, atol=1e-6)
    assert w.min() >= 0 and w.max() <= 1.0
    for row in w:                                           # one contiguous run per band (what the kernel's plan needs)
        nz = np.nonzero(row)[0]
        assert nz.size >= 1 and np.all(np.diff(nz) == 1) and nz.max() < 256
    assert np.all(w[:, :4] == 0)                            # nothing below 125 Hz (bin width 31.25 Hz)
    ws = pa.mel_filterbank(22050, 2048, n_mels=128)         # librosa defaults: Slaney area normalisation
    np.testing.assert_allclose(ws, pats_oracle.filterbank(22050, 2048, 128, 0.0, 11025.0, True), rtol=0, atol=1e-7)
    with pytest.raises(NotImplementedError):
        pa.mel_filterbank(16000, 512, norm=1)


def test_centred_window(pa):
    w = pa.centred_window(400, 512)
    assert w.shape == (512,) and np.all(w[:56] == 0) and np.all(w[456:] == 0)
    np.testing.assert_array_equal(w[56:456], mel_oracle.hann(400))


def test_oracle_relation_to_the_vggish_path():
    """A 512-sample librosa frame at t*160 windows samples [t*160+56, t*160+456): its magnitudes equal the VGGish
    path's (frames of 400 at hop 160) on the waveform advanced by 56 samples."""
    y = synth.wav_clip(3, 4000)
    a = pats_oracle.stft_mag_uncentred(y, 512, 160, 400)
    b = mel_oracle.stft_mag(y[56:], 512, 160, 400)
    assert a.shape[0] == 1 + (4000 - 512) // 160
    np.testing.assert_allclose(a, b[:a.shape[0]], rtol=0, atol=1e-9)
    z = pats_oracle.log_mel_400(np.zeros(1000, np.float32))
    assert z.shape == (4, 64) and np.all(z == np.log(1e-6))


def test_no_cpu_fallback_and_unsupported(pa):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        pa.log_mel_400(np.zeros(2000, np.float32), 16000)
    with pytest.raises(RuntimeError):
        pa.log_mel_512(np.zeros(4096, np.float32), 44100)
    assert pa.Audio().fs("audio/log_mel_400") == 103 and pa.Audio().fs_map["log_mel_512"] == 89


def test_log_mel_512_oracle_known_answers():
    """Frame count 1 + N // 512; reflect padding makes frame 0 symmetric about sample 0; silence -> log(eps); a pure
    tone lands in the band whose centre is nearest."""
    sr = 44100
    y = synth.wav_clip(11, 6000)
    ref = pats_oracle.log_mel_512(y, sr)
    assert ref.shape == (1 + 6000 // 512, 128)
    assert np.all(pats_oracle.log_mel_512(np.zeros(3000), sr) == np.log(1e-10))
    t = np.arange(20000) / sr
    tone = np.sin(2 * np.pi * 2000.0 * t)
    band = pats_oracle.log_mel_512(tone, sr)[10].argmax()
    centres = pats_oracle.slaney_hz(np.linspace(0, pats_oracle.slaney_mel(sr / 2), 130))[1:-1]
    assert abs(centres[band] - 2000.0) == np.abs(centres - 2000.0).min()
    zc = pats_oracle.stft_power_centred(y, 2048, 512, "constant")
    zr = pats_oracle.stft_power_centred(y, 2048, 512, "reflect")
    np.testing.assert_allclose(zc[3:-3], zr[3:-3], rtol=1e-12)          # interior frames never see the padding
    assert not np.allclose(zc[0], zr[0])


def test_resample_oracle_known_answers():
    """The restated resampler: output length ceil(n * ratio); a sine below both Nyquist rates survives a 44.1 -> 16 kHz
    conversion to within the filter's gain error (resampy's integer table step: < 0.5 %); equal rates are the identity;
    the Kaiser-windowed sinc table starts at the roll-off and ends at zero."""
    win, bits = pats_oracle.sinc_table(**pats_oracle.KAISER_BEST)
    assert bits == 512 and win.size == 64 * 512 + 1 and abs(win[0] - 0.9475937167399596) < 1e-15 and abs(win[-1]) < 1e-7
    n, sr = 5000, 44100
    t = np.arange(n) / sr
    y = np.sin(2 * np.pi * 440 * t)
    out = pats_oracle.resample(y, sr, 16000)
    assert out.size == int(np.ceil(n * 16000 / sr))
    ref = np.sin(2 * np.pi * 440 * np.arange(out.size) / 16000)
    assert np.abs(out[200:-200] - ref[200:-200]).max() < 5e-3
    assert np.array_equal(pats_oracle.resample(y, sr, sr), y)
    assert pats_oracle.resample(y[:1000], 8000, 16000).size == 2000
