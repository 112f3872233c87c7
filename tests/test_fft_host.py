"""CPU: the FFT factorisation used by the log-mel kernel (csrc/fft_math.cuh + the lane/partner
index algebra of csrc/logmel.cu) emulated on the host and checked against the oracle."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import mel_oracle, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "fft_host_test.cu")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("fft") / "fft_host_test")
    subprocess.run([nvcc, "-O1", "-o", exe, SRC], check=True, capture_output=True)
    return exe


@pytest.mark.parametrize("kind,n", [("noise", 4000), ("tone", 8000), ("int16", 4000)])
def test_host_fft_matches_oracle(harness, tmp_path, kind, n):
    wav = synth.wav_clip(11, n, kind).astype(np.float32)
    hann = mel_oracle.hann(400)
    wav.tofile(tmp_path / "wav.f32")
    hann.tofile(tmp_path / "hann.f64")
    subprocess.run([harness, str(tmp_path / "wav.f32"), str(n), "400", "160", str(tmp_path / "hann.f64"),
                    str(tmp_path / "out.f32")], check=True)
    ref = mel_oracle.stft_mag(wav, 512, 160, 400)
    got = np.fromfile(tmp_path / "out.f32", dtype=np.float32).reshape(ref.shape)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-6 * scale, (np.abs(got - ref).max(), scale)
    # and through the mel projection + log: the D8 tolerance of the product test
    w = mel_oracle.mel_matrix(64, 257, 16000, 125, 7500)
    lm_ref = np.log(ref @ w + 0.01)
    lm_got = np.log(got.astype(np.float64) @ w + 0.01)
    assert np.all(np.abs(lm_got - lm_ref) <= 1e-4 * np.maximum(1.0, np.abs(lm_ref)))


def test_host_fft2048_index_algebra(tmp_path):
    """melspec2048_kernel's 16 x 16 x 4 split, exchange layouts and untangle (tests/host/fft2048_host_test.cu emulates
    the kernel's data flow with the same fft_math.cuh functions) against a direct fp64 DFT."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fft2048_host_test")
    subprocess.run([nvcc, "-O1", "-o", exe, os.path.join(ROOT, "tests", "host", "fft2048_host_test.cu")], check=True,
                   capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout
