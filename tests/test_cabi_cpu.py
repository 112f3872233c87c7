"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/a2m_b200.h
declares; host-side logic of the drop-in modules (no compute calls)."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest
import torch

from oracle import mel_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "a2m_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(a2m_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    cabi = importlib.import_module(pkg.__name__ + "._cabi")
    lib = pkg.load_library()
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), "liba2m_b200.so does not export %s" % name
        assert name in cabi.SIGNATURES, "no ctypes signature for %s" % name
    assert set(cabi.SIGNATURES) == set(names)
    assert lib.a2m_version() == 100
    assert ctypes.sizeof(cabi.Metrics) == 64


def test_no_cpu_fallback(pkg):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pkg.install_dropin()
    from pose_video.mel_features import log_mel_spectrogram
    from motion_evaluation import compute_pck
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        log_mel_spectrogram(np.zeros(1000, np.float32), audio_sample_rate=16000)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        compute_pck(np.zeros((1, 2, 52), np.float32), np.zeros((1, 2, 52), np.float32))


def test_argument_errors_without_gpu(pkg):
    lib = pkg.load_library()
    out = ctypes.c_void_p()
    hann = np.ones(400)
    w = np.zeros((257, 64))
    rc = lib.a2m_mel_plan_create(400, 160, 768, 64, hann.ctypes.data_as(ctypes.c_void_p),      # not a power of two
                                 w.ctypes.data_as(ctypes.c_void_p), 0.01, 0, ctypes.byref(out))
    assert rc == -3 and b"not supported" in lib.a2m_last_error()
    rc = lib.a2m_mel_plan_create(600, 160, 512, 64, hann.ctypes.data_as(ctypes.c_void_p),
                                 w.ctypes.data_as(ctypes.c_void_p), 0.01, 0, ctypes.byref(out))
    assert rc == -1 and b"window" in lib.a2m_last_error()
    assert lib.a2m_eval_l1_pck_f32(None, None, 4, 64, 0.2, None, None, None, None) == -1
    # entry points added for the "next" rows reject bad arguments before touching the device
    rc = lib.a2m_mel_plan_create_ex(400, 160, 512, 64, hann.ctypes.data_as(ctypes.c_void_p),
                                    w.ctypes.data_as(ctypes.c_void_p), 0.01, 7, 0, ctypes.byref(out))
    assert rc == -1 and b"log_mode" in lib.a2m_last_error()
    win = np.ones(2048)
    w2 = np.zeros((1025, 128))
    rc = lib.a2m_melspec_plan_create(1024, 512, 128, 2, 1, win.ctypes.data_as(ctypes.c_void_p),
                                     w2.ctypes.data_as(ctypes.c_void_p), 1e-10, 1, 0, ctypes.byref(out))
    assert rc == -3 and b"not supported" in lib.a2m_last_error()
    rc = lib.a2m_melspec_plan_create(2048, 512, 128, 3, 1, win.ctypes.data_as(ctypes.c_void_p),
                                     w2.ctypes.data_as(ctypes.c_void_p), 1e-10, 1, 0, ctypes.byref(out))
    assert rc == -1 and b"power" in lib.a2m_last_error()
    w2[10, 0] = w2[12, 0] = 1.0                            # a band whose support is not one contiguous run of bins
    rc = lib.a2m_melspec_plan_create(2048, 512, 128, 2, 1, win.ctypes.data_as(ctypes.c_void_p),
                                     w2.ctypes.data_as(ctypes.c_void_p), 1e-10, 1, 0, ctypes.byref(out))
    assert rc == -3 and b"contiguous" in lib.a2m_last_error()
    assert lib.a2m_melspec_num_frames(None, 4096) == -1
    assert lib.a2m_motion_smoothness_f32(None, 4, 64, 200, 0, None, None) == -1      # more than 128 features
    assert lib.a2m_motion_smoothness_f32(None, 4, 64, 104, 0, None, None) == -1      # no accumulator
    assert lib.a2m_model_set_output_denorm(None, None, None, None) == -1
    assert lib.a2m_model_timeline_begin(None, 4, 64, 64, 2) == -1


def test_host_tables_match_oracle(pkg):
    mf = pkg.install_dropin()["pose_video.mel_features"]
    np.testing.assert_array_equal(mf.periodic_hann(400), mel_oracle.hann(400))
    np.testing.assert_array_equal(mf.spectrogram_to_mel_matrix(64, 257, 16000, 125, 7500),
                                  mel_oracle.mel_matrix(64, 257, 16000, 125, 7500))
    np.testing.assert_array_equal(mf.spectrogram_to_mel_matrix(), mel_oracle.mel_matrix())
    x = np.arange(1000.0)
    np.testing.assert_array_equal(mf.frame(x, 400, 160), mel_oracle.frames(x, 400, 160))
    assert mf.frame(np.zeros(399), 400, 160).shape == (0, 400)
    with pytest.raises(ValueError):
        mf.frame(np.zeros(100), 400, 160)
    with pytest.raises(ValueError):
        mf.spectrogram_to_mel_matrix(64, 257, 16000, 125, 9000)
    assert mf._geometry(16000, 0.025, 0.010) == (400, 160, 512)
    assert mf.hertz_to_mel(700.0) == pytest.approx(1127.0 * np.log(2.0))


def test_audio_repr_registry(pkg):
    ar = pkg.install_dropin()["pose_video.audio_repr"]
    assert ar.get_repr("log_mel_spect") is ar.log_mel_spectograms
    assert ar.get_repr(ar.RAW) is ar.raw_repr and ar.SR == 16000
    with pytest.raises(KeyError):
        ar.get_repr("nope")


def test_mel_schedule_reproduces_the_filterbank_and_is_conflict_free(pkg):
    """The plan-time schedule of the 512-point kernel's mel sum (a2m_mel_schedule_host, no GPU): every weight of the
    reference's filterbank (and of the Slaney one of log_mel_400) is visited exactly once -- band c = R[c] + F[c + 1]
    reproduces 0.5 * mag . W to fp32 rounding --, the 16 lanes of a step read bins in 16 different banks, and a
    matrix that is not a triangular filterbank is refused (the general kernel takes it)."""
    import importlib
    from oracle import mel_oracle
    lib = pkg.load_library()
    pa = importlib.import_module(pkg.__name__ + ".pats_audio")

    def schedule(w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        cap = 256
        b, uv, rs = np.zeros(cap * 16, np.int32), np.zeros(cap * 32, np.float32), np.zeros(12, np.int32)
        nr, dl = ctypes.c_int(), ctypes.c_int()
        n = lib.a2m_mel_schedule_host(w.ctypes.data_as(ctypes.c_void_p), w.shape[1], cap, b.ctypes.data_as(ctypes.c_void_p),
                                      uv.ctypes.data_as(ctypes.c_void_p), rs.ctypes.data_as(ctypes.c_void_p),
                                      ctypes.byref(nr), ctypes.byref(dl))
        if n < 0:
            return n, None, None, None, None
        return n, b[:n * 16].reshape(n, 16), uv[:n * 32].reshape(n, 16, 2), rs[:nr.value], dl.value

    cases = [mel_oracle.mel_matrix(64, 257, 16000, 125, 7500), mel_oracle.mel_matrix(40, 257, 16000, 60, 7000),
             mel_oracle.mel_matrix(128, 257, 22050, 20, 10000), mel_oracle.mel_matrix(20, 257, 16000, 125, 3800),
             pa.mel_filterbank(16000, 512, n_mels=64, fmin=125.0, fmax=7500.0, norm=None).T.astype(np.float64)]
    rng = np.random.default_rng(0)
    for w in cases:
        n_mel = w.shape[1]
        n, b, uv, rs, dist = schedule(w)
        assert n == rs.sum() and np.all(rs % 2 == 0)
        for t in range(n):
            banks = [k % 16 for k in b[t]]                                  # idle slots read a zero entry of a free bank
            assert len(set(banks)) == 16 and np.all((b[t] <= 256) | (b[t] >= 272)), (t, b[t])
        visited = b[b <= 256]
        assert len(set(visited.tolist())) == len(visited) == np.count_nonzero(w.any(axis=1))
        mag = np.append(rng.random(257), np.zeros(31))
        r_sum, f_sum = np.zeros(n_mel + 17), np.zeros(n_mel + 17)
        step = 0
        for r, nt in enumerate(rs):
            for _ in range(nt):
                for lane in range(16):
                    g = n_mel if (dist and r == len(rs) - 1) else 16 * r + lane
                    r_sum[g] += mag[b[step, lane]] * uv[step, lane, 0]
                    f_sum[g] += mag[b[step, lane]] * uv[step, lane, 1]
                step += 1
        got = r_sum[:n_mel] + f_sum[1:n_mel + 1]
        np.testing.assert_allclose(got, 0.5 * (mag[:257] @ w), rtol=0, atol=2e-7)
    dense = np.zeros((257, 8))
    dense[10:20, 0] = dense[10:20, 1] = dense[10:20, 2] = 1.0            # a bin feeding three bands
    assert schedule(dense)[0] == -3 and b"triangular" in lib.a2m_last_error()
