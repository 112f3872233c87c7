"""CPU: the HDF5-free clip container and the window gatherer against the reference's index arithmetic
(pats/data_loading/dataUtils.py:585-620, 648-654), restated inline from the shipped configuration."""
import importlib

import numpy as np
import pytest

PKG = "audio-to-motion-generation_b200"


@pytest.fixture(scope="module")
def cs():
    return importlib.import_module(PKG + ".clip_store")


def reference_windows(data, fs, fs_new, time, window_hop):
    """MiniData.update_idx_list + __getitem__ for one modality, line by line."""
    window = int(time * fs)
    fs_ratio = round(fs / fs_new)
    if not window_hop:
        splits = np.r_[range(0, data.shape[0] - window, int(window))]
    else:
        splits = np.r_[range(0, data.shape[0] - window, int(window_hop * fs_ratio))]
    return [data[s:e:fs_ratio] for s, e in zip(splits, splits + window)]


def test_round_trip_and_alignment(cs, tmp_path):
    rng = np.random.default_rng(0)
    arrays = {"pose/data": rng.normal(size=(148, 104)).astype(np.float32),
              "audio/log_mel_512": rng.normal(size=(849, 128)).astype(np.float32),
              "audio/pcm": rng.integers(-3000, 3000, 70001).astype(np.int16),
              "meta/empty": np.zeros((0, 7), np.float64)}
    path = cs.write_clip_store(str(tmp_path / "interval.a2mclip"), arrays)
    store = cs.ClipStore(path)
    assert store.keys() == list(arrays)
    for k, a in arrays.items():
        got = store[k]
        assert got.shape == a.shape and got.dtype == a.dtype and np.array_equal(np.asarray(got), a)
        if a.size:
            assert got.offset % 256 == 0
    with pytest.raises(ValueError):
        open(tmp_path / "junk", "wb").write(b"not a clip store at all")
        cs.ClipStore(str(tmp_path / "junk"))
    with pytest.raises(TypeError):
        cs.write_clip_store(str(tmp_path / "bad"), {"x": np.zeros(3, np.complex64)})


def test_windows_follow_the_reference_arithmetic(cs, tmp_path):
    """Shipped configuration (version5_model_train.py:200-205): time 4.3 s, fs_new 15, window_hop 5; poses at 15 fps ->
    window 64, ratio 1, hop 5; log_mel_512 at 89 fps -> window 382, ratio 6 (64 steps), hop 30."""
    rng = np.random.default_rng(1)
    pose = rng.normal(size=(300, 104)).astype(np.float32)
    mel = rng.normal(size=(1786, 128)).astype(np.float32)
    path = cs.write_clip_store(str(tmp_path / "i.a2mclip"), {"pose/data": pose, "audio/log_mel_512": mel})
    w = cs.ClipWindows(path, ["pose/data", "audio/log_mel_512"], [15, 15], time=4.3, window_hop=5)
    ref_p = reference_windows(pose, 15, 15, 4.3, 5)
    ref_m = reference_windows(mel, 89, 15, 4.3, 5)
    assert cs.FS_MAP["audio/log_mel_512"] == 89 and cs.window_index(1786, 89, 15, 4.3, 5)[1:] == (382, 6)
    assert len(w) == min(len(ref_p), len(ref_m)) and len(w) > 40
    for i in (0, 1, len(w) // 2, len(w) - 1):
        item = w[i]
        assert np.array_equal(item["pose/data"], ref_p[i]) and item["pose/data"].shape == (64, 104)
        assert np.array_equal(item["audio/log_mel_512"], ref_m[i]) and item["audio/log_mel_512"].shape == (64, 128)
    vp, vm = w.views("pose/data"), w.views("audio/log_mel_512")
    assert vp.shape == (len(w), 64, 104) and vm.shape == (len(w), 64, 128)
    assert np.array_equal(vm, np.stack(ref_m[:len(w)])) and np.array_equal(vp, np.stack(ref_p[:len(w)]))
    assert not vm.flags.owndata                                  # a view, nothing gathered
    # no hop: back-to-back windows (dataUtils.py:611-613)
    w0 = cs.ClipWindows(path, ["pose/data"], [15], time=4.3, window_hop=0)
    ref0 = reference_windows(pose, 15, 15, 4.3, 0)
    assert len(w0) == len(ref0) and np.array_equal(w0[len(w0) - 1]["pose/data"], ref0[-1])
    with pytest.raises(AssertionError):
        cs.window_index(300, 15, 15, 4.3, 64)                    # hop must be below the window
    with pytest.raises(IndexError):
        w[len(w)]
