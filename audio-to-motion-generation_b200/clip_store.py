"""HDF5-free on-disk clip format and the reference's window gatherer.

The reference keeps one HDF5 file per interval (``pats/data_loading/dataUtils.py:MiniData``: datasets such as
``pose/data [n, 104]`` and ``audio/log_mel_512 [m, 128]``), loads each dataset whole and serves sliding windows by index
arithmetic (:585-620) and a strided slice (:648-654).  h5py is not a dependency of this package, and an HDF5 chunk
cannot be handed to ``cudaMemcpy``; this module defines a flat container that can:

    magic "A2MCLIP1" | u32 n_entries | u32 header bytes | entries: (u16 name length, name, u8 dtype code, u8 ndim,
    u64 shape[ndim], u64 byte offset) ... | raw little-endian C-order arrays, each starting on a 256-byte boundary

so every array is an ``np.memmap`` slice that goes to pinned memory / the GPU without parsing.  ``ClipWindows``
restates the reference's window arithmetic on top of it and serves the windows of all modalities of one interval --
as strided VIEWS (host: numpy, device: torch) the way the native generator consumes them (no gather kernel is needed:
the first encoder kernel reads a strided slice in place), or materialised like ``MiniData.__getitem__`` does.
"""
import os
import struct

import numpy as np

MAGIC = b"A2MCLIP1"
_DTYPES = {0: np.dtype("<f4"), 1: np.dtype("<f8"), 2: np.dtype("<i2"), 3: np.dtype("<i4"), 4: np.dtype("<i8"),
           5: np.dtype("u1")}
_CODES = {v: k for k, v in _DTYPES.items()}
ALIGN = 256

# sampling rates of the reference's modalities (pats/data_loading/audio.py:175-181 fs_map, skeleton.py: 15 fps poses)
FS_MAP = {"pose/data": 15, "pose/normalize": 15, "audio/log_mel_512": int(45.6 * 1000 / 512),
          "audio/log_mel_400": int(16.52 * 1000 / 160), "audio/silence": 15}


def write_clip_store(path, arrays):
    """arrays: {name: ndarray} (e.g. {"pose/data": [n, 104] f32, "audio/log_mel_512": [m, 128] f32}) -> file."""
    entries, blobs = [], []
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
        if np.dtype(dt) not in _CODES:
            raise TypeError("clip store: unsupported dtype %s for %r" % (a.dtype, name))
        entries.append((name.encode("utf-8"), _CODES[np.dtype(dt)], a.shape))
        blobs.append(a.astype(dt, copy=False))
    header = len(MAGIC) + 8 + sum(2 + len(n) + 2 + 8 * len(shape) + 8 for n, _, shape in entries)
    offset = -(-header // ALIGN) * ALIGN
    table, offsets = b"", []
    for (n, code, shape), blob in zip(entries, blobs):
        offsets.append(offset)
        table += struct.pack("<H", len(n)) + n + struct.pack("<BB", code, len(shape))
        table += struct.pack("<%dQ" % len(shape), *shape) + struct.pack("<Q", offset)
        offset = -(-(offset + blob.nbytes) // ALIGN) * ALIGN
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", len(entries), header) + table)
        for off, blob in zip(offsets, blobs):
            f.seek(off)
            f.write(blob.tobytes())
        f.truncate(max(offset, f.tell()))
    return path


class ClipStore:
    """Read-only view of a clip file: ``store[name]`` is an np.memmap of the array (no copy, no parsing)."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(len(MAGIC) + 8)
            if head[:len(MAGIC)] != MAGIC:
                raise ValueError("%s is not a clip store (bad magic)" % path)
            n, header = struct.unpack("<II", head[len(MAGIC):])
            table = f.read(header - len(head))
        self._entries, pos = {}, 0
        for _ in range(n):
            (ln,) = struct.unpack_from("<H", table, pos); pos += 2
            name = table[pos:pos + ln].decode("utf-8"); pos += ln
            code, ndim = struct.unpack_from("<BB", table, pos); pos += 2
            shape = struct.unpack_from("<%dQ" % ndim, table, pos); pos += 8 * ndim
            (off,) = struct.unpack_from("<Q", table, pos); pos += 8
            self._entries[name] = (_DTYPES[code], tuple(int(s) for s in shape), int(off))
        size = os.path.getsize(path)
        for name, (dt, shape, off) in self._entries.items():
            if off % ALIGN or off + int(np.prod(shape, dtype=np.int64)) * dt.itemsize > size:
                raise ValueError("%s: entry %r lies outside the file" % (path, name))

    def keys(self):
        return list(self._entries)

    def __contains__(self, name):
        return name in self._entries

    def shape(self, name):
        return self._entries[name][1]

    def __getitem__(self, name):
        dt, shape, off = self._entries[name]
        if int(np.prod(shape, dtype=np.int64)) == 0:
            return np.empty(shape, dtype=dt)
        return np.memmap(self.path, dtype=dt, mode="r", offset=off, shape=shape)


def window_index(length, fs, fs_new, time, window_hop=0):
    """The reference's index arithmetic for one modality (MiniData.update_idx_list, dataUtils.py:585-620):
    window = int(time * fs); fs_ratio = round(fs / fs_new); starts = range(0, length - window, window_hop * fs_ratio or
    window).  Returns (starts ndarray, window, fs_ratio); item i is data[starts[i] : starts[i] + window : fs_ratio]."""
    window = int(time * fs)
    if not window_hop < window:
        raise AssertionError("hop size {} must be less than window size {}".format(window_hop, window))
    fs_ratio = round(fs / fs_new)
    step = int(window) if not window_hop else int(window_hop * fs_ratio)
    return np.r_[range(0, length - window, step)].astype(np.int64), window, fs_ratio


class ClipWindows:
    """Sliding windows over the modalities of one interval (the role of MiniData): ``len(w)`` is the minimum window
    count over the modalities (dataUtils.py:623-624); ``w[i]`` returns {modality: data[start:end:fs_ratio]} like
    ``MiniData.__getitem__`` (:637-654, float32); ``w.views(modality)`` returns ALL windows of a modality as one strided
    view [n_windows, steps, features] without copying -- numpy on the host, or torch on ``device`` after one upload of
    the whole array (what ``SelfAttention_G.forward`` / ``forward_windows`` read in place)."""

    def __init__(self, store, modalities, fs_new, time, window_hop=0, fs_map=None):
        self.store = store if isinstance(store, ClipStore) else ClipStore(store)
        self.modalities = list(modalities)
        fs_map = dict(FS_MAP, **(fs_map or {}))
        self.index = {}
        for m, fn in zip(self.modalities, fs_new):
            if m not in fs_map:
                raise KeyError("no sampling rate known for modality %r" % m)
            self.index[m] = window_index(self.store.shape(m)[0], fs_map[m], fn, time, window_hop)

    def __len__(self):
        return min(len(self.index[m][0]) for m in self.modalities)

    def __getitem__(self, idx):
        if not -len(self) <= idx < len(self):
            raise IndexError(idx)
        item = {}
        for m in self.modalities:
            starts, window, ratio = self.index[m]
            s = int(starts[idx])
            item[m] = np.asarray(self.store[m][s:s + window:ratio], dtype=np.float32)
        return item

    def views(self, modality, device=None):
        starts, window, ratio = self.index[modality]
        n = len(self)
        data = self.store[modality]
        steps = len(range(0, window, ratio))
        hop = int(starts[1] - starts[0]) if len(starts) > 1 else 0
        if device is None:
            a = np.asarray(data)
            if n == 0:
                return np.empty((0, steps) + a.shape[1:], a.dtype)
            row = a.strides[0]
            return np.lib.stride_tricks.as_strided(a, (n, steps) + a.shape[1:], (hop * row, ratio * row) + a.strides[1:],
                                                   writeable=False)
        import torch
        t = torch.from_numpy(np.ascontiguousarray(data)).to(device)
        if n == 0:
            return t.new_empty((0, steps) + tuple(t.shape[1:]))
        row = t.stride(0)
        return t.as_strided((n, steps) + tuple(t.shape[1:]), (hop * row, ratio * row) + tuple(t.stride()[1:]))
