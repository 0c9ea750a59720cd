"""Stand-in for the `data` module the reference imports (train_biear.py:13, evaluate_biear.py:12) but does not ship.

The wire format is the H5 layout written by create_h5_data/data_h5_save.py:72-81:
    x1, x2 : (n, 16000) float32 left / right waveform        x3 : (n, 100) float32 CC feature
    y      : (n, 56) float32 = 8 sectors x [presence, angle_norm, 5-way one-hot distance]  (data_save.py:75-119)
    passive files additionally hold x4, x5 : (n, 19, 100) sub-band phases.
h5py is not available in this image, so a path is read as follows:
    <path>        if it exists and h5py imports     -> the real H5 file
    <path>.npz    (same datasets, numpy archive)     -> loaded with numpy
    otherwise                                        -> FileNotFoundError / ImportError.  Only with the explicit opt-in
                                                        BIEAR_ALLOW_SYNTHETIC=1 (tests, smoke runs of the reference's
                                                        scripts without the corpora): a deterministic synthetic set (size
                                                        BIEAR_SYNTH_CLIPS, default 256) generated from a hash of the
                                                        path, announced with a loud warning.
Datasets are fork-safe (plain numpy arrays) for DataLoader(num_workers=4).
"""
import os
import zlib

import numpy as np
import torch
from torch.utils.data import Dataset

N_SECTORS, N_DIST_CLASS, FS = 8, 5, 16000


def _synthetic(path, n, passive):
    rs = np.random.RandomState(zlib.crc32(os.path.basename(str(path)).encode()) & 0x7FFFFFFF)
    src = rs.standard_normal((n, FS + 64)).astype(np.float32)
    for i in range(1, src.shape[1]):                       # AR(1) tilt, speech-like spectrum
        src[:, i] += 0.9 * src[:, i - 1]
    itd = rs.randint(-12, 13, size=n)
    gain = rs.uniform(0.5, 1.0, size=n).astype(np.float32)
    idx = (32 - itd)[:, None] + np.arange(FS)[None, :]
    x1 = src[:, 32:32 + FS]
    x2 = gain[:, None] * np.take_along_axis(src, idx, axis=1)
    peak = np.maximum(np.abs(x1).max(1), np.abs(x2).max(1))[:, None]
    x1, x2 = x1 / peak, x2 / peak
    y = np.zeros((n, N_SECTORS, 2 + N_DIST_CLASS), np.float32)
    for i in range(n):
        for s in rs.choice(N_SECTORS, size=rs.randint(1, 4), replace=False):
            y[i, s, 0] = 1.0
            y[i, s, 1] = rs.uniform()
            y[i, s, 2 + rs.randint(N_DIST_CLASS)] = 1.0
    out = {"x1": x1.astype(np.float32), "x2": x2.astype(np.float32),
           "x3": rs.uniform(-1, 1, size=(n, 100)).astype(np.float32), "y": y.reshape(n, -1)}
    if passive:
        out["x1"] = rs.standard_normal((n, 19, 100)).astype(np.float32)
        out["x2"] = rs.standard_normal((n, 19, 100)).astype(np.float32)
        out["x4"] = rs.uniform(-np.pi, np.pi, size=(n, 19, 100)).astype(np.float32)
        out["x5"] = rs.uniform(-np.pi, np.pi, size=(n, 19, 100)).astype(np.float32)
    return out


def load_arrays_from_h5(path, keys=("x1", "x2", "x3", "y"), passive=False):
    path = str(path)
    if os.path.exists(path + ".npz"):          # numpy twin of the H5 wire format (written by biear_b200.precompute)
        z = np.load(path + ".npz")
        return {k: z[k] for k in keys}
    problem = None
    if os.path.exists(path):
        try:
            import h5py
        except ImportError as e:
            problem = ImportError(f"{path} exists but h5py is not importable ({e}); install h5py or provide {path}.npz")
        else:
            with h5py.File(path, "r") as f:
                return {k: np.asarray(f[k]) for k in keys}
    else:
        problem = FileNotFoundError(f"dataset not found: {path} (nor {path}.npz)")
    if os.environ.get("BIEAR_ALLOW_SYNTHETIC") == "1":
        n = int(os.environ.get("BIEAR_SYNTH_CLIPS", "256"))
        import warnings
        msg = f"biear_b200.dropin.data: {problem}; BIEAR_ALLOW_SYNTHETIC=1 -> using {n} SYNTHETIC clips, metrics are meaningless"
        warnings.warn(msg, stacklevel=2)
        print("[WARNING] " + msg, flush=True)
        return _synthetic(path, n, passive)
    raise problem


class DeepEarH5Dataset_Active(Dataset):
    """Items are (wavL, wavR, x3, y) float32 tensors (train_biear.py:457)."""

    def __init__(self, h5_path):
        a = load_arrays_from_h5(h5_path)
        self.x1, self.x2, self.x3, self.y = a["x1"], a["x2"], a["x3"], a["y"]

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return (torch.from_numpy(self.x1[i]), torch.from_numpy(self.x2[i]), torch.from_numpy(self.x3[i]),
                torch.from_numpy(self.y[i]))


class DeepEarH5Dataset(Dataset):
    """Passive features: items are (x1, x2, x3, x4, x5, y) (create_h5_data/data_save.py:261-262)."""

    def __init__(self, h5_path):
        a = load_arrays_from_h5(h5_path, keys=("x1", "x2", "x3", "x4", "x5", "y"), passive=True)
        self.a = a

    def __len__(self):
        return len(self.a["y"])

    def __getitem__(self, i):
        return tuple(torch.from_numpy(self.a[k][i]) for k in ("x1", "x2", "x3", "x4", "x5", "y"))
