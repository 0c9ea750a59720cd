"""Feature precompute for the passive mode (BASELINE.json config 3; reference: create_h5_data/precompute_h5.py ->
data_save.py:122-164 / data_h5_save.py:72-81): waveforms in, the arrays of the reference's H5 wire format out.

    active-wav format   x1 = wavL (n,16000)  x2 = wavR  x3 = CC (n,100)                      y = labels (n,56)
    passive format      x1 = log-energy L (n,19,100)  x2 = log-energy R  x3 = CC  x4 = phase L  x5 = phase R   y

The reference computes x3 on the CPU in float64 (np.correlate over all 31 999 lags, ProcessPool over files); here the CC
kernel, the STFT and the fixed-Q band GEMM run on the GPU over chunks of clips, with the next chunk's H2D copy overlapping
the current chunk's kernels.  Output is an .npz archive (and an H5 file with the same dataset names when h5py is
importable); biear_b200/dropin/data.py reads both.

    python -m biear_b200.precompute --synthetic 4096 --format passive --out /tmp/feats.npz
"""
from __future__ import annotations

import argparse
import time
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .frontend import BinauralAdaptiveGammatoneFB


@torch.no_grad()
def precompute(wav_l: np.ndarray, wav_r: np.ndarray, labels: Optional[np.ndarray] = None, fmt: str = "passive",
               chunk: int = 1024, device: str = "cuda:0", fs: int = 16000, n_bands: int = 100,
               max_lag_ms: float = 3.0) -> Dict[str, np.ndarray]:
    """wav_l / wav_r: (n, nsamp) float32 host arrays.  Returns the wire-format arrays (see module docstring)."""
    assert wav_l.shape == wav_r.shape and wav_l.ndim == 2
    assert fmt in ("active", "passive")
    dev = torch.device(device)
    n = wav_l.shape[0]
    fb = BinauralAdaptiveGammatoneFB(fs=fs, Nbands=n_bands, fixed_frontend_q=True).to(dev).eval()
    out = {"x3": np.empty((n, n_bands), np.float32)}
    if fmt == "passive":
        for k in ("x1", "x2", "x4", "x5"):
            out[k] = np.empty((n, fb.timesteps, n_bands), np.float32)
    else:
        out["x1"], out["x2"] = wav_l.astype(np.float32, copy=False), wav_r.astype(np.float32, copy=False)
    if labels is not None:
        out["y"] = np.asarray(labels, np.float32)
    nsamp = wav_l.shape[1]
    pin = [(torch.empty((chunk, nsamp), dtype=torch.float32).pin_memory(),
            torch.empty((chunk, nsamp), dtype=torch.float32).pin_memory()) for _ in range(2)]
    devbuf = [(torch.empty((chunk, nsamp), dtype=torch.float32, device=dev),
               torch.empty((chunk, nsamp), dtype=torch.float32, device=dev)) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    starts = list(range(0, n, chunk))

    def stage(j):
        lo = starts[j]
        m = min(chunk, n - lo)
        a, b = pin[j % 2]
        a[:m].copy_(torch.from_numpy(np.ascontiguousarray(wav_l[lo:lo + m], np.float32)))
        b[:m].copy_(torch.from_numpy(np.ascontiguousarray(wav_r[lo:lo + m], np.float32)))
        copy_stream.wait_stream(main)            # the device buffer's previous consumer has been enqueued
        with torch.cuda.stream(copy_stream):
            devbuf[j % 2][0][:m].copy_(a[:m], non_blocking=True)
            devbuf[j % 2][1][:m].copy_(b[:m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        return ev, m

    pending = stage(0) if starts else None
    for j, lo in enumerate(starts):
        ev, m = pending
        main.wait_event(ev)
        wl, wr = devbuf[j % 2][0][:m], devbuf[j % 2][1][:m]
        cc = ops.cc_feature(wl, wr, fs, n_bands, max_lag_ms)
        if fmt == "passive":
            o = fb.forward_features(wl, wr, want_phase=True, want_logenergy=True)
            res = (o["logYL"], o["logYR"], o["phaseL"], o["phaseR"])
        if j + 1 < len(starts):
            pending = stage(j + 1)               # overlaps the kernels just enqueued
        out["x3"][lo:lo + m] = cc.cpu().numpy()
        if fmt == "passive":
            for k, t in zip(("x1", "x2", "x4", "x5"), res):
                out[k][lo:lo + m] = t.cpu().numpy()
    return out


def save(arrays: Dict[str, np.ndarray], path: str):
    """.npz always; additionally <path>.h5 with the same dataset names (data_h5_save.py:72-81) when h5py is available."""
    np.savez(path if path.endswith(".npz") else path + ".npz", **arrays)
    try:
        import h5py
    except ImportError:
        return
    with h5py.File(path.replace(".npz", "") + ".h5", "w") as f:
        for k, v in arrays.items():
            f.create_dataset(k, data=v)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--in", dest="inp", help=".npz with wavL, wavR (n, nsamp) [and y]")
    ap.add_argument("--synthetic", type=int, default=0, help="generate this many synthetic clips instead of --in")
    ap.add_argument("--format", default="passive", choices=["active", "passive"])
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    if a.synthetic:
        rs = np.random.RandomState(0)
        wl = rs.uniform(-1, 1, size=(a.synthetic, 16000)).astype(np.float32)
        wr = np.roll(wl, 5, axis=1) * 0.8
        y = None
    else:
        z = np.load(a.inp)
        wl, wr, y = z["wavL"], z["wavR"], (z["y"] if "y" in z.files else None)
    t0 = time.perf_counter()
    arrays = precompute(wl, wr, y, fmt=a.format, chunk=a.chunk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    save(arrays, a.out)
    print(f"{wl.shape[0]} clips in {dt:.3f} s = {wl.shape[0] / dt:.0f} clips/s (host arrays in, host arrays out) -> {a.out}")


if __name__ == "__main__":
    main()
