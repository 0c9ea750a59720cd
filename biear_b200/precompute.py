"""Feature precompute for the passive mode (BASELINE.json config 3; reference: create_h5_data/precompute_h5.py ->
data_save.py:122-164 / data_h5_save.py:72-81): waveforms in, the arrays of the reference's H5 wire format out.

    active-wav format   x1 = wavL (n,16000)  x2 = wavR  x3 = CC (n,100)                      y = labels (n,56)
    passive format      x1 = log-energy L (n,19,100)  x2 = log-energy R  x3 = CC  x4 = phase L  x5 = phase R   y

The reference computes x3 on the CPU in float64 (np.correlate over all 31 999 lags, ProcessPool over files); here the CC
kernel, the STFT and the fixed-Q band GEMM run on the GPU over chunks of clips; the next chunk's staging (thread pool) and
host->device copy and the previous chunk's device->host copy (into pinned output arrays) overlap the current chunk's kernels.  Output is an .npz archive (and an H5 file with the same dataset names when h5py is
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


_STAGE_THREADS = 8
_pinned = {}     # (tag, shape) -> pinned host buffers, kept across calls: page-locking hundreds of MB costs ~100 ms


def _pinned_pair(tag, shape):
    key = (tag, tuple(shape))
    if key not in _pinned:
        if len(_pinned) > 16:
            _pinned.clear()
        _pinned[key] = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(2)]
    return _pinned[key]


def _parallel_copy(pool, dst: np.ndarray, src: np.ndarray):
    """dst[:] = src over row ranges on a thread pool (numpy releases the GIL in large copies): the staging memcpy into the
    pinned bounce buffers -- 128 KB per clip -- is what bounds this tool when it runs on one thread (~10 GB/s)."""
    m = src.shape[0]
    step = max(1, (m + _STAGE_THREADS - 1) // _STAGE_THREADS)
    futs = [pool.submit(np.copyto, dst[lo:lo + step], src[lo:lo + step]) for lo in range(0, m, step)]
    for f in futs:
        f.result()


@torch.no_grad()
def precompute(wav_l: np.ndarray, wav_r: np.ndarray, labels: Optional[np.ndarray] = None, fmt: str = "passive",
               chunk: int = 1024, device: str = "cuda:0", fs: int = 16000, n_bands: int = 100,
               max_lag_ms: float = 3.0) -> Dict[str, np.ndarray]:
    """wav_l / wav_r: (n, nsamp) float32 host arrays.  Returns the wire-format arrays (see module docstring).

    Per chunk of clips, concurrently: the staging of chunk j+1 into pinned bounce buffers (thread pool) and its
    host->device copy, the kernels of chunk j, the device->host copy of chunk j's features into pinned buffers and the
    copy-out of chunk j-1 into the result arrays.  The pinned buffers are chunk-sized and kept across calls.  The tool is
    bounded by moving the waveforms (128 KB per clip) through host memory and PCIe."""
    from concurrent.futures import ThreadPoolExecutor
    assert wav_l.shape == wav_r.shape and wav_l.ndim == 2
    assert fmt in ("active", "passive")
    dev = torch.device(device)
    n, nsamp = wav_l.shape
    wav_l = np.asarray(wav_l, np.float32)
    wav_r = np.asarray(wav_r, np.float32)
    chunk = max(1, min(chunk, n))
    fb = BinauralAdaptiveGammatoneFB(fs=fs, Nbands=n_bands, fixed_frontend_q=True).to(dev).eval()
    fb.graph_replay = False                       # chunk shapes vary (last chunk); plain launches
    T = fb.timesteps
    keys = ("x3",) + (("x1", "x2", "x4", "x5") if fmt == "passive" else ())
    shape_of = lambda k, rows: (rows, n_bands) if k == "x3" else (rows, T, n_bands)
    out = {k: np.empty(shape_of(k, n), np.float32) for k in keys}
    stage_out = {k: _pinned_pair("out." + k, shape_of(k, chunk)) for k in keys}
    bounce = [_pinned_pair("in.L", (chunk, nsamp)), _pinned_pair("in.R", (chunk, nsamp))]
    main = torch.cuda.current_stream(dev)
    h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    bounce_free = [None, None]                    # event after which the host may overwrite bounce buffers j % 2
    devbuf = [(torch.empty((chunk, nsamp), dtype=torch.float32, device=dev),
               torch.empty((chunk, nsamp), dtype=torch.float32, device=dev)) for _ in range(2)]
    starts = list(range(0, n, chunk))
    with ThreadPoolExecutor(max_workers=_STAGE_THREADS) as pool:

        def stage(j):
            lo = starts[j]
            m = min(chunk, n - lo)
            if bounce_free[j % 2] is not None:
                bounce_free[j % 2].synchronize()  # its previous host->device copy has finished
            _parallel_copy(pool, bounce[0][j % 2].numpy()[:m], wav_l[lo:lo + m])
            _parallel_copy(pool, bounce[1][j % 2].numpy()[:m], wav_r[lo:lo + m])
            h2d.wait_stream(main)                 # the device buffer's previous consumer has been enqueued
            with torch.cuda.stream(h2d):
                devbuf[j % 2][0][:m].copy_(bounce[0][j % 2][:m], non_blocking=True)
                devbuf[j % 2][1][:m].copy_(bounce[1][j % 2][:m], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            bounce_free[j % 2] = ev
            return ev, m

        def copy_out(j, m, ev):                   # features of chunk j: pinned buffers -> result arrays
            ev.synchronize()
            lo = starts[j]
            for k in keys:
                _parallel_copy(pool, out[k][lo:lo + m], stage_out[k][j % 2].numpy()[:m])

        pending = stage(0) if starts else None
        prev = None                               # (j, m, event) of the chunk whose features are on their way to the host
        for j, lo in enumerate(starts):
            ev, m = pending
            main.wait_event(ev)
            wl, wr = devbuf[j % 2][0][:m], devbuf[j % 2][1][:m]
            res = {"x3": ops.cc_feature(wl, wr, fs, n_bands, max_lag_ms)}
            if fmt == "passive":
                o = fb.forward_features(wl, wr, want_phase=True, want_logenergy=True)
                res.update(x1=o["logYL"], x2=o["logYR"], x4=o["phaseL"], x5=o["phaseR"])
            done = torch.cuda.Event()
            done.record(main)
            d2h.wait_event(done)
            with torch.cuda.stream(d2h):          # (stage_out[.][j % 2] was copied out by the host two chunks ago)
                for k in keys:
                    stage_out[k][j % 2][:m].copy_(res[k], non_blocking=True)
                    res[k].record_stream(d2h)
                out_ev = torch.cuda.Event()
                out_ev.record()
            if j + 1 < len(starts):
                pending = stage(j + 1)            # overlaps the kernels and copies just enqueued
            if prev is not None:
                copy_out(*prev)
            prev = (j, m, out_ev)
        if prev is not None:
            copy_out(*prev)
        torch.cuda.synchronize(dev)
    if fmt == "active":
        out["x1"], out["x2"] = wav_l, wav_r
    if labels is not None:
        out["y"] = np.asarray(labels, np.float32)
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
