"""Torch-facing wrappers of the C ABI (device pointers + the current CUDA stream go straight through).

PyTorch is used here for device memory, streams and autograd bookkeeping only; every numeric result
comes from the sm_100a kernels in ``biear_b200/csrc``.  All functions raise on CPU tensors: there is
no fallback path.
"""
from __future__ import annotations

from ctypes import c_void_p
from functools import lru_cache
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, timing

DEFAULT_CUTOFF = 6.0   # Gaussian support half-width in units of bw (1.5e-8 on Y, 4e-7 on dY/dQ)


def _ptr(t: Optional[torch.Tensor]):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _stream(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError(f"biear_b200: {name} must be a CUDA tensor (got {t.device}); there is no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"biear_b200: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"biear_b200: {name} must be contiguous")


_seed_state = {}   # device index -> int64 device tensor: dropout seed counter used while a CUDA graph is being captured


def _prepare(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    _lib.ensure_init(idx)
    if idx not in _seed_state and not torch.cuda.is_current_stream_capturing():
        base = int(torch.randint(0, 2 ** 62, (1,)).item())
        _seed_state[idx] = torch.tensor([base], dtype=torch.int64, device=torch.device("cuda", idx))
    return _lib.load()


def _captured_seed(device) -> torch.Tensor:
    """Dropout seed for a forward that is being recorded into a CUDA graph: a device-side counter is advanced on
    the stream and snapshotted, so every replay of the graph draws fresh masks and the backward of the same replay
    sees the seed its forward used."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _seed_state:
        raise RuntimeError("biear_b200: run the front-end once eagerly (warm-up) before capturing it in a CUDA graph")
    st = _seed_state[idx]
    st.add_(0x632BE59BD9B4E019)          # odd increment: a full-period walk over the 64-bit seeds
    return st.clone()


# ------------------------------------------------------------------------------------------------
# STFT
# ------------------------------------------------------------------------------------------------
def stft(wav, win_fn: torch.Tensor, fs: int, timesteps: int, win: int, hop: int, n_fft: int) -> torch.Tensor:
    """(rows, nsamp) fp32 -> X (rows, T, n_fft//2+1) complex64.  model_torch.py:289-312, 334-335.
    `wav` may be a list of equally shaped (B, nsamp) tensors (the ears): their spectra are written one after the other
    into ONE output tensor (E*B, T, F) without concatenating the waveforms first."""
    wavs = list(wav) if isinstance(wav, (list, tuple)) else [wav]
    for w in wavs:
        if w.dim() != 2:
            raise ValueError(f"Expected wav_1s (B,N), got {tuple(w.shape)}")
        if w.shape != wavs[0].shape:
            raise ValueError(f"waveform shapes differ: {tuple(w.shape)} vs {tuple(wavs[0].shape)}")
        _need_cuda(w, "wav")
    _need_cuda(win_fn, "win_fn")
    dev = wavs[0].device
    with torch.cuda.device(dev):
        lib = _prepare(dev)
        rows, nsamp = wavs[0].shape
        nbins = n_fft // 2 + 1
        xr = torch.empty((len(wavs) * rows, timesteps, nbins, 2), dtype=torch.float32, device=dev)
        for e, w in enumerate(wavs):
            out = c_void_p(xr.data_ptr() + 4 * e * rows * timesteps * nbins * 2)
            _lib.check(lib.biear_stft_fwd(_ptr(w), rows, nsamp, nsamp, _ptr(win_fn), fs, timesteps, win, hop,
                                          n_fft, out, _stream(dev)), "biear_stft_fwd")
    return torch.view_as_complex(xr)


def pcm16_to_f32(pcm: torch.Tensor, out: Optional[torch.Tensor] = None, scale: float = 1.0 / 32768.0) -> torch.Tensor:
    """int16 PCM (any shape, CUDA, contiguous) -> float32 waveform pcm * scale, on the current stream (biear_pcm16_to_f32).
    The reference's harness divides int16-range input by 32768 itself (train_biear.py:463-467)."""
    _need_cuda(pcm, "pcm", torch.int16)
    dev = pcm.device
    if out is None:
        out = torch.empty(pcm.shape, dtype=torch.float32, device=dev)
    else:
        _need_cuda(out, "out")
        if out.shape != pcm.shape:
            raise ValueError(f"out {tuple(out.shape)} and pcm {tuple(pcm.shape)} differ")
    with torch.cuda.device(dev):
        lib = _prepare(dev)
        _lib.check(lib.biear_pcm16_to_f32(_ptr(pcm), _ptr(out), pcm.numel(), float(scale), _stream(dev)), "biear_pcm16_to_f32")
    return out


def stft_pair(wav_a: torch.Tensor, wav_b: torch.Tensor, win_fn: torch.Tensor, fs: int, timesteps: int, win: int, hop: int,
              n_fft: int, ready: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Both ears in one launch, frame-major work order: X (2B, T, F) complex64 (rows of wav_a first).  `ready` ((2B*T + 4,)
    int32, cleared beforehand; the last entries are the kernel's work counter) receives a 1 per finished (row, frame): the streaming hand-over to the recurrence kernel
    (BiearSeqParams.x_ready), which may then run CONCURRENTLY with this launch."""
    for w in (wav_a, wav_b):
        if w.dim() != 2:
            raise ValueError(f"Expected wav_1s (B,N), got {tuple(w.shape)}")
        _need_cuda(w, "wav")
    if wav_a.shape != wav_b.shape:
        raise ValueError(f"waveform shapes differ: {tuple(wav_a.shape)} vs {tuple(wav_b.shape)}")
    _need_cuda(win_fn, "win_fn")
    dev = wav_a.device
    with torch.cuda.device(dev):
        lib = _prepare(dev)
        rows, nsamp = wav_a.shape
        nbins = n_fft // 2 + 1
        xr = torch.empty((2 * rows, timesteps, nbins, 2), dtype=torch.float32, device=dev)
        if ready is not None:
            assert ready.dtype == torch.int32 and ready.numel() == 2 * rows * timesteps + 4 and ready.is_contiguous()
        _lib.check(lib.biear_stft_fwd_pair(_ptr(wav_a), _ptr(wav_b), rows, nsamp, nsamp, _ptr(win_fn), fs, timesteps, win, hop,
                                           n_fft, _ptr(xr), _ptr(ready), _stream(dev)), "biear_stft_fwd_pair")
    return torch.view_as_complex(xr)


# ------------------------------------------------------------------------------------------------
# band stage
# ------------------------------------------------------------------------------------------------
def band_forward(xr: torch.Tensor, t: Optional[int], q: torch.Tensor, fc: torch.Tensor, df: float,
                 cutoff: float = DEFAULT_CUTOFF, want_phase: bool = True, want_jac: bool = True):
    """Band energies (+ phase, + exact Jacobians wrt Q).

    xr: (rows, T, F, 2) fp32 view of the spectra.
    t is an int:  one frame; q is (rows, N); outputs are (rows, N).
    t is None:    all frames; q is (N,) [broadcast, fixed-Q path] or (rows, T, N); outputs (rows, T, N).
    Returns (Y, phase|None, dYdQ|None, dPdQ|None).
    """
    _need_cuda(xr, "X")
    _need_cuda(q, "Q")
    _need_cuda(fc, "fc")
    rows, T, F, two = xr.shape
    assert two == 2
    N = fc.numel()
    dev = xr.device
    with torch.cuda.device(dev):
        lib = _prepare(dev)
        if t is not None:
            if q.shape != (rows, N):
                raise ValueError(f"Q must be ({rows},{N}), got {tuple(q.shape)}")
            items, x_off, x_stride, q_stride = rows, t * F * 2, T * F * 2, N
            out_shape = (rows, N)
        else:
            items, x_off, x_stride = rows * T, 0, F * 2
            if q.dim() == 1:
                q_stride = 0
            elif q.shape == (rows, T, N):
                q_stride = N
            else:
                raise ValueError(f"Q must be ({N},) or ({rows},{T},{N}), got {tuple(q.shape)}")
            out_shape = (rows, T, N)
        y = torch.empty(out_shape, dtype=torch.float32, device=dev)
        ph = torch.empty(out_shape, dtype=torch.float32, device=dev) if want_phase else None
        dy = torch.empty(out_shape, dtype=torch.float32, device=dev) if want_jac else None
        dp = torch.empty(out_shape, dtype=torch.float32, device=dev) if (want_jac and want_phase) else None
        x_ptr = c_void_p(xr.data_ptr() + 4 * x_off)
        _lib.check(lib.biear_band_fwd(x_ptr, x_stride, _ptr(q), q_stride, _ptr(fc), items, N, F, float(df),
                                      float(cutoff), _ptr(y), N, _ptr(ph), N, _ptr(dy), _ptr(dp), N,
                                      _stream(dev)), "biear_band_fwd")
    return y, ph, dy, dp


FIXED_VARIANT = "tc"   # the shipped path: tcgen05 / TMEM 3xTF32 GEMM (csrc/band_fixed_tc.cu), picked by measured throughput
                       # (profiles/r1_other_configs.txt).  variant="ffma" (fp32 FFMA2 GEMM, csrc/band_fixed.cu) is a
                       # test hook: the parity tests cross-check the two; nothing in the package selects it.


def band_fixed_forward(xr: torch.Tensor, q: torch.Tensor, fc: torch.Tensor, df: float,
                       cutoff: float = DEFAULT_CUTOFF, want_phase: bool = True, variant: Optional[str] = None):
    """Fixed-Q band stage for ALL (row, frame) items as one dense contraction (biear_band_fixed_fwd[_tc]).
    xr (rows, T, F, 2), q (N,) shared by every item -> Y (rows, T, N), phase (rows, T, N) | None."""
    _need_cuda(xr, "X")
    _need_cuda(q, "Q")
    _need_cuda(fc, "fc")
    rows, T, F, _ = xr.shape
    N = fc.numel()
    if q.shape != (N,):
        raise ValueError(f"Q must be ({N},), got {tuple(q.shape)}")
    variant = variant or FIXED_VARIANT
    if variant not in ("tc", "ffma"):
        raise ValueError(f"unknown fixed-Q variant {variant!r}")
    dev = xr.device
    with torch.cuda.device(dev):
        lib = _prepare(dev)
        y = torch.empty((rows, T, N), dtype=torch.float32, device=dev)
        ph = torch.empty((rows, T, N), dtype=torch.float32, device=dev) if want_phase else None
        if variant == "tc":
            work = torch.empty(int(lib.biear_band_fixed_tc_workspace_floats(F)), dtype=torch.float32, device=dev)
            _lib.check(lib.biear_band_fixed_fwd_tc(_ptr(xr), 2 * F, _ptr(q), _ptr(fc), rows * T, N, F, float(df),
                                                   float(cutoff), _ptr(y), N, _ptr(ph), N, _ptr(work), _stream(dev)),
                       "biear_band_fixed_fwd_tc")
        else:
            work = torch.empty(int(lib.biear_band_fixed_workspace_floats(F)), dtype=torch.float32, device=dev)
            _lib.check(lib.biear_band_fixed_fwd(_ptr(xr), 2 * F, _ptr(q), _ptr(fc), rows * T, N, F, float(df),
                                                float(cutoff), _ptr(y), N, _ptr(ph), N, _ptr(work), _stream(dev)),
                       "biear_band_fixed_fwd")
    return y, ph


def band_backward(xr: torch.Tensor, t: Optional[int], q: torch.Tensor, fc: torch.Tensor, df: float,
                  g_y: Optional[torch.Tensor], g_phase: Optional[torch.Tensor],
                  cutoff: float = DEFAULT_CUTOFF) -> torch.Tensor:
    """dL/dQ by recomputation (nothing saved by the forward).  Shapes as in band_forward; q must be
    per-item here (the broadcast fixed-Q path has no gradient)."""
    _need_cuda(xr, "X")
    _need_cuda(q, "Q")
    rows, T, F, _ = xr.shape
    N = fc.numel()
    dev = xr.device
    for name, g in (("gY", g_y), ("gphase", g_phase)):
        if g is not None:
            _need_cuda(g, name)
    with torch.cuda.device(dev):
        lib = _prepare(dev)
        if t is not None:
            items, x_off, x_stride = rows, t * F * 2, T * F * 2
        else:
            items, x_off, x_stride = rows * T, 0, F * 2
        dq = torch.empty_like(q)
        x_ptr = c_void_p(xr.data_ptr() + 4 * x_off)
        _lib.check(lib.biear_band_bwd(x_ptr, x_stride, _ptr(q), N, _ptr(fc), items, N, F, float(df), float(cutoff),
                                      _ptr(g_y), N, _ptr(g_phase), N, _ptr(dq), N, 0, _stream(dev)),
                   "biear_band_bwd")
    return dq


class BandFrame(torch.autograd.Function):
    """One frame of the adaptive band stage as an autograd node: (Q_t) -> (Y_t, phase_t).

    mode "jacobian": the forward also emits dY/dQ and dphase/dQ (2 floats per band) and the backward
    is one fused multiply-add; mode "recompute": nothing is saved and the backward re-runs the band
    kernel in its dL/dQ form.  Both give the closed-form gradient of SURVEY.md A.3.
    """

    @staticmethod
    def forward(ctx, q, xr, t, fc, df, cutoff, want_phase, mode):
        ctx.set_materialize_grads(False)
        need_grad = ctx.needs_input_grad[0]      # (before contiguous(): inside forward() a copy never requires grad)
        q = q.contiguous()
        jac = need_grad and mode == "jacobian"
        y, ph, dy, dp = band_forward(xr, t, q, fc, df, cutoff, want_phase, jac)
        ctx.mode = mode
        ctx.meta = (t, df, cutoff)
        if need_grad:
            if jac:
                ctx.save_for_backward(dy, dp) if dp is not None else ctx.save_for_backward(dy)
            else:
                ctx.save_for_backward(xr, q, fc)
        if ph is None:
            ph = y.new_empty(0)
            ctx.mark_non_differentiable(ph)
        return y, ph

    @staticmethod
    def backward(ctx, g_y, g_ph):
        if g_y is None and g_ph is None:
            return (None,) * 8
        if ctx.mode == "jacobian":
            saved = ctx.saved_tensors
            dq = None
            if g_y is not None:
                dq = g_y * saved[0]
            if g_ph is not None and len(saved) > 1:
                dq = g_ph * saved[1] if dq is None else torch.addcmul(dq, g_ph, saved[1])
        else:
            xr, q, fc = ctx.saved_tensors
            t, df, cutoff = ctx.meta
            dq = band_backward(xr, t, q, fc, df,
                               g_y.contiguous() if g_y is not None else None,
                               g_ph.contiguous() if g_ph is not None else None, cutoff)
        return (dq,) + (None,) * 7


# ------------------------------------------------------------------------------------------------
# cross-correlation feature
# ------------------------------------------------------------------------------------------------
@lru_cache(maxsize=32)
def _cc_tables(nsamp: int, fs: float, num_lags: int, max_lag_ms: float):
    """Integer lag range the reference keeps and np.interp's (left index, weight) for every output point,
    computed with the same float64 expressions as utils.py:408-418."""
    max_lag_sec = max_lag_ms * 1e-3
    reach = int(np.ceil(abs(max_lag_sec) * fs)) + 2
    lo, hi = max(-(nsamp - 1), -reach), min(nsamp - 1, reach)
    lag_idx = np.arange(lo, hi + 1)
    lag_sec = lag_idx / fs
    keep = np.logical_and(lag_sec >= -max_lag_sec, lag_sec <= max_lag_sec)
    ks = lag_idx[keep]
    if ks.size == 0:
        raise ValueError("cross-correlation: empty lag range")
    xp = lag_sec[keep]
    target = np.linspace(-max_lag_sec, max_lag_sec, num_lags)
    j = np.searchsorted(xp, target, side="right") - 1
    idx = np.clip(j, 0, max(len(xp) - 2, 0))
    if len(xp) > 1:
        frac = (target - xp[idx]) / (xp[idx + 1] - xp[idx])
    else:
        frac = np.zeros_like(target)
    left = target <= xp[0]
    right = target >= xp[-1]
    idx = np.where(left, 0, np.where(right, len(xp) - 1, idx))
    frac = np.where(left | right, 0.0, frac)
    return int(ks[0]), int(ks[-1]), idx.astype(np.int32), frac.astype(np.float32)


_cc_dev_cache = {}


def cc_feature(wav_l: torch.Tensor, wav_r: torch.Tensor, fs: float = 16000, num_lags: int = 100,
               max_lag_ms: float = 3.0) -> torch.Tensor:
    """(B, nsamp) x2 fp32 -> (B, num_lags) fp32.  utils.py:390-420 (compute_cross_correlation_feature)."""
    if wav_l.dim() != 2 or wav_l.shape != wav_r.shape:
        raise ValueError(f"expected two (B,N) waveforms, got {tuple(wav_l.shape)} and {tuple(wav_r.shape)}")
    _need_cuda(wav_l, "wavL")
    _need_cuda(wav_r, "wavR")
    dev = wav_l.device
    B, nsamp = wav_l.shape
    key = (nsamp, float(fs), int(num_lags), float(max_lag_ms), dev.index)
    if key not in _cc_dev_cache:
        k_min, k_max, idx, frac = _cc_tables(nsamp, float(fs), int(num_lags), float(max_lag_ms))
        _cc_dev_cache[key] = (k_min, k_max, torch.from_numpy(idx).to(dev), torch.from_numpy(frac).to(dev))
    k_min, k_max, idx_d, frac_d = _cc_dev_cache[key]
    with torch.cuda.device(dev):
        lib = _prepare(dev)
        out = torch.empty((B, num_lags), dtype=torch.float32, device=dev)
        _lib.check(lib.biear_cc_fwd(_ptr(wav_l), _ptr(wav_r), B, nsamp, nsamp, k_min, k_max, _ptr(idx_d),
                                    _ptr(frac_d), num_lags, _ptr(out), _stream(dev)), "biear_cc_fwd")
    return out


# ------------------------------------------------------------------------------------------------
# Q regularisers (train_biear.py:476-490)
# ------------------------------------------------------------------------------------------------
_qreg_ws = {}   # device index -> workspace (its counter word zeroed once; the kernel leaves it zero)


def _qreg_workspace(dev, lib):
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _qreg_ws.get(idx)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("biear_b200: run q_regularizers once eagerly (warm-up) before capturing it in a CUDA graph")
        ws = torch.zeros(int(lib.biear_q_regularizers_workspace_floats()), dtype=torch.float32, device=dev)
        _qreg_ws[idx] = ws
    return ws


class QRegularizers(torch.autograd.Function):
    """loss = w_reg * mean((logQ - logQ0)^2) + w_smooth * mean(diff_band(logQ)^2) on Q = (QL + QR) / 2, value and
    d loss / dQ from one kernel (biear_q_regularizers); the backward is one scaling by the upstream gradient."""

    @staticmethod
    def forward(ctx, ql, qr, q0, w_reg, w_smooth):
        _need_cuda(ql, "QL")
        _need_cuda(q0, "Q0")
        if qr is not None:
            _need_cuda(qr, "QR")
            if qr.shape != ql.shape:
                raise ValueError(f"QL {tuple(ql.shape)} and QR {tuple(qr.shape)} differ")
        N = ql.shape[-1]
        if q0.numel() != N:
            raise ValueError(f"Q0 has {q0.numel()} bands, Q has {N}")
        dev = ql.device
        need = ql.requires_grad or (qr is not None and qr.requires_grad)
        with torch.cuda.device(dev):
            lib = _prepare(dev)
            ws = _qreg_workspace(dev, lib)
            out = torch.empty(3, dtype=torch.float32, device=dev)
            gq = torch.empty_like(ql) if need else None
            _lib.check(lib.biear_q_regularizers(_ptr(ql), _ptr(qr), _ptr(q0), ql.numel() // N, N, float(w_reg),
                                                float(w_smooth), _ptr(out), _ptr(gq), _ptr(ws), _stream(dev)),
                       "biear_q_regularizers")
        ctx.gq = gq
        ctx.two = qr is not None
        loss, reg_q, reg_smooth = out[0], out[1], out[2]
        ctx.mark_non_differentiable(reg_q, reg_smooth)
        return loss, reg_q, reg_smooth

    @staticmethod
    def backward(ctx, g, _g1, _g2):
        if ctx.gq is None or g is None:
            return None, None, None, None, None
        grad = ctx.gq * g
        return grad, (grad if ctx.two else None), None, None, None


def q_regularizers(ql: torch.Tensor, qr: Optional[torch.Tensor], q0: torch.Tensor, w_reg: float, w_smooth: float):
    """The two Q regularisers of train_biear.py:476-490 on Q = (QL + QR) / 2 (qr None: Q = QL).
    Returns (w_reg * reg_q + w_smooth * reg_smooth [differentiable], reg_q, reg_smooth) as 0-d tensors."""
    return QRegularizers.apply(ql, qr, q0, w_reg, w_smooth)


# ------------------------------------------------------------------------------------------------
# fused adaptive recurrence (dual front-end)
# ------------------------------------------------------------------------------------------------
WEIGHT_NAMES = ("w_ih", "w_hh", "b_ih", "b_hh", "w1", "b1", "ln1_g", "ln1_b", "w2", "b2", "ln2_g", "ln2_b", "w3", "b3")
HID = 128


def _fill(params, **tensors):
    for k, v in tensors.items():
        setattr(params, k, v.data_ptr() if v is not None else None)


def fused_supported(n_bands: int, n_bins: int) -> bool:
    """Can the persistent recurrence kernels take this geometry (shared-memory budget)?"""
    return bool(_lib.load().biear_adaptive_supported(int(n_bands), int(n_bins)))


def single_supported(n_bands: int, n_bins: int) -> bool:
    """The same for the single-controller variant (csrc/seq_single.cu)."""
    return bool(_lib.load().biear_single_supported(int(n_bands), int(n_bins)))


_resident = {}


def resident_clusters(n_bands: int, n_bins: int, device=None) -> int:
    """How many clusters (tiles) of the persistent forward kernel fit on the device at once
    (cudaOccupancyMaxActiveClusters; 33 on a B200).  Cached per (device, geometry)."""
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    key = (idx, int(n_bands), int(n_bins))
    if key not in _resident:
        from ctypes import byref, c_int
        f, b = c_int(0), c_int(0)
        with torch.cuda.device(idx):
            _lib.check(_lib.load().biear_adaptive_occupancy(int(n_bands), int(n_bins), byref(f), byref(b)),
                       "biear_adaptive_occupancy")
        _resident[key] = int(f.value)
    return _resident[key]


CLUSTER_CTAS = 4   # CTAs per cluster of the persistent recurrence kernels (kCS in csrc/seq_dev.cuh)


def tile_rows() -> int:
    """Rows per tile (R) of the "tile layout" tensors (G, T-1, tiles, D, R), see include/biear_b200.h."""
    return int(_lib.load().biear_adaptive_tile_rows())


WGRAD_VARIANT = "tc"   # the shipped path: tcgen05 / TMEM 3xTF32 (wgrad_tc_kernel); variant="ffma" is a test hook and the
                       # shape fallback below


def ctrl_wgrad(jobs, variant: Optional[str] = None):
    """Weight gradients for a list of jobs in two launches (biear_ctrl_wgrad[_tc]).

    A job is (a, do, bm, di, chunks, bias[, dw_out, db_out[, dw2_out, scale2]]).  a (G, chunks', Da, R) and bm (G, chunks', Db, R) are
    tile-layout operands (views with arbitrary group / chunk strides are fine); the first `do` / `di` features and the
    first `chunks` chunks are used.  di == 0 selects the diagonal form dW[g][o] = sum a*bm (LayerNorm weight).
    Without dw_out / db_out, fresh dense outputs dW (G,do,di) | (G,do) and db (G,do) (if bias) are allocated; with them
    the job writes into the given (possibly strided: a row block of a larger tensor) views of shape (G,do,di) / (G,do).
    dw2_out (same shape / strides as dw_out) receives scale2 * dW as well.  Returns [(dW, db | None), ...].
    """
    a0 = jobs[0][0]
    G, R, dev = a0.shape[0], a0.shape[3], a0.device
    lib = _prepare(dev)
    arr = (_lib.WgradJob * len(jobs))()
    outs = []
    for j, job in enumerate(jobs):
        a, do, bm, di, chunks, want_bias = job[:6]
        dw = job[6] if len(job) > 6 else None
        db = job[7] if len(job) > 7 else None
        assert a.shape[0] == G and bm.shape[0] == G and a.shape[3] == R and bm.shape[3] == R
        assert a.stride(3) == 1 and a.stride(2) == R and bm.stride(3) == 1 and bm.stride(2) == R
        if dw is None:
            dw = torch.empty((G, do, di) if di > 0 else (G, do), dtype=torch.float32, device=dev)
        if db is None and want_bias:
            db = torch.empty((G, do), dtype=torch.float32, device=dev)
        assert dw.stride(-1) == 1 and (db is None or db.stride(-1) == 1)
        outs.append((dw, db))
        q = arr[j]
        q.A, q.a_group_stride, q.a_chunk_stride, q.Do = a.data_ptr(), a.stride(0), a.stride(1), do
        q.Bm, q.b_group_stride, q.b_chunk_stride, q.Di = bm.data_ptr(), bm.stride(0), bm.stride(1), di
        q.chunks, q.dW, q.db = chunks, dw.data_ptr(), (db.data_ptr() if db is not None else None)
        dw2 = job[8] if len(job) > 8 else None
        if dw2 is not None:
            assert di > 0 and dw2.shape == dw.shape and dw2.stride() == dw.stride()
            q.dW2, q.scale2 = dw2.data_ptr(), float(job[9])
        q.dw_group_stride = dw.stride(0)
        q.dw_row_stride = dw.stride(1) if di > 0 else 1
        q.db_group_stride = db.stride(0) if db is not None else 0
    variant = variant or WGRAD_VARIANT
    if variant == "tc" and any(job[3] > 128 or job[3] % 4 for job in jobs):
        variant = "ffma"                      # wider / odd inputs: the FFMA kernel takes any shape
    sizer, runner = ((lib.biear_wgrad_scratch_floats_tc, lib.biear_ctrl_wgrad_tc) if variant == "tc"
                     else (lib.biear_wgrad_scratch_floats, lib.biear_ctrl_wgrad))
    n = int(sizer(arr, len(jobs), G, R))
    if n < 0:
        _lib.check(-1, "biear_wgrad_scratch_floats")
    scratch = torch.empty(max(n, 4), dtype=torch.float32, device=dev)
    _lib.check(runner(arr, len(jobs), G, R, _ptr(scratch), _stream(dev)), "biear_ctrl_wgrad")
    return outs


class PreparedSequence:
    """What biear_adaptive_prepare leaves behind for one step: the packed weight images (workspace), the GRU state
    tensor with its zeroed step 0, the cleared fallback flags, the snapshotted dropout seed -- and the event that marks
    the preparation launch on the stream it ran on."""
    __slots__ = ("work", "H", "flags", "seed_dev", "event", "stream", "key", "launched", "x_ready")


def adaptive_prepare(weights, B: int, T: int, N: int, training: bool, stream: Optional[torch.cuda.Stream] = None,
                     launch: bool = True, streamed_spectra: bool = False, ears: Optional[int] = None):
    """Run the spectra-independent part of a recurrence step (weight-image packing for the forward AND the backward
    kernel, H[:, 0] = 0, flags = 0, dropout-seed snapshot) as one launch, on `stream` if given: the front-end issues
    it on a forked stream so that it overlaps the STFT.  Pass the result to adaptive_sequence(prep=...).
    streamed_spectra=True also allocates and clears the (G*B*T) ready flags of the streaming spectra hand-over
    (stft_pair(ready=prep.x_ready) then runs next to the recurrence instead of before it).
    launch=False only allocates (the buffers are poisoned): biear_adaptive_fwd / _bwd then prepare on their own stream,
    the path a C caller that never calls biear_adaptive_prepare takes (testing).
    weights: dict name -> list of the G controllers' tensors."""
    G = len(weights[WEIGHT_NAMES[0]])
    E = G if ears is None else int(ears)          # ears == 2 with one controller: the single-controller front-end
    single = (G == 1 and E == 2)
    assert E == G or single
    ws = [w.detach() for k in WEIGHT_NAMES for w in weights[k]]
    dev = ws[0].device
    for i, w in enumerate(ws):
        _need_cuda(w, WEIGHT_NAMES[i // G])
    f32 = dict(dtype=torch.float32, device=dev)
    S = max(T - 1, 1)
    TILE = tile_rows()
    tiles = (B + TILE - 1) // TILE
    cur = torch.cuda.current_stream(dev)
    run = stream if stream is not None else cur
    out = PreparedSequence()
    with torch.cuda.device(dev):
        lib = _prepare(dev)
        if stream is not None:
            stream.wait_stream(cur)                      # the weights (e.g. an optimizer step) are ordered before us
        with torch.cuda.stream(run):
            # h_t lives at step index t+1 of H; H[:, 0] = 0 is "h_{-1}", so H[:, :S] are the GRU's previous states
            out.H = torch.empty((G, S + 1, tiles, HID, TILE), **f32)
            out.flags = torch.empty((S * G + 1,), dtype=torch.int32, device=dev)
            out.work = torch.empty(int(lib.biear_single_workspace_floats(N) if single
                                       else lib.biear_adaptive_workspace_floats(G, N)), **f32)
            out.seed_dev = _captured_seed(dev) if (training and torch.cuda.is_current_stream_capturing()) else None
            out.x_ready = torch.empty((G * B * T + 4,), dtype=torch.int32, device=dev) \
                if (streamed_spectra and launch and not single) else None
            prm = _lib.SeqParams()
            prm.G, prm.E, prm.B, prm.T, prm.N, prm.F, prm.Kin = G, E, B, T, N, 2, ws[0].shape[1]
            _fill(prm, workspace=out.work, H=out.H, flags=out.flags, x_ready=out.x_ready)
            for i, name in enumerate(WEIGHT_NAMES):
                arr = getattr(prm, name)
                for g in range(G):
                    arr[g] = ws[i * G + g].data_ptr()
            from ctypes import byref
            if launch:
                _lib.check(lib.biear_adaptive_prepare(byref(prm), c_void_p(run.cuda_stream)), "biear_adaptive_prepare")
            else:
                out.H.fill_(float("nan"))
                out.flags.fill_(1)
                out.work.fill_(float("nan"))
            out.launched = launch
            out.event = torch.cuda.Event()
            out.event.record(run)
    out.stream = run
    out.key = (G, E, B, T, N, tuple(w.data_ptr() for w in ws))
    return out


class AdaptiveSequence(torch.autograd.Function):
    """The whole 19-frame Q recurrence of the dual front-end as one autograd node.

    forward : ONE persistent cluster kernel (biear_adaptive_fwd) carries every 32-row tile through all frames;
    backward: ONE persistent cluster kernel (biear_adaptive_bwd) runs the chain t = T-2 .. 0 and leaves the
              per-sample pre-activation gradients, from which biear_ctrl_wgrad forms the weight gradients.
    Inputs : xr (E*B,T,F,2), fc/q0/dq (N), then the 14 weight tensors of every controller, name-major
             (w_ih of controller 0, w_ih of controller 1, w_hh of controller 0, ...): the parameters themselves, no
             stacking copy; the gradients come back in the same order.
    Outputs: PER EAR (B,T,N) views of the contiguous (E*B,T,N) result buffers, ear-major:
             Y_0..Y_{E-1}, Q_0.., phase_0.. [empty when want_phase is False], logY_0.. [clamp(log(Y + 1e-8), +-12);
             empty when want_logy is False].  Returning the ears as separate outputs means autograd hands the backward
             the per-ear gradients as they are (the C ABI takes one pointer per ear) instead of scattering them into
             zero-filled full-size tensors and adding those.
    """

    @staticmethod
    def forward(ctx, xr, fc, q0, dq, relative, training, want_phase, cutoff, df, seed, strict, want_logy, G, E, prep, *weights):
        ctx.set_materialize_grads(False)
        _need_cuda(xr, "X")
        dev = xr.device
        rows, T, F, _ = xr.shape
        N = fc.numel()
        assert len(weights) == G * len(WEIGHT_NAMES) and 1 <= G <= _lib.MAX_CTRL and (E == G or (G == 1 and E == 2))
        B = rows // E
        weights = tuple(w.detach().contiguous() for w in weights)
        for i, w in enumerate(weights):
            _need_cuda(w, WEIGHT_NAMES[i // G])
        Kin = weights[0].shape[1]
        need_grad = any(ctx.needs_input_grad[15:])
        f32 = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            lib = _prepare(dev)
            Y = torch.empty((rows, T, N), **f32)
            Q = torch.empty((G * B, T, N), **f32)
            D = torch.empty((G * B, T, N), **f32)
            P = torch.empty((rows, T, N), **f32) if want_phase else None
            LX = torch.empty((rows, T, N), **f32) if want_logy else None
            dY = torch.empty((rows, T, N), **f32)
            dP = torch.empty((rows, T, N), **f32) if want_phase else None
            S = max(T - 1, 1)
            TILE = tile_rows()
            tiles = (B + TILE - 1) // TILE
            wdict = {name: list(weights[i * G:(i + 1) * G]) for i, name in enumerate(WEIGHT_NAMES)}
            if prep is None:
                prep = adaptive_prepare(wdict, B, T, N, training, ears=E)
            elif prep.key != (G, E, B, T, N, tuple(w.data_ptr() for w in weights)):
                raise ValueError("biear_b200: adaptive_prepare was called for another geometry / other weights")
            cur = torch.cuda.current_stream(dev)
            if prep.stream != cur:                       # prepared on a forked stream: join it here
                cur.wait_event(prep.event)
                for t_ in (prep.work, prep.H, prep.flags, prep.seed_dev, prep.x_ready):
                    if t_ is not None:
                        t_.record_stream(cur)
            H, flags, work, seed_dev = prep.H, prep.flags, prep.work, prep.seed_dev
            sv = {k: torch.empty((G, S, tiles, d, TILE), **f32) for k, d in
                  (("gates", 4 * HID), ("xh1", HID), ("d1", HID), ("xh2", HID), ("d2", HID), ("rstd", 2),
                   ("yc", N if E == G else 4 * N))}       # single controller: the whole 4N-wide controller input
            prm = _lib.SeqParams()
            prm.G, prm.E, prm.B, prm.T, prm.N, prm.F, prm.Kin = G, E, B, T, N, F, Kin
            prm.relative, prm.training, prm.seed, prm.force_strict = int(relative), int(training), int(seed), int(strict)
            prm.prepared = int(prep.launched)
            prm.df, prm.cutoff, prm.q_min, prm.q_max = float(df), float(cutoff), 0.05, 30.0
            _fill(prm, fc=fc, q0=q0, dq=dq, X=xr, Y=Y, phase=P, dYdQ=dY, dPdQ=dP, Q=Q, delta=D, flags=flags,
                  workspace=work, H=H, seed_ptr=seed_dev, logY=LX, x_ready=prep.x_ready, **sv)
            for i, name in enumerate(WEIGHT_NAMES):
                arr = getattr(prm, name)
                for g in range(G):
                    arr[g] = weights[i * G + g].data_ptr()
            from ctypes import byref
            _lib.check(lib.biear_adaptive_fwd(byref(prm), _stream(dev)), "biear_adaptive_fwd")
        ctx.prm = prm
        # owners of every pointer in prm (the per-ear outputs are views of Y / Q / P / LX, which are not outputs
        # themselves, so holding them here closes no reference cycle)
        ctx.keep = (xr, fc, q0, dq, weights, Y, Q, P, LX, D, dY, dP, sv, flags, work, H, seed_dev, prep.x_ready)
        ctx.has_phase = P is not None
        ctx.has_logy = LX is not None
        ctx.dims = (G, E, B, T, N, Kin, tiles, TILE)
        empty = Y.new_empty(0)
        ears = lambda t, k=E: tuple(t[e * B:(e + 1) * B] for e in range(k)) if t is not None else tuple(empty for _ in range(k))
        outs = ears(Y) + ears(Q, G) + ears(P) + ears(LX)
        nd = [o for o in outs if o.numel() == 0]
        if not need_grad:
            nd = list(outs)
        if nd:
            ctx.mark_non_differentiable(*nd)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        from ctypes import byref
        xr, fc, q0, dq, weights, Y, Q, P, LX, D, dY, dP, sv, flags, work, H, _seed_dev, _x_ready = ctx.keep
        G, E, B, T, N, Kin, tiles, TILE = ctx.dims
        none13 = (None,) * 15
        gY, gQ, gP, gLX = grads[:E], grads[E:E + G], grads[E + G:2 * E + G], grads[2 * E + G:3 * E + G]
        gY, gQ, gP, gLX = list(gY), list(gQ), list(gP), list(gLX)
        if not ctx.has_phase:
            gP = [None] * E
        if not ctx.has_logy:
            gLX = [None] * E
        if T < 2 or all(g is None for g in gY + gQ + gP + gLX):
            return none13 + (None,) * (G * len(WEIGHT_NAMES))
        with timing.span("frontend.backward", B):
            return AdaptiveSequence._backward(ctx, gY, gQ, gP, gLX)

    @staticmethod
    def _backward(ctx, gY, gQ, gP, gLX):
        from ctypes import byref
        xr, fc, q0, dq, weights, Y, Q, P, LX, D, dY, dP, sv, flags, work, H, _seed_dev, _x_ready = ctx.keep
        G, E, B, T, N, Kin, tiles, TILE = ctx.dims
        none13 = (None,) * 15
        dev = Y.device
        f32 = dict(dtype=torch.float32, device=dev)
        S = T - 1
        cont = lambda lst: [g.contiguous() if g is not None else None for g in lst]
        gY, gQ, gP, gLX = cont(gY), cont(gQ), cont(gP), cont(gLX)
        with torch.cuda.device(dev):
            lib = _prepare(dev)
            wk = {k: torch.empty((G, S, tiles, d, TILE), **f32) for k, d in
                  (("GG", 4 * HID), ("G_a1", HID), ("G_v1", HID), ("G_a2", HID), ("G_v2", HID), ("G_pre", N))}
            prm = ctx.prm
            _fill(prm, **wk)
            for name, lst in (("gY", gY), ("gQ", gQ), ("gP", gP), ("gLogY", gLX)):
                arr = getattr(prm, name)
                for g in range(_lib.MAX_CTRL):
                    arr[g] = lst[g].data_ptr() if (g < len(lst) and lst[g] is not None) else None
            _lib.check(lib.biear_adaptive_bwd(byref(prm), _stream(dev)), "biear_adaptive_bwd")
            for name in ("gY", "gQ", "gP", "gLogY"):
                arr = getattr(prm, name)
                for g in range(_lib.MAX_CTRL):
                    arr[g] = None

            # ---- weight gradients: split-K GEMMs over all (step, tile) chunks, off the serial chain ----
            K = S * tiles
            fl = lambda t: t.view(G, K, t.shape[3], TILE)
            GG = fl(wk["GG"])
            Hk = H.view(G, (S + 1) * tiles, HID, TILE)      # chunk = (step, tile); group stride covers S+1 steps
            h_prev, h_cur = Hk[:, :K], Hk[:, tiles:]
            d_w_hh = torch.empty((G, 3 * HID, HID), **f32)
            d_b_hh = torch.empty((G, 3 * HID), **f32)
            fold = Kin == 2 * N                   # feat = [yc, 0.2 yc.detach()]: dW_ih[:, N:] = 0.2 dW_ih[:, :N]
            d_w_ih = torch.empty((G, 3 * HID, Kin), **f32)
            ih_out = (d_w_ih[:, :, :N], None, d_w_ih[:, :, N:], 0.2) if fold else (d_w_ih[:, :, :N],)
            yc_t = fl(sv["yc"])
            if not fold:   # single controller: weight_ih is (384, 4N) over [cL, mL, cR, mR]; one job per N-wide column block
                assert Kin == 4 * N and yc_t.shape[2] == 4 * N
                ctrl_wgrad([(GG, 3 * HID, yc_t[:, :, j * N:(j + 1) * N], N, K, False, d_w_ih[:, :, j * N:(j + 1) * N])
                            for j in range(1, 4)])
            (a, d_b_ih), _, _, (d_w1, d_b1), (d_w2, d_b2), (d_w3, d_b3), (d_g1, d_be1), (d_g2, d_be2) = ctrl_wgrad([
                (GG, 3 * HID, yc_t[:, :, :N], N, K, True) + ih_out,                               # dL/dW_ih (both halves), b_ih
                (GG, 2 * HID, h_prev, HID, K, True, d_w_hh[:, :2 * HID], d_b_hh[:, :2 * HID]),     # r, z rows of W_hh / b_hh
                (GG[:, :, 3 * HID:], HID, h_prev, HID, K, True, d_w_hh[:, 2 * HID:], d_b_hh[:, 2 * HID:]),   # n rows: dL/d(W_hn h + b_hn)
                (fl(wk["G_a1"]), HID, h_cur, HID, K, True),
                (fl(wk["G_a2"]), HID, fl(sv["d1"]), HID, K, True),
                (fl(wk["G_pre"]), N, fl(sv["d2"]), HID, K, True),
                (fl(wk["G_v1"]), HID, fl(sv["xh1"]), 0, K, True),                                  # LayerNorm 1 weight / bias
                (fl(wk["G_v2"]), HID, fl(sv["xh2"]), 0, K, True),
            ])
        stacked = (d_w_ih, d_w_hh, d_b_ih, d_b_hh, d_w1, d_b1, d_g1, d_be1, d_w2, d_b2, d_g2, d_be2, d_w3, d_b3)
        grads = tuple(t[g] if t is not None else None for t in stacked for g in range(G))
        return none13 + grads


def adaptive_sequence(xr, fc, q0, dq, weights, relative: bool, training: bool, want_phase: bool,
                      cutoff: float, df: float, seed: int = 0, strict: bool = False, want_logy: bool = False,
                      prep: Optional[PreparedSequence] = None, ears: Optional[int] = None):
    """weights: dict name -> list of the G controllers' tensors (WEIGHT_NAMES).  Returns Y, Q, phase|None[, logY], each a
    LIST with one (B,T,N) tensor per ear / controller.
    ears=2 with ONE controller selects the single-controller front-end (model_torch.py:695-776): x holds both ears,
    Y / phase / logY come back per ear, Q once.
    strict=True skips the fast pass and runs the batch-global-fallback replay pass only (testing).
    prep: result of adaptive_prepare (same weights / geometry) issued earlier, typically on a forked stream."""
    G = len(weights[WEIGHT_NAMES[0]])
    E = G if ears is None else int(ears)
    outs = AdaptiveSequence.apply(xr, fc, q0, dq, relative, training, want_phase, cutoff, df, seed, strict,
                                  want_logy, G, E, prep, *[w for k in WEIGHT_NAMES for w in weights[k]])
    y, q, ph, lx = list(outs[:E]), list(outs[E:E + G]), list(outs[E + G:2 * E + G]), list(outs[2 * E + G:3 * E + G])
    if want_logy:
        return y, q, (ph if want_phase else None), lx
    return y, q, (ph if want_phase else None)


# ------------------------------------------------------------------------------------------------
# back-end: the per-sector heads (csrc/heads.cu)
# ------------------------------------------------------------------------------------------------
HEAD_TENSOR_NAMES = ("shared.0.weight", "shared.0.bias") + tuple(
    f"{b}.{i}.{w}" for b in ("sound", "aoa", "dist") for i in (0, 2, 4) for w in ("weight", "bias"))
_head_tables = {}   # (device index, data_ptrs) -> int64 device tensor of parameter pointers


def _head_table(params, dev) -> torch.Tensor:
    """Device table of the heads' parameter pointers (rebuilt only when a parameter's storage moves)."""
    ptrs = tuple(p.data_ptr() for p in params)
    key = (dev.index, ptrs)
    tab = _head_tables.get(key)
    if tab is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("biear_b200: run the model once eagerly (warm-up) before capturing it in a CUDA graph")
        if len(_head_tables) > 16:
            _head_tables.clear()
        tab = torch.tensor(ptrs, dtype=torch.int64).to(dev)
        _head_tables[key] = tab
    return tab


class SectorHeads(torch.autograd.Function):
    """All sector heads (model_torch.py:869-906 and the loop at :941-955 / :1096-1110) as one forward and one backward
    launch.  Inputs: body (B,D), then the S * 20 head parameters in state-dict order (the nn.Linear parameters
    themselves).  Outputs: sound logits (B,S), aoa in (0,1) (B,S), distance logits (B,S,C)."""

    @staticmethod
    def forward(ctx, body, S, C, training, seed, *params):
        _need_cuda(body, "body")
        dev = body.device
        body = body.contiguous()
        B, D = body.shape
        with torch.cuda.device(dev):
            lib = _prepare(dev)
            per = int(lib.biear_heads_tensors_per_head())
            assert len(params) == S * per
            ps = tuple(p.detach() for p in params)
            for p_ in ps:
                _need_cuda(p_, "head parameter")
                assert p_.is_contiguous()
            table = _head_table(ps, dev)
            f32 = dict(dtype=torch.float32, device=dev)
            sound, aoa, dist = torch.empty((B, S), **f32), torch.empty((B, S), **f32), torch.empty((B, S, C), **f32)
            seed_dev = _captured_seed(dev) if (training and torch.cuda.is_current_stream_capturing()) else None
            prm = _lib.HeadsParams()
            prm.B, prm.S, prm.D, prm.C, prm.training, prm.seed = B, S, D, C, int(training), int(seed)
            _fill(prm, seed_ptr=seed_dev, body=body, wptr=table, sound=sound, aoa=aoa, dist=dist)
            from ctypes import byref
            _lib.check(lib.biear_heads_fwd(byref(prm), _stream(dev)), "biear_heads_fwd")
        ctx.prm = prm
        ctx.keep = (body, ps, table, seed_dev, sound, aoa, dist)
        ctx.shapes = [tuple(p_.shape) for p_ in ps]
        return sound, aoa, dist

    @staticmethod
    def backward(ctx, g_sound, g_aoa, g_dist):
        from ctypes import byref
        body, ps, table, seed_dev, sound, aoa, dist = ctx.keep
        dev = body.device
        B, D = body.shape
        prm = ctx.prm
        S, C = prm.S, prm.C
        cont = lambda g: g.contiguous() if g is not None else None
        g_sound, g_aoa, g_dist = cont(g_sound), cont(g_aoa), cont(g_dist)
        with torch.cuda.device(dev):
            lib = _prepare(dev)
            flat = int(lib.biear_heads_flat_floats(D, C))
            tiles = (B + int(lib.biear_heads_tile_rows()) - 1) // int(lib.biear_heads_tile_rows())
            f32 = dict(dtype=torch.float32, device=dev)
            d_body_part, dw_part = torch.empty((S, B, D), **f32), torch.empty((tiles, S, flat), **f32)
            d_body, dw = torch.empty((B, D), **f32), torch.empty((S, flat), **f32)
            _fill(prm, g_sound=g_sound, g_aoa=g_aoa, g_dist=g_dist, d_body_part=d_body_part, dw_part=dw_part,
                  d_body=d_body, dw=dw)
            _lib.check(lib.biear_heads_bwd(byref(prm), _stream(dev)), "biear_heads_bwd")
            _fill(prm, g_sound=None, g_aoa=None, g_dist=None, d_body_part=None, dw_part=None, d_body=None, dw=None)
        grads = []
        per = len(ctx.shapes) // S
        for s_ in range(S):
            off = 0
            for j in range(per):
                shp = ctx.shapes[s_ * per + j]
                n = 1
                for d_ in shp:
                    n *= d_
                grads.append(dw[s_, off:off + n].view(shp))
                off += n
            assert off == flat
        return (d_body, None, None, None, None) + tuple(grads)


def sector_heads(body: torch.Tensor, head_modules, training: bool):
    """head_modules: the S SubHead modules (shared / sound / aoa / dist as in model_torch.py:869-906)."""
    params = []
    for h in head_modules:
        sd = dict(h.named_parameters())
        params += [sd[k] for k in HEAD_TENSOR_NAMES]
    S = len(head_modules)
    C = head_modules[0].dist[4].out_features
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
    return SectorHeads.apply(body, S, C, bool(training), seed, *params)



# ------------------------------------------------------------------------------------------------
# back-end: the recurrence of one GRU layer (csrc/gru.cu)
# ------------------------------------------------------------------------------------------------
@lru_cache(maxsize=None)
def gru_supported(hidden: int) -> bool:
    """Can the persistent GRU-layer kernels take this hidden width (shared-memory budget, H % 4 == 0)?"""
    return bool(_lib.load().biear_gru_supported(int(hidden)))


_gru_sides = {}   # (device index, id of the stream the layer runs on) -> two side streams for its parameter gradients


def _gru_side_streams(dev, cur):
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), cur.cuda_stream)
    st = _gru_sides.get(key)
    if st is None:
        if len(_gru_sides) > 64:
            _gru_sides.clear()
        st = _gru_sides[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return st


class GruLayer(torch.autograd.Function):
    """One ``nn.GRU`` layer (batch_first, h_0 = 0; model_torch.py:834-835, 842-843) with the 19-step recurrence as ONE
    launch forward and ONE backward instead of the library's per-step GEMM + cell kernels.  The input projection and the
    weight gradients are plain GEMMs over all frames at once (library calls).  Inputs: x (B,T,I) and the layer's own
    parameters (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0); output: the hidden sequence (B,T,H)."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh):
        from ctypes import byref
        _need_cuda(x, "x")
        dev = x.device
        B, T, I = x.shape
        H = w_hh.shape[1]
        w_ih, w_hh, b_ih, b_hh = (p.detach() for p in (w_ih, w_hh, b_ih, b_hh))
        for p_ in (w_ih, w_hh, b_ih, b_hh):
            _need_cuda(p_, "GRU parameter")
        assert w_ih.shape == (3 * H, I) and w_hh.shape == (3 * H, H)
        with torch.cuda.device(dev):
            lib = _prepare(dev)
            f32 = dict(dtype=torch.float32, device=dev)
            gi = torch.addmm(b_ih, x.reshape(B * T, I), w_ih.t())                   # (B*T, 3H)
            h_seq, h_prev = torch.empty((B, T, H), **f32), torch.empty((B, T, H), **f32)
            gates = torch.empty((B, T, 4, H), **f32)
            ws = torch.empty(int(lib.biear_gru_workspace_floats(H)), **f32)
            prm = _lib.GruParams()
            prm.B, prm.T, prm.H, prm.I = B, T, H, I
            _fill(prm, gi=gi, w_hh=w_hh, b_hh=b_hh, h_seq=h_seq, h_prev=h_prev, gates=gates, workspace=ws)
            _lib.check(lib.biear_gru_fwd(byref(prm), _stream(dev)), "biear_gru_fwd")
        ctx.prm = prm
        ctx.keep = (x, w_ih, w_hh, h_prev, gates)           # what the backward reads (the output itself is not among it)
        return h_seq

    @staticmethod
    def backward(ctx, g_seq):
        from ctypes import byref
        x, w_ih, w_hh, h_prev, gates = ctx.keep
        dev = x.device
        B, T, I = x.shape
        H = w_hh.shape[1]
        prm = ctx.prm
        g_seq = g_seq.contiguous()
        with torch.cuda.device(dev):
            lib = _prepare(dev)
            f32 = dict(dtype=torch.float32, device=dev)
            dgi, dgh = torch.empty((B * T, 3 * H), **f32), torch.empty((B * T, 3 * H), **f32)
            _fill(prm, dh_seq=g_seq, dgi=dgi, dgh=dgh)
            _lib.check(lib.biear_gru_bwd(byref(prm), _stream(dev)), "biear_gru_bwd")
            _fill(prm, dh_seq=None, dgi=None, dgh=None)
            need = ctx.needs_input_grad
            # dL/dx continues the chain on this stream.  While a step is being captured, the four parameter gradients (two
            # GEMMs with a 4864-long contraction and few output tiles, two column sums) go to two side streams and become
            # parallel branches of the graph, joined before the node returns (full step 2.51 -> 2.41 ms); issued eagerly
            # the launches are host-bound and the extra event calls cost more than the overlap gives, so they stay in line.
            cur = torch.cuda.current_stream(dev)
            sides = _gru_side_streams(dev, cur) if torch.cuda.is_current_stream_capturing() else (cur, cur)
            dw_ih = torch.empty((3 * H, I), **f32) if need[1] else None
            db_ih = torch.empty((3 * H,), **f32) if need[3] else None
            dw_hh = torch.empty((3 * H, H), **f32) if need[2] else None
            db_hh = torch.empty((3 * H,), **f32) if need[4] else None
            used = []
            if need[1] or need[3]:
                if sides[0] is not cur:
                    sides[0].wait_stream(cur)
                with torch.cuda.stream(sides[0]):
                    if need[1]:
                        torch.mm(dgi.t(), x.reshape(B * T, I), out=dw_ih)
                    if need[3]:
                        torch.sum(dgi, dim=0, out=db_ih)
                used.append(sides[0])
            if need[2] or need[4]:
                if sides[1] is not cur:
                    sides[1].wait_stream(cur)
                with torch.cuda.stream(sides[1]):
                    if need[2]:
                        torch.mm(dgh.t(), h_prev.view(B * T, H), out=dw_hh)
                    if need[4]:
                        torch.sum(dgh, dim=0, out=db_hh)
                used.append(sides[1])
            dx = (dgi @ w_ih).view(B, T, I) if need[0] else None
            for st in used:
                if st is not cur:
                    cur.wait_stream(st)
        return dx, dw_ih, dw_hh, db_ih, db_hh


def gru_layer(x: torch.Tensor, gru: "torch.nn.GRU") -> torch.Tensor:
    """Hidden sequence of a one-layer, unidirectional, batch_first ``nn.GRU`` with zero initial state."""
    assert gru.num_layers == 1 and not gru.bidirectional and gru.batch_first and gru.bias
    return GruLayer.apply(x.contiguous(), gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0)
