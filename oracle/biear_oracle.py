"""CPU oracle for the BiEAR active-mode binaural front-end.

TEST INFRASTRUCTURE ONLY.  Nothing under ``biear_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker (or as the timed CPU port), never as a product code path.

This is a functional (weights-passed-as-dict) restatement of the reference's
algorithm, written from SURVEY.md Appendix A and the reference sources it
cites.  It is dtype-generic: call it with float64 tensors to get the
"truth" the fp32 results are conditioned against, or with float32 tensors to
mirror the reference's own arithmetic.  Autograd supplies the backward pass,
exactly as it does in the reference (which has no explicit backward code).

Parity status: PINNED against the reference itself.  ``tests/golden/make_golden.py``
imports ``/root/reference/model_torch.py`` and ``utils.py`` in the build
container, runs them on seeded inputs and commits the outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
here against those vectors.  (The reference ships no tests or golden vectors of
its own)

Reference map (paths relative to /root/reference):
  erb_constants            model_torch.py:19-34
  deltaq_profile           model_torch.py:36-51
  frame_clip               model_torch.py:289-312
  stft_frames              model_torch.py:232, 334-335
  band_weights             model_torch.py:340-343
  band_energy              model_torch.py:345-346
  gru_cell / q_mlp         model_torch.py:256-267, 366-367 (torch.nn.GRU / Sequential semantics)
  next_q                   model_torch.py:369-380
  adaptive_fb_forward      model_torch.py:314-386
  fixed_fb_forward         model_torch.py:451-487
  auralnet_fb_forward      model_torch.py:161-195
  single_controller_forward model_torch.py:695-776
  subband_phase            model_torch.py:1039-1063
  log_energy               model_torch.py:1080-1083
  cc_feature               utils.py:390-420
  dq_closed_form           (derived) checked against autograd
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

Q_MIN = 0.05
Q_MAX = 30.0


# ----------------------------------------------------------------------------
# constants
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class FrontEndConfig:
    """Static shape/constant set of one front-end instance (model_torch.py:500-515)."""
    fs: int = 16000
    timesteps: int = 19
    n_fft: int = 1024
    n_bands: int = 100
    fmin: float = 50.0
    fmax: Optional[float] = None
    hop_ratio: float = 1.0
    deltaq_base: float = 2.0
    deltaq_low: float = 0.5
    deltaq_high: float = 1.0
    deltaq_mode: str = "absolute"

    @property
    def win(self) -> int:
        return int(round(self.fs / self.timesteps))

    @property
    def hop(self) -> int:
        return max(1, int(round(self.win * self.hop_ratio)))

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def fmax_eff(self) -> float:
        return self.fs / 2 * 0.9 if self.fmax is None else self.fmax


def _erb_rate(f):
    return 21.4 * np.log10(4.37 * f / 1000.0 + 1.0)


def erb_constants(n_bands=100, fmin=50.0, fmax=7200.0, erb_factor=1.019):
    """Centre frequencies uniform in ERB-rate and their Q0 = fc / (1.019 ERB(fc)); float64 numpy."""
    e = np.linspace(_erb_rate(fmin), _erb_rate(fmax), n_bands)
    fc = (10.0 ** (e / 21.4) - 1.0) * 1000.0 / 4.37
    q0 = fc / (erb_factor * 24.7 * (4.37 * fc / 1000.0 + 1.0))
    return fc, q0


def deltaq_profile(fc32: np.ndarray, base=2.0, low=0.5, high=1.0) -> np.ndarray:
    """Per-band Delta-Q: linear in normalised ERB-rate of the *float32* fc, clamp >= 1e-3, float32."""
    e = _erb_rate(fc32.astype(np.float32))
    e = (e - e.min()) / (e.max() - e.min() + 1e-12)
    mult = (low + (high - low) * e).astype(np.float32)
    return np.maximum(np.float32(base) * mult, np.float32(1e-3)).astype(np.float32)


def constants(cfg: FrontEndConfig, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Buffers the reference registers: fc, Q0, f_fft, deltaQ_vec, hann window (all stored as fp32 first)."""
    fc64, q064 = erb_constants(cfg.n_bands, cfg.fmin, cfg.fmax_eff)
    fc32 = fc64.astype(np.float32)
    q032 = q064.astype(np.float32)
    dq = deltaq_profile(fc32, cfg.deltaq_base, cfg.deltaq_low, cfg.deltaq_high)
    f_fft = torch.linspace(0, cfg.fs / 2, cfg.n_bins)  # fp32, as the reference
    win = torch.hann_window(cfg.win)  # periodic, fp32
    return {
        "fc": torch.from_numpy(fc32).to(dtype),
        "Q0": torch.from_numpy(q032).to(dtype),
        "deltaQ_vec": torch.from_numpy(dq).to(dtype),
        "f_fft": f_fft.to(dtype),
        "win_fn": win.to(dtype),
    }


# ----------------------------------------------------------------------------
# STFT
# ----------------------------------------------------------------------------
def frame_clip(wav: torch.Tensor, cfg: FrontEndConfig) -> torch.Tensor:
    """(B,Nsamp) -> (B,T,win): pad/truncate to fs samples, hop-strided frames, pad/truncate to T frames."""
    if wav.dim() != 2:
        raise ValueError(f"Expected wav_1s (B,N), got {tuple(wav.shape)}")
    n = wav.shape[1]
    target = cfg.fs
    if n < target:
        wav = torch.nn.functional.pad(wav, (0, target - n))
    else:
        wav = wav[:, :target]
    if target < cfg.win:
        wav = torch.nn.functional.pad(wav, (0, cfg.win - target))
    frames = wav.unfold(1, cfg.win, cfg.hop)
    t = cfg.timesteps
    if frames.shape[1] >= t:
        frames = frames[:, :t]
    else:
        frames = torch.nn.functional.pad(frames, (0, 0, 0, t - frames.shape[1]))
    return frames


def stft_frames(wav: torch.Tensor, cfg: FrontEndConfig, win_fn: torch.Tensor) -> torch.Tensor:
    """(B,Nsamp) -> X (B,T,F) complex: Hann-windowed frames, rfft zero-padded/truncated to n_fft."""
    return torch.fft.rfft(frame_clip(wav, cfg) * win_fn, n=cfg.n_fft)


# ----------------------------------------------------------------------------
# band stage
# ----------------------------------------------------------------------------
def band_weights(q: torch.Tensor, fc: torch.Tensor, f_fft: torch.Tensor, sanitize=True) -> torch.Tensor:
    """Q (...,N) -> row-normalised Gaussian weights W (...,N,F)."""
    bw = (fc / (q + 1e-8)).unsqueeze(-1) + 1e-8
    g = torch.exp(-0.5 * ((f_fft - fc.unsqueeze(-1)) / bw) ** 2)
    w = g / (g.sum(dim=-1, keepdim=True) + 1e-8)
    if sanitize:
        w = torch.nan_to_num(w, nan=0.0, posinf=0.0, neginf=0.0)
    return w


def band_energy(xmag: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """|X| (B,F), W (B,N,F) -> Y (B,N)."""
    y = torch.einsum("bf,bnf->bn", xmag, w)
    return torch.nan_to_num(y, nan=0.0, posinf=0.0, neginf=0.0)


# ----------------------------------------------------------------------------
# controller
# ----------------------------------------------------------------------------
def gru_cell(p: Dict[str, torch.Tensor], x: torch.Tensor, h: Optional[torch.Tensor]) -> torch.Tensor:
    """One torch.nn.GRU step (gate order r,z,n; n uses r*(W_hn h + b_hn))."""
    w_ih, w_hh = p["q_rnn.weight_ih_l0"], p["q_rnn.weight_hh_l0"]
    b_ih, b_hh = p["q_rnn.bias_ih_l0"], p["q_rnn.bias_hh_l0"]
    hid = w_hh.shape[1]
    if h is None:
        h = x.new_zeros(x.shape[0], hid)
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    i_r, i_z, i_n = gi.split(hid, dim=1)
    h_r, h_z, h_n = gh.split(hid, dim=1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def _layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def q_mlp(p: Dict[str, torch.Tensor], h: torch.Tensor, drop_masks=None) -> torch.Tensor:
    """Linear-LN-SiLU-Dropout x2 + Linear (Sequential indices 0,1,4,5,8).  drop_masks: optional
    pair of pre-scaled keep masks (train mode); None = eval mode."""
    a = h @ p["q_out.0.weight"].t() + p["q_out.0.bias"]
    a = torch.nn.functional.silu(_layer_norm(a, p["q_out.1.weight"], p["q_out.1.bias"]))
    if drop_masks is not None:
        a = a * drop_masks[0]
    a = a @ p["q_out.4.weight"].t() + p["q_out.4.bias"]
    a = torch.nn.functional.silu(_layer_norm(a, p["q_out.5.weight"], p["q_out.5.bias"]))
    if drop_masks is not None:
        a = a * drop_masks[1]
    return a @ p["q_out.8.weight"].t() + p["q_out.8.bias"]


def next_q(delta: torch.Tensor, q0: torch.Tensor, dq: torch.Tensor, mode: str) -> torch.Tensor:
    if mode == "relative":
        q = q0 * (1.0 + dq * delta)
    else:
        q = q0 + dq * delta
    return torch.clamp(q, Q_MIN, Q_MAX)


# ----------------------------------------------------------------------------
# filterbanks
# ----------------------------------------------------------------------------
def adaptive_fb_forward(wav, p, cfg: FrontEndConfig, c=None, q_hook=None, taps=None):
    """Monaural adaptive FB.  Returns Y (B,T,N), Q (B,T,N), X (B,T,F) complex.

    q_hook(t, Q_t) -> Q_t lets tests tap (or make a leaf of) the per-frame Q actually used.
    taps: optional list that receives the pre-tanh controller output of every step (grad retained).
    """
    dtype = wav.dtype
    c = c or constants(cfg, dtype)
    x_all = stft_frames(wav, cfg, c["win_fn"])
    b = wav.shape[0]
    q = c["Q0"].unsqueeze(0).expand(b, -1)
    h = None
    ys, qs = [], []
    for t in range(cfg.timesteps):
        if q_hook is not None:
            q = q_hook(t, q)
        w = band_weights(q, c["fc"], c["f_fft"])
        y = band_energy(x_all[:, t].abs(), w)
        ys.append(y)
        qs.append(q)
        yc = torch.log1p(torch.clamp(y, min=0.0))
        feat = torch.cat([yc, 0.2 * yc.detach()], dim=-1)
        h = gru_cell(p, feat, h)
        pre = q_mlp(p, h)
        if taps is not None:
            if pre.requires_grad:
                pre.retain_grad()
            taps.append(pre)
        delta = torch.tanh(pre)
        q = next_q(delta, c["Q0"], c["deltaQ_vec"], cfg.deltaq_mode)
        if not bool(torch.isfinite(q).all()):
            q = c["Q0"].unsqueeze(0).expand(b, -1)
            h = None
    return torch.stack(ys, 1), torch.stack(qs, 1), x_all


def fixed_fb_forward(wav, cfg: FrontEndConfig, c=None):
    """Monaural fixed FB: Q == clamp(Q0) for every frame."""
    c = c or constants(cfg, wav.dtype)
    x_all = stft_frames(wav, cfg, c["win_fn"])
    b = wav.shape[0]
    q = torch.clamp(c["Q0"], Q_MIN, Q_MAX).unsqueeze(0).expand(b, -1)
    ys = []
    for t in range(cfg.timesteps):
        w = band_weights(q, c["fc"], c["f_fft"])
        ys.append(band_energy(x_all[:, t].abs(), w))
    y = torch.stack(ys, 1)
    return y, q.unsqueeze(1).expand(-1, cfg.timesteps, -1), x_all


def auralnet_fb_forward(wav, cfg: FrontEndConfig, c=None):
    """AuralNet fixed FB: one shared (N,F) weight matrix, all frames in one contraction."""
    c = c or constants(cfg, wav.dtype)
    x_all = stft_frames(wav, cfg, c["win_fn"])
    q = torch.clamp(c["Q0"], Q_MIN, Q_MAX)
    w = band_weights(q, c["fc"], c["f_fft"])
    y = torch.einsum("btf,nf->btn", x_all.abs(), w)
    return torch.nan_to_num(y, nan=0.0, posinf=0.0, neginf=0.0)


def single_controller_forward(wav_l, wav_r, p, cfg: FrontEndConfig, c=None):
    """One controller, shared Q for both ears, carried Y memory (beta=0.8)."""
    c = c or constants(cfg, wav_l.dtype)
    xl = stft_frames(wav_l, cfg, c["win_fn"])
    xr = stft_frames(wav_r, cfg, c["win_fn"])
    b = wav_l.shape[0]
    q = c["Q0"].unsqueeze(0).expand(b, -1)
    h = None
    ml = wav_l.new_zeros(b, cfg.n_bands)
    mr = wav_l.new_zeros(b, cfg.n_bands)
    yls, yrs, qs = [], [], []
    for t in range(cfg.timesteps):
        w = band_weights(q, c["fc"], c["f_fft"])
        yl = band_energy(xl[:, t].abs(), w)
        yr = band_energy(xr[:, t].abs(), w)
        yls.append(yl)
        yrs.append(yr)
        qs.append(q)
        cl = torch.log1p(torch.clamp(yl, min=0.0))
        cr = torch.log1p(torch.clamp(yr, min=0.0))
        feat = torch.cat([cl, ml, cr, mr], dim=-1)
        h = gru_cell(p, feat, h)
        delta = torch.tanh(q_mlp(p, h))
        q = next_q(delta, c["Q0"], c["deltaQ_vec"], cfg.deltaq_mode)
        if not bool(torch.isfinite(q).all()):
            q = c["Q0"].unsqueeze(0).expand(b, -1)
            h = None
        ml = 0.8 * ml + 0.2 * cl.detach()
        mr = 0.8 * mr + 0.2 * cr.detach()
    qq = torch.stack(qs, 1)
    return torch.stack(yls, 1), torch.stack(yrs, 1), qq, qq, xl, xr


def binaural_forward(wav_l, wav_r, p_l, p_r, cfg: FrontEndConfig, fixed=False, c=None):
    """Dual binaural FB: two independent monaural FBs -> (YL,YR,QL,QR,XL,XR)."""
    if fixed:
        yl, ql, xl = fixed_fb_forward(wav_l, cfg, c)
        yr, qr, xr = fixed_fb_forward(wav_r, cfg, c)
    else:
        yl, ql, xl = adaptive_fb_forward(wav_l, p_l, cfg, c)
        yr, qr, xr = adaptive_fb_forward(wav_r, p_r, cfg, c)
    return yl, yr, ql, qr, xl, xr


# ----------------------------------------------------------------------------
# features downstream of the FB
# ----------------------------------------------------------------------------
def subband_phase(x_all, q_all, f_fft, fc, eps_mag=1e-3):
    """phase (B,T,N) of Z = sum_f W(Q) X, W rebuilt per frame without nan_to_num."""
    out = []
    for t in range(x_all.shape[1]):
        w = band_weights(q_all[:, t], fc, f_fft, sanitize=False)
        z = torch.einsum("bnf,bf->bn", torch.complex(w, torch.zeros_like(w)), x_all[:, t])
        zn = z / torch.clamp(z.abs(), min=eps_mag)
        out.append(torch.atan2(zn.imag, zn.real))
    return torch.stack(out, 1)


def log_energy(y):
    return torch.clamp(torch.log(y + 1e-8), -12.0, 12.0)


def q_regularizers(q: torch.Tensor, q0: torch.Tensor):
    """The two Q regularisers of the training loss (train_biear.py:476-490) on Q = model.last_Q = (QL + QR) / 2
    (model_torch.py:1076-1078): returns (reg_q, reg_smooth); loss += REG_Q_W * reg_q + REG_SMOOTH_W * reg_smooth."""
    log_q = torch.log(q + 1e-8)
    log_q0 = torch.log(q0.view(1, 1, -1) + 1e-8)
    reg_q = ((log_q - log_q0) ** 2).mean()
    reg_smooth = ((log_q[:, :, 1:] - log_q[:, :, :-1]) ** 2).mean()
    return reg_q, reg_smooth


def cc_feature(left: np.ndarray, right: np.ndarray, fs=16000, num_lags=100, max_lag_ms=3.0) -> np.ndarray:
    """Broadband interaural cross-correlation cropped to +-max_lag, max-abs normalised, resampled to
    num_lags points by linear interpolation.  float64 arithmetic, float32 result."""
    l = left.astype(np.float64)
    r = right.astype(np.float64)
    l = l - l.mean()
    r = r - r.mean()
    n = len(l)
    max_lag_sec = max_lag_ms * 1e-3
    lag_idx = np.arange(-(n - 1), n)
    lag_sec = lag_idx / fs
    keep = np.logical_and(lag_sec >= -max_lag_sec, lag_sec <= max_lag_sec)
    ks = lag_idx[keep]
    cc = np.empty(len(ks), dtype=np.float64)
    for i, k in enumerate(ks):
        # c[k] = sum_n l[n+k] r[n]  (np.correlate(l, r, "full") convention)
        if k >= 0:
            cc[i] = np.dot(l[k:], r[: n - k])
        else:
            cc[i] = np.dot(l[: n + k], r[-k:])
    cc = cc / (np.max(np.abs(cc)) + 1e-8)
    target = np.linspace(-max_lag_sec, max_lag_sec, num_lags)
    return np.interp(target, lag_sec[keep], cc).astype(np.float32)


def cc_feature_batch(wav_l: np.ndarray, wav_r: np.ndarray, fs=16000, num_lags=100, max_lag_ms=3.0):
    return np.stack([cc_feature(a, b, fs, num_lags, max_lag_ms) for a, b in zip(wav_l, wav_r)])


# ----------------------------------------------------------------------------
# closed-form backward into Q (what the CUDA backward implements)
# ----------------------------------------------------------------------------
def band_moments(x_t: torch.Tensor, q_t: torch.Tensor, fc, f_fft):
    """Per-frame forward quantities and the three u^2-moments used by the closed-form dQ.
    x_t (B,F) complex, q_t (B,N).  Returns dict of (B,N) tensors (z, z2 complex)."""
    bw = (fc / (q_t + 1e-8)).unsqueeze(-1) + 1e-8
    u = (f_fft - fc.unsqueeze(-1)) / bw
    g = torch.exp(-0.5 * u * u)
    w = g / (g.sum(-1, keepdim=True) + 1e-8)
    a = x_t.abs().unsqueeze(1)
    wu2 = w * u * u
    xr = x_t.real.unsqueeze(1)
    xi = x_t.imag.unsqueeze(1)
    return {
        "Y": (a * w).sum(-1),
        "Z": torch.complex((xr * w).sum(-1), (xi * w).sum(-1)),
        "m2": wu2.sum(-1),
        "a2": (a * wu2).sum(-1),
        "z2": torch.complex((xr * wu2).sum(-1), (xi * wu2).sum(-1)),
        "bw": bw.squeeze(-1),
    }


def dq_closed_form(m: Dict[str, torch.Tensor], q_t, fc, g_y=None, g_phase=None):
    """dL/dQ_t from upstream gY_t and gphase_t."""
    kappa = -fc / ((q_t + 1e-8) ** 2 * m["bw"])
    out = torch.zeros_like(q_t)
    if g_y is not None:
        out = out + g_y * kappa * (m["a2"] - m["Y"] * m["m2"])
    if g_phase is not None:
        dz = kappa * (m["z2"] - m["Z"] * m["m2"])
        z = m["Z"]
        out = out + g_phase * (z.real * dz.imag - z.imag * dz.real) / (z.real ** 2 + z.imag ** 2)
    return out


# ----------------------------------------------------------------------------
# deterministic synthetic data + weights (shared by golden generation, tests, bench)
# ----------------------------------------------------------------------------
def synth_binaural(batch: int, seed: int = 1234, n: int = 16000) -> Tuple[np.ndarray, np.ndarray]:
    """AR(1)-tilted noise with per-clip integer ITD in [-12,12] and ILD gain in [0.5,1], joint
    peak-normalised to 1.  numpy RandomState => stable across versions."""
    rs = np.random.RandomState(seed)
    src = rs.standard_normal((batch, n + 64))
    s = np.empty_like(src)
    acc = np.zeros(batch)
    for i in range(src.shape[1]):
        acc = 0.9 * acc + src[:, i]
        s[:, i] = acc
    itd = rs.randint(-12, 13, size=batch)
    ild = rs.uniform(0.5, 1.0, size=batch)
    noise = 0.01 * rs.standard_normal((batch, n))
    wl = s[:, 32:32 + n]
    wr = np.stack([ild[b] * s[b, 32 - itd[b]:32 - itd[b] + n] for b in range(batch)]) + noise
    peak = np.maximum(np.abs(wl).max(1), np.abs(wr).max(1))[:, None]
    return (wl / peak).astype(np.float32), (wr / peak).astype(np.float32)


def controller_shapes(n_bands=100, in_mult=2, hid=128):
    return {
        "q_rnn.weight_ih_l0": (3 * hid, in_mult * n_bands),
        "q_rnn.weight_hh_l0": (3 * hid, hid),
        "q_rnn.bias_ih_l0": (3 * hid,),
        "q_rnn.bias_hh_l0": (3 * hid,),
        "q_out.0.weight": (hid, hid), "q_out.0.bias": (hid,),
        "q_out.1.weight": (hid,), "q_out.1.bias": (hid,),
        "q_out.4.weight": (hid, hid), "q_out.4.bias": (hid,),
        "q_out.5.weight": (hid,), "q_out.5.bias": (hid,),
        "q_out.8.weight": (n_bands, hid), "q_out.8.bias": (n_bands,),
    }


def synth_controller(seed: int, n_bands=100, in_mult=2, hid=128, out_std=0.02) -> Dict[str, np.ndarray]:
    """Seeded controller weights: U(-1/sqrt(hid), 1/sqrt(hid)) like torch's default init, LayerNorm
    gains near 1, and a NON-zero last layer (std out_std) so that Q actually moves."""
    rs = np.random.RandomState(seed)
    k = 1.0 / math.sqrt(hid)
    out = {}
    for name, shape in controller_shapes(n_bands, in_mult, hid).items():
        if name in ("q_out.1.weight", "q_out.5.weight"):
            v = 1.0 + 0.1 * rs.standard_normal(shape)
        elif name in ("q_out.1.bias", "q_out.5.bias"):
            v = 0.1 * rs.standard_normal(shape)
        elif name == "q_out.8.weight":
            v = out_std * rs.standard_normal(shape)
        elif name == "q_out.8.bias":
            v = 0.5 * out_std * rs.standard_normal(shape)
        else:
            v = rs.uniform(-k, k, size=shape)
        out[name] = v.astype(np.float32)
    return out


def to_torch(d: Dict[str, np.ndarray], dtype=torch.float32, requires_grad=False):
    return {k: torch.from_numpy(v).to(dtype).requires_grad_(requires_grad) for k, v in d.items()}
