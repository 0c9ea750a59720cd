"""Drop-in namespace for the reference's ``model_torch`` module.

``train_biear.py`` / ``evaluate_biear.py`` do ``from model_torch import build_model, build_model_active,
N_SECTORS, N_DIST_CLASS`` (train_biear.py:12, evaluate_biear.py:11).  Putting this package directory first on
``sys.path`` (route (i)) resolves that import here, and the scripts run unchanged with the
front-end executing on the sm_100a kernels of this package.

What is native and what is plain PyTorch:
  * the binaural front-end (``bifb``) -- STFT, Gaussian band stage, Q controllers, sub-band phase, and their
    backward -- is this package's CUDA path (frontend.py / ops.py / csrc);
  * the back-end (ILD/IPD GRU encoders, body MLP, 8 sub-heads; model_torch.py:828-960, 1088-1110) keeps the
    reference's ``torch.nn`` module tree, so that state-dict keys, default initialisation order (same parameters
    under the same seed) and the optimiser's parameter groups are identical; on CUDA tensors the two GRU layers of
    each encoder run their recurrences through csrc/gru.cu (one launch per layer and direction) and the eight sector
    heads through csrc/heads.cu (one launch forward, one backward); the body MLP, the input projections and the
    weight-gradient products are plain library GEMMs.

The one structural change in ``DeepEarActiveWaveform.forward``: the sub-band phase comes out of the same band
pass that produces Y (``bifb.forward_features``) instead of a second W(Q) rebuild from (X, Q)
(model_torch.py:1039-1063, 1085-1086), and the four ``isfinite(...).all()`` host synchronisations
(:1071-1074) collapse into one device-side reduction that is only read back when it fails.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from .frontend import (AuralNetGammatoneFB, BinauralAdaptiveGammatoneFB,  # noqa: F401  (re-exported)
                       BinauralAdaptiveGammatoneFB_SingleController, FramewiseAdaptiveGammatoneFB,
                       FramewiseFixedGammatoneFB, erb_hz, erb_rate, erb_spaced_fc_and_q, inv_erb_rate,
                       make_deltaQ_profile)

N_SECTORS = 8
N_DIST_CLASS = 5
DATA_DIM = 100
LATENT_DIM = 100


def _zero_nonfinite(x):
    return torch.nan_to_num(x, nan=0.0, posinf=0.0, neginf=0.0)


PARALLEL_BRANCHES = True    # independent sub-networks of the back-end (two encoders, eight sector heads) on forked streams
_branch_streams = {}


def _fork_join(fns, ref: torch.Tensor):
    """Run independent closures on forked CUDA streams and join them on the current one.  Each branch is a few dozen
    tiny kernels (the eight sector heads alone are ~200 launches forward); side by side -- also inside a captured CUDA
    graph, where the forks become parallel branches, and in the backward, which autograd runs on the forward's streams --
    they cost the depth of one branch instead of the sum.  Results (and dropout masks: the generator offset advances in
    host order) are identical to running them one after the other.  CPU tensors: plain sequential calls."""
    if not (PARALLEL_BRANCHES and ref.is_cuda and len(fns) > 1):
        return [fn() for fn in fns]
    dev = ref.device
    cur = torch.cuda.current_stream(dev)
    pool = _branch_streams.setdefault(dev.index if dev.index is not None else torch.cuda.current_device(), [])
    while len(pool) < len(fns):
        pool.append(torch.cuda.Stream(device=dev))
    outs = []
    for fn, st in zip(fns, pool):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            outs.append(fn())
    for o, st in zip(outs, pool):
        cur.wait_stream(st)
        for t in (o if isinstance(o, (tuple, list)) else (o,)):
            t.record_stream(cur)
    return outs


class _PairEncoder(nn.Module):
    """LayerNorm -> GRU(in->hidden) -> GRU(hidden->latent) -> mean over frames (model_torch.py:828-867).
    Sub-classes define the interaural feature built from the left/right inputs."""

    def __init__(self, input_dim=DATA_DIM, hidden_dim=200, latent_dim=LATENT_DIM):
        super().__init__()
        self.in_norm = nn.LayerNorm(input_dim)
        self.gru1 = nn.GRU(input_dim, hidden_dim, batch_first=True)
        self.gru2 = nn.GRU(hidden_dim, latent_dim, batch_first=True)

    native_gru = True         # False: torch.nn.GRU on the GPU too (the cross-check; what CPU tensors always take)

    def interaural(self, xL, xR):
        raise NotImplementedError

    def _layer(self, gru, x):
        # each layer's 19-step recurrence as one launch forward and one backward (csrc/gru.cu) instead of the library's
        # per-step GEMM + cell kernels; same parameters, same results to fp32 rounding
        if x.is_cuda and self.native_gru and x.dtype == torch.float32 and ops.gru_supported(gru.hidden_size):
            return ops.gru_layer(x, gru)
        return gru(x)[0]

    def forward(self, xL, xR):
        seq = self._layer(self.gru1, self.in_norm(self.interaural(xL, xR)))
        seq = self._layer(self.gru2, seq)
        return _zero_nonfinite(seq.mean(dim=1))


class ILDEncoder(_PairEncoder):
    """Level difference of the log band energies, clamped to +-10 (model_torch.py:828-846)."""

    def interaural(self, xL, xR):
        return torch.clamp(_zero_nonfinite(xL - xR), -10.0, 10.0)


class IPDEncoder(_PairEncoder):
    """Wrapped phase difference (model_torch.py:848-867)."""

    def interaural(self, xL, xR):
        d = xL - xR
        return _zero_nonfinite(torch.atan2(torch.sin(d), torch.cos(d)))


def _mlp_head(out_dim):
    return nn.Sequential(nn.Linear(100, 50), nn.ReLU(), nn.Linear(50, 10), nn.ReLU(), nn.Linear(10, out_dim))


class SubHead(nn.Module):
    """Per-sector head: presence logit, normalised angle in (0,1), distance-class logits (model_torch.py:869-906)."""

    def __init__(self, body_dim=200, n_dist_class=N_DIST_CLASS):
        super().__init__()
        self.shared = nn.Sequential(nn.Linear(body_dim, 100), nn.ReLU(), nn.Dropout(0.2))
        self.sound = _mlp_head(1)
        self.aoa = _mlp_head(1)
        self.dist = _mlp_head(n_dist_class)

    def forward(self, body_feat):
        h = self.shared(body_feat)
        return self.sound(h), torch.sigmoid(self.aoa(h)), self.dist(h)


def _body(feat_dim):
    return nn.Sequential(nn.Linear(feat_dim, 512), nn.ReLU(), nn.Dropout(0.2),
                         nn.Linear(512, 400), nn.ReLU(), nn.Dropout(0.2),
                         nn.Linear(400, 200), nn.ReLU(), nn.Dropout(0.2))


class _BackEnd(nn.Module):
    """Encoders + CC projection + body + sector heads shared by the passive and the active model.
    Attribute names and creation order follow model_torch.py:908-960 / 1006-1026."""

    def _init_backend(self, use_cc, data_dim, latent_dim, n_sectors, n_dist_class):
        self.use_cc = use_cc
        self.encoder_ild = ILDEncoder(input_dim=data_dim, hidden_dim=200, latent_dim=latent_dim)
        self.encoder_ipd = IPDEncoder(input_dim=data_dim, hidden_dim=200, latent_dim=latent_dim)
        if use_cc:
            self.cc_proj = nn.Linear(data_dim, latent_dim)
        self.body = _body(2 * latent_dim + (latent_dim if use_cc else 0))
        self.subheads = nn.ModuleList([SubHead(200, n_dist_class=n_dist_class) for _ in range(n_sectors)])
        self.native_heads = True      # False: the torch.nn modules (the cross-check; what CPU tensors always take)

    def _backend(self, x1, x2, x3, ph_l, ph_r):
        branches = [lambda: self.encoder_ild(x1, x2), lambda: self.encoder_ipd(ph_l, ph_r)]
        if self.use_cc:
            branches.append(lambda: self.cc_proj(x3))
        feats = _fork_join(branches, x1)
        return self._heads(self.body(torch.cat(feats, dim=-1)))

    def _heads(self, body):
        if body.is_cuda and self.native_heads and body.dtype == torch.float32 and body.shape[1] % 4 == 0 \
                and body.shape[1] <= 200:
            # all sector heads in one launch forward, one backward (csrc/heads.cu)
            return ops.sector_heads(body, list(self.subheads), self.training)
        outs = _fork_join([lambda h=head: h(body) for head in self.subheads], body)     # host tensors: torch.nn
        sound = torch.cat([o[0] for o in outs], dim=1)
        aoa = torch.cat([o[1] for o in outs], dim=1)
        dist = torch.stack([o[2] for o in outs], dim=1)
        return sound, aoa, dist


class DeepEarTorchILD(_BackEnd):
    """Passive model on precomputed features: forward(x1, x2, x3, x4, x5) (model_torch.py:908-960)."""

    def __init__(self, use_cc: bool = True, data_dim: int = DATA_DIM, latent_dim: int = LATENT_DIM,
                 n_sectors: int = N_SECTORS, n_dist_class: int = N_DIST_CLASS):
        super().__init__()
        self._init_backend(use_cc, data_dim, latent_dim, n_sectors, n_dist_class)

    def forward(self, x1, x2, x3, x4, x5):
        return self._backend(x1, x2, x3, x4, x5)


class DeepEarActiveWaveform(_BackEnd):
    """Active model: raw binaural waveforms -> front-end -> back-end (model_torch.py:965-1112).

    forward(wavL_1s, wavR_1s, x3=None) -> (sound_logits (B,S), aoa_pred (B,S), dist_logits (B,S,C)).
    """

    def __init__(self, fs=16000, timesteps=19, n_fft=1024, n_bands=DATA_DIM, use_cc=True, latent_dim=LATENT_DIM,
                 n_sectors=N_SECTORS, n_dist_class=N_DIST_CLASS, fb_alpha: float = 0.2,
                 fixed_frontend_q: bool = False, deltaQ_base: float = 2.0, deltaQ_low_factor: float = 0.5,
                 deltaQ_high_factor: float = 1.0, deltaQ_mode: str = "absolute", bifb_class=None):
        super().__init__()
        self.fb_alpha = fb_alpha
        cls = bifb_class if bifb_class is not None else BinauralAdaptiveGammatoneFB
        self.bifb = cls(fs=fs, timesteps=timesteps, n_fft=n_fft, Nbands=n_bands, alpha=fb_alpha,
                        fixed_frontend_q=bool(fixed_frontend_q), deltaQ_base=deltaQ_base,
                        deltaQ_low_factor=deltaQ_low_factor, deltaQ_high_factor=deltaQ_high_factor,
                        deltaQ_mode=deltaQ_mode)
        self._init_backend(use_cc, n_bands, latent_dim, n_sectors, n_dist_class)
        self.last_QL = self.last_QR = self.last_Q = None

    # model_torch.py:1032-1037, 1071-1074 raise RuntimeError when YL / YR / QL / QR hold NaN / Inf -- four host
    # synchronisations per forward there.  Here the four counts are ONE device-side reduction whose result travels to
    # pinned host memory behind the step; the host looks at it when the NEXT forward starts (or on check_finite()), so
    # the forward path has no host synchronisation and can be captured in a CUDA graph.
    #   finite_check = "deferred" (default) | "sync" (the reference's immediate raise, one host sync) | "off"
    finite_check = "deferred"
    _FINITE_NAMES = ("YL", "YR", "QL", "QR")

    def _raise_if_bad(self, counts):
        for name, n in zip(self._FINITE_NAMES, counts):
            if n:
                raise RuntimeError(f"[NaN/Inf] {name} has {int(n)} non-finite values.")

    def check_finite(self):
        """Raise now if an earlier forward produced non-finite YL / YR / QL / QR (waits for that forward)."""
        pend = getattr(self, "_finite_pending", None)
        if pend is not None:
            host, ev = pend
            self._finite_pending = None
            ev.synchronize()
            self._raise_if_bad(host.tolist())

    def _assert_finite(self, tensors):
        if self.finite_check == "off":
            return
        bad = torch.stack([(~torch.isfinite(t)).sum() for _, t in tensors])
        if torch.cuda.is_current_stream_capturing():
            self.finite_flags = bad                       # static tensor of the graph: the owner of the graph may read it
            return
        if self.finite_check == "sync":
            self._raise_if_bad(bad.tolist())
            return
        self.check_finite()                               # the previous forward's verdict (long finished: no stall)
        host = getattr(self, "_finite_host", None)
        if host is None:
            host = self._finite_host = torch.zeros(len(tensors), dtype=torch.int64).pin_memory()
        host.copy_(bad, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._finite_pending = (host, ev)

    def forward(self, wavL_1s, wavR_1s, x3=None):
        wavL_1s = wavL_1s.float()
        wavR_1s = wavR_1s.float()
        if hasattr(self.bifb, "forward_features"):
            o = self.bifb.forward_features(wavL_1s, wavR_1s, want_phase=True, want_logenergy=True)
            YL, YR, QL, QR, ph_l, ph_r = o["YL"], o["YR"], o["QL"], o["QR"], o["phaseL"], o["phaseR"]
        else:   # a foreign bifb_class that only implements the reference's 6-tuple protocol (model_torch.py:1069):
            # log energies in PyTorch, sub-band phase from (X, Q) through the band kernel (one pass, all frames)
            from . import ops
            from .frontend import _log_energy
            YL, YR, QL, QR, XL, XR = self.bifb(wavL_1s, wavR_1s)
            fc = self.bifb.fc.to(YL.device)
            df = float(self.bifb.f_fft[1] - self.bifb.f_fft[0])
            phs = []
            for X, Q in ((XL, QL), (XR, QR)):
                xr = torch.view_as_real(X.contiguous())
                T_ = xr.shape[1]
                ph = torch.stack([ops.BandFrame.apply(Q[:, t], xr, t, fc, df, ops.DEFAULT_CUTOFF, True, "jacobian")[1]
                                  for t in range(T_)], dim=1)
                phs.append(ph)
            ph_l, ph_r = phs
            o = {"logYL": _log_energy(YL), "logYR": _log_energy(YR)}
        self._assert_finite((("YL", YL), ("YR", YR), ("QL", QL), ("QR", QR)))
        self.last_QL, self.last_QR, self.last_Q = QL, QR, 0.5 * (QL + QR)
        x1, x2 = o["logYL"], o["logYR"]          # clamp(log(Y + 1e-8), +-12), fused into the band stage (:1080-1083)
        if self.use_cc:
            if x3 is None:
                x3 = torch.zeros(wavL_1s.size(0), DATA_DIM, device=wavL_1s.device)
            x3 = x3.float()
        return self._backend(x1, x2, x3, ph_l, ph_r)


def _sinusoidal_positions(T: int, d_model: int, device) -> torch.Tensor:
    """(T, d_model) transformer position table, sin on even / cos on odd features (model_torch.py:56-67)."""
    pos = torch.arange(T, dtype=torch.float32, device=device).unsqueeze(1)
    freq = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32, device=device) * (-math.log(10000.0) / max(d_model, 1)))
    table = torch.zeros(T, d_model, dtype=torch.float32, device=device)
    table[:, 0::2] = torch.sin(pos * freq)
    table[:, 1::2] = torch.cos(pos * freq)
    return table


class AuralNetAttentionBlock(nn.Module):
    """Linear(d_in -> d_model) + sinusoidal positions + 2 pre-norm transformer encoder layers (4 heads, GELU, FFN 4 x
    d_model, dropout 0.1): (B,T,d_in) -> (B,T,d_model) (model_torch.py:779-823).  Library modules: the comparison model is
    not on the hot path."""

    def __init__(self, d_in: int = 64, d_model: int = 128, n_heads: int = 4, n_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        self.d_model = d_model
        self.proj = nn.Linear(d_in, d_model)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=n_heads, dim_feedforward=4 * d_model, dropout=dropout,
                                           activation="gelu", batch_first=True, norm_first=True)
        self.encoder = nn.TransformerEncoder(layer, num_layers=n_layers)

    def forward(self, x):
        h = self.proj(x)
        return self.encoder(h + _sinusoidal_positions(x.shape[1], self.d_model, x.device).unsqueeze(0))


class AuralNetActiveWaveform(_BackEnd):
    """The AuralNet-style comparison model (model_torch.py:1113-1247): same inputs / outputs / heads as the active model,
    fixed filterbank per ear + attention aggregation instead of the adaptive front-end and the GRU encoders.  The two
    filterbanks run on this package's kernels (STFT + the tcgen05 fixed-Q band GEMM), the sector heads on csrc/heads.cu;
    the attention blocks and the body are torch.nn (library) modules with the reference's module tree."""

    def __init__(self, fs: int = 16000, use_cc: bool = True, n_bands: int = DATA_DIM, timesteps: int = 19,
                 hop_ratio: float = 1.0, n_fft: int = 1024, d_model: int = 128, n_sectors: int = N_SECTORS,
                 n_dist_class: int = N_DIST_CLASS):
        super().__init__()
        self.use_cc = use_cc
        self.fb_L = AuralNetGammatoneFB(fs=fs, n_bands=n_bands, timesteps=timesteps, hop_ratio=hop_ratio, n_fft=n_fft)
        self.fb_R = AuralNetGammatoneFB(fs=fs, n_bands=n_bands, timesteps=timesteps, hop_ratio=hop_ratio, n_fft=n_fft)
        self.bifb = None                       # (the scripts' front / back parameter split finds nothing: :1147-1149)
        self.attn_L = AuralNetAttentionBlock(d_in=n_bands, d_model=d_model)
        self.attn_R = AuralNetAttentionBlock(d_in=n_bands, d_model=d_model)
        self.attn_diff = AuralNetAttentionBlock(d_in=n_bands, d_model=d_model)
        self.cc_proj = nn.Linear(DATA_DIM, d_model) if use_cc else None
        self.body = _body(3 * d_model + (d_model if use_cc else 0))
        self.subheads = nn.ModuleList([SubHead(200, n_dist_class=n_dist_class) for _ in range(n_sectors)])
        self.native_heads = True
        self.last_Q = None                     # no adaptive Q here: the training loop's Q regularisers see None

    def forward(self, wavL_1s, wavR_1s, x3=None):
        wl = torch.clamp(wavL_1s.float(), -1.0, 1.0)
        wr = torch.clamp(wavR_1s.float(), -1.0, 1.0)
        xl = torch.clamp(torch.log(self.fb_L(wl) + 1e-8), -12.0, 12.0)
        xr = torch.clamp(torch.log(self.fb_R(wr) + 1e-8), -12.0, 12.0)
        feats = [self.attn_L(xl).mean(dim=1), self.attn_R(xr).mean(dim=1), self.attn_diff(xl - xr).mean(dim=1)]
        if self.use_cc:
            if x3 is None:
                x3 = torch.zeros(wl.size(0), DATA_DIM, device=wl.device)
            feats.append(self.cc_proj(x3.float()))
        return self._heads(self.body(torch.cat(feats, dim=-1)))


def build_model(use_cc: bool = True, data_dim: int = DATA_DIM, latent_dim: int = LATENT_DIM,
                n_sectors: int = N_SECTORS, n_dist_class: int = N_DIST_CLASS) -> nn.Module:
    """model_torch.py:1252-1265."""
    return DeepEarTorchILD(use_cc=use_cc, data_dim=data_dim, latent_dim=latent_dim, n_sectors=n_sectors,
                           n_dist_class=n_dist_class)


def _build_active(bifb_class, use_cc, fs, timesteps, n_fft, data_dim, latent_dim, n_sectors, n_dist_class, fb_alpha,
                  fixed_frontend_q, deltaQ_base, deltaQ_low_factor, deltaQ_high_factor, deltaQ_mode):
    return DeepEarActiveWaveform(fs=fs, timesteps=timesteps, n_fft=n_fft, n_bands=data_dim, use_cc=use_cc,
                                 latent_dim=latent_dim, n_sectors=n_sectors, n_dist_class=n_dist_class,
                                 fb_alpha=fb_alpha, fixed_frontend_q=bool(fixed_frontend_q), deltaQ_base=deltaQ_base,
                                 deltaQ_low_factor=deltaQ_low_factor, deltaQ_high_factor=deltaQ_high_factor,
                                 deltaQ_mode=deltaQ_mode, bifb_class=bifb_class)


def build_model_active(use_cc=True, fs=16000, timesteps=19, n_fft=1024, data_dim=DATA_DIM, latent_dim=LATENT_DIM,
                       n_sectors=N_SECTORS, n_dist_class=N_DIST_CLASS, fb_alpha: float = 0.2,
                       fixed_frontend_q: bool = False, deltaQ_base: float = 2.0, deltaQ_low_factor: float = 0.5,
                       deltaQ_high_factor: float = 1.0, deltaQ_mode: str = "absolute") -> nn.Module:
    """model_torch.py:1303-1334 (dual controllers)."""
    return _build_active(None, use_cc, fs, timesteps, n_fft, data_dim, latent_dim, n_sectors, n_dist_class, fb_alpha,
                         fixed_frontend_q, deltaQ_base, deltaQ_low_factor, deltaQ_high_factor, deltaQ_mode)


def build_model_active_single_controller(use_cc=True, fs=16000, timesteps=19, n_fft=1024, data_dim=DATA_DIM,
                                         latent_dim=LATENT_DIM, n_sectors=N_SECTORS, n_dist_class=N_DIST_CLASS,
                                         fb_alpha: float = 0.2, fixed_frontend_q: bool = False,
                                         deltaQ_base: float = 2.0, deltaQ_low_factor: float = 0.5,
                                         deltaQ_high_factor: float = 1.0, deltaQ_mode: str = "absolute") -> nn.Module:
    """model_torch.py:1267-1300 (one shared controller for both ears)."""
    return _build_active(BinauralAdaptiveGammatoneFB_SingleController, use_cc, fs, timesteps, n_fft, data_dim,
                         latent_dim, n_sectors, n_dist_class, fb_alpha, fixed_frontend_q, deltaQ_base,
                         deltaQ_low_factor, deltaQ_high_factor, deltaQ_mode)


def build_model_auralnet_active(use_cc: bool = True, fs: int = 16000, n_bands: int = DATA_DIM, timesteps: int = 19,
                                hop_ratio: float = 1.0, n_fft: int = 1024, d_model: int = 128,
                                n_sectors: int = N_SECTORS, n_dist_class: int = N_DIST_CLASS) -> nn.Module:
    """model_torch.py:1337-1367 (the AuralNet-style comparison model)."""
    return AuralNetActiveWaveform(fs=fs, use_cc=use_cc, n_bands=n_bands, timesteps=timesteps, hop_ratio=hop_ratio,
                                  n_fft=n_fft, d_model=d_model, n_sectors=n_sectors, n_dist_class=n_dist_class)
