"""Host-side mirror of the reference's binaural front-end modules (model_torch.py:70-776), running on the
sm_100a kernels of this package.

Same class names, constructor keywords, forward signatures, buffers and state-dict keys as the reference,
so `DeepEarActiveWaveform(bifb_class=...)`, `train_biear.py`'s parameter groups / freeze helpers and
checkpoints written by either implementation work unchanged.  What differs is the execution plan:

  * the spectra X do not depend on Q, so framing + Hann + rFFT for every (ear, clip, frame) is ONE
    launch up front instead of 19 x 2 cuFFT calls inside the frame loop;
  * one band kernel per frame serves both ears, never materialises the (B,100,513) weight tensor and
    emits band energy, sub-band phase and the exact dY/dQ, dphase/dQ Jacobians in the same pass, so the
    reference's second W rebuild (_subband_phase_from_X) and autograd's W-sized backward disappear;
  * the two per-ear Q controllers run as one batched (ear-stacked) chain;
  * the batch-global `isfinite(Q).all()` guard (model_torch.py:378-380) is evaluated on the device, so
    the frame loop has no host synchronisation.

All compute requires CUDA tensors; there is no CPU fallback.
"""
from __future__ import annotations

import os

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops, timing

Q_MIN, Q_MAX = 0.05, 30.0


# ------------------------------------------------------------------------------------------------
# ERB constants (model_torch.py:19-51), float64 numpy on the host exactly as the reference computes them
# ------------------------------------------------------------------------------------------------
def erb_hz(f_hz):
    return 24.7 * (4.37 * f_hz / 1000.0 + 1.0)


def erb_rate(f_hz):
    return 21.4 * np.log10(4.37 * f_hz / 1000.0 + 1.0)


def inv_erb_rate(e):
    return (10 ** (e / 21.4) - 1.0) * 1000.0 / 4.37


def erb_spaced_fc_and_q(N=100, fmin=50.0, fmax=7200.0, erb_factor=1.019):
    e = np.linspace(erb_rate(fmin), erb_rate(fmax), N)
    fc = inv_erb_rate(e)
    return fc, fc / (erb_factor * erb_hz(fc))


def make_deltaQ_profile(fc_hz: torch.Tensor, deltaQ_base: float = 2.0, low_factor: float = 0.5,
                        high_factor: float = 1.0) -> torch.Tensor:
    e = erb_rate(fc_hz.detach().cpu().numpy())
    e = (e - e.min()) / (e.max() - e.min() + 1e-12)
    mult = torch.tensor(low_factor + (high_factor - low_factor) * e, dtype=torch.float32, device=fc_hz.device)
    return torch.clamp(deltaQ_base * mult, min=1e-3)


# ------------------------------------------------------------------------------------------------
# shared geometry / buffers
# ------------------------------------------------------------------------------------------------
class _FilterbankBase(nn.Module):
    """Frame geometry and the fc / Q0 / f_fft buffers every filterbank variant registers."""

    def _init_geometry(self, fs, timesteps, n_fft, Nbands, fmin, fmax, hop_ratio, window_buffer=True):
        self.fs = fs
        self.timesteps = timesteps
        self.n_fft = n_fft
        self.Nbands = Nbands
        self.win = int(round(fs / timesteps))
        self.hop = max(1, int(round(self.win * hop_ratio)))
        if window_buffer:
            self.register_buffer("win_fn", torch.hann_window(self.win), persistent=False)
        self.register_buffer("f_fft", torch.linspace(0, fs / 2, n_fft // 2 + 1))
        if fmax is None:
            fmax = fs / 2 * 0.9
        fc_np, q0_np = erb_spaced_fc_and_q(Nbands, fmin, fmax, erb_factor=1.019)
        self.register_buffer("fc", torch.tensor(fc_np, dtype=torch.float32))
        self.register_buffer("Q0", torch.tensor(q0_np, dtype=torch.float32))

    @property
    def df(self) -> float:
        return (self.fs / 2) / (self.n_fft // 2)

    def _spectra(self, wavs: List[torch.Tensor]) -> torch.Tensor:
        """[(B,Nsamp)] * E -> X (E*B, T, F) complex64: both ears in one launch into one output tensor (no waveform concat)."""
        for w in wavs:
            if w.dim() != 2:
                raise ValueError(f"Expected wav_1s (B,N), got {w.shape}")
            if w.requires_grad:
                raise RuntimeError("biear_b200: gradients with respect to the waveform are not implemented")
        ws = [w.float().contiguous() for w in wavs]
        if len(ws) == 2 and ws[0].shape == ws[1].shape:      # both ears in one launch (18 % faster than two)
            return ops.stft_pair(ws[0], ws[1], self.win_fn, self.fs, self.timesteps, self.win, self.hop, self.n_fft)
        return ops.stft(ws, self.win_fn, self.fs, self.timesteps, self.win, self.hop, self.n_fft)


_side_streams = {}


def _side_stream(device) -> torch.cuda.Stream:
    """One auxiliary stream per device for work that is independent of the recurrence (the CC feature)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _side_streams:
        _side_streams[idx] = torch.cuda.Stream(device=device)
    return _side_streams[idx]


def _make_controller(in_features: int, Nbands: int):
    """Same modules, same construction order (=> same default init under a given seed) as
    model_torch.py:256-267, 283-287."""
    q_rnn = nn.GRU(input_size=in_features, hidden_size=128, batch_first=True)
    q_out = nn.Sequential(
        nn.Linear(128, 128), nn.LayerNorm(128), nn.SiLU(), nn.Dropout(p=0.1),
        nn.Linear(128, 128), nn.LayerNorm(128), nn.SiLU(), nn.Dropout(p=0.1),
        nn.Linear(128, Nbands),
    )
    nn.init.zeros_(q_out[-1].weight)
    nn.init.zeros_(q_out[-1].bias)
    return q_rnn, q_out


def _next_q(delta, q0, dq, mode):
    if mode == "relative":
        q = q0 * (1.0 + dq * delta)
    else:
        q = q0 + dq * delta
    return torch.clamp(q, Q_MIN, Q_MAX)


def _controller_weights(ctrl_mods):
    """name -> list of the controllers' parameters themselves (the C ABI takes one pointer per controller)."""
    st = lambda f: [f(m) for m in ctrl_mods]
    return {"w_ih": st(lambda m: m.q_rnn.weight_ih_l0), "w_hh": st(lambda m: m.q_rnn.weight_hh_l0),
            "b_ih": st(lambda m: m.q_rnn.bias_ih_l0), "b_hh": st(lambda m: m.q_rnn.bias_hh_l0),
            "w1": st(lambda m: m.q_out[0].weight), "b1": st(lambda m: m.q_out[0].bias),
            "ln1_g": st(lambda m: m.q_out[1].weight), "ln1_b": st(lambda m: m.q_out[1].bias),
            "w2": st(lambda m: m.q_out[4].weight), "b2": st(lambda m: m.q_out[4].bias),
            "ln2_g": st(lambda m: m.q_out[5].weight), "ln2_b": st(lambda m: m.q_out[5].bias),
            "w3": st(lambda m: m.q_out[8].weight), "b3": st(lambda m: m.q_out[8].bias)}


CHAIN_ENGINE = None   # test hook: tests/chain_engine.py installs the per-frame band kernel + PyTorch controller cross-check here


def _adaptive_chain(x: torch.Tensor, ears: int, ctrl_mods, fc, q0, dq_vec, dq_mode, df, training,
                    shared: bool, want_phase: bool, band_mode: str, cutoff: float, engine: str = "fused",
                    want_logy: bool = False, prep=None):
    """The 19-step Q recurrence for `ears` ears of B clips.

    x: (ears*B, T, F) complex64, ear-major.  ctrl_mods: G controller-owning modules.
      dual  : ears == G (each ear has its own controller and its own Q)       model_torch.py:314-386
      single: ears == 2, G == 1, shared=True (one Q for both ears, carried Y memory)  model_torch.py:695-776
    Returns per-ear LISTS of (B,T,N) tensors: Y [ears], Q [G], phase [ears] or None, logY = clamp(log(Y + 1e-8), +-12)
    [ears] or None.
    """
    rows, T, Fbins = x.shape
    B = rows // ears
    G = len(ctrl_mods)
    N = fc.numel()
    xr = torch.view_as_real(x)
    if engine in ("fused", "fused-strict"):
        single = shared and G == 1 and ears == 2
        if not (single or (not shared and G == ears)) or not \
                (ops.single_supported(N, Fbins) if single else ops.fused_supported(N, Fbins)):
            raise NotImplementedError(f"biear_b200: the fused recurrence kernels cover the dual and the single-controller "
                                      f"front-end with at most 128 bands and spectra that fit the shared memory of an SM; "
                                      f"got Nbands={N}, {Fbins} bins (there is no PyTorch fallback)")
        w = _controller_weights(ctrl_mods)
        # CPU generator: no device sync.  (Under CUDA-graph capture the kernels read a device-side seed instead.)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if training else 0
        res = ops.adaptive_sequence(xr, fc, q0, dq_vec, w, dq_mode == "relative", training, want_phase, cutoff,
                                    df, seed, strict=(engine == "fused-strict"), want_logy=want_logy, prep=prep,
                                    ears=ears)
        return res if want_logy else res + (None,)
    if engine == "chain" and CHAIN_ENGINE is not None:
        return CHAIN_ENGINE(x, ears, ctrl_mods, fc, q0, dq_vec, dq_mode, df, training, shared, want_phase, band_mode, cutoff,
                            want_logy)
    raise NotImplementedError(f"biear_b200: engine {engine!r} is not available (the per-frame cross-check engine lives in "
                              "tests/chain_engine.py); the product path is the fused recurrence")


def _log_energy(y: torch.Tensor) -> torch.Tensor:
    """model_torch.py:1080-1083."""
    return torch.clamp(torch.log(y + 1e-8), -12.0, 12.0)


STREAM_SPECTRA = True     # STFT next to the recurrence (flag hand-over) whenever the co-residency conditions below hold
STFT_PRODUCER_SMS = 16    # SMs that must remain free of recurrence CTAs for the streamed STFT producer
FIXED_ENGINE = "gemm"   # "gemm": shared-weight dense contraction (csrc/band_fixed.cu); "item": per-item band kernel,
                        # bit-identical to the adaptive path at Q == Q0 (kept for that identity and as a cross-check)


def _fixed_bands(x: torch.Tensor, fc, q_fixed, df, want_phase, cutoff):
    """Fixed-Q path: every (row, frame) item in one launch.  All items share one Q vector, so the band weights are
    one (N x F) matrix and the stage is a GEMM (model_torch.py:451-487 rebuilds that matrix 19 times per call)."""
    xr = torch.view_as_real(x)
    if FIXED_ENGINE == "gemm" and fc.numel() <= 128:
        return ops.band_fixed_forward(xr, q_fixed.contiguous(), fc, df, cutoff, want_phase)
    y, ph, _, _ = ops.band_forward(xr, None, q_fixed.contiguous(), fc, df, cutoff, want_phase, False)
    return y, ph


# ------------------------------------------------------------------------------------------------
# monaural filterbanks
# ------------------------------------------------------------------------------------------------
class FramewiseAdaptiveGammatoneFB(_FilterbankBase):
    """model_torch.py:200-386.  forward(wav_1s (B,Nsamp)) -> Y (B,T,N), Q (B,T,N), X (B,T,F) complex."""

    def __init__(self, fs, timesteps=19, n_fft=1024, Nbands=100, Q0=8.0, fmin=50.0, fmax=None, hop_ratio=1.0,
                 alpha=0.2, deltaQ_base: float = 2.0, deltaQ_low_factor: float = 0.5,
                 deltaQ_high_factor: float = 1.0, deltaQ_mode: str = "absolute"):
        super().__init__()
        self.alpha = alpha
        self._init_geometry(fs, timesteps, n_fft, Nbands, fmin, fmax, hop_ratio)
        self.register_buffer("deltaQ_vec", make_deltaQ_profile(self.fc, deltaQ_base, deltaQ_low_factor,
                                                               deltaQ_high_factor))
        self.deltaQ_mode = deltaQ_mode.lower()
        self.q_rnn, self.q_out = _make_controller(2 * Nbands, Nbands)
        self.Q_min, self.Q_max = Q_MIN, Q_MAX
        self.freeze_Q = False
        self.band_mode = "jacobian"
        self.engine = "fused"
        self.cutoff = ops.DEFAULT_CUTOFF

    def forward(self, wav_1s: torch.Tensor):
        x = self._spectra([wav_1s])
        if self.freeze_Q:
            y, _ = _fixed_bands(x, self.fc, self.Q0, self.df, False, self.cutoff)
            return y, self.Q0.view(1, 1, -1).expand(x.shape[0], self.timesteps, -1), x
        y, q, _, _ = _adaptive_chain(x, 1, [self], self.fc, self.Q0, self.deltaQ_vec, self.deltaQ_mode, self.df,
                                     self.training, False, False, self.band_mode, self.cutoff,
                                     self.engine)
        return y[0], q[0], x


class FramewiseFixedGammatoneFB(_FilterbankBase):
    """model_torch.py:391-487: Q == clamp(Q0) for every frame, no controller, no parameters."""

    def __init__(self, fs, timesteps=19, n_fft=1024, Nbands=100, fmin=50.0, fmax=None, hop_ratio=1.0):
        super().__init__()
        self._init_geometry(fs, timesteps, n_fft, Nbands, fmin, fmax, hop_ratio)
        self.Q_min, self.Q_max = Q_MIN, Q_MAX
        self.cutoff = ops.DEFAULT_CUTOFF

    def forward(self, wav_1s: torch.Tensor):
        x = self._spectra([wav_1s])
        qf = torch.clamp(self.Q0, self.Q_min, self.Q_max)
        y, _ = _fixed_bands(x, self.fc, qf, self.df, False, self.cutoff)
        return y, qf.view(1, 1, -1).expand(x.shape[0], self.timesteps, -1), x


class AuralNetGammatoneFB(_FilterbankBase):
    """model_torch.py:70-195: fixed-Q filterbank in batched form; forward(wav) -> Y (B,T,N) only.  Constructor arguments
    in the reference's order and names (``n_bands``, not ``Nbands``: model_torch.py:89-98)."""

    def __init__(self, fs: int = 16000, n_bands: int = 100, fmin: float = 50.0, fmax: float = None, timesteps: int = 19,
                 hop_ratio: float = 1.0, n_fft: int = 1024):
        super().__init__()
        if int(timesteps) <= 0:
            raise ValueError(f"[AuralNetGammatoneFB] timesteps must be > 0, got {timesteps}")
        self._init_geometry(fs, int(timesteps), n_fft, n_bands, fmin, fmax, float(hop_ratio))
        self.n_bands = n_bands
        self.hop_ratio = float(hop_ratio)
        self.Q_min, self.Q_max = Q_MIN, Q_MAX
        self.cutoff = ops.DEFAULT_CUTOFF

    def forward(self, wav_1s: torch.Tensor):
        if wav_1s.dim() != 2:
            raise ValueError(f"[AuralNetGammatoneFB] Expected wav_1s (B,N), got {wav_1s.shape}")
        x = self._spectra([wav_1s])
        y, _ = _fixed_bands(x, self.fc, torch.clamp(self.Q0, self.Q_min, self.Q_max), self.df, False, self.cutoff)
        return y


# ------------------------------------------------------------------------------------------------
# binaural filterbanks
# ------------------------------------------------------------------------------------------------
GRAPH_REPLAY_DEFAULT = True   # value of BinauralAdaptiveGammatoneFB.graph_replay at construction (tests that pin RNG
                              # streams or count launches construct their modules with False)


class _FeatureTuple(nn.Module):
    """forward_features as a tensors-in / tuple-of-tensors-out module (what torch.cuda.make_graphed_callables captures)."""

    def __init__(self, bifb, flags):
        super().__init__()
        self.bifb = bifb
        self.flags = flags
        self.keys = None

    def forward(self, wl, wr):
        o = self.bifb._forward_features(wl, wr, *self.flags)
        if self.keys is None:
            self.keys = tuple(o.keys())
        return tuple(o[k] for k in self.keys)


class _GraphCache:
    """Per-module cache of captured front-end calls, keyed by (device, batch, samples, train / eval, grad mode, requested
    outputs, parameter storage).  The reference's scripts call the model eagerly, ~35 launches and ~60 allocations per
    front-end forward + backward, which is host-bound (measured 2.2 ms per step at batch 256 against 0.9 ms of kernels);
    the third call with an unchanged key captures the forward -- and, in grad mode, its backward -- as CUDA graphs
    (torch.cuda.make_graphed_callables: static input / output / saved-state buffers, autograd node that replays the backward
    graph) and every later call copies the two waveforms in and replays.  Dropout stays random (the kernels read their
    Philox seed from a device counter that the graph advances).  Contract of the graphed path: the returned tensors are
    views of static buffers, valid until the next forward with the same key; run backward before that forward."""
    CAPTURE_AFTER = 2        # eager calls with a key before it is captured (allocator / lazy-init warm-up)

    def __init__(self):
        self.seen = {}
        self.graphed = {}

    def __deepcopy__(self, memo):          # copies of a module start with an empty cache (graphs are not copyable)
        return _GraphCache()

    def __getstate__(self):                # ... and so do pickled ones
        return {}

    def __setstate__(self, state):
        self.seen, self.graphed = {}, {}

    def lookup(self, bifb, wl, wr, flags):
        params = tuple(bifb.parameters())
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        key = (wl.device.index, tuple(wl.shape), wl.dtype, wr.dtype, bifb.training, need_grad, flags,
               getattr(bifb, "engine", None), bifb.fixed_frontend_q, getattr(bifb, "freeze_Q", None),
               getattr(getattr(bifb, "fb_L", None), "freeze_Q", None), getattr(getattr(bifb, "fb_R", None), "freeze_Q", None),
               tuple((p.data_ptr(), p.requires_grad) for p in params))
        if wl.dtype != torch.float32 or wr.dtype != torch.float32 or wl.requires_grad or wr.requires_grad:
            return None
        entry = self.graphed.get(key)
        if entry is None:
            n = self.seen.get(key, 0)
            self.seen[key] = n + 1
            if n < self.CAPTURE_AFTER:
                return None
            entry = self._capture(bifb, wl, wr, flags, need_grad)
            self.graphed[key] = entry
        fn, mod = entry
        outs = fn(wl.contiguous(), wr.contiguous())
        if need_grad and timing.enabled:
            node = next((o.grad_fn for o in outs if o.grad_fn is not None), None)
            timing.time_backward_node("frontend.backward", wl.shape[0], node)
        return dict(zip(mod.keys, outs))

    @staticmethod
    def _capture(bifb, wl, wr, flags, need_grad):
        mod = _FeatureTuple(bifb, flags)
        return _GraphedCall(mod, wl, wr, need_grad), mod


class _GraphedCall:
    """forward (and, in grad mode, backward) of a _FeatureTuple as CUDA graphs over static buffers, exposed as ONE autograd
    node.  Same scheme as torch.cuda.make_graphed_callables, but captured with capture_error_mode="thread_local": the
    reference's training loop runs a DataLoader pin-memory thread whose CUDA calls would invalidate a "global" capture."""

    def __init__(self, mod, wl, wr, need_grad):
        dev = wl.device
        self.s_in = (wl.detach().clone(), wr.detach().clone())
        named = [(n, p) for n, p in mod.named_parameters() if p.requires_grad] if need_grad else []
        self.params = [p for _, p in named]
        self.need_grad = bool(self.params)
        # The graphs are recorded against fresh leaf ALIASES of the parameters (same storage, so every replay reads the
        # current weights): the parameters' own AccumulateGrad nodes may still be alive from an eager backward on another
        # stream (e.g. the previous step's loss tensor keeps them), and routing a captured gradient into such a node would
        # make that stream depend on the capturing one, which CUDA refuses.
        alias = {n: p.detach().requires_grad_(True) for n, p in named}
        aliases = list(alias.values())
        run = (lambda: torch.func.functional_call(mod, alias, self.s_in)) if self.need_grad else (lambda: mod(*self.s_in))
        grad_ctx = torch.enable_grad if self.need_grad else torch.no_grad
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), grad_ctx():          # one more eager pass on the capture side (allocator warm-up)
            outs = run()
            if self.need_grad:
                req = [o for o in outs if o.requires_grad]
                torch.autograd.grad(req, aliases, [torch.zeros_like(o) for o in req], allow_unused=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.fwd = torch.cuda.CUDAGraph()
        with grad_ctx(), torch.cuda.graph(self.fwd, capture_error_mode="thread_local"):
            outs = run()
        self.s_out = tuple(outs)
        self.req_idx = [i for i, o in enumerate(outs) if o.requires_grad] if self.need_grad else []
        if self.need_grad:
            req = [outs[i] for i in self.req_idx]
            self.s_gout = tuple(torch.zeros_like(o) for o in req)
            self.bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.bwd, pool=self.fwd.pool(), capture_error_mode="thread_local"):
                gin = torch.autograd.grad(req, aliases, self.s_gout, allow_unused=True)
            self.s_gin = tuple(gin)
        call = self

        class _Node(torch.autograd.Function):
            @staticmethod
            def forward(ctx, a, b, *params):
                call.s_in[0].copy_(a, non_blocking=True)
                call.s_in[1].copy_(b, non_blocking=True)
                call.fwd.replay()
                outs_ = tuple(o.detach() for o in call.s_out)
                ctx.mark_non_differentiable(*[o for i, o in enumerate(outs_) if i not in call.req_idx])
                return outs_

            @staticmethod
            def backward(ctx, *grads):
                for dst, i in zip(call.s_gout, call.req_idx):
                    g = grads[i]
                    if g is None:
                        dst.zero_()
                    elif g.data_ptr() != dst.data_ptr():
                        dst.copy_(g, non_blocking=True)
                call.bwd.replay()
                # Fresh copies (autograd may adopt a returned tensor as .grad, and the static buffers are rewritten by the
                # next step): ONE concatenating launch, handed out as views of that new buffer.
                live = [g for g in call.s_gin if g is not None]
                flat = torch.cat([g.reshape(-1) for g in live])
                out, off = [], 0
                for g in call.s_gin:
                    if g is None:
                        out.append(None)
                    else:
                        out.append(flat[off:off + g.numel()].view_as(g))
                        off += g.numel()
                return (None, None) + tuple(out)

        self.node = _Node

    def __call__(self, wl, wr):
        if self.need_grad:
            return self.node.apply(wl, wr, *self.params)
        self.s_in[0].copy_(wl, non_blocking=True)
        self.s_in[1].copy_(wr, non_blocking=True)
        self.fwd.replay()
        return self.s_out


class BinauralAdaptiveGammatoneFB(nn.Module):
    """model_torch.py:492-573 (dual: two independent monaural filterbanks).

    forward(wavL_1s, wavR_1s) -> (YL, YR, QL, QR, XL, XR), as the reference.
    forward_features(...) additionally returns the sub-band phases computed in the same band pass
    (what DeepEarActiveWaveform._subband_phase_from_X, model_torch.py:1039-1063, recomputes from X and Q).
    """

    def __init__(self, fs=16000, timesteps=19, n_fft=1024, Nbands=100, fmin=50.0, fmax=None, hop_ratio=1.0,
                 alpha: float = 0.2, fixed_frontend_q: bool = False, deltaQ_base: float = 2.0,
                 deltaQ_low_factor: float = 0.5, deltaQ_high_factor: float = 1.0, deltaQ_mode: str = "absolute"):
        super().__init__()
        self.controller_mode = "dual"
        self.fs = fs
        self.timesteps = timesteps
        self.n_fft = n_fft
        self.Nbands = Nbands
        self.alpha = float(alpha)
        fc_np, q0_np = erb_spaced_fc_and_q(Nbands, fmin, (fs / 2 * 0.9) if fmax is None else fmax, erb_factor=1.019)
        self.register_buffer("fc", torch.tensor(fc_np, dtype=torch.float32))
        self.register_buffer("Q0", torch.tensor(q0_np, dtype=torch.float32))
        self.register_buffer("f_fft", torch.linspace(0, fs / 2, n_fft // 2 + 1))
        self.fixed_frontend_q = bool(fixed_frontend_q)
        if self.fixed_frontend_q:
            mk = lambda: FramewiseFixedGammatoneFB(fs=fs, timesteps=timesteps, n_fft=n_fft, Nbands=Nbands, fmin=fmin,
                                                   fmax=fmax, hop_ratio=hop_ratio)
        else:
            mk = lambda: FramewiseAdaptiveGammatoneFB(
                fs=fs, timesteps=timesteps, n_fft=n_fft, Nbands=Nbands, Q0=8.0, fmin=fmin, fmax=fmax,
                hop_ratio=hop_ratio, alpha=alpha, deltaQ_base=deltaQ_base, deltaQ_low_factor=deltaQ_low_factor,
                deltaQ_high_factor=deltaQ_high_factor, deltaQ_mode=deltaQ_mode)
        self.fb_L = mk()
        self.fb_R = mk()
        self.freeze_Q = False   # kept for compatibility; like the reference it is not propagated to fb_L/fb_R
        # Transparent CUDA-graph replay of the front-end's forward and backward for callers that issue it eagerly (the
        # reference's train_biear.py / evaluate_biear.py): see _GraphCache.  Outputs then live in static buffers that the
        # next forward of the same shape / mode overwrites.  Set to False for plain eager launches.
        self.graph_replay = GRAPH_REPLAY_DEFAULT
        self._graphs = _GraphCache()
        # "fused": the whole recurrence in one persistent cluster kernel per direction (csrc/seq.cu);
        # "fused-strict": only its batch-global-fallback replay pass (testing);
        # ("chain": the per-frame cross-check engine of tests/chain_engine.py, when the tests have installed it)
        self.engine = "fused"

    def forward_features(self, wavL_1s: torch.Tensor, wavR_1s: torch.Tensor, want_phase: bool = True,
                         want_cc: bool = False, cc_max_lag_ms: float = 3.0, want_logenergy: bool = False):
        """Everything the back-end consumes, in one call: band energies, Q, spectra, sub-band phases and -- with
        want_cc -- the broadband cross-correlation feature x3 (utils.py:390-420), which the reference precomputes
        offline.  The CC kernel is independent of the recurrence and runs on a forked stream next to it (the
        persistent recurrence kernels occupy 128 of the 148 SMs).  With want_logenergy the log band energies
        clamp(log(Y + 1e-8), +-12) (model_torch.py:1080-1083) come out of the band stage's epilogue as "logYL" / "logYR"
        and their gradient is folded into the backward kernel."""
        mode = "train" if (self.training and torch.is_grad_enabled()) else "eval"
        with timing.span(f"frontend.forward[{mode}]", wavL_1s.shape[0] if wavL_1s.dim() == 2 else 0):
            flags = (bool(want_phase), bool(want_cc), float(cc_max_lag_ms), bool(want_logenergy))
            if self.graph_replay and wavL_1s.is_cuda and wavL_1s.dim() == 2 and wavL_1s.shape == wavR_1s.shape \
                    and not torch.cuda.is_current_stream_capturing():
                hit = self._graphs.lookup(self, wavL_1s, wavR_1s, flags)
                if hit is not None:
                    return hit
            return self._forward_features(wavL_1s, wavR_1s, *flags)

    def _forward_features(self, wavL_1s, wavR_1s, want_phase, want_cc, cc_max_lag_ms, want_logenergy):
        fb = self.fb_L
        if wavL_1s.shape != wavR_1s.shape:
            raise ValueError(f"wavL {tuple(wavL_1s.shape)} and wavR {tuple(wavR_1s.shape)} differ")
        if wavL_1s.dim() != 2:
            raise ValueError(f"Expected wav_1s (B,N), got {tuple(wavL_1s.shape)}")
        if not (wavL_1s.is_cuda and wavR_1s.is_cuda):
            raise RuntimeError(f"biear_b200: waveforms must be CUDA tensors (got {wavL_1s.device}, {wavR_1s.device}); "
                               "there is no CPU path")
        cc = None
        prep = None
        B = wavL_1s.shape[0]
        frozen = (not self.fixed_frontend_q) and self.fb_L.freeze_Q and self.fb_R.freeze_Q
        fused = (not self.fixed_frontend_q) and not frozen and self.engine in ("fused", "fused-strict") \
            and self.fb_L.freeze_Q == self.fb_R.freeze_Q and ops.fused_supported(fb.Nbands, fb.n_fft // 2 + 1)
        if want_cc or fused:
            cur = torch.cuda.current_stream(wavL_1s.device)
            side = _side_stream(wavL_1s.device)
        stft_done = None
        if fused:
            # The spectra-independent part of the recurrence step (weight images, zero state, flags, dropout seed) runs on
            # the forked stream, and BEHIND it, on the same lowest-priority stream, the STFT of both ears in frame-major
            # order with per-(row, frame) ready flags: the recurrence kernel is launched right after the preparation and
            # waits, frame by frame, for the spectra it needs -- the STFT then runs on the SMs the persistent kernel
            # leaves idle instead of in front of it (when the caller's stream has a higher priority, e.g. GraphedStep).
            for w in (wavL_1s, wavR_1s):
                if w.requires_grad:
                    raise RuntimeError("biear_b200: gradients with respect to the waveform are not implemented")
            # (only when the whole batch is resident at once: with several waves of clusters the first wave would wait
            # for the frames of every row)
            # Co-residency: the recurrence kernel spin-waits on flags the STFT kernel (another stream) publishes, so the
            # STFT must be able to run NEXT TO it.  Required: every cluster of the recurrence is resident at once, and
            # -- counted on the device's actual SM count, the 227 KB CTAs of the recurrence own their SM -- at least
            # STFT_PRODUCER_SMS SMs stay free for the producer.  Anything else (small parts, MIG slices, very large
            # batches) takes the STFT-first order.  The STFT is always LAUNCHED first and never waits for anything, so a
            # serialising tool (ncu, compute-sanitizer) runs it to completion before the consumer starts.
            tiles = 2 * ((B + ops.tile_rows() - 1) // ops.tile_rows())
            sms = torch.cuda.get_device_properties(wavL_1s.device).multi_processor_count
            streamed = STREAM_SPECTRA and tiles <= ops.resident_clusters(fb.Nbands, fb.n_fft // 2 + 1, wavL_1s.device) \
                and tiles * ops.CLUSTER_CTAS + STFT_PRODUCER_SMS <= sms
            prep = ops.adaptive_prepare(_controller_weights([self.fb_L, self.fb_R]), B, fb.timesteps, fb.Nbands,
                                        self.training, stream=side, streamed_spectra=streamed)
            if streamed:
                with torch.cuda.stream(side):
                    x = ops.stft_pair(wavL_1s.float().contiguous(), wavR_1s.float().contiguous(), fb.win_fn, fb.fs,
                                      fb.timesteps, fb.win, fb.hop, fb.n_fft, ready=prep.x_ready)
                    stft_done = torch.cuda.Event()
                    stft_done.record(side)
        if stft_done is None:
            x = fb._spectra([wavL_1s, wavR_1s])
        if want_cc:
            # The CC kernel is forked BEHIND the STFT: it becomes eligible together with the recurrence kernel, and when
            # the caller's stream has a higher priority than the (lowest-priority) side stream -- GraphedStep captures on
            # such a stream -- the recurrence's 128 persistent CTAs are placed first and the CC kernel's CTAs run on the 20
            # SMs they leave idle instead of delaying them.
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                cc = ops.cc_feature(wavL_1s.float().contiguous(), wavR_1s.float().contiguous(), fb.fs, fb.Nbands,
                                    cc_max_lag_ms)
        if self.fixed_frontend_q or frozen:
            qf = fb.Q0 if frozen else torch.clamp(fb.Q0, Q_MIN, Q_MAX)
            y, ph = _fixed_bands(x, fb.fc, qf, fb.df, want_phase, fb.cutoff)
            q = qf.view(1, 1, -1).expand(2 * B, fb.timesteps, -1)
            lx = _log_energy(y) if want_logenergy else None
            halves = lambda t: [t[:B], t[B:]] if t is not None else None
            y, q, ph, lx = halves(y), halves(q), halves(ph), halves(lx)
        else:
            if self.fb_L.freeze_Q != self.fb_R.freeze_Q:
                raise NotImplementedError("freeze_Q on one ear only")
            engine = self.engine
            y, q, ph, lx = _adaptive_chain(x, 2, [self.fb_L, self.fb_R], fb.fc, fb.Q0, fb.deltaQ_vec, fb.deltaQ_mode,
                                           fb.df, self.training, False, want_phase, fb.band_mode, fb.cutoff, engine,
                                           want_logy=want_logenergy, prep=prep if engine == self.engine else None)
        if stft_done is not None:           # every later reader of X (the caller included) is ordered behind the STFT
            cur.wait_event(stft_done)
            x.record_stream(cur)
        out = {"YL": y[0], "YR": y[1], "QL": q[0], "QR": q[1], "XL": x[:B], "XR": x[B:]}     # per-ear lists
        if want_phase:
            out["phaseL"], out["phaseR"] = ph
        if want_logenergy:
            out["logYL"], out["logYR"] = lx
        if cc is not None:
            cur.wait_stream(side)
            cc.record_stream(cur)
            out["cc"] = cc
        return out

    def forward(self, wavL_1s: torch.Tensor, wavR_1s: torch.Tensor):
        o = self.forward_features(wavL_1s, wavR_1s, want_phase=False)
        return o["YL"], o["YR"], o["QL"], o["QR"], o["XL"], o["XR"]


class BinauralAdaptiveGammatoneFB_SingleController(_FilterbankBase):
    """model_torch.py:579-776: one controller (GRU(4N->128) + MLP) sets a single Q for both ears; the
    controller sees [YLc, YLmem, YRc, YRmem] with a carried memory (beta = 0.8)."""

    def __init__(self, fs=16000, timesteps=19, n_fft=1024, Nbands=100, fmin=50.0, fmax=None, hop_ratio=1.0,
                 alpha: float = 0.2, fixed_frontend_q: bool = False, deltaQ_base: float = 2.0,
                 deltaQ_low_factor: float = 0.5, deltaQ_high_factor: float = 1.0, deltaQ_mode: str = "absolute"):
        super().__init__()
        self.controller_mode = "single"
        self.alpha = float(alpha)
        self.fixed_frontend_q = bool(fixed_frontend_q)
        self._init_geometry(fs, timesteps, n_fft, Nbands, fmin, fmax, hop_ratio)
        self.register_buffer("deltaQ_vec", make_deltaQ_profile(self.fc, deltaQ_base, deltaQ_low_factor,
                                                               deltaQ_high_factor))
        self.deltaQ_mode = deltaQ_mode.lower()
        self.Q_min, self.Q_max = Q_MIN, Q_MAX
        self.freeze_Q = False
        self.band_mode = "jacobian"
        self.cutoff = ops.DEFAULT_CUTOFF
        self.graph_replay = GRAPH_REPLAY_DEFAULT      # see _GraphCache
        self._graphs = _GraphCache()
        self.engine = "fused"                         # ("chain": the cross-check engine of tests/chain_engine.py)
        if not self.fixed_frontend_q:
            self.q_rnn, self.q_out = _make_controller(4 * Nbands, Nbands)
            self.fb_L = self.fb_R = None
        else:   # no controller parameters exist; the reference delegates to two fixed filterbanks
            self.q_rnn = self.q_out = None
            mk = lambda: FramewiseFixedGammatoneFB(fs=fs, timesteps=timesteps, n_fft=n_fft, Nbands=Nbands, fmin=fmin,
                                                   fmax=fmax, hop_ratio=hop_ratio)
            self.fb_L = mk()
            self.fb_R = mk()

    def forward_features(self, wavL_1s, wavR_1s, want_phase: bool = True, want_logenergy: bool = False):
        mode = "train" if (self.training and torch.is_grad_enabled()) else "eval"
        with timing.span(f"frontend.forward[{mode}]", wavL_1s.shape[0] if wavL_1s.dim() == 2 else 0):
            flags = (bool(want_phase), False, 3.0, bool(want_logenergy))
            if self.graph_replay and wavL_1s.is_cuda and wavL_1s.dim() == 2 and wavL_1s.shape == wavR_1s.shape \
                    and not torch.cuda.is_current_stream_capturing():
                hit = self._graphs.lookup(self, wavL_1s, wavR_1s, flags)
                if hit is not None:
                    return hit
            return self._forward_features(wavL_1s, wavR_1s, *flags)

    def _forward_features(self, wavL_1s, wavR_1s, want_phase, want_cc, cc_max_lag_ms, want_logenergy):
        x = self._spectra([wavL_1s, wavR_1s])
        B = wavL_1s.shape[0]
        if self.fixed_frontend_q or self.freeze_Q:
            qf = torch.clamp(self.Q0, Q_MIN, Q_MAX) if self.fixed_frontend_q else self.Q0
            y, ph = _fixed_bands(x, self.fc, qf, self.df, want_phase, self.cutoff)
            q = qf.view(1, 1, -1).expand(B, self.timesteps, -1)
            y, ph = [y[:B], y[B:]], ([ph[:B], ph[B:]] if ph is not None else None)
            lx = None
        else:
            engine = self.engine
            y, q, ph, lx = _adaptive_chain(x, 2, [self], self.fc, self.Q0, self.deltaQ_vec, self.deltaQ_mode, self.df,
                                           self.training, True, want_phase, self.band_mode, self.cutoff, engine,
                                           want_logy=want_logenergy)
            q = q[0]
        out = {"YL": y[0], "YR": y[1], "QL": q, "QR": q, "XL": x[:B], "XR": x[B:]}
        if want_phase:
            out["phaseL"], out["phaseR"] = ph
        if want_logenergy:
            out["logYL"], out["logYR"] = lx if lx is not None else (_log_energy(y[0]), _log_energy(y[1]))
        return out

    def forward(self, wavL_1s, wavR_1s):
        o = self.forward_features(wavL_1s, wavR_1s, want_phase=False)
        return o["YL"], o["YR"], o["QL"], o["QR"], o["XL"], o["XR"]
