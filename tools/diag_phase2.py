"""Diagnostic (GPU): split the phase / dphase/dQ error into FFT error and band-kernel error."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import biear_oracle as orc
from biear_b200 import ops
cfg = orc.FrontEndConfig(deltaq_base=1.0, deltaq_low=0.3, deltaq_high=5.0, deltaq_mode="relative")
c64 = orc.constants(cfg, torch.float64); c32 = orc.constants(cfg)
B = 16
wl, _ = orc.synth_binaural(B, seed=9)
x64 = orc.stft_frames(torch.from_numpy(wl).double(), cfg, c64["win_fn"])
x_cpu32 = orc.stft_frames(torch.from_numpy(wl), cfg, c32["win_fn"])            # pocketfft fp32 (reference path)
x_gpu = ops.stft(torch.from_numpy(wl).cuda(), c32["win_fn"].cuda(), cfg.fs, cfg.timesteps, cfg.win, cfg.hop, cfg.n_fft)
x_cufft = torch.fft.rfft(orc.frame_clip(torch.from_numpy(wl), cfg).cuda() * c32["win_fn"].cuda(), n=1024)
def xerr(x): 
    x = x.cpu().to(torch.complex128)
    return float((x - x64).abs().max() / x64.abs().max()), float(((x - x64).abs() / x64.abs()).median())
print("X error (max-abs-norm, median elementwise rel): ours", xerr(x_gpu), "pocketfft32", xerr(x_cpu32), "cufft", xerr(x_cufft), "rounded", xerr(x64.to(torch.complex64)))
rs = np.random.RandomState(0)
q = (c64["Q0"] * torch.from_numpy(np.exp(0.5 * rs.standard_normal((B, 100))))).clamp(0.05, 30)
one = torch.ones_like(q)
res = {}
for t in (2, 4, 9, 15):
    m64 = orc.band_moments(x64[:, t], q, c64["fc"], c64["f_fft"])
    dp64 = orc.dq_closed_form(m64, q, c64["fc"], g_phase=one)
    ph64 = torch.atan2(m64["Z"].imag, m64["Z"].real)
    w1 = (m64["Z"].abs() / m64["Y"]); w2 = w1 ** 2
    def score(ph, dp):
        d = (ph.double() - ph64).abs() % (2 * np.pi); d = torch.minimum(d, 2 * np.pi - d)
        return float((d * w1).max()), float(((dp.double() - dp64).abs() * w2).max() / (dp64.abs() * w2).max()), float((dp.double() - dp64).abs().max() / dp64.abs().max())
    def kern(x):
        xr = torch.view_as_real(x.to(torch.complex64)).contiguous().cuda()
        y, ph, dy, dp = ops.band_forward(xr, t, q.float().cuda(), c32["fc"].cuda(), 15.625, 6.0, True, True)
        return score(ph.cpu(), dp.cpu())
    def ref(x):
        q32 = q.float().requires_grad_(True)
        ph = orc.subband_phase(x[:, t:t+1].to(torch.complex64).cpu(), q32.unsqueeze(1), c32["f_fft"], c32["fc"])[:, 0]
        ph.sum().backward()
        return score(ph.detach(), q32.grad)
    for name, f in (("kernel(X ours)", lambda: kern(x_gpu)), ("kernel(X exact-rounded)", lambda: kern(x64)),
                    ("kernel(X pocketfft)", lambda: kern(x_cpu32)), ("ref32(X pocketfft)", lambda: ref(x_cpu32)),
                    ("ref32(X ours)", lambda: ref(x_gpu)), ("ref32(X exact-rounded)", lambda: ref(x64))):
        res.setdefault(name, []).append(f())
for k, v in res.items():
    v = np.array(v)
    print(f"{k:26s} phase*|Z|/Y {v[:,0].max():.2e}   dP/dQ*|Z|^2 {v[:,1].max():.2e}   dP/dQ max-norm {v[:,2].max():.2e}")
