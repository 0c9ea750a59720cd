"""Diagnostic (GPU): error of Z / phase / dphase/dQ of the band kernel vs fp64, next to the fp32 torch formula."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import biear_oracle as orc
from biear_b200 import ops
cfg = orc.FrontEndConfig(deltaq_base=1.0, deltaq_low=0.3, deltaq_high=5.0, deltaq_mode="relative")
c64 = orc.constants(cfg, torch.float64); c32 = orc.constants(cfg)
wl, _ = orc.synth_binaural(16, seed=9)
x64 = orc.stft_frames(torch.from_numpy(wl).double(), cfg, c64["win_fn"])
rs = np.random.RandomState(0)
q = (c64["Q0"] * torch.from_numpy(np.exp(0.5 * rs.standard_normal((16, 100))))).clamp(0.05, 30)
t = 4
m64 = orc.band_moments(x64[:, t], q, c64["fc"], c64["f_fft"])
x32 = x64.to(torch.complex64)
m32 = orc.band_moments(x32[:, t], q.float(), c32["fc"], c32["f_fft"])
one = torch.ones_like(q)
dp64 = orc.dq_closed_form(m64, q, c64["fc"], g_phase=one)
# reference-style autograd in fp32
q32 = q.float().requires_grad_(True)
ph32 = orc.subband_phase(x32[:, t:t+1], q32.unsqueeze(1), c32["f_fft"], c32["fc"])[:, 0]
ph32.sum().backward()
dp_ref32 = q32.grad.double()
xr = torch.view_as_real(x32).contiguous().cuda()
Y = m64["Y"]
for cutoff in (6.0, 8.0, 0.0):
    y, ph, dy, dp = ops.band_forward(xr, t, q.float().cuda(), c32["fc"].cuda(), 15.625, cutoff, True, True)
    dp = dp.cpu().double()
    print(f"cutoff {cutoff}: dP/dQ err ours {float(((dp-dp64).abs()).max()/dp64.abs().max()):.3e}  ref32 {float(((dp_ref32-dp64).abs()).max()/dp64.abs().max()):.3e}")
    w = (m64['Z'].abs()/m64['Z'].abs().max())**2
    print(f"    weighted |Z|^2: ours {float(((dp-dp64).abs()*w).max()/(dp64.abs()*w).max()):.3e} ref32 {float(((dp_ref32-dp64).abs()*w).max()/(dp64.abs()*w).max()):.3e}")
    e_o = (dp-dp64).abs(); e_r = (dp_ref32-dp64).abs()
    print("    median elementwise rel err ours %.3e ref %.3e ; mean ratio ours/ref %.2f" % (float((e_o/dp64.abs()).median()), float((e_r/dp64.abs()).median()), float((e_o/(e_r+1e-30)).median())))
    d = (ph.cpu().double() - torch.atan2(m64['Z'].imag, m64['Z'].real)).abs(); d = torch.minimum(d, 2*np.pi-d)
    dr = (ph32.detach().double() - torch.atan2(m64['Z'].imag, m64['Z'].real)).abs(); dr = torch.minimum(dr, 2*np.pi-dr)
    print(f"    phase err ours max {float(d.max()):.3e} med {float(d.median()):.3e}; ref32 max {float(dr.max()):.3e} med {float(dr.median()):.3e}")
    ey = ((y.cpu().double()-Y).abs()/Y).max(); print(f"    Y elementwise rel err {float(ey):.3e}; dY/dQ err {float((dy.cpu().double()-orc.dq_closed_form(m64,q,c64['fc'],g_y=one)).abs().max()/orc.dq_closed_form(m64,q,c64['fc'],g_y=one).abs().max()):.3e}")
