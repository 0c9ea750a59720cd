"""First-light check of the tcgen05 fixed-Q band kernel against the FFMA variant and a float64 contraction, plus timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from biear_b200 import ops
from oracle import biear_oracle as orc
dev = torch.device("cuda", 0)
cfg = orc.FrontEndConfig()
c = orc.constants(cfg, dtype=torch.float64)
fc, q0, f_fft = c["fc"], c["Q0"], c["f_fft"]
for rows in (1, 7, 40, 512):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn((rows, 19, 513, 2), generator=g) * torch.linspace(3.0, 0.2, 513).view(1, 1, -1, 1)
    xr = x.to(dev).contiguous()
    q = torch.clamp(q0, 0.05, 30.0).float().to(dev)
    y_tc, p_tc = ops.band_fixed_forward(xr, q, fc.float().to(dev), float(cfg.fs / 2 / 512), 6.0, True, variant="tc")
    y_ff, p_ff = ops.band_fixed_forward(xr, q, fc.float().to(dev), float(cfg.fs / 2 / 512), 6.0, True, variant="ffma")
    torch.cuda.synchronize()
    w = orc.band_weights(torch.clamp(q0, 0.05, 30.0).view(1, -1), fc, f_fft, sanitize=True)[0]          # (N, F) float64
    xc = torch.view_as_complex(x.double().contiguous()).reshape(-1, 513)
    y64 = xc.abs() @ w.T
    z64 = xc @ w.T.to(torch.complex128)
    ref = y64.reshape(rows, 19, -1).numpy()
    e_tc = np.abs(y_tc.cpu().numpy() - ref).max() / np.abs(ref).max()
    e_ff = np.abs(y_ff.cpu().numpy() - ref).max() / np.abs(ref).max()
    ph64 = torch.atan2(z64.imag, z64.real).reshape(rows, 19, -1).numpy()
    wgt = (z64.abs() / z64.abs().max()).reshape(rows, 19, -1).numpy()
    def perr(p):
        d = np.abs(p.cpu().numpy() - ph64) % (2 * np.pi)
        return float((np.minimum(d, 2 * np.pi - d) * wgt).max())
    print(f"rows {rows:4d}: Y err tc {e_tc:.2e} ffma {e_ff:.2e}; weighted phase err tc {perr(p_tc):.2e} ffma {perr(p_ff):.2e}")
for B in (256, 1024):
    xr = torch.randn((2 * B, 19, 513, 2), device=dev)
    q = torch.clamp(q0, 0.05, 30.0).float().to(dev); fcd = fc.float().to(dev)
    for variant in ("ffma", "tc"):
        for _ in range(3):
            ops.band_fixed_forward(xr, q, fcd, 15.625, 6.0, True, variant=variant)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.band_fixed_forward(xr, q, fcd, 15.625, 6.0, True, variant=variant)
        e1.record(); torch.cuda.synchronize()
        print(f"batch {B} variant {variant}: {e0.elapsed_time(e1) * 100:.1f} us per call (weights + GEMM)")
