"""Timing of the two fixed-Q band-stage variants (weights kernel + GEMM) at a few batch sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biear_b200 import ops
from oracle import biear_oracle as orc
dev = torch.device("cuda", 0)
c = orc.constants(orc.FrontEndConfig())
q = torch.clamp(c["Q0"], 0.05, 30.0).float().to(dev); fcd = c["fc"].float().to(dev)
for B in [int(a) for a in sys.argv[1:]] or (256, 1024):
    xr = torch.randn((2 * B, 19, 513, 2), device=dev)
    for variant in ("ffma", "tc"):
        for _ in range(3):
            ops.band_fixed_forward(xr, q, fcd, 15.625, 6.0, True, variant=variant)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.band_fixed_forward(xr, q, fcd, 15.625, 6.0, True, variant=variant)
        e1.record(); torch.cuda.synchronize()
        print(f"batch {B} variant {variant}: {e0.elapsed_time(e1) * 100:.1f} us per call (weights + GEMM)")
