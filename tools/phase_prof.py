"""Per-phase cycle breakdown of the persistent recurrence kernels (block 0), using the diagnostic library
(`make -C biear_b200/csrc prof`).  BIEAR_B200_LIB selects it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["BIEAR_B200_LIB"] = os.path.join(ROOT, "biear_b200", "lib", "libbiear_b200_prof.so")
sys.path.insert(0, ROOT)
import ctypes
import numpy as np, torch
import bench
from biear_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
model = bench.build_frontend(dev)
rs = np.random.RandomState(3)
up = {k: torch.from_numpy(rs.standard_normal((B, 19, 100)).astype(np.float32)).to(dev) for k in ("gYL", "gYR", "gPL", "gPR")}
up["gC"] = torch.from_numpy(rs.standard_normal((B, 100)).astype(np.float32)).to(dev)
step, _ = bench.make_step(model, up)
wl, wr = bench.synth_binaural(B, 1234)
wl, wr = torch.from_numpy(wl).to(dev), torch.from_numpy(wr).to(dev)
lib = _lib.load()
buf = (ctypes.c_ulonglong * 40)()
for _ in range(3):
    step(wl, wr)
torch.cuda.synchronize()
_lib.check(lib.biear_debug_phase_cycles(buf), "phase cycles")
reps = 10
for _ in range(reps):
    step(wl, wr)
torch.cuda.synchronize()
_lib.check(lib.biear_debug_phase_cycles(buf), "phase cycles")
names = [["loop head", "spectra ready (wait + convert)", "band stage (incl. waiting for the band token)",
          "chain barrier + push + hand-over #1", "GRU + #2", "Linear 1 + #3", "LayerNorm 1", "Linear 2 + #4", "LayerNorm 2",
          "Linear 3 + Q + #5"],
         ["loop head", "dL/dpre + push + #1", "Linear 3^T + #2", "LayerNorm 2 bwd", "Linear 2^T + #3", "LayerNorm 1 bwd",
          "Linear 1^T + GRU bwd + #4", "last phase: wait for hand-over #5", "last: issue next step's loads",
          "last: K = 384 transposed products", "last: Y loads + finish_pre", "last: k-split reduction", "last: dh + dL/dpre push"]]
for k, title in enumerate(("seq_fwd2_kernel (chain 0 of block 0)", "seq_bwd_kernel")):
    vals = [buf[k * 16 + i] / reps for i in range(len(names[k]))]
    tot = sum(vals)
    print(f"{title}: {tot:.0f} cycles per launch (block 0) = {tot / 1.965e3:.0f} us at 1965 MHz")
    for n, v in zip(names[k], vals):
        print(f"   {n:36s} {v:10.0f} cyc  {100 * v / tot:5.1f}%  {v / 1.965e3 / (19 if k == 0 else 18):6.2f} us/frame")

