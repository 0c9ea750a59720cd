"""CUDA-event timing of the pieces of one front-end step at batch B (default 256): STFT, forward recurrence,
loss + backward recurrence + weight gradients, CC, and the pinned H2D copy of one batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from biear_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
model = bench.build_frontend(dev)
if len(sys.argv) > 3:
    model.eval() if sys.argv[3] == "eval" else None
rs = np.random.RandomState(3)
up = {k: torch.from_numpy(rs.standard_normal((B, 19, 100)).astype(np.float32)).to(dev) for k in ("gYL", "gYR", "gPL", "gPR")}
wl, wr = bench.synth_binaural(B, 1234)
hl, hr = torch.from_numpy(wl).pin_memory(), torch.from_numpy(wr).pin_memory()
wl, wr = hl.to(dev), hr.to(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = {}
def timed(name, fn):
    a, b = ev(), ev()
    a.record(); out = fn(); b.record()
    acc.setdefault(name, []).append((a, b))
    return out
for it in range(reps + 3):
    if it == 3:
        acc.clear()
    for p in model.parameters():
        p.grad = None
    timed("h2d", lambda: (hl.to(dev, non_blocking=True), hr.to(dev, non_blocking=True)))
    x = timed("stft", lambda: model.fb_L._spectra([wl, wr]))
    o = timed("forward (stft + recurrence)", lambda: model.forward_features(wl, wr))
    cc = timed("cc", lambda: ops.cc_feature(wl, wr))
    loss = timed("loss", lambda: (up["gYL"] * torch.log(o["YL"] + 1e-8)).mean() + (up["gYR"] * torch.log(o["YR"] + 1e-8)).mean()
                 + (up["gPL"] * o["phaseL"]).mean() + (up["gPR"] * o["phaseR"]).mean() + (o["QL"] * o["QR"]).mean())
    timed("backward (recurrence + wgrad)", lambda: loss.backward())
torch.cuda.synchronize()
for k, v in acc.items():
    ms = [a.elapsed_time(b) for a, b in v]
    print(f"{k:34s} median {np.median(ms)*1e3:9.1f} us   min {np.min(ms)*1e3:9.1f} us   max {np.max(ms)*1e3:9.1f} us")
t0 = time.perf_counter()
for _ in range(10):
    hl.to(dev, non_blocking=True); hr.to(dev, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"pinned H2D of one batch ({2*B*16000*4/1e6:.1f} MB): {dt*1e3:.2f} ms = {2*B*16000*4/dt/1e9:.1f} GB/s")
