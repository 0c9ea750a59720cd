"""The reference's FORMULATION (oracle/biear_oracle.py: materialised W(Q) per frame, 19-step Python loop, autograd) run
eagerly on the same B200 in PyTorch -- the competitor SURVEY.md 8(d) config 2 asks for ("PyTorch eager CUDA").  The
reference itself is not on the GPU box; the oracle is its functional restatement, pinned to it by golden vectors.
Measurement tool only (never a product path).  Front-end fwd+bwd with phase, no CC (numpy on the host in the reference)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import biear_oracle as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
cfg = orc.FrontEndConfig(deltaq_base=1.0, deltaq_low=0.3, deltaq_high=5.0, deltaq_mode="relative")
c = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in orc.constants(cfg).items()}
wl, wr = orc.synth_binaural(B, seed=1234)
tl, tr = torch.from_numpy(wl).to(dev), torch.from_numpy(wr).to(dev)
pl = {k: v.to(dev).requires_grad_(True) for k, v in orc.to_torch(orc.synth_controller(11)).items()}
pr = {k: v.to(dev).requires_grad_(True) for k, v in orc.to_torch(orc.synth_controller(12)).items()}
rs = np.random.RandomState(3)
up = {k: torch.from_numpy(rs.standard_normal((B, 19, 100)).astype(np.float32)).to(dev) for k in ("a", "b", "c", "d")}

def step():
    for p in list(pl.values()) + list(pr.values()):
        p.grad = None
    yl, yr, ql, qr, xl, xr = orc.binaural_forward(tl, tr, pl, pr, cfg, c=c)
    phl = orc.subband_phase(xl, ql, c["f_fft"], c["fc"])
    phr = orc.subband_phase(xr, qr, c["f_fft"], c["fc"])
    loss = (up["a"] * orc.log_energy(yl)).mean() + (up["b"] * orc.log_energy(yr)).mean() + (up["c"] * phl).mean() + (up["d"] * phr).mean()
    loss.backward()
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
print(f"reference formulation, PyTorch eager on {torch.cuda.get_device_name(0)}, batch {B}, front-end fwd+bwd + phase: "
      f"{dt * 1e3:.1f} ms/step = {B / dt:.0f} audio-s/s; peak memory {torch.cuda.max_memory_allocated() / 1e9:.2f} GB")
