"""One forward of the fixed-Q front-end + phase + CC at batch B (default 1024), for profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import biear_b200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
fb = biear_b200.BinauralAdaptiveGammatoneFB(fixed_frontend_q=True).to(dev).eval()
g = torch.Generator(device="cpu").manual_seed(0)
wl = (torch.rand((B, 16000), generator=g) * 2 - 1).to(dev)
wr = (torch.rand((B, 16000), generator=g) * 2 - 1).to(dev)
with torch.no_grad():
    for _ in range(3):
        o = fb.forward_features(wl, wr, want_phase=True, want_cc=True, want_logenergy=True)
torch.cuda.synchronize()
print("ok", float(o["logYL"].sum()))
