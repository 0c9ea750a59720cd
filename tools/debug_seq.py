"""Per-frame comparison of the fused engines against the chain engine (debugging aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import biear_b200
from oracle import biear_oracle as orc
DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
kw = dict(deltaQ_base=1.0, deltaQ_low_factor=0.3, deltaQ_high_factor=5.0, deltaQ_mode="relative")
def mk(engine):
    torch.manual_seed(0)
    m = biear_b200.BinauralAdaptiveGammatoneFB(alpha=0.0, **kw)
    for fb, s in ((m.fb_L, 11), (m.fb_R, 12)):
        fb.load_state_dict({k: torch.from_numpy(v) for k, v in orc.synth_controller(s, out_std=0.02).items()}, strict=False)
    m = m.to(DEV).eval(); m.engine = engine
    return m
wl, wr = orc.synth_binaural(B, seed=1234)
tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
rs = np.random.RandomState(5)
up = {k: torch.from_numpy(rs.standard_normal((B, 19, 100)).astype(np.float32)).to(DEV) for k in ("a", "b", "c", "d", "e")}
res = {}
for eng in ("chain", "fused", "fused-strict"):
    m = mk(eng)
    o = m.forward_features(tl, tr)
    ((up["a"] * torch.log(o["YL"] + 1e-8)).sum() + (up["b"] * torch.log(o["YR"] + 1e-8)).sum() + (up["c"] * o["QL"]).sum()
     + (up["d"] * o["QR"]).sum() + 1e-3 * (up["e"] * o["phaseL"]).sum()).backward()
    torch.cuda.synchronize()
    res[eng] = (o, {n: p.grad.clone() for n, p in m.named_parameters()})
ref = res["chain"]
for eng in ("fused", "fused-strict"):
    o, g = res[eng]
    for k in ("YL", "QL", "YR", "QR", "phaseL"):
        d = (o[k] - ref[0][k]).abs().detach()
        per_t = (d.amax(dim=(0, 2)) / ref[0][k].detach().abs().amax()).cpu().numpy()
        print(eng, k, "per-frame err:", " ".join(f"{v:.1e}" for v in per_t))
        if k == "QL":
            per_n = d[:, 1].amax(0).cpu().numpy()
            print("   frame1 per-band err:", " ".join(f"{v:.0e}" for v in per_n))
            per_b = d[:, 1].amax(1).cpu().numpy()
            print("   frame1 per-row err:", " ".join(f"{v:.0e}" for v in per_b[:40]))
    for n in g:
        e = float((g[n] - ref[1][n]).abs().max() / ref[1][n].abs().max().clamp_min(1e-30))
        print(f"   grad {n}: {e:.2e}")
