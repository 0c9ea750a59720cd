"""Development check of the fused single-controller kernels: fused vs the per-frame chain engine (same weights / clips),
then a timing at batch 256."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import biear_b200 as bb
from oracle import biear_oracle as orc
from tests.common import CONFIG_SINGLE
from tests import chain_engine
chain_engine.install()

DEV = "cuda:0"
kw = dict(deltaQ_base=CONFIG_SINGLE["deltaq_base"], deltaQ_low_factor=CONFIG_SINGLE["deltaq_low"],
          deltaQ_high_factor=CONFIG_SINGLE["deltaq_high"], deltaQ_mode=CONFIG_SINGLE["deltaq_mode"])


def load_ctrl(m, w):
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    m.load_state_dict(sd, strict=False)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def run(B, engine, train=False, want_phase=True):
    torch.manual_seed(0)
    m = bb.BinauralAdaptiveGammatoneFB_SingleController(**kw)
    load_ctrl(m, orc.synth_controller(31, in_mult=4))
    m = m.to(DEV)
    m.train(train)
    m.engine = engine
    m.graph_replay = False
    wl, wr = orc.synth_binaural(B, seed=77)
    tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
    o = m.forward_features(tl, tr, want_phase=want_phase, want_logenergy=True)
    g = torch.Generator().manual_seed(5)
    keys = ("YL", "YR", "QL", "phaseL", "phaseR", "logYL", "logYR") if want_phase else ("YL", "YR", "QL", "logYL", "logYR")
    ups = {k: torch.randn(o[k].shape, generator=g).to(DEV) for k in keys if k in o}
    loss = sum((ups[k] * o[k]).sum() for k in ups)
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.clone() for n, p in m.named_parameters()}
    return o, grads


for B, wp in ((2, False), (8, False), (33, False), (33, True)):
    of, gf = run(B, "fused", want_phase=wp)
    oc, gc = run(B, "chain", want_phase=wp)
    print(f"B={B} phase in the loss: {wp}")
    for k in (("YL", "YR", "QL", "phaseL", "phaseR", "logYL") if wp else ("YL", "YR", "QL", "logYL")):
        print(f"  {k:8s} fused vs chain {rel(of[k], oc[k]):.2e}")
    worst = max(rel(gf[n], gc[n]) for n in gf)
    print("  worst weight-gradient difference", f"{worst:.2e}", {n: f"{rel(gf[n], gc[n]):.1e}" for n in gf})
    ofs, gfs = run(B, "fused-strict", want_phase=wp)
    print("  strict vs fast: Y", rel(ofs["YL"], of["YL"]), "Q", rel(ofs["QL"], of["QL"]),
          "grads", max(rel(gfs[n], gf[n]) for n in gf))

# timing
B = 256
torch.manual_seed(0)
m = bb.BinauralAdaptiveGammatoneFB_SingleController(**kw)
load_ctrl(m, orc.synth_controller(31, in_mult=4))
m = m.to(DEV).train()
m.graph_replay = False
wl, wr = orc.synth_binaural(B, seed=78)
tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
for engine in ("fused", "chain"):
    m.engine = engine
    def step():
        o = m.forward_features(tl, tr, want_phase=True, want_logenergy=True)
        (o["logYL"].sum() + o["logYR"].sum() + o["phaseL"].sum() + o["phaseR"].sum() + o["QL"].sum()).backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"engine {engine}: {e0.elapsed_time(e1) / 10:.3f} ms per fwd+bwd at batch {B} (eager issue)")
