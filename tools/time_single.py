"""One warm-up + three fwd+bwd steps of the fused single-controller front-end at batch 256 (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import biear_b200 as bb
from oracle import biear_oracle as orc
from tests.common import CONFIG_SINGLE

DEV = "cuda:0"
kw = dict(deltaQ_base=CONFIG_SINGLE["deltaq_base"], deltaQ_low_factor=CONFIG_SINGLE["deltaq_low"],
          deltaQ_high_factor=CONFIG_SINGLE["deltaq_high"], deltaQ_mode=CONFIG_SINGLE["deltaq_mode"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
m = bb.BinauralAdaptiveGammatoneFB_SingleController(**kw)
m.load_state_dict({k: torch.from_numpy(v) for k, v in orc.synth_controller(31, in_mult=4).items()}, strict=False)
m = m.to(DEV).train()
m.graph_replay = False
wl, wr = orc.synth_binaural(B, seed=78)
tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
for _ in range(4):
    o = m.forward_features(tl, tr, want_phase=True, want_logenergy=True)
    (o["logYL"].sum() + o["logYR"].sum() + o["phaseL"].sum() + o["phaseR"].sum() + o["QL"].sum()).backward()
torch.cuda.synchronize()
print("done")
