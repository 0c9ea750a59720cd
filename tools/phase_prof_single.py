"""Per-phase cycle breakdown of the single-controller forward kernel (block 0), diagnostic library (`make prof`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["BIEAR_B200_LIB"] = os.path.join(ROOT, "biear_b200", "lib", "libbiear_b200_prof.so")
sys.path.insert(0, ROOT)
import ctypes
import torch
import biear_b200 as bb
from biear_b200 import _lib
from oracle import biear_oracle as orc
from tests.common import CONFIG_SINGLE
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kw = dict(deltaQ_base=CONFIG_SINGLE["deltaq_base"], deltaQ_low_factor=CONFIG_SINGLE["deltaq_low"],
          deltaQ_high_factor=CONFIG_SINGLE["deltaq_high"], deltaQ_mode=CONFIG_SINGLE["deltaq_mode"])
m = bb.BinauralAdaptiveGammatoneFB_SingleController(**kw)
m.load_state_dict({k: torch.from_numpy(v) for k, v in orc.synth_controller(31, in_mult=4).items()}, strict=False)
m = m.to("cuda:0").train()
m.graph_replay = False
wl, wr = orc.synth_binaural(B, seed=78)
tl, tr = torch.from_numpy(wl).cuda(), torch.from_numpy(wr).cuda()
lib = _lib.load()
buf = (ctypes.c_ulonglong * 16)()
def step():
    o = m.forward_features(tl, tr, want_phase=True, want_logenergy=True)
    (o["logYL"].sum() + o["logYR"].sum() + o["phaseL"].sum() + o["phaseR"].sum() + o["QL"].sum()).backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.check(lib.biear_debug_phase_cycles_single(buf), "phase cycles")
reps = 10
for _ in range(reps):
    step()
torch.cuda.synchronize()
_lib.check(lib.biear_debug_phase_cycles_single(buf), "phase cycles")
names = ["loop tail / head", "state + spectra ready", "band stage (4 items, 16 warps)", "prefetch issue + push + hand-over #1",
         "GRU (W_ih ring, K = 4N + 128) + #2", "memory update + Linear 1 + #3", "LayerNorm 1", "Linear 2 + #4", "LayerNorm 2",
         "Linear 3 + Q + #5"]
vals = [buf[i] / reps for i in range(len(names))]
tot = sum(vals)
print(f"seq1_fwd_kernel (block 0): {tot:.0f} cycles per launch = {tot / 1.965e3:.0f} us at 1965 MHz, batch {B}")
for n, v in zip(names, vals):
    print(f"   {n:42s} {v:10.0f} cyc  {100 * v / tot:5.1f}%  {v / 1.965e3 / 19:6.2f} us/frame")
