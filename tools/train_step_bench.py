"""BASELINE.json config 4: one FULL active-mode training step -- front-end + ILD/IPD encoders + body + 8 sector heads
(biear_b200.model_torch.build_model_active), the reference's losses and Q regularisers (train_biear.py:417-431, 476-490),
the two global-norm clips (:523-525) and Adam with the two parameter groups (:617-621) -- captured as one CUDA graph
(forward + backward) followed by clip + optimizer.step, batch 256 per GPU.  With torchrun the gradients are all-reduced
(one flat bucket) before the clips.  Prints ms/step and audio-s/s.

Back-end: the GRU recurrences (csrc/gru.cu) and the sector heads (csrc/heads.cu) are this package's kernels, the body MLP
and the input / weight-gradient products library GEMMs.  argv: [batch] [steps] [libgru]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import bench
from biear_b200 import GraphedStep, model_torch as mt
from biear_b200.dist import FlatGradAllReducer

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
if len(sys.argv) > 3 and sys.argv[3] == "libgru":      # A/B: the encoders' GRU layers through torch.nn.GRU (cuDNN)
    mt._PairEncoder.native_gru = False
torch.manual_seed(0)
model = mt.build_model_active(use_cc=True, fb_alpha=0.0, **bench.CONFIG_YAML)
with torch.no_grad():
    for fb in (model.bifb.fb_L, model.bifb.fb_R):
        torch.nn.init.normal_(fb.q_out[-1].weight, std=0.02)
model = model.to(dev).train()
fb_params = list(model.bifb.parameters()); be_params = [p for n, p in model.named_parameters() if not n.startswith("bifb.")]
params = fb_params + be_params
opt = torch.optim.Adam([{"params": fb_params, "lr": 5e-5}, {"params": be_params, "lr": 1e-4}], weight_decay=1e-5, eps=1e-7,
                       capturable=True, fused=True)
wl, wr = bench.synth_binaural(B, 1234 + rank)
wl, wr = torch.from_numpy(wl).to(dev), torch.from_numpy(wr).to(dev)
rs = np.random.RandomState(5)
y = torch.zeros(B, 8, 7, device=dev)
y[..., 0] = torch.from_numpy((rs.uniform(size=(B, 8)) < 0.25).astype(np.float32)).to(dev)
y[..., 1] = torch.from_numpy(rs.uniform(size=(B, 8)).astype(np.float32)).to(dev)
dist_cls = torch.from_numpy(rs.randint(0, 5, size=(B, 8))).to(dev)
log_q0 = torch.log(model.bifb.Q0 + 1e-8).view(1, 1, -1)
pos_w = torch.tensor(3.0, device=dev)

def loss_fn(a, b):
    from biear_b200 import ops
    x3 = ops.cc_feature(a, b)
    sound, aoa, dl = model(a, b, x3)
    pres = y[..., 0]
    l_sound = F.binary_cross_entropy_with_logits(sound, pres, pos_weight=pos_w)
    l_aoa = (F.smooth_l1_loss(aoa, y[..., 1], beta=0.02, reduction="none") * pres).sum() / pres.sum().clamp_min(1.0)
    l_dist = (F.cross_entropy(dl.reshape(-1, 5), dist_cls.reshape(-1), reduction="none") * pres.reshape(-1)).sum() / pres.sum().clamp_min(1.0)
    reg = ops.q_regularizers(model.last_QL, model.last_QR, model.bifb.Q0, 1e-3, 1e-3)[0]      # train_biear.py:476-490
    return 0.2 * l_sound + 0.45 * l_aoa + 0.35 * l_dist + reg

model._assert_finite = lambda tensors: None      # the finiteness read-back is a host sync: not capturable (checked eagerly below)
step = GraphedStep(loss_fn, (wl, wr), params, warmup=3, flat_grads=world > 1)
red = FlatGradAllReducer(params, flat=step.flat) if world > 1 else None

def update():
    torch.nn.utils.clip_grad_norm_(fb_params, 0.2, foreach=True)
    torch.nn.utils.clip_grad_norm_(be_params, 3.0, foreach=True)
    opt.step()

EAGER_UPDATE = os.environ.get("BIEAR_EAGER_UPDATE") == "1"
update_graph = None

def full_step():
    loss = step(wl, wr)
    if red is not None:
        red()                      # one flat-bucket NCCL all-reduce between the two graphs
    if update_graph is None:
        update()
    else:
        update_graph.replay()
    return loss

for _ in range(5):
    full_step()
torch.cuda.synchronize()
if not EAGER_UPDATE:
    # the two global-norm clips + Adam (capturable) as a second CUDA graph over the static gradient tensors: ~300 tiny
    # foreach launches that are host-bound when issued eagerly
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        update()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    update_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(update_graph, pool=step.pool()):
        update()
    for _ in range(3):
        full_step()
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = full_step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g0.record()
for _ in range(steps):
    step(wl, wr)
g1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"full active training step, batch {B} x {world} GPU(s): {ms:.3f} ms/step = {B * world / ms * 1e3:.0f} audio-s/s "
          f"(graph replay fwd+bwd alone {g0.elapsed_time(g1) / steps:.3f} ms; clips + Adam "
          f"{'eager' if update_graph is None else 'as a second graph'}; loss {float(loss):.4f}; "
          f"{step.launches_per_replay} launches of ours per step)")
if dist is not None:
    dist.destroy_process_group()
