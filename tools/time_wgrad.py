"""Timing of the controller weight-gradient kernels (both variants) on the benchmark's operand shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biear_b200 import ops
dev = torch.device("cuda", 0)
G, S, tiles, R, N, H = 2, 18, 16, 16, 100, 128
K = S * tiles
g = torch.Generator(device="cpu").manual_seed(0)
mk = lambda d: torch.randn((G, K + tiles, d, R), generator=g).to(dev)
GG, yc, hp, a1, d1, a2, d2, pre, v1, x1 = mk(512), mk(N), mk(H), mk(H), mk(H), mk(H), mk(H), mk(N), mk(H), mk(H)
jobs = [(GG, 384, yc, N, K, True), (GG, 256, hp, H, K, True), (GG[:, :, 384:], 128, hp, H, K, True), (a1, H, hp[:, tiles:], H, K, True),
        (a2, H, d1, H, K, True), (pre, N, d2, H, K, True), (v1, H, x1, 0, K, True), (a2, H, x1, 0, K, True)]
import itertools
for variant, jobs in itertools.chain((("ffma", jobs), ("tc", jobs), ("tc", jobs[:6]), ("tc", jobs[:1]), ("ffma", jobs[6:]))):
    print(f"{len(jobs)} job(s):", end=" ")
    for _ in range(3):
        ops.ctrl_wgrad(jobs, variant=variant)
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.ctrl_wgrad(jobs, variant=variant)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(gr):
        keep = ops.ctrl_wgrad(jobs, variant=variant)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"wgrad {variant}: {e0.elapsed_time(e1) * 50:.1f} us per call (partial + reduce, graph replay, operands L2-warm)")
