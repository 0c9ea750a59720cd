"""BASELINE.json config 3: throughput of the fixed-Q front-end + sub-band phase + CC feature (the passive / precompute
path), forward only, over batch sizes; resident inputs, CUDA events around CUDA-graph replays; plus the end-to-end
precompute (host arrays in, host arrays out) at the largest size.  Writes a table to stdout."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np, torch
import biear_b200
from biear_b200 import ops, precompute
dev = torch.device("cuda", 0)
A_FIXED = 128000 + 15200 + 15200 + 400         # both ears' wav in; Y, phase (both ears each), CC out (bytes per clip)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0
fb = biear_b200.BinauralAdaptiveGammatoneFB(fixed_frontend_q=True).to(dev).eval()
print(f"{'batch':>6s} {'us/batch':>10s} {'clips/s':>12s} {'GB/s (alg.)':>12s} {'of HBM peak':>12s}")
for B in (64, 256, 1024, 4096):
    g = torch.Generator(device="cpu").manual_seed(B)
    ins = [(torch.rand((B, 16000), generator=g) * 2 - 1).to(dev) for _ in range(4)]
    def run(i):
        wl, wr = ins[i % 4], ins[(i + 1) % 4]
        o = fb.forward_features(wl, wr, want_phase=True, want_cc=True, want_logenergy=True)
        return o["logYL"], o["phaseR"], o["cc"]
    with torch.no_grad():
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run(0)
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        graphs = []
        for i in range(4):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep = run(i)
            graphs.append((gr, keep))
        for gr, _ in graphs:
            gr.replay()
        torch.cuda.synchronize()
        reps = max(8, 8192 // B)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            graphs[r % 4][0].replay()
        e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    cps = B / (us * 1e-6)
    print(f"{B:6d} {us:10.1f} {cps:12.0f} {cps * A_FIXED / 1e9:12.1f} {cps * A_FIXED / 1e9 / peak:12.3f}")
    del graphs
n = 4096
rs = np.random.RandomState(0)
wl = rs.uniform(-1, 1, size=(n, 16000)).astype(np.float32); wr = np.roll(wl, 5, axis=1) * 0.8
precompute.precompute(wl[:1024], wr[:1024])
torch.cuda.synchronize(); t0 = time.perf_counter()
out = precompute.precompute(wl, wr, fmt="passive", chunk=1024)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"precompute.precompute, {n} clips host->host, passive format: {dt * 1e3:.1f} ms = {n / dt:.0f} clips/s")
