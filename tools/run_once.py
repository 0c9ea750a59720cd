"""One forward+backward of the dual adaptive front-end at batch B (default 256), for profiling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
model = bench.build_frontend(dev)
rs = np.random.RandomState(3)
up = {k: torch.from_numpy(rs.standard_normal((B, 19, 100)).astype(np.float32)).to(dev) for k in ("gYL", "gYR", "gPL", "gPR")}
up["gC"] = torch.from_numpy(rs.standard_normal((B, 100)).astype(np.float32)).to(dev)
step, params = bench.make_step(model, up)
wl, wr = bench.synth_binaural(B, 1234)
wl, wr = torch.from_numpy(wl).to(dev), torch.from_numpy(wr).to(dev)
for _ in range(reps):
    loss = step(wl, wr)
torch.cuda.synchronize()
print("loss", float(loss))
