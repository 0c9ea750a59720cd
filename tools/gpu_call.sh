cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out

timeout 300 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -2
timeout 300 python - <<'PY' 2>&1 | tail -5
import time, numpy as np, torch
from biear_b200 import precompute as pc
rs = np.random.RandomState(0)
n = 8192
wl = rs.uniform(-1, 1, size=(n, 16000)).astype(np.float32); wr = np.roll(wl, 5, axis=1) * 0.8
for fmt in ("passive", "active"):
    pc.precompute(wl[:2048], wr[:2048], None, fmt=fmt)           # warm-up (tables, allocator)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = pc.precompute(wl, wr, None, fmt=fmt)
    dt = time.perf_counter() - t0
    print(f"precompute {fmt}: {n} clips in {dt*1e3:.1f} ms = {n/dt:.0f} clips/s (host arrays in, host arrays out)")
from concurrent.futures import ThreadPoolExecutor
dst = np.empty((1024, 16000), np.float32)
with ThreadPoolExecutor(8) as pool:
    pc._parallel_copy(pool, dst, wl[:1024]); t0 = time.perf_counter(); pc._parallel_copy(pool, dst, wl[1024:2048]); dt = time.perf_counter() - t0
t0 = time.perf_counter(); np.copyto(dst, wl[2048:3072]); d1 = time.perf_counter() - t0
print(f"staging copy: thread pool {dst.nbytes/dt/1e9:.1f} GB/s, one thread {dst.nbytes/d1/1e9:.1f} GB/s")
PY
