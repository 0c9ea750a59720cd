"""Device time of the GRU-layer recurrence (csrc/gru.cu) next to torch.nn.GRU (cuDNN) at the encoder shapes
(model_torch.py:828-867): forward and forward + backward of one layer, CUDA events, batch 256 x 19 frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biear_b200 import ops

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = 19
PROFILE = len(sys.argv) > 2 and sys.argv[2] == "profile"      # under ncu: a few eager native steps of the wide layer, nothing else


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for I, H in ((100, 200), (200, 100)):
    torch.manual_seed(0)
    gru = torch.nn.GRU(I, H, batch_first=True).to(dev)
    x = torch.randn(B, T, I, device=dev, requires_grad=True)
    up = torch.randn(B, T, H, device=dev)

    def run(native, backward):
        y = ops.gru_layer(x, gru) if native else gru(x)[0]
        if backward:
            x.grad = None
            for p in gru.parameters():
                p.grad = None
            y.backward(up)

    if PROFILE:
        for _ in range(4):
            run(True, True)
        torch.cuda.synchronize()
        sys.exit(0)
    for native in (True, False):
        f = timed(lambda: run(native, False))
        fb = timed(lambda: run(native, True))
        print(f"GRU({I}->{H}) B={B} T={T} {'native ' if native else 'library'}: forward {f:7.1f} us   forward+backward {fb:7.1f} us"
              f"   (eager launches, host-issue bound where short)")
    # the same under a CUDA graph (what the captured training step sees)
    for native in (True, False):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run(native, True)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        x.grad = None
        for p in gru.parameters():
            p.grad = None
        with torch.cuda.graph(g):
            run(native, True)
        t = timed(g.replay)
        print(f"GRU({I}->{H}) B={B} T={T} {'native ' if native else 'library'}: forward+backward as one graph {t:7.1f} us")

# the two recurrence launches alone (C ABI, CUDA events): what is left of the numbers above is the library GEMMs
from ctypes import byref, c_void_p
from biear_b200 import _lib
lib = _lib.load()
for I, H in ((100, 200), (200, 100)):
    f32 = dict(dtype=torch.float32, device=dev)
    gi = torch.randn(B, T, 3 * H, **f32)
    w_hh, b_hh = torch.randn(3 * H, H, **f32) * 0.05, torch.randn(3 * H, **f32) * 0.05
    h_seq, h_prev, gates = torch.empty(B, T, H, **f32), torch.empty(B, T, H, **f32), torch.empty(B, T, 4, H, **f32)
    dh, dgi, dgh = torch.randn(B, T, H, **f32), torch.empty(B, T, 3 * H, **f32), torch.empty(B, T, 3 * H, **f32)
    ws = torch.empty(int(lib.biear_gru_workspace_floats(H)), **f32)
    prm = _lib.GruParams()
    prm.B, prm.T, prm.H, prm.I = B, T, H, I
    ops._fill(prm, gi=gi, w_hh=w_hh, b_hh=b_hh, h_seq=h_seq, h_prev=h_prev, gates=gates, workspace=ws, dh_seq=dh, dgi=dgi, dgh=dgh)
    st = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    tf = timed(lambda: _lib.check(lib.biear_gru_fwd(byref(prm), st), "fwd"))
    tb = timed(lambda: _lib.check(lib.biear_gru_bwd(byref(prm), st), "bwd"))
    print(f"GRU(.->{H}) B={B} T={T}: pack + gru_fwd_kernel {tf:6.1f} us ({tf / T:.2f} us/step)   gru_bwd_kernel {tb:6.1f} us ({tb / T:.2f} us/step)")
