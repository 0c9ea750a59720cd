"""GPU parity of the GRU-layer recurrence kernels (csrc/gru.cu) against ``torch.nn.GRU`` itself -- the reference's
encoders ARE nn.GRU (model_torch.py:834-835, 842-843, 855-856, 863-864), so the library module evaluated in float64 on
the host is the oracle here: hidden sequence, dL/dx and the four parameter gradients, ragged batches, both encoder
widths, the whole encoders in float64, and a captured-graph replay."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 3e-5      # max-norm relative; fp32 kernels against a float64 evaluation (contract: 1e-4)


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def _case(B, T, I, H, seed):
    torch.manual_seed(seed)
    gru = torch.nn.GRU(I, H, batch_first=True)
    with torch.no_grad():
        for p in gru.parameters():            # larger than the default init: gates leave the linear region
            p.mul_(2.0)
    x = torch.randn(B, T, I)
    up = torch.randn(B, T, H)
    return gru, x, up


@pytest.mark.parametrize("B,T,I,H", [(256, 19, 100, 200), (256, 19, 200, 100), (33, 19, 100, 200), (1, 19, 100, 200),
                                     (17, 1, 8, 232), (5, 3, 12, 8), (70, 7, 36, 132)])
def test_gru_layer_against_float64_nn_gru(B, T, I, H):
    _need_gpu()
    from biear_b200 import ops
    gru, x, up = _case(B, T, I, H, seed=B + H)
    assert ops.gru_supported(H)
    ref = torch.nn.GRU(I, H, batch_first=True).double()
    ref.load_state_dict({k: v.double() for k, v in gru.state_dict().items()})
    xr = x.double().requires_grad_(True)
    yr = ref(xr)[0]
    (yr * up.double()).sum().backward()
    gru = gru.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    yg = ops.gru_layer(xg, gru)
    assert yg.shape == (B, T, H)
    (yg * up.to(DEV)).sum().backward()
    assert _rel(yg, yr) <= TOL
    assert _rel(xg.grad, xr.grad) <= TOL
    for (n, p), (_, q) in zip(gru.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and _rel(p.grad, q.grad) <= TOL, (n, _rel(p.grad, q.grad))
    # bit-identical when repeated (fixed summation order everywhere)
    y2 = ops.gru_layer(xg.detach(), gru)
    assert torch.equal(y2, yg.detach())


def test_gru_layer_element_wise_and_saturated_gates():
    """Element-wise (on |ref| > 1e-3 max) with inputs that drive the gates into saturation: gradient elements that are
    sums of cancelling terms carry fp32 rounding of the terms, in torch.nn.GRU's own fp32 evaluation just as in ours, so
    the bound is 1e-4 or three times the library's fp32-vs-float64 element-wise distance, whichever is larger."""
    _need_gpu()
    from biear_b200 import ops
    gru, x, up = _case(40, 19, 100, 200, seed=5)
    x = x * 6.0

    def host(dtype):
        m = torch.nn.GRU(100, 200, batch_first=True).to(dtype)
        m.load_state_dict({k: v.to(dtype) for k, v in gru.state_dict().items()})
        xi = x.detach().clone().to(dtype).requires_grad_(True)
        yi = m(xi)[0]
        (yi * up.to(dtype)).sum().backward()
        return [yi.detach().double(), xi.grad.double()] + [p.grad.double() for p in m.parameters()]

    want, lib32 = host(torch.float64), host(torch.float32)
    gru = gru.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    yg = ops.gru_layer(xg, gru)
    (yg * up.to(DEV)).sum().backward()
    ours = [yg.detach().double().cpu(), xg.grad.double().cpu()] + [p.grad.double().cpu() for p in gru.parameters()]
    for name, o, w, l in zip(("h", "dx", "dW_ih", "dW_hh", "db_ih", "db_hh"), ours, want, lib32):
        big = w.abs() > 1e-3 * w.abs().max()
        e_ours = float(((o - w).abs() / w.abs())[big].max())
        e_lib = float(((l - w).abs() / w.abs())[big].max())
        print(f"{name}: element-wise ours {e_ours:.2e}  torch fp32 {e_lib:.2e}")
        assert e_ours <= max(1e-4, 3.0 * e_lib), (name, e_ours, e_lib)


@pytest.mark.parametrize("kind", ["ild", "ipd"])
def test_encoders_native_gru_against_float64(kind):
    """ILD / IPD encoder (LayerNorm -> GRU(100->200) -> GRU(200->100) -> mean; model_torch.py:828-867) with the native
    recurrences against the same module evaluated in float64 on the host (CPU tensors take the torch.nn path): output and
    every parameter / input gradient."""
    _need_gpu()
    import copy
    from biear_b200 import model_torch as mt
    torch.manual_seed(2)
    enc = (mt.ILDEncoder if kind == "ild" else mt.IPDEncoder)().train()
    ref = copy.deepcopy(enc).double()
    a, b = torch.randn(96, 19, 100) * 3.0, torch.randn(96, 19, 100) * 3.0
    up = torch.randn(96, 100)
    ar, br = a.double().requires_grad_(True), b.double().requires_grad_(True)
    out_r = ref(ar, br)
    (out_r * up.double()).sum().backward()
    enc = enc.to(DEV)
    assert enc.native_gru
    ag, bg = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    from biear_b200 import _lib
    n0 = _lib.load().biear_launch_count()
    out = enc(ag, bg)
    (out * up.to(DEV)).sum().backward()
    assert _lib.load().biear_launch_count() - n0 == 2 * 3        # per layer: pack + forward, backward
    assert _rel(out, out_r) <= TOL
    assert _rel(ag.grad, ar.grad) <= TOL and _rel(bg.grad, br.grad) <= TOL
    for (n, p), (_, q) in zip(enc.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad, q.grad) <= TOL, (n, _rel(p.grad, q.grad))


def test_native_gru_backward_in_eval_mode_and_graph_replay():
    """The native recurrence has no train-mode restriction in its backward (cuDNN's RNN backward has), and a captured
    forward + backward replays to the eager result with new inputs."""
    _need_gpu()
    from biear_b200 import ops
    gru, x, up = _case(64, 19, 100, 200, seed=11)
    gru = gru.to(DEV).eval()
    up = up.to(DEV)
    xs = [x.to(DEV), torch.randn_like(x).to(DEV)]

    def step(inp):
        for p in gru.parameters():
            p.grad = None
        y = ops.gru_layer(inp, gru)
        (y * up).sum().backward()
        return y.detach().clone()       # (no reference to the autograd graph survives the call: its AccumulateGrad nodes
                                        #  belong to the stream they were created on and must not leak into the capture)

    want = []
    for inp in xs:
        y = step(inp)
        want.append((y, [p.grad.clone() for p in gru.parameters()]))
    static = xs[0].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step(static)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    for p in gru.parameters():
        p.grad = None
    with torch.cuda.graph(graph):
        y_static = ops.gru_layer(static, gru)
        (y_static * up).sum().backward()
    for inp, (y_want, g_want) in zip(xs, want):
        static.copy_(inp)
        graph.replay()
        torch.cuda.synchronize()
        assert _rel(y_static, y_want) <= 1e-6
        for p, g in zip(gru.parameters(), g_want):
            assert _rel(p.grad, g) <= 1e-6      # (the library GEMMs may pick another split under capture)
