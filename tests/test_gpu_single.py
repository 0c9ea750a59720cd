"""GPU parity of the FUSED single-controller front-end (csrc/seq_single.cu forward, seq_bwd_kernel<single> backward)
against the CPU oracle's restatement of model_torch.py:695-776 -- forward and all 14 controller weight gradients, on
multi-tile batches with a ragged last tile; its sub-band phase against the float64 oracle next to the fp32 oracle's own
error; the batch-global non-finite fallback (strict replay pass); train-mode dropout consistency (finite difference);
CUDA-graph replay.  The golden-vector case (reference outputs, batch 2) is tests/test_gpu_parity.py::test_single_controller."""
import numpy as np
import pytest
import torch

from oracle import biear_oracle as orc
from tests.common import CONFIG_SINGLE, RTOL, assert_close, cfg_single, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _kw(d):
    return dict(deltaQ_base=d["deltaq_base"], deltaQ_low_factor=d["deltaq_low"], deltaQ_high_factor=d["deltaq_high"],
                deltaQ_mode=d["deltaq_mode"])


def _model(seed=31, std=0.02, engine="fused", n_bands=100, **over):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import biear_b200
    from biear_b200 import _lib
    _lib.load()
    torch.manual_seed(0)
    m = biear_b200.BinauralAdaptiveGammatoneFB_SingleController(Nbands=n_bands, **{**_kw(CONFIG_SINGLE), **over})
    w = orc.synth_controller(seed, n_bands=n_bands, in_mult=4, out_std=std)
    res = m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=False)
    assert not res.unexpected_keys
    m = m.to(DEV).eval()
    m.engine = engine
    m.graph_replay = False
    return m, w


def _np(t):
    return t.detach().float().cpu().numpy()


def _upstream(batch, seed, n=100):
    rs = np.random.RandomState(seed)
    return {k: rs.standard_normal((batch, 19, n)).astype(np.float32) for k in ("gYL", "gYR", "gQ", "gPL", "gPR", "gLL", "gLR")}


def _loss(o, u, phase_w=0.0):
    loss = (u["gYL"] * torch.log(o["YL"] + 1e-8)).sum() + (u["gYR"] * torch.log(o["YR"] + 1e-8)).sum() + (u["gQ"] * o["QL"]).sum()
    if "logYL" in o:
        loss = loss + (u["gLL"] * o["logYL"]).sum() + (u["gLR"] * o["logYR"]).sum()
    if phase_w:
        loss = loss + phase_w * ((u["gPL"] * o["phaseL"]).sum() + (u["gPR"] * o["phaseR"]).sum())
    return loss


def _oracle(wl, wr, w, up, cfg, dtype=torch.float32, chunk=24):
    """Oracle forward + backward in row chunks (clips are independent; weight gradients add up)."""
    p = orc.to_torch(w, dtype=dtype, requires_grad=True)
    outs = {k: [] for k in ("YL", "YR", "Q")}
    for b0 in range(0, wl.shape[0], chunk):
        sl = slice(b0, b0 + chunk)
        tl, tr = torch.from_numpy(wl[sl]).to(dtype), torch.from_numpy(wr[sl]).to(dtype)
        yl, yr, q, _, _, _ = orc.single_controller_forward(tl, tr, p, cfg)
        u = {k: torch.from_numpy(v[sl]).to(dtype) for k, v in up.items()}
        lx = lambda y: torch.clamp(torch.log(y + 1e-8), -12.0, 12.0)
        loss = (u["gYL"] * torch.log(yl + 1e-8)).sum() + (u["gYR"] * torch.log(yr + 1e-8)).sum() + (u["gQ"] * q).sum() \
            + (u["gLL"] * lx(yl)).sum() + (u["gLR"] * lx(yr)).sum()
        loss.backward()
        for k, v in (("YL", yl), ("YR", yr), ("Q", q)):
            outs[k].append(v.detach().numpy())
    return {k: np.concatenate(v) for k, v in outs.items()}, {k: v.grad.numpy() for k, v in p.items()}


@pytest.mark.parametrize("engine", ["fused", "fused-strict"])
@pytest.mark.parametrize("batch,std", [(5, 0.02), (33, 0.05), (70, 0.02)])
def test_single_fused_forward_backward_against_oracle(batch, std, engine):
    """Ragged batches over one / three / five 16-row tiles (8-clip clusters, the last ones partly or wholly padding)."""
    m, w = _model(std=std, engine=engine)
    wl, wr = orc.synth_binaural(batch, seed=77)
    up = _upstream(batch, 5)
    ref_o, ref_g = _oracle(wl, wr, w, up, cfg_single())
    u = {k: torch.from_numpy(v).to(DEV) for k, v in up.items()}
    from biear_b200 import _lib
    _lib.reset_launch_count()
    o = m.forward_features(torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV), want_phase=False, want_logenergy=True)
    _loss(o, u).backward()
    torch.cuda.synchronize()
    # stft, prepare, forward (fast + replay launch), backward, two pairs of weight-gradient launches: nothing per frame
    assert 0 < _lib.launch_count() <= 10, _lib.launch_count()
    assert_close(_np(o["YL"]), ref_o["YL"], RTOL, "YL")
    assert_close(_np(o["YR"]), ref_o["YR"], RTOL, "YR")
    assert_close(_np(o["QL"]), ref_o["Q"], RTOL, "Q")
    assert o["QR"] is o["QL"] or torch.equal(o["QR"], o["QL"])
    worst = 0.0
    for name, prm in m.named_parameters():
        e = rel_err(_np(prm.grad), ref_g[name])
        worst = max(worst, e)
        assert e <= RTOL, (name, e)
    print(f"[single fused, B={batch}, {engine}] worst weight-gradient error vs fp32 oracle {worst:.2e}")


def test_single_fused_matches_chain_engine_bitwise_semantics():
    """Same weights / clips through the per-frame band kernel + PyTorch controller (the cross-check engine): forward to
    rounding, and fast pass == strict replay pass bit for bit (same arithmetic, same order)."""
    wl, wr = orc.synth_binaural(19, seed=3)
    tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
    res = {}
    for engine in ("chain", "fused", "fused-strict"):
        m, _ = _model(engine=engine)
        with torch.no_grad():
            res[engine] = m.forward_features(tl, tr, want_phase=True)
    for k in ("YL", "YR", "QL"):
        assert_close(_np(res["fused"][k]), _np(res["chain"][k]), 2e-5, k)
        assert torch.equal(res["fused"][k], res["fused-strict"][k]), k
    assert torch.equal(res["fused"]["phaseL"], res["fused-strict"]["phaseL"])


def test_single_fused_phase_against_float64():
    """The sub-band phase is ill-conditioned in fp32 in the reference itself: compare with the float64
    oracle, wrap-aware, next to the fp32 oracle's own error."""
    batch = 6
    m, w = _model()
    wl, wr = orc.synth_binaural(batch, seed=12)
    with torch.no_grad():
        o = m.forward_features(torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV), want_phase=True)
    errs = {}
    for dtype in (torch.float32, torch.float64):
        cfg = cfg_single()
        c = orc.constants(cfg, dtype)
        p = orc.to_torch(w, dtype=dtype)
        with torch.no_grad():
            yl, yr, q, _, xl, xr = orc.single_controller_forward(torch.from_numpy(wl).to(dtype), torch.from_numpy(wr).to(dtype), p, cfg)
            errs[dtype] = (orc.subband_phase(xl, q, c["f_fft"], c["fc"]).double().numpy(),
                           orc.subband_phase(xr, q, c["f_fft"], c["fc"]).double().numpy())
    wrap = lambda d: np.minimum(np.abs(d), 2 * np.pi - np.abs(d))
    for side, ours in enumerate((_np(o["phaseL"]).astype(np.float64), _np(o["phaseR"]).astype(np.float64))):
        ref64, ref32 = errs[torch.float64][side], errs[torch.float32][side]
        e_ours, e_ref = wrap(ours - ref64), wrap(ref32 - ref64)
        print(f"[single fused] phase side {side}: ours vs fp64 max {e_ours.max():.2e} mean {e_ours.mean():.2e}; "
              f"fp32 oracle vs fp64 max {e_ref.max():.2e} mean {e_ref.mean():.2e}")
        assert e_ours.mean() <= max(3.0 * e_ref.mean(), 1e-6)
        assert np.quantile(e_ours, 0.99) <= max(3.0 * np.quantile(e_ref, 0.99), 1e-5)


def test_single_fused_nonfinite_fallback_is_batch_global():
    """model_torch.py:766-768: any non-finite Q_{t+1} resets Q to Q0 and drops the GRU state for the WHOLE batch (the
    carried memory keeps running).  An Inf in one W_ih column makes only a silent clip produce NaN (0 * Inf); every other
    clip must fall back with it.  Fast pass -> flag -> strict replay, against the chain engine and the oracle."""
    batch = 21
    wl, wr = orc.synth_binaural(batch, seed=5)
    wl[17] = 0.0
    wr[17] = 0.0
    tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
    out = {}
    for engine in ("chain", "fused"):
        m, w = _model(engine=engine, std=0.05)
        with torch.no_grad():
            m.q_rnn.weight_ih_l0[5, 17] = float("inf")
            out[engine] = m.forward_features(tl, tr, want_phase=False)
    q0 = _np(m.Q0)
    assert np.array_equal(_np(out["fused"]["QL"]), np.broadcast_to(q0, (batch, 19, 100)))
    for k in ("YL", "YR", "QL"):
        assert torch.isfinite(out["fused"][k]).all()
        assert_close(_np(out["fused"][k]), _np(out["chain"][k]), 2e-5, k)
    # and without the poisoned weight the same clips do adapt
    m, _ = _model(engine="fused", std=0.05)
    with torch.no_grad():
        o = m.forward_features(tl, tr, want_phase=False)
    assert float((o["QL"] - torch.from_numpy(q0).to(DEV)).abs().max()) > 1e-3


def test_single_fused_train_mode_gradient_is_consistent():
    """Dropout on: the backward regenerates the forward's Philox masks (central finite difference of the seed-pinned loss)."""
    batch = 5
    m, _ = _model(std=0.05)
    m.train()
    wl, wr = orc.synth_binaural(batch, seed=8)
    tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
    u = {k: torch.from_numpy(v).to(DEV) for k, v in _upstream(batch, 3).items()}
    params = list(m.parameters())

    def loss_fn():
        torch.manual_seed(123)
        return _loss(m.forward_features(tl, tr, want_phase=False, want_logenergy=True), u).double()

    loss_fn().backward()
    g = [p.grad.clone() for p in params]
    torch.manual_seed(7)
    dirs = [torch.randn_like(p) * p.detach().abs().mean().clamp_min(1e-3) for p in params]
    eps = 2e-3
    with torch.no_grad():
        for p, d in zip(params, dirs):
            p.add_(eps * d)
        lp = loss_fn()
        for p, d in zip(params, dirs):
            p.sub_(2 * eps * d)
        lm = loss_fn()
        for p, d in zip(params, dirs):
            p.add_(eps * d)
    fd = float((lp - lm) / (2 * eps))
    an = float(sum((gi.double() * di.double()).sum() for gi, di in zip(g, dirs)))
    assert abs(fd - an) <= 0.05 * abs(an) + 1e-3, (fd, an)
    # eval and train differ (masks are applied), two train calls with different seeds differ
    torch.manual_seed(1)
    a = m.forward_features(tl, tr, want_phase=False)["QL"]
    torch.manual_seed(2)
    b = m.forward_features(tl, tr, want_phase=False)["QL"]
    assert not torch.equal(a, b)


@pytest.mark.parametrize("nb", [32, 64, 128])
def test_single_fused_band_count_sweep(nb):
    """BASELINE config 5: other band counts (other W_ih ring geometries: 4N = 128 / 256 / 512 input columns)."""
    batch = 9
    m, w = _model(n_bands=nb, std=0.05)
    wl, wr = orc.synth_binaural(batch, seed=21)
    up = _upstream(batch, 4, nb)
    ref_o, ref_g = _oracle(wl, wr, w, up, cfg_single(n_bands=nb))
    u = {k: torch.from_numpy(v).to(DEV) for k, v in up.items()}
    o = m.forward_features(torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV), want_phase=False, want_logenergy=True)
    _loss(o, u).backward()
    assert_close(_np(o["YL"]), ref_o["YL"], RTOL, "YL")
    assert_close(_np(o["QL"]), ref_o["Q"], RTOL, "Q")
    for name, prm in m.named_parameters():
        assert rel_err(_np(prm.grad), ref_g[name]) <= RTOL, (nb, name, rel_err(_np(prm.grad), ref_g[name]))


def test_single_fused_graph_replay_matches_eager():
    """The drop-in's transparent CUDA-graph replay (third call with an unchanged key) returns what the eager call returns."""
    batch = 16
    m, _ = _model(std=0.05)
    m.graph_replay = True
    wl, wr = orc.synth_binaural(batch, seed=2)
    tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
    u = {k: torch.from_numpy(v).to(DEV) for k, v in _upstream(batch, 6).items()}
    res = []
    for _ in range(5):
        for p in m.parameters():
            p.grad = None
        o = m.forward_features(tl, tr, want_phase=True, want_logenergy=True)
        _loss(o, u, 1e-3).backward()
        res.append(({k: o[k].clone() for k in ("YL", "YR", "QL", "phaseL")}, [p.grad.clone() for p in m.parameters()]))
    for k in res[0][0]:
        assert torch.equal(res[0][0][k], res[-1][0][k]), k
    for a, b in zip(res[0][1], res[-1][1]):
        assert torch.equal(a, b)
