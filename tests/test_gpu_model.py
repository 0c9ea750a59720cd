"""The full active model through the drop-in namespace (biear_b200.model_torch) on the GPU, against golden vectors
produced by the unmodified reference (tests/golden/make_model_golden.py): logits / predictions and gradients of a
fixed scalar loss.  This is BASELINE.json config 4's model (front-end + ILD/IPD encoders + heads); the back-end is plain
PyTorch on both sides, so any difference comes from the front-end kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import biear_oracle as orc
from tests.common import RTOL, rel_err, sub
from tests.golden.make_model_golden import GRAD_KEYS, loss_weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG_YAML = dict(deltaQ_base=1.0, deltaQ_low_factor=0.3, deltaQ_high_factor=5.0, deltaQ_mode="relative")


@pytest.fixture(scope="module")
def golden_model():
    return np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))


def _build():
    from biear_b200 import model_torch as mt
    torch.manual_seed(0)                     # same default initialisation as the reference under this seed
    m = mt.build_model_active(use_cc=True, fb_alpha=0.0, fixed_frontend_q=False, **CONFIG_YAML)
    for fb, s in ((m.bifb.fb_L, 11), (m.bifb.fb_R, 12)):
        fb.load_state_dict({k: torch.from_numpy(v) for k, v in orc.synth_controller(s).items()}, strict=False)
    return m.to(DEV).eval()


def test_full_active_model_against_reference(golden_model):
    from biear_b200 import ops
    g = golden_model
    m = _build()
    wl, wr = orc.synth_binaural(3, seed=1234)
    tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
    x3 = ops.cc_feature(tl, tr)                                   # our CC kernel feeds cc_proj
    assert float((x3.cpu() - torch.from_numpy(g["x3"])).abs().max()) <= 1e-6
    with torch.backends.cudnn.flags(enabled=False):   # cuDNN refuses RNN backward in eval mode; the native GRU does not
        sound, aoa, dist = m(tl, tr, x3)
    assert sound.shape == (3, 8) and aoa.shape == (3, 8) and dist.shape == (3, 8, 5)
    for name, t in (("sound", sound), ("aoa", aoa), ("dist", dist)):
        ref32, ref64 = g[f"model32.{name}"], g[f"model64.{name}"]
        e = rel_err(t.detach().cpu().numpy(), ref32)
        assert e <= max(RTOL, 3 * rel_err(ref32, ref64)), (name, e)
    assert m.last_Q is not None and m.last_Q.shape == (3, 19, 100)
    ws, wa, wd = (torch.from_numpy(a).to(DEV) for a in loss_weights(3))
    ((ws * sound).sum() + (wa * aoa).sum() + (wd * dist).sum()).backward()
    params = dict(m.named_parameters())
    for k in GRAD_KEYS:
        ref32, ref64 = g[f"model32.grad.{k}"], g[f"model64.grad.{k}"]
        e = rel_err(sub(params[k].grad.cpu().numpy()), ref32)
        # gradients that pass through the sub-band phase inherit its fp32 conditioning
        assert e <= max(RTOL, 5 * rel_err(ref32, ref64)), (k, e, rel_err(ref32, ref64))


def test_full_model_raises_on_nonfinite_and_cpu_input():
    m = _build()
    wl, wr = orc.synth_binaural(2, seed=5)
    with pytest.raises(RuntimeError):                              # no CPU path by design
        m(torch.from_numpy(wl), torch.from_numpy(wr))
    tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
    with torch.no_grad():
        # NaN recurrent weights: W_hh is unused on a fresh GRU state, so the step after every reset yields a finite Q,
        # the next one NaN -> fallback to Q0 + state reset (model_torch.py:378-380), and so on: even frames are Q0
        m.bifb.fb_L.q_rnn.weight_hh_l0.fill_(float("nan"))
        s, a, d = m(tl, tr, None)
    assert torch.isfinite(s).all() and torch.isfinite(a).all() and torch.isfinite(d).all()
    q0 = m.bifb.Q0.view(1, 1, -1)
    assert torch.equal(m.last_QL[:, 0::2], q0.expand_as(m.last_QL[:, 0::2]))
    assert float((m.last_QL[:, 1::2] - q0).abs().max()) > 1e-3
    assert float((m.last_QR[:, 2::2] - q0).abs().max()) > 1e-3       # the other ear is unaffected


def test_precompute_wire_format_against_oracle():
    """biear_b200.precompute (BASELINE config 3): chunked host->host feature precompute in the reference's H5 wire
    format, against the CPU oracle (fixed-Q filterbank + phase + CC) on a few clips, with a chunk size that does not
    divide the clip count."""
    from biear_b200 import precompute
    wl, wr = orc.synth_binaural(7, seed=21)
    y = np.arange(7 * 56, dtype=np.float32).reshape(7, 56)
    out = precompute.precompute(wl, wr, y, fmt="passive", chunk=3)
    assert {k: v.shape for k, v in out.items()} == {"x1": (7, 19, 100), "x2": (7, 19, 100), "x3": (7, 100), "x4": (7, 19, 100),
                                                    "x5": (7, 19, 100), "y": (7, 56)}
    cfg = orc.FrontEndConfig()
    yl, ql, xl = orc.fixed_fb_forward(torch.from_numpy(wl), cfg)
    c = orc.constants(cfg)
    assert rel_err(out["x1"], orc.log_energy(yl).numpy()) <= RTOL
    assert float(np.max(np.abs(out["x3"] - orc.cc_feature_batch(wl, wr)))) <= 1e-6
    # phase against the float64 truth, weighted by its conditioning abs(Z)/Y (see tests/test_gpu_parity.py)
    c64 = orc.constants(cfg, torch.float64)
    x64 = orc.stft_frames(torch.from_numpy(wl).double(), cfg, c64["win_fn"])
    q64 = torch.clamp(c64["Q0"], orc.Q_MIN, orc.Q_MAX).expand(7, -1)
    mom = [orc.band_moments(x64[:, t], q64, c64["fc"], c64["f_fft"]) for t in range(19)]
    z = torch.stack([m_["Z"] for m_ in mom], 1)
    wgt = (z.abs() / torch.stack([m_["Y"] for m_ in mom], 1)).numpy()
    d = np.abs(out["x4"].astype(np.float64) - torch.atan2(z.imag, z.real).numpy()) % (2 * np.pi)
    assert float(np.max(np.minimum(d, 2 * np.pi - d) * wgt)) <= 1e-6
    act = precompute.precompute(wl, wr, y, fmt="active", chunk=4)
    assert act["x1"].shape == (7, 16000) and np.array_equal(act["x3"], out["x3"]) and np.array_equal(act["x1"], wl)


def test_auralnet_comparison_model_against_reference():
    """build_model_auralnet_active (model_torch.py:1113-1247, 1337-1367) through the drop-in namespace: same default
    initialisation under the seed, outputs and gradients against the unmodified reference's vectors
    (tests/golden/make_auralnet_golden.py).  The two filterbanks run on the STFT + fixed-Q band kernels, the heads on
    csrc/heads.cu."""
    from biear_b200 import _lib, model_torch as mt
    from tests.golden.make_auralnet_golden import GRAD_KEYS as A_KEYS, inputs
    g = np.load(os.path.join(ROOT, "tests", "golden", "auralnet_golden.npz"))
    torch.manual_seed(0)
    m = mt.build_model_auralnet_active().to(DEV).eval()
    assert m.bifb is None and m.last_Q is None
    wl, wr, x3 = (torch.from_numpy(a).to(DEV) for a in inputs())
    from torch.nn.attention import SDPBackend, sdpa_kernel
    n0 = _lib.launch_count()
    with sdpa_kernel(SDPBackend.MATH):               # plain fp32 matmuls in the library attention (no reduced-precision kernel)
        sound, aoa, dist = m(wl, wr, x3)
    assert _lib.launch_count() - n0 >= 3             # STFT + band GEMM per ear, the heads: the native path ran
    for name, t in (("sound", sound), ("aoa", aoa), ("dist", dist)):
        ref32, ref64 = g[f"a32.{name}"], g[f"a64.{name}"]
        e = rel_err(t.detach().cpu().numpy(), ref32)
        assert e <= max(RTOL, 3 * rel_err(ref32, ref64)), (name, e)
    ws, wa, wd = (torch.from_numpy(a).to(DEV) for a in loss_weights(3))
    with sdpa_kernel(SDPBackend.MATH):
        ((ws * sound).sum() + (wa * aoa).sum() + (wd * dist).sum()).backward()
    params = dict(m.named_parameters())
    for k in A_KEYS:
        ref32, ref64 = g[f"a32.grad.{k}"], g[f"a64.grad.{k}"]
        e = rel_err(sub(params[k].grad.cpu().numpy()), ref32)
        assert e <= max(RTOL, 3 * rel_err(ref32, ref64)), (k, e)
    # the reference's constructor keywords (n_bands, not Nbands) and its errors
    fb = mt.AuralNetGammatoneFB(fs=16000, n_bands=64, fmin=80.0, fmax=7000.0, timesteps=19, hop_ratio=1.0, n_fft=1024)
    assert fb.n_bands == 64 and fb.fc.shape == (64,)
    with pytest.raises(ValueError):
        mt.AuralNetGammatoneFB(timesteps=0)
    with pytest.raises(ValueError):
        fb.to(DEV)(wl[0])
