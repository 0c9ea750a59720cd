"""Shared helpers for the parity tests (mirrors tests/golden/make_golden.py's conventions)."""
import numpy as np
import torch

from oracle import biear_oracle as orc

SUB_STEP = 11
SUB_MIN = 2000

CONFIG_YAML = dict(deltaq_base=1.0, deltaq_low=0.3, deltaq_high=5.0, deltaq_mode="relative")
CONFIG_SINGLE = dict(deltaq_base=2.0, deltaq_low=0.5, deltaq_high=5.0, deltaq_mode="absolute")

# Tolerance contract (BASELINE.json north_star): max relative error <= 1e-4 in fp32 on filterbank
# outputs, CC features and dQ.  "relative" = max-abs error normalised by the max-abs of the reference
# tensor; the element-wise form is checked on elements above 1e-3 of the max.
RTOL = 1e-4


def sub(a, step=SUB_STEP):
    a = np.asarray(a)
    return a.reshape(-1)[::step].copy() if a.size > SUB_MIN else a.copy()


def upstream(batch, seed=3, t=19, n=100):
    rs = np.random.RandomState(seed)
    return {k: rs.standard_normal((batch, t, n)).astype(np.float32)
            for k in ("gYL", "gYR", "gPL", "gPR", "gQL", "gQR")}


def rel_err(ours, ref):
    ours = np.asarray(ours, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = np.max(np.abs(ref))
    return float(np.max(np.abs(ours - ref)) / (den if den > 0 else 1.0))


def elem_rel_err(ours, ref, floor=1e-3):
    ours = np.asarray(ours, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    m = np.abs(ref) > floor * np.max(np.abs(ref))
    if not m.any():
        return 0.0
    return float(np.max(np.abs(ours[m] - ref[m]) / np.abs(ref[m])))


def wrap_err(ours, ref):
    """Max wrap-aware phase difference in radians."""
    d = np.abs(np.asarray(ours, np.float64) - np.asarray(ref, np.float64)) % (2 * np.pi)
    return float(np.max(np.minimum(d, 2 * np.pi - d)))


def assert_close(ours, ref, tol=RTOL, what=""):
    e = rel_err(ours, ref)
    assert e <= tol, f"{what}: max-abs-normalised error {e:.3e} > {tol:.1e}"


def cfg_yaml(**kw):
    return orc.FrontEndConfig(**{**CONFIG_YAML, **kw})


def cfg_single(**kw):
    return orc.FrontEndConfig(**{**CONFIG_SINGLE, **kw})


def loss_a(yl, yr, ql, qr, up):
    return (up["gYL"] * torch.log(yl + 1e-8)).sum() + (up["gYR"] * torch.log(yr + 1e-8)).sum() \
        + (up["gQL"] * ql).sum() + (up["gQR"] * qr).sum()


def oracle_dual_chunked(wl, wr, w_l, w_r, up, cfg, dtype=torch.float32, chunk=32, want_phase=False):
    """The oracle's dual front-end, forward + backward of loss A (+ optionally a phase term), run in row chunks (rows are
    independent; the weight gradients are sums over rows, so they accumulate across chunks) to bound host memory at the
    benchmark batch.  Returns ({YL,YR,QL,QR[,PL,PR]} numpy, {"L.<param>" / "R.<param>": grad numpy})."""
    pl = orc.to_torch(w_l, dtype=dtype, requires_grad=True)
    pr = orc.to_torch(w_r, dtype=dtype, requires_grad=True)
    c = orc.constants(cfg, dtype)
    outs = {k: [] for k in ("YL", "YR", "QL", "QR", "PL", "PR")}
    B = wl.shape[0]
    for lo in range(0, B, chunk):
        sl = slice(lo, min(B, lo + chunk))
        tl = torch.from_numpy(wl[sl]).to(dtype)
        tr = torch.from_numpy(wr[sl]).to(dtype)
        u = {k: torch.from_numpy(v[sl]).to(dtype) for k, v in up.items()}
        yl, yr, ql, qr, xl, xr = orc.binaural_forward(tl, tr, pl, pr, cfg, c=c)
        loss = loss_a(yl, yr, ql, qr, u)
        if want_phase:
            phl = orc.subband_phase(xl, ql, c["f_fft"], c["fc"])
            phr = orc.subband_phase(xr, qr, c["f_fft"], c["fc"])
            loss = loss + 1e-3 * ((u["gPL"] * phl).sum() + (u["gPR"] * phr).sum())
            outs["PL"].append(phl.detach().numpy())
            outs["PR"].append(phr.detach().numpy())
        loss.backward()
        for k, v in (("YL", yl), ("YR", yr), ("QL", ql), ("QR", qr)):
            outs[k].append(v.detach().numpy())
    res = {k: np.concatenate(v) for k, v in outs.items() if v}
    grads = {f"{side}.{k}": p.grad.numpy() for side, prm in (("L", pl), ("R", pr)) for k, p in prm.items()}
    return res, grads
