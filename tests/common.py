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
