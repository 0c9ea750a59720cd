"""The drop-in boundary: the `model_torch` namespace the reference's scripts import, the `data`
stand-in, and -- when the reference tree is present (this container, not the GPU box) -- the unmodified
train_biear.py running end to end against them (passive mode: the back-end is plain PyTorch, so it runs on CPU)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FIXTURE = os.path.join(ROOT, "tests", "golden", "model_namespace.json")

BUILDS = {
    "active_dual": ("build_model_active", dict(use_cc=True, fb_alpha=0.0, deltaQ_base=1.0, deltaQ_low_factor=0.3,
                                               deltaQ_high_factor=5, deltaQ_mode="relative")),
    "active_fixed": ("build_model_active", dict(fixed_frontend_q=True)),
    "active_single": ("build_model_active_single_controller", dict(deltaQ_base=2.0, deltaQ_low_factor=0.5,
                                                                   deltaQ_high_factor=5)),
    "passive": ("build_model", dict(use_cc=True)),
    "active_auralnet": ("build_model_auralnet_active", dict(use_cc=True)),
}


def _summary(ns, name):
    fn, kw = BUILDS[name]
    torch.manual_seed(0)
    m = getattr(ns, fn)(**kw)
    sd = m.state_dict()
    return {"keys": {k: list(v.shape) for k, v in sd.items()},
            "sum": {k: float(v.double().sum()) for k, v in sd.items()},
            "n_params": sum(p.numel() for p in m.parameters())}


def test_namespace_matches_reference_fixture():
    """State-dict keys, shapes, default initialisation under seed 0 and parameter counts equal the reference's
    (fixture written by tests/golden/make_namespace_fixture.py from /root/reference/model_torch.py)."""
    from biear_b200 import model_torch as ours
    with open(FIXTURE) as f:
        ref = json.load(f)
    assert ours.N_SECTORS == 8 and ours.N_DIST_CLASS == 5 and ours.DATA_DIM == 100
    for name in BUILDS:
        got = _summary(ours, name)
        assert got["keys"] == ref[name]["keys"], name
        assert got["n_params"] == ref[name]["n_params"], name
        for k, v in ref[name]["sum"].items():
            assert abs(got["sum"][k] - v) <= 1e-6 * max(1.0, abs(v)), (name, k)
    assert ref["active_dual"]["n_params"] == 1634780 and ref["active_fixed"]["n_params"] == 1288468


def test_dropin_modules_resolve():
    code = ("import sys; sys.path.insert(0, %r); import model_torch, data, visualize_q; "
            "from model_torch import build_model, build_model_active, N_SECTORS, N_DIST_CLASS; "
            "ds = data.DeepEarH5Dataset_Active('/nonexistent/anechoic_val_active_wav.h5'); "
            "w = ds[0]; assert [tuple(t.shape) for t in w] == [(16000,), (16000,), (100,), (56,)], w; "
            "assert abs(float(max(w[0].abs().max(), w[1].abs().max())) - 1.0) < 1e-6; print(len(ds))"
            ) % os.path.join(ROOT, "biear_b200", "dropin")
    env = dict(os.environ, BIEAR_SYNTH_CLIPS="8", BIEAR_ALLOW_SYNTHETIC="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip().splitlines()[-1] == "8" and "SYNTHETIC" in out.stdout   # the loud warning comes first


def test_data_shim_refuses_missing_datasets_by_default():
    """A mistyped path (or an H5 file without h5py) must not silently train on synthetic data."""
    code = ("import sys; sys.path.insert(0, %r); import data\n"
            "try:\n    data.DeepEarH5Dataset_Active('/nonexistent/anechoic_val_active_wav.h5')\n"
            "except FileNotFoundError as e:\n    print('refused')\n") % os.path.join(ROOT, "biear_b200", "dropin")
    env = {k: v for k, v in os.environ.items() if k != "BIEAR_ALLOW_SYNTHETIC"}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "refused"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_biear.py")), reason="reference tree not present")
def test_reference_train_script_runs_unchanged_passive():
    env = dict(os.environ, BIEAR_SYNTH_CLIPS="48", BIEAR_ALLOW_SYNTHETIC="1", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"),
                          os.path.join(REF, "train_biear.py"), "Active=false", "BATCH_SIZE=16"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Training finished." in out.stdout and "Test metrics:" in out.stdout
