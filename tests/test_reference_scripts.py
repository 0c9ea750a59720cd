"""The reference's entry scripts, BYTE-UNCHANGED, in ACTIVE mode on the CUDA front-end:
train_biear.py (train_biear.py:12-14, 346-356, 457-490, 523-525: builders, model(wavL, wavR, x3), Q regularisers on
model.last_Q / model.bifb.Q0, the two clip_grad_norm_ groups) for one epoch, then evaluate_biear.py (:11, 157-193) on the
checkpoint that run wrote.  The scripts come from /root/reference (build container) or from its byte-identical staging
oracle/_ref (GPU box; sha256 verified against the manifest) and are executed by tools/run_reference_pipeline.py."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scripts_available():
    return any(os.path.exists(os.path.join(b, "train_biear.py"))
               for b in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")))


def test_staged_reference_is_byte_identical():
    """oracle/_ref (when staged) holds what the manifest recorded, and the manifest says staged == source."""
    from oracle import stage_ref
    if not stage_ref.available():
        pytest.skip("oracle/_ref not staged")
    man = stage_ref.verify()
    assert {"model_torch.py", "utils.py", "train_biear.py", "evaluate_biear.py"} <= set(man)
    for rel, rec in man.items():
        assert rec["sha256"] == rec["source_sha256"], rel
        if os.path.exists(os.path.join("/root/reference", rel)):      # build container: compare with the live source too
            import hashlib
            with open(os.path.join("/root/reference", rel), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == rec["sha256"], rel


@pytest.mark.gpu
@pytest.mark.skipif(not _scripts_available(), reason="reference scripts neither at /root/reference nor staged in oracle/_ref")
def test_reference_train_and_evaluate_scripts_run_unchanged_active(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_pipeline.py"), "--clips", "192",
                          "--batch", "64", "--log-dir", str(tmp_path)], capture_output=True, text=True, timeout=1500)
    train_log = (tmp_path / "r2_train_biear_unchanged.log").read_text()
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:] + train_log[-3000:]
    assert "Training finished." in train_log and "Test metrics:" in train_log
    assert "[biear_b200 timing] frontend.forward[train]" in train_log and "frontend.backward" in train_log   # the CUDA path ran
    ev = tmp_path / "r2_evaluate_biear_unchanged.log"
    if "skipping it" in out.stdout:
        pytest.skip("evaluate_biear.py's hard-coded absolute paths cannot be created here")
    eval_log = ev.read_text()
    assert "Missing keys: 0" in eval_log and "Unexpected keys: 0" in eval_log
    assert "Saved metrics to" in eval_log or "overall" in eval_log.lower()
