"""Run the reference's train_biear.py and then evaluate_biear.py BYTE-UNCHANGED, in ACTIVE mode, against biear_b200's
drop-in modules on a GPU.

    python tools/run_reference_pipeline.py [--clips 6144] [--batch 256] [--log-dir gpurun_out]

What it does (nothing of the reference is modified; its files come from /root/reference or, on the GPU box, from the
byte-identical staging in oracle/_ref -- sha256 checked against the manifest):
  1. writes a synthetic active-wav dataset in the reference's H5 wire format (x1, x2 waveforms, x3 = CC computed by our
     GPU precompute path, y labels) as <ROOT>/anechoic_{train,val,test1}_active_wav.h5.npz (h5py is not in the image);
  2. runs train_biear.py through tools/run_reference_script.py with Active=true, EPOCHS=1, BATCH_SIZE=<batch>;
  3. materialises the absolute paths evaluate_biear.py hard-codes (its CHECKPOINT_PATH and test H5, parsed from the
     script text) as symlinks / files pointing at the run of step 2 and a test dataset;
  4. runs evaluate_biear.py the same way.
Both logs are written to <log-dir>/r2_train_biear_unchanged.log and r2_evaluate_biear_unchanged.log; BIEAR_TIMING=1 makes
the drop-in namespace print the front-end's device / host time per training step at exit.
"""
import argparse
import glob
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def find_scripts():
    for base in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if os.path.exists(os.path.join(base, "train_biear.py")) and os.path.exists(os.path.join(base, "conf", "config.yaml")):
            return base
    raise FileNotFoundError("neither /root/reference nor the staged oracle/_ref holds train_biear.py")


def check_unchanged(base):
    man = os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")
    if not os.path.exists(man):
        return "no manifest (scripts taken from /root/reference directly)"
    with open(man) as f:
        manifest = json.load(f)
    for rel in ("train_biear.py", "evaluate_biear.py"):
        with open(os.path.join(base, rel), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        assert sha == manifest[rel]["source_sha256"], f"{rel} differs from the reference"
    return "sha256 of train_biear.py / evaluate_biear.py equal the reference's"


def make_dataset(path, n, seed):
    """Synthetic active-wav dataset in the H5 wire format (data_h5_save.py:72-81), CC from the GPU precompute path."""
    import bench
    from biear_b200 import precompute
    wl, wr = bench.synth_binaural(n, seed=seed)
    rs = np.random.RandomState(seed + 1)
    y = np.zeros((n, 8, 7), np.float32)
    for i in range(n):
        for s in rs.choice(8, size=rs.randint(1, 4), replace=False):
            y[i, s, 0] = 1.0
            y[i, s, 1] = rs.uniform()
            y[i, s, 2 + rs.randint(5)] = 1.0
    arrays = precompute.precompute(wl, wr, y.reshape(n, -1), fmt="active", chunk=1024)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path + ".npz", **arrays)
    return path


def run_script(script, log, env, *overrides):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), script, *overrides]
    with open(log, "w") as f:
        f.write("$ " + " ".join(cmd) + "\n")
        f.flush()
        rc = subprocess.run(cmd, stdout=f, stderr=subprocess.STDOUT, env=env).returncode
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=6144)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--log-dir", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--skip-evaluate", action="store_true")
    a = ap.parse_args()
    base = find_scripts()
    note = check_unchanged(base)
    os.makedirs(a.log_dir, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="biear_pipeline_")
    data_root = os.path.join(scratch, "data")
    for name, n, seed in (("train", a.clips, 100), ("val", a.batch, 200), ("test1", a.batch, 300)):
        make_dataset(os.path.join(data_root, f"anechoic_{name}_active_wav.h5"), n, seed)
    env = dict(os.environ, BIEAR_TIMING="1")
    env.pop("BIEAR_ALLOW_SYNTHETIC", None)
    runs = os.path.join(scratch, "runs")
    train_log = os.path.join(a.log_dir, "r2_train_biear_unchanged.log")
    rc = run_script(os.path.join(base, "train_biear.py"), train_log, env, "Active=true", "EPOCHS=1",
                    f"BATCH_SIZE={a.batch}", f"ROOT={data_root}", f"RUNS_ROOT={runs}", "PRINT_EVERY=4")
    with open(train_log, "a") as f:
        f.write(f"\n[pipeline] {note}; exit code {rc}\n")
    print(f"train_biear.py (unchanged, Active=true, batch {a.batch}, {a.clips} clips): exit {rc} -> {train_log}")
    if rc != 0:
        sys.stdout.write(open(train_log).read()[-3000:])
        return rc
    run_dirs = glob.glob(os.path.join(runs, "*", "checkpoints", "best.pth"))
    assert run_dirs, "train_biear.py wrote no checkpoints/best.pth"
    run_dir = os.path.dirname(os.path.dirname(run_dirs[0]))
    if a.skip_evaluate:
        return 0
    # ---- evaluate_biear.py: materialise its hard-coded absolute paths ---------------------------------------------
    text = open(os.path.join(base, "evaluate_biear.py")).read()
    ckpt = re.search(r'^CHECKPOINT_PATH\s*=\s*"([^"]+)"', text, re.M).group(1)
    test_h5 = re.search(r'^\s*ROOT\s*=\s*"([^"]+)"', text, re.M).group(1) + "/anechoic_test2_active_wav.h5"
    eval_log = os.path.join(a.log_dir, "r2_evaluate_biear_unchanged.log")
    made = []
    try:
        target_run = os.path.dirname(os.path.dirname(ckpt))
        os.makedirs(os.path.dirname(target_run), exist_ok=True)
        if not os.path.lexists(target_run):
            os.symlink(run_dir, target_run)
            made.append(target_run)
        make_dataset(test_h5, a.batch * 2, 400)
        made.append(test_h5 + ".npz")
        if not os.path.exists(test_h5):
            with open(test_h5, "wb") as f:          # the script only checks that the path exists; data.py reads <path>.npz
                f.write(b"placeholder: the dataset is in the .npz next to this file (h5py is not installed)\n")
            made.append(test_h5)
    except OSError as e:
        print(f"cannot materialise evaluate_biear.py's hard-coded paths here ({e}); skipping it")
        return 0
    rc = run_script(os.path.join(base, "evaluate_biear.py"), eval_log, env)
    with open(eval_log, "a") as f:
        f.write(f"\n[pipeline] {note}; checkpoint = the best.pth written by the train_biear.py run above "
                f"({os.path.basename(run_dir)}); exit code {rc}\n")
    print(f"evaluate_biear.py (unchanged): exit {rc} -> {eval_log}")
    for p in made:
        try:
            os.remove(p)
        except OSError:
            pass
    if rc != 0:
        sys.stdout.write(open(eval_log).read()[-3000:])
    shutil.rmtree(scratch, ignore_errors=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
