"""Run one of the reference's entry scripts UNCHANGED against biear_b200's drop-in modules.

    python tools/run_reference_script.py /root/reference/train_biear.py [KEY=VALUE ...]

The script is copied byte-identically into a scratch directory (never into this repo) next to a
conf/config.yaml made of the reference's own config with the KEY=VALUE overrides applied (defaults here:
EPOCHS=1, BATCH_SIZE=32, ROOT/RUNS_ROOT inside the scratch directory), biear_b200/dropin (model_torch, data,
visualize_q) is put first on sys.path, and the copy is executed with runpy as __main__ (INTEGRATION.md route ii).
Active: true needs a CUDA device (the front-end has no CPU path); Active: false runs anywhere.
"""
import os
import runpy
import shutil
import sys
import tempfile

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(argv):
    script = os.path.abspath(argv[0])
    ref_root = os.path.dirname(script)
    scratch = tempfile.mkdtemp(prefix="biear_dropin_")
    shutil.copy(script, os.path.join(scratch, os.path.basename(script)))
    with open(os.path.join(ref_root, "conf", "config.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg.update({"EPOCHS": 1, "BATCH_SIZE": 32, "ROOT": os.path.join(scratch, "data"),
                "RUNS_ROOT": os.path.join(scratch, "runs")})
    for kv in argv[1:]:
        k, v = kv.split("=", 1)
        cfg[k] = yaml.safe_load(v)
    os.makedirs(os.path.join(scratch, "conf"))
    with open(os.path.join(scratch, "conf", "config.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    sys.path.insert(0, os.path.join(ROOT, "biear_b200", "dropin"))
    sys.path.insert(1, ROOT)
    os.chdir(scratch)
    sys.argv = [os.path.join(scratch, os.path.basename(script))]
    runpy.run_path(sys.argv[0], run_name="__main__")
    print("scratch directory:", scratch)


if __name__ == "__main__":
    main(sys.argv[1:])
