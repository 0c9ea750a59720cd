"""Recipe for oracle/_ref/: stage the UNMODIFIED reference so that it can be executed where /root/reference does
not exist (the GPU box).  TEST / BASELINE INFRASTRUCTURE ONLY -- nothing under biear_b200/ imports it.

    python oracle/stage_ref.py            # run by __graft_entry__.build() when /root/reference is present

The reference is pure Python (no build system), so "building" it means placing byte-identical copies of the files the
hot path lives in -- and of the two entry scripts that must run unchanged against the drop-in -- into the git-ignored
oracle/_ref/ (never into history; it travels to the GPU box like a built .so):

    model_torch.py, utils.py            the reference implementation itself (CPU arm of bench.py: kind "reference";
                                        the PyTorch-eager-on-B200 competitor; live cross-check of the oracle port)
    train_biear.py, evaluate_biear.py   executed byte-unchanged by tools/run_reference_script.py / tests
    conf/*.yaml                         their configuration files
    MANIFEST.json                       sha256 of every staged file next to the sha256 of its source

Users: bench.py (--impl reference, cpu_baseline, gpu_eager_reference), tests/test_reference_scripts.py,
tools/run_reference_script.py.  load_reference() imports the staged model_torch / utils with empty stand-ins for
utils.py's absent, unused top-level imports (librosa, gammatone).
"""
import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ("model_torch.py", "utils.py", "train_biear.py", "evaluate_biear.py", "visualize_q.py",
         "conf/config.yaml", "conf/config_single_ctrl.yaml", "conf/config_auralnet_deepear.yaml")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(src=REF_SRC, dst=REF_DST):
    """Copy FILES byte for byte; returns the manifest.  No-op (returns None) when the reference tree is absent."""
    if not os.path.isdir(src):
        return None
    manifest = {}
    for rel in FILES:
        s = os.path.join(src, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = {"sha256": _sha(d), "source_sha256": _sha(s), "bytes": os.path.getsize(d)}
        assert manifest[rel]["sha256"] == manifest[rel]["source_sha256"], rel
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def available(dst=REF_DST):
    return os.path.exists(os.path.join(dst, "model_torch.py")) and os.path.exists(os.path.join(dst, "utils.py"))


def verify(dst=REF_DST):
    """The staged files still are what the manifest recorded (nobody edited them)."""
    with open(os.path.join(dst, "MANIFEST.json")) as f:
        manifest = json.load(f)
    for rel, rec in manifest.items():
        assert _sha(os.path.join(dst, rel)) == rec["source_sha256"], f"oracle/_ref/{rel} differs from the reference"
    return manifest


_loaded = None


def load_reference(dst=REF_DST, model=True):
    """(model_torch, utils) modules of the staged reference, imported under private names so that they can never
    shadow / be shadowed by the drop-in namespace of the same name.  model=False: utils only (no torch import)."""
    global _loaded
    if _loaded is not None and (not model or _loaded[0] is not None):
        return _loaded
    if not available(dst):
        raise FileNotFoundError("oracle/_ref is not staged (run `python oracle/stage_ref.py` where /root/reference exists)")
    verify(dst)
    import importlib.util
    for name in ("librosa", "gammatone", "gammatone.gtgram"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["gammatone"].gtgram = sys.modules["gammatone.gtgram"]
    if not hasattr(sys.modules["gammatone.gtgram"], "gtgram"):
        sys.modules["gammatone.gtgram"].gtgram = None
    mods = []
    for name in ("model_torch", "utils"):
        if name == "model_torch" and not model:
            mods.append(None)
            continue
        spec = importlib.util.spec_from_file_location(f"biear_reference_{name}", os.path.join(dst, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    _loaded = tuple(mods)
    return _loaded


class CcPool:
    """N persistent oracle/cc_server.py workers; map(wav_l, wav_r) -> (B, num_lags) float32 through the reference's own
    compute_cross_correlation_feature, clips dealt round-robin, one feeder thread per worker (pipe I/O releases the GIL)."""

    def __init__(self, workers):
        import subprocess
        env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
        self.procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "cc_server.py")], stdin=subprocess.PIPE,
                                       stdout=subprocess.PIPE, env=env) for _ in range(workers)]
        for p in self.procs:
            assert p.stdout.read(1) == b"R", "cc_server did not start"
        from concurrent.futures import ThreadPoolExecutor
        self.threads = ThreadPoolExecutor(max_workers=workers)

    def _run(self, w, wav_l, wav_r, idx, fs, num_lags, max_lag_ms, out):
        import struct
        import numpy as np
        p = self.procs[w]
        for i in idx:
            a, b = np.ascontiguousarray(wav_l[i], np.float32), np.ascontiguousarray(wav_r[i], np.float32)
            p.stdin.write(struct.pack("<qdqd", a.size, float(fs), int(num_lags), float(max_lag_ms)))
            p.stdin.write(a.tobytes())
            p.stdin.write(b.tobytes())
            p.stdin.flush()
            out[i] = np.frombuffer(p.stdout.read(4 * num_lags), np.float32)

    def submit(self, wav_l, wav_r, fs=16000, num_lags=100, max_lag_ms=3.0):
        """Start computing; returns a function that waits and returns the (B, num_lags) array."""
        import numpy as np
        n = len(self.procs)
        out = np.empty((len(wav_l), num_lags), np.float32)
        futs = [self.threads.submit(self._run, w, wav_l, wav_r, range(w, len(wav_l), n), fs, num_lags, max_lag_ms, out)
                for w in range(n)]

        def wait():
            for f in futs:
                f.result()
            return out
        return wait

    def close(self):
        import struct
        for p in self.procs:
            try:
                p.stdin.write(struct.pack("<qdqd", 0, 0.0, 0, 0.0))
                p.stdin.flush()
                p.stdin.close()
            except OSError:
                pass
        for p in self.procs:
            p.wait(timeout=10)
        self.threads.shutdown()


def cc_worker(args):
    """compute_cross_correlation_feature of the staged reference on one clip (utils.py:390-420); importable by name so that
    a spawn-context process pool (the CPU baseline's stand-in for data_save.py:213-221's pool over files) can run it."""
    left, right, fs, num_lags, max_lag_ms = args
    _, ref_utils = load_reference()
    return ref_utils.compute_cross_correlation_feature(left, right, fs, num_lags, max_lag_ms)


if __name__ == "__main__":
    m = stage()
    if m is None:
        print(f"{REF_SRC} not present: nothing staged")
    else:
        print(f"staged {len(m)} files into {REF_DST}")
