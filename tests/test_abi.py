"""The drop-in boundary without a GPU: the C-ABI library loads, exports every symbol include/biear_b200.h declares, the
ctypes mirror matches the C structs byte for byte, argument validation fails loudly before any CUDA call, and the host
logic around the kernels (np.interp tables of the CC feature, CPU-tensor rejection) behaves."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "biear_b200.h")


def _declared():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(biear_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from biear_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/biear_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    version = int(re.search(r"#define BIEAR_ABI_VERSION (\d+)", open(HEADER).read()).group(1))
    assert lib.biear_abi_version() == version == _lib.ABI_VERSION


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof from the real header (gcc) against the ctypes mirror."""
    from biear_b200 import _lib
    probe = tmp_path / "probe.c"
    fields_seq = ["G", "seed", "df", "fc", "w_ih", "b3", "X", "Q", "delta", "gates", "H", "flags", "gY", "GG", "workspace", "seed_ptr", "gLogY", "prepared"]
    fields_job = ["A", "Do", "Bm", "Di", "chunks", "dW", "db", "dw_group_stride", "db_group_stride", "dW2", "scale2"]
    fields_heads = ["B", "C", "training", "seed", "seed_ptr", "body", "wptr", "dist", "g_sound", "d_body_part", "dw"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
             'printf("%zu %zu\\n", sizeof(BiearSeqParams), sizeof(BiearWgradJob));']
    lines += [f'printf("%zu\\n", offsetof(BiearSeqParams, {f}));' for f in fields_seq]
    lines += [f'printf("%zu\\n", offsetof(BiearWgradJob, {f}));' for f in fields_job]
    lines += ['printf("%zu\\n", sizeof(BiearHeadsParams));']
    lines += [f'printf("%zu\\n", offsetof(BiearHeadsParams, {f}));' for f in fields_heads]
    fields_gru = ["B", "I", "gi", "w_hh", "b_hh", "h_seq", "h_prev", "gates", "dh_seq", "dgi", "dgh", "workspace"]
    lines += ['printf("%zu\\n", sizeof(BiearGruParams));']
    lines += [f'printf("%zu\\n", offsetof(BiearGruParams, {f}));' for f in fields_gru]
    lines += ["return 0;}"]
    probe.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-o", str(exe), str(probe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert [int(out[0]), int(out[1])] == [ctypes.sizeof(_lib.SeqParams), ctypes.sizeof(_lib.WgradJob)]
    offs = [int(x) for x in out[2:]]
    want = [getattr(_lib.SeqParams, f).offset for f in fields_seq] + [getattr(_lib.WgradJob, f).offset for f in fields_job]
    want += [ctypes.sizeof(_lib.HeadsParams)] + [getattr(_lib.HeadsParams, f).offset for f in fields_heads]
    want += [ctypes.sizeof(_lib.GruParams)] + [getattr(_lib.GruParams, f).offset for f in fields_gru]
    assert offs == want


def test_argument_validation_needs_no_gpu():
    from biear_b200 import _lib
    lib = _lib.load()
    rc = lib.biear_band_fwd(None, 0, None, 0, None, 4, 0, 513, 15.625, 6.0, None, 0, None, 0, None, None, 0, None)
    assert rc == -1 and b"bad shape" in lib.biear_last_error()
    assert lib.biear_band_fwd(None, 0, None, 0, None, 0, 100, 513, 15.625, 6.0, None, 0, None, 0, None, None, 0, None) == 0   # empty batch
    assert lib.biear_stft_fwd(None, 1, 16000, 16000, None, 16000, 19, 842, 842, 512, None, None) == -1        # n_fft != 1024
    prm = _lib.SeqParams()
    prm.G, prm.E, prm.B, prm.T, prm.N, prm.F, prm.Kin = 3, 3, 4, 19, 100, 513, 200
    assert lib.biear_adaptive_fwd(ctypes.byref(prm), None) == -1 and b"controllers" in lib.biear_last_error()
    prm.G = prm.E = 2
    prm.N = 200
    assert lib.biear_adaptive_fwd(ctypes.byref(prm), None) == -1
    assert lib.biear_adaptive_supported(100, 513) == 1 and lib.biear_adaptive_supported(128, 513) == 1
    assert lib.biear_adaptive_supported(129, 513) == 0
    assert lib.biear_adaptive_tile_rows() in (16, 32)
    assert lib.biear_adaptive_workspace_floats(2, 100) > 0
    assert lib.biear_gru_supported(200) == 1 and lib.biear_gru_supported(100) == 1                          # the two encoder layers
    assert lib.biear_gru_supported(232) == 1 and lib.biear_gru_supported(236) == 0        # shared-memory budget
    assert lib.biear_gru_supported(202) == 0 and lib.biear_gru_supported(260) == 0 and lib.biear_gru_workspace_floats(202) == -1
    assert lib.biear_gru_workspace_floats(200) == 4 * 600 * 52
    gp = _lib.GruParams()
    gp.B, gp.T, gp.H, gp.I = 4, 19, 202, 100
    assert lib.biear_gru_fwd(ctypes.byref(gp), None) == -1 and b"bad geometry" in lib.biear_last_error()
    gp.H = 200
    assert lib.biear_gru_fwd(ctypes.byref(gp), None) == -1 and b"null pointer" in lib.biear_last_error()
    job = (_lib.WgradJob * 1)()
    assert lib.biear_wgrad_scratch_floats(job, 1, 2, 16) == -1                                              # null operands
    assert lib.biear_wgrad_scratch_floats(job, 9, 2, 16) == -1                                              # too many jobs


def test_cc_interp_tables_follow_numpy_interp():
    """ops._cc_tables (host, float64) reproduces np.interp over the cropped lag axis (utils.py:408-418)."""
    from biear_b200.ops import _cc_tables
    rs = np.random.RandomState(0)
    for nsamp, fs, lags, ms in ((16000, 16000.0, 100, 3.0), (16000, 16000.0, 64, 1.0), (16000, 16000.0, 128, 5.0),
                                (40, 16000.0, 100, 3.0), (16000, 44100.0, 100, 3.0)):
        k_min, k_max, idx, frac = _cc_tables(nsamp, fs, lags, ms)
        ks = np.arange(k_min, k_max + 1)
        assert np.all(np.abs(ks / fs) <= ms * 1e-3 + 1e-15) and (k_min - 1) / fs < -ms * 1e-3 or k_min == -(nsamp - 1)
        c = rs.standard_normal(len(ks))
        ref = np.interp(np.linspace(-ms * 1e-3, ms * 1e-3, lags), ks / fs, c)
        right = np.minimum(idx + 1, len(ks) - 1)
        ours = c[idx] * (1.0 - frac.astype(np.float64)) + c[right] * frac.astype(np.float64)
        assert np.max(np.abs(ours - ref)) <= 2e-7


def test_no_cpu_fallback():
    import biear_b200
    from biear_b200 import ops
    fb = biear_b200.BinauralAdaptiveGammatoneFB(fixed_frontend_q=True)
    w = torch.zeros(2, 16000)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fb(w, w)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.cc_feature(w, w)
    with pytest.raises(ValueError):
        fb(w[0], w[0])
    # the reference's own error for the active model's finiteness check is kept by the drop-in namespace
    from biear_b200 import model_torch as mt
    assert mt.N_SECTORS == 8 and mt.N_DIST_CLASS == 5 and mt.DATA_DIM == 100


def test_3xtf32_split_keeps_fp32_accuracy():
    """The tensor-core kernels (band_fixed_tc.cu, wgrad_tc_kernel) feed TF32 operands split as x = hi + lo, hi = the top
    19 bits (what tf32_hi() in csrc/tc_dev.cuh computes), lo = x - hi, and form hi*hi + lo*hi + hi*lo with fp32
    accumulation.  Emulated here in numpy (lo additionally truncated to TF32, as the tensor core reads it): the result
    stays within a few 1e-6 of the float64 dot product, where plain TF32 operands are off by ~1e-3 -- the reason the
    split exists (parity contract 1e-4)."""
    rs = np.random.RandomState(0)

    def tf32(x):
        return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

    a = rs.standard_normal((64, 513)).astype(np.float32) * np.linspace(3.0, 0.2, 513, dtype=np.float32)
    w = np.abs(rs.standard_normal((100, 513))).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    a_hi, w_hi = tf32(a), tf32(w)
    a_lo, w_lo = tf32(a - a_hi), tf32(w - w_hi)
    assert np.array_equal(a_hi + (a - a_hi), a)                       # the split is exact in fp32
    split = (a_lo.astype(np.float64) @ w_hi.astype(np.float64).T + a_hi.astype(np.float64) @ w_lo.astype(np.float64).T
             + a_hi.astype(np.float64) @ w_hi.astype(np.float64).T).astype(np.float32)
    plain = (a_hi.astype(np.float64) @ w_hi.astype(np.float64).T).astype(np.float32)
    scale = np.abs(ref).max()
    assert np.abs(split - ref).max() / scale <= 2e-6
    assert np.abs(plain - ref).max() / scale >= 1e-5                  # plain TF32 is far outside what the contract allows


def test_umma_tile_layout_is_a_bijection():
    """tc_tile_off() of csrc/tc_dev.cuh (K-major, no swizzle: core matrices of 8 rows x 16 bytes, chunk-major): every
    (row, k) of a 128 x 16 tile maps to a distinct float slot of the 8 KB tile, 16-byte groups stay together."""
    lbo, sbo = 16 * 128, 128
    off = lambda r, kk: (kk >> 2) * (lbo // 4) + (r >> 3) * (sbo // 4) + (r & 7) * 4 + (kk & 3)
    slots = {off(r, kk) for r in range(128) for kk in range(16)}
    assert slots == set(range(128 * 16))
    for r in range(128):
        for c in range(4):
            assert [off(r, 4 * c + i) for i in range(4)] == list(range(off(r, 4 * c), off(r, 4 * c) + 4))
            assert off(r, 4 * c) % 4 == 0
