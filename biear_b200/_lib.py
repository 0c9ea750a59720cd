"""ctypes binding of ``biear_b200/lib/libbiear_b200.so`` (the C ABI declared in ``include/biear_b200.h``).

The library is the product's only compute path: there is no CPU or PyTorch fallback behind these
calls.  If the shared object is missing (not built) or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BIEAR_B200_LIB") or os.path.join(_HERE, "lib", "libbiear_b200.so")
ABI_VERSION = 20

_p = c_void_p
_i = c_int
_l = c_int64
_f = c_float



MAX_CTRL = 2   # BIEAR_MAX_CTRL


class SeqParams(Structure):
    """struct BiearSeqParams of include/biear_b200.h (field for field)."""
    _fields_ = (
        [(n, c_int32) for n in ("G", "E", "B", "T", "N", "F", "Kin", "relative", "training", "force_strict")]
        + [("seed", c_uint64)]
        + [(n, c_float) for n in ("df", "cutoff", "q_min", "q_max")]
        + [(n, c_void_p) for n in ("fc", "q0", "dq")]
        + [(n, c_void_p * MAX_CTRL) for n in (
            "w_ih", "w_hh", "b_ih", "b_hh", "w1", "b1", "ln1_g", "ln1_b", "w2", "b2", "ln2_g", "ln2_b", "w3", "b3")]
        + [(n, c_void_p) for n in ("X", "Y", "phase", "dYdQ", "dPdQ", "Q", "delta",
                                   "gates", "xh1", "d1", "xh2", "d2", "rstd", "yc", "H", "flags")]
        + [(n, c_void_p * MAX_CTRL) for n in ("gY", "gP", "gQ")]
        + [(n, c_void_p) for n in ("GG", "G_a1", "G_v1", "G_a2", "G_v2", "G_pre", "workspace", "seed_ptr", "logY")]
        + [("gLogY", c_void_p * MAX_CTRL)]
        + [("prepared", c_int32), ("x_ready", c_void_p)]
    )


class WgradJob(Structure):
    """struct BiearWgradJob of include/biear_b200.h."""
    _fields_ = [("A", c_void_p), ("a_group_stride", c_int64), ("a_chunk_stride", c_int64), ("Do", c_int32),
                ("Bm", c_void_p), ("b_group_stride", c_int64), ("b_chunk_stride", c_int64), ("Di", c_int32),
                ("chunks", c_int64), ("dW", c_void_p), ("db", c_void_p),
                ("dw_group_stride", c_int64), ("dw_row_stride", c_int64), ("db_group_stride", c_int64),
                ("dW2", c_void_p), ("scale2", c_float)]


class HeadsParams(Structure):
    """struct BiearHeadsParams of include/biear_b200.h."""
    _fields_ = [("B", c_int32), ("S", c_int32), ("D", c_int32), ("C", c_int32), ("training", c_int32),
                ("seed", c_uint64), ("seed_ptr", c_void_p), ("body", c_void_p), ("wptr", c_void_p),
                ("sound", c_void_p), ("aoa", c_void_p), ("dist", c_void_p),
                ("g_sound", c_void_p), ("g_aoa", c_void_p), ("g_dist", c_void_p),
                ("d_body_part", c_void_p), ("dw_part", c_void_p), ("d_body", c_void_p), ("dw", c_void_p)]


class GruParams(Structure):
    """struct BiearGruParams of include/biear_b200.h."""
    _fields_ = [("B", c_int32), ("T", c_int32), ("H", c_int32), ("I", c_int32), ("gi", c_void_p), ("w_hh", c_void_p),
                ("b_hh", c_void_p), ("h_seq", c_void_p), ("h_prev", c_void_p), ("gates", c_void_p), ("dh_seq", c_void_p), ("dgi", c_void_p),
                ("dgh", c_void_p), ("workspace", c_void_p)]


WGRAD_MAX_JOBS = 8

# name -> (restype, argtypes); mirrors include/biear_b200.h one to one
SIGNATURES = {
    "biear_abi_version": (_i, []),
    "biear_last_error": (c_char_p, []),
    "biear_launch_count": (_l, []),
    "biear_reset_launch_count": (None, []),
    "biear_init": (_i, []),
    "biear_stft_fwd": (_i, [_p, _l, _l, _l, _p, _i, _i, _i, _i, _i, _p, _p]),
    "biear_pcm16_to_f32": (_i, [_p, _p, _l, _f, _p]),
    "biear_stft_fwd_pair": (_i, [_p, _p, _l, _l, _l, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "biear_band_fwd": (_i, [_p, _l, _p, _l, _p, _l, _i, _i, _f, _f, _p, _l, _p, _l, _p, _p, _l, _p]),
    "biear_band_bwd": (_i, [_p, _l, _p, _l, _p, _l, _i, _i, _f, _f, _p, _l, _p, _l, _p, _l, _i, _p]),
    "biear_band_fixed_workspace_floats": (_l, [_i]),
    "biear_band_fixed_fwd": (_i, [_p, _l, _p, _p, _l, _i, _i, _f, _f, _p, _l, _p, _l, _p, _p]),
    "biear_band_fixed_tc_workspace_floats": (_l, [_i]),
    "biear_band_fixed_fwd_tc": (_i, [_p, _l, _p, _p, _l, _i, _i, _f, _f, _p, _l, _p, _l, _p, _p]),
    "biear_cc_fwd": (_i, [_p, _p, _l, _l, _l, _i, _i, _p, _p, _i, _p, _p]),
    "biear_adaptive_prepare": (_i, [POINTER(SeqParams), _p]),
    "biear_adaptive_fwd": (_i, [POINTER(SeqParams), _p]),
    "biear_adaptive_bwd": (_i, [POINTER(SeqParams), _p]),
    "biear_adaptive_workspace_floats": (_l, [_i, _i]),
    "biear_debug_phase_cycles": (_i, [_p]),
    "biear_debug_phase_cycles_single": (_i, [_p]),
    "biear_adaptive_occupancy": (_i, [_i, _i, POINTER(c_int), POINTER(c_int)]),
    "biear_adaptive_tile_rows": (_i, []),
    "biear_adaptive_supported": (_i, [_i, _i]),
    "biear_single_supported": (_i, [_i, _i]),
    "biear_single_workspace_floats": (_l, [_i]),
    "biear_wgrad_scratch_floats": (_l, [POINTER(WgradJob), _i, _i, _i]),
    "biear_ctrl_wgrad": (_i, [POINTER(WgradJob), _i, _i, _i, _p, _p]),
    "biear_wgrad_scratch_floats_tc": (_l, [POINTER(WgradJob), _i, _i, _i]),
    "biear_ctrl_wgrad_tc": (_i, [POINTER(WgradJob), _i, _i, _i, _p, _p]),
    "biear_heads_tile_rows": (_i, []),
    "biear_heads_tensors_per_head": (_i, []),
    "biear_heads_flat_floats": (_l, [_i, _i]),
    "biear_heads_fwd": (_i, [POINTER(HeadsParams), _p]),
    "biear_heads_bwd": (_i, [POINTER(HeadsParams), _p]),
    "biear_gru_supported": (_i, [_i]),
    "biear_gru_workspace_floats": (_l, [_i]),
    "biear_gru_fwd": (_i, [POINTER(GruParams), _p]),
    "biear_gru_bwd": (_i, [POINTER(GruParams), _p]),
    "biear_q_regularizers_workspace_floats": (_l, []),
    "biear_q_regularizers": (_i, [_p, _p, _p, _l, _i, _f, _f, _p, _p, _p, _p]),
}

_lib = None
_inited_devices = set()


class BiearLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and type its entry points.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BiearLibraryError(
            f"{LIB_PATH} not found: build it with `make -C biear_b200/csrc` (or "
            f"`python -c 'import __graft_entry__ as g; g.build()'`).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:   # stale build
            raise BiearLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    got = lib.biear_abi_version()
    if got != ABI_VERSION:
        raise BiearLibraryError(f"{LIB_PATH} has ABI version {got}, host code expects {ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load().biear_last_error().decode("utf-8", "replace")
        raise BiearLibraryError(f"{what} failed with code {code}: {msg}")


def ensure_init(device_index: int):
    """Per-device one-time setup; must run outside CUDA-graph capture (it uploads a table)."""
    if device_index in _inited_devices:
        return
    check(load().biear_init(), "biear_init")
    _inited_devices.add(device_index)


def launch_count() -> int:
    return int(load().biear_launch_count())


def reset_launch_count():
    load().biear_reset_launch_count()
