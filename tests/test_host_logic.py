"""Host-side logic of the round-2 additions, no GPU needed: the parameter order / flat gradient layout the sector-heads
kernel assumes, the index arithmetic of the single-controller weight images (input permutation, bank skew, W_ih ring
chunking) restated in Python against csrc/seq_single.cu / seq_dev.cuh, the sizing entry points, and the loud failures of
the product path (no PyTorch controller behind the front-end, no CPU tensors)."""
import re

import numpy as np
import pytest
import torch

from biear_b200 import _lib, frontend, model_torch as mt, ops


def test_head_parameter_order_is_the_state_dict_order():
    head = mt.SubHead(200, n_dist_class=5)
    names = [n for n, _ in head.named_parameters()]
    assert tuple(names) == ops.HEAD_TENSOR_NAMES
    lib = _lib.load()
    assert lib.biear_heads_tensors_per_head() == len(names) == 20
    # the flat gradient buffer of one head is the concatenation of its parameters in that order
    assert lib.biear_heads_flat_floats(200, 5) == sum(p.numel() for p in head.parameters()) == 36857
    for d, c in ((200, 5), (64, 2), (200, 8), (4, 1)):
        h = mt.SubHead(d, n_dist_class=c)
        assert lib.biear_heads_flat_floats(d, c) == sum(p.numel() for p in h.parameters())
    assert lib.biear_heads_flat_floats(200, 9) == -1          # more distance classes than the kernel's output tile holds
    assert lib.biear_heads_tile_rows() == 32


def test_heads_reject_bad_geometry_and_cpu_tensors():
    lib = _lib.load()
    prm = _lib.HeadsParams()
    prm.B, prm.S, prm.D, prm.C = 4, 8, 202, 5                 # body width must be a multiple of 4 and <= 200
    from ctypes import byref
    assert lib.biear_heads_fwd(byref(prm), None) == -1 and b"bad geometry" in lib.biear_last_error()
    heads = [mt.SubHead(200) for _ in range(2)]
    with pytest.raises(Exception):
        ops.sector_heads(torch.zeros(3, 200), heads, False)    # CPU tensor: no fallback behind the op
    # ... while the model's back-end takes host tensors through the torch.nn modules (the passive CPU path of the drop-in)
    m = mt.build_model(use_cc=True).eval()
    x = torch.zeros(2, 19, 100)
    s, a, d = m(x, x, torch.zeros(2, 100), x, x)
    assert s.shape == (2, 8) and a.shape == (2, 8) and d.shape == (2, 8, 5)


def _torch_col(k, n):
    """csrc/seq_single.cu torch_col: kernel input order mL | mR | cL | cR -> column of weight_ih (cL | mL | cR | mR)."""
    part, b = divmod(k, n)
    return {0: 1, 1: 3, 2: 0, 3: 2}[part] * n + b


@pytest.mark.parametrize("n", [100, 32, 64, 128, 7])
def test_single_controller_image_geometry(n):
    lib = _lib.load()
    # input permutation: a bijection of the 4N columns that sends the kernel's block order to torch's
    cols = [_torch_col(k, n) for k in range(4 * n)]
    assert sorted(cols) == list(range(4 * n))
    assert cols[:n] == list(range(n, 2 * n)) and cols[n:2 * n] == list(range(3 * n, 4 * n))   # mL, mR (first: they do not
    assert cols[2 * n:3 * n] == list(range(n)) and cols[3 * n:] == list(range(2 * n, 3 * n))   # depend on the frame); cL, cR
    src = open(_lib.os.path.join(_lib.os.path.dirname(_lib.__file__), "csrc", "seq_single.cu")).read()
    assert "part == 0 ? 1 : (part == 1 ? 3 : (part == 2 ? 0 : 2))" in src
    # bank skew of the weight images: for every k the 32 unit columns are permuted, and the 32 lanes of a warp
    # (4 units x 8 k-splits, k = j*8 + ks) hit 32 distinct banks
    for k in range(16):
        assert sorted((u + 4 * (k & 7)) & 31 for u in range(32)) == list(range(32))
    for u0 in range(0, 32, 4):
        banks = {((u + 4 * ks) & 31) for u in range(u0, u0 + 4) for ks in range(8)}
        assert len(banks) == 32
    # ring geometry: padded input width, chunk rows divide it, every chunk is a whole number of 8-wide k-split rounds
    kp = (4 * n + 7) & ~7
    m = kp // 8
    ck = 8 * next(d for d in (5, 4, 3, 2, 1) if m % d == 0)
    assert kp % ck == 0 and ck % 8 == 0 and (ck * 96 * 4) % 16 == 0
    # workspace: 4 CTAs x (resident forward image + streamed W_ih image + resident backward image + two W_ih^T column slices)
    fres = 128 * 96 + 3 * 128 * 32
    bres = n * 32 + 2 * 128 * 32 + 384 * 32
    assert lib.biear_single_workspace_floats(n) == 4 * (fres + kp * 96 + bres + 2 * 384 * 32)
    assert lib.biear_single_supported(n, 513) == 1
    assert lib.biear_single_supported(129, 513) == 0


def test_constants_in_python_match_the_kernel_sources():
    src = open(_lib.os.path.join(_lib.os.path.dirname(_lib.__file__), "csrc", "heads.cu")).read()
    assert float(re.search(r"kDropP = ([0-9.]+)f", src).group(1)) == mt.SubHead(200).shared[2].p
    for name, val in (("kH1", 100), ("kH2", 50), ("kH3", 10)):
        assert re.search(rf"{name} = {val}\b", src)
    single = open(_lib.os.path.join(_lib.os.path.dirname(_lib.__file__), "csrc", "seq_single.cu")).read()
    assert "__fmul_rn(0.8f" in single and "__fmul_rn(0.2f" in single      # beta = 0.8, model_torch.py:716, 770-771


def test_no_pytorch_controller_behind_the_front_end():
    assert frontend.CHAIN_ENGINE is not None          # installed by tests/conftest.py for the cross-check tests ...
    assert not hasattr(frontend, "_ControllerStack")  # ... the product module itself holds no PyTorch controller
    from tests import chain_engine
    assert frontend.CHAIN_ENGINE is chain_engine.run
    m = frontend.BinauralAdaptiveGammatoneFB_SingleController()
    assert m.engine == "fused"
    with pytest.raises(Exception):                     # CPU tensors: the front-end raises, by design
        m(torch.zeros(2, 16000), torch.zeros(2, 16000))
    saved = frontend.CHAIN_ENGINE
    frontend.CHAIN_ENGINE = None
    try:
        with pytest.raises(NotImplementedError):
            frontend._adaptive_chain(torch.zeros(2, 19, 513, dtype=torch.complex64), 1, [], torch.zeros(100), torch.zeros(100),
                                     torch.zeros(100), "relative", 15.625, False, False, False, "jacobian", 6.0, "chain")
    finally:
        frontend.CHAIN_ENGINE = saved


def test_gradient_averaging_op_choice():
    """dist.FlatGradAllReducer: NCCL averages inside the collective (no scaling launch), other backends sum then scale."""
    import inspect
    from biear_b200 import dist as bdist
    src = inspect.getsource(bdist.FlatGradAllReducer.__call__)
    assert "ReduceOp.AVG" in src and 'get_backend(self.group) == "nccl"' in src and "mul_(1.0 / world)" in src
