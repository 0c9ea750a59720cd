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


def test_gru_backward_partial_sum_exchange_layout():
    """Index arithmetic of gru_bwd_kernel (csrc/gru.cu) emulated thread by thread in numpy for one cluster: every CTA forms
    the partial sums of W_hh^T dgh over ITS gate rows for all H units in 4 x 4 tiles and sends each unit's partial to the CTA
    that owns the unit (slot = sender CTA x gate-row half, layout [slot][row group][unit][4 rows]); the owner adds its 8
    slots.  Checks that every slot element is written exactly once per step (the mbarrier's expected byte count) and that
    dL/dx, dW_hh equal torch.nn.GRU's autograd in float64 -- including H / 4 not a multiple of 4, where a tile's four
    units belong to two CTAs."""
    import numpy as np
    import torch

    def run(B, T, I, H):
        torch.manual_seed(H)
        cs, HU = 4, H // 4
        O = 3 * HU
        gru = torch.nn.GRU(I, H, batch_first=True).double()
        x = torch.randn(B, T, I, dtype=torch.double, requires_grad=True)
        up = torch.randn(B, T, H, dtype=torch.double)
        (gru(x)[0] * up).sum().backward()
        w_ih, w_hh, b_ih, b_hh = (p.detach().numpy() for p in (gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0))
        sig = lambda v: 1 / (1 + np.exp(-v))
        gi = (x.detach().numpy().reshape(B * T, I) @ w_ih.T + b_ih).reshape(B, T, 3 * H)
        h = np.zeros((B, H))
        gates, hprev = np.zeros((B, T, 4, H)), np.zeros((B, T, H))
        for t in range(T):                      # what gru_fwd_kernel saves
            gh = h @ w_hh.T + b_hh
            r, z = sig(gi[:, t, :H] + gh[:, :H]), sig(gi[:, t, H:2 * H] + gh[:, H:2 * H])
            n = np.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
            hprev[:, t] = h
            h = (1 - z) * n + z * h
            gates[:, t] = np.stack([r, z, n, gh[:, 2 * H:]], 1)
        part_buf = 8 * HU * 16
        w_s = [np.concatenate([w_hh[g * H + c * HU:g * H + (c + 1) * HU] for g in range(3)], 0).reshape(-1) for c in range(cs)]
        d_s = [np.zeros(O * 16) for _ in range(cs)]
        part = [np.full(2 * part_buf, np.nan) for _ in range(cs)]
        direct = np.zeros((cs, H, 4))
        dgi, dgh = np.zeros((B, T, 3 * H)), np.zeros((B, T, 3 * H))
        it = 0
        for t in range(T - 1, -1, -1):
            for c in range(cs):
                for tid in range(H):            # owner role: (unit u, rows 4 rg ..)
                    rg, u = divmod(tid, HU)
                    unit = c * HU + u
                    carry = np.zeros(4)
                    if it > 0:
                        pb = ((it - 1) & 1) * part_buf + (rg * HU + u) * 4
                        for slot in range(8):
                            carry += part[c][pb + slot * 16 * HU: pb + slot * 16 * HU + 4]
                    d = np.zeros((3, 4))
                    for i in range(4):
                        row, dirn = rg * 4 + i, 0.0
                        if row < B:
                            dh = up[row, t, unit].item() + carry[i] + direct[c, tid, i]
                            r, z, n, hn = gates[row, t, :, unit]
                            dnp = dh * (1 - z) * (1 - n * n)
                            dirn = dh * z
                            d[:, i] = dnp * hn * r * (1 - r), dh * (hprev[row, t, unit] - n) * z * (1 - z), dnp * r
                            dgi[row, t, [unit, H + unit, 2 * H + unit]] = d[0, i], d[1, i], dnp
                            dgh[row, t, [unit, H + unit, 2 * H + unit]] = d[:, i]
                        direct[c, tid, i] = dirn
                    for g in range(3):
                        d_s[c][(g * HU + u) * 16 + rg * 4:(g * HU + u) * 16 + rg * 4 + 4] = d[g]
            if t == 0:
                break
            hits = [np.zeros(2 * part_buf, int) for _ in range(cs)]
            for c in range(cs):
                for tid in range(2 * H):        # product role: units 4 jg .. 4 jg + 3 of all H, gate-row half ks
                    ks, tile = divmod(tid, H)
                    rg, jg = divmod(tile, HU)
                    acc = np.zeros((4, 4))
                    for o in (range(0, O // 2) if ks == 0 else range(O // 2, O)):
                        acc += np.outer(w_s[c][o * H + 4 * jg:o * H + 4 * jg + 4], d_s[c][o * 16 + rg * 4:o * 16 + rg * 4 + 4])
                    base = (it & 1) * part_buf + ((c * 2 + ks) * 4 + rg) * HU * 4
                    for i in range(4):
                        dest, uu = divmod(4 * jg + i, HU)
                        part[dest][base + uu * 4:base + uu * 4 + 4] = acc[i]
                        hits[dest][base + uu * 4:base + uu * 4 + 4] += 1
            for c in range(cs):                 # exactly the expected transaction bytes, every element once
                buf = slice((it & 1) * part_buf, (it & 1) * part_buf + part_buf)
                assert hits[c][buf].min() == 1 and hits[c][buf].max() == 1 and hits[c].sum() == part_buf
            it += 1
        dx = (dgi.reshape(B * T, 3 * H) @ w_ih).reshape(B, T, I)
        assert np.abs(dx - x.grad.numpy()).max() <= 1e-12
        assert np.abs(dgh.reshape(-1, 3 * H).T @ hprev.reshape(-1, H) - gru.weight_hh_l0.grad.numpy()).max() <= 1e-12
        assert np.abs(dgi.sum((0, 1)) - gru.bias_ih_l0.grad.numpy()).max() <= 1e-12

    run(5, 4, 6, 8)        # H / 4 = 2
    run(16, 5, 6, 12)      # H / 4 = 3: tiles straddle CTAs
    run(9, 3, 5, 20)       # H / 4 = 5
