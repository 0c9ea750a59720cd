"""The per-frame CROSS-CHECK engine of the adaptive recurrence: one `band_kernel` launch per frame (ops.BandFrame, the
product's per-item band stage) + the Q controller as a plain PyTorch chain (baddbmm / layer_norm / dropout), i.e. the
reference's formulation of model_torch.py:351-380 / 748-771 around our band kernel.  Test infrastructure only: the product
path is the fused persistent kernels; tests/conftest.py installs this as biear_b200.frontend.CHAIN_ENGINE so that
`module.engine = "chain"` selects it."""
from typing import Optional

import torch
import torch.nn.functional as F

from biear_b200 import frontend, ops
from biear_b200.frontend import _log_energy, _next_q


class _ControllerStack:
    """Weights of G structurally identical controllers stacked along a leading group axis, so that the
    per-ear chains of the dual front-end run as one batched chain (autograd un-stacks the gradients)."""

    def __init__(self, mods):
        st = lambda f: torch.stack([f(m) for m in mods])
        self.w_ih = st(lambda m: m.q_rnn.weight_ih_l0).transpose(1, 2)   # (G, in, 384)
        self.w_hh = st(lambda m: m.q_rnn.weight_hh_l0).transpose(1, 2)   # (G, 128, 384)
        self.b_ih = st(lambda m: m.q_rnn.bias_ih_l0).unsqueeze(1)        # (G, 1, 384)
        self.b_hh = st(lambda m: m.q_rnn.bias_hh_l0).unsqueeze(1)
        self.lin = []
        for i in (0, 4, 8):
            self.lin.append((st(lambda m: m.q_out[i].weight).transpose(1, 2),
                             st(lambda m: m.q_out[i].bias).unsqueeze(1)))
        self.ln = []
        for i in (1, 5):
            self.ln.append((st(lambda m: m.q_out[i].weight).unsqueeze(1), st(lambda m: m.q_out[i].bias).unsqueeze(1),
                            mods[0].q_out[i].eps))
        self.p_drop = mods[0].q_out[3].p
        self.hid = mods[0].q_rnn.hidden_size

    def step(self, feat: torch.Tensor, h: Optional[torch.Tensor], training: bool):
        """feat (G,B,in), h (G,B,128) or None -> (pre-tanh output (G,B,N), new h).  torch.nn.GRU gate
        order (r, z, n) and n = tanh(i_n + r * (W_hn h + b_hn)); q_out as model_torch.py:257-267."""
        gi = torch.baddbmm(self.b_ih, feat, self.w_ih)
        if h is None:
            gh = self.b_hh.expand(-1, feat.shape[1], -1)
        else:
            gh = torch.baddbmm(self.b_hh, h, self.w_hh)
        i_r, i_z, i_n = gi.split(self.hid, dim=-1)
        h_r, h_z, h_n = gh.split(self.hid, dim=-1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h_new = (1.0 - z) * n if h is None else (1.0 - z) * n + z * h
        a = h_new
        for k in range(2):
            w, b = self.lin[k]
            g, beta, eps = self.ln[k]
            a = torch.baddbmm(b, a, w)
            a = F.layer_norm(a, (a.shape[-1],), None, None, eps) * g + beta
            a = F.silu(a)
            a = F.dropout(a, self.p_drop, training)
        w, b = self.lin[2]
        return torch.baddbmm(b, a, w), h_new




def run(x, ears, ctrl_mods, fc, q0, dq_vec, dq_mode, df, training, shared, want_phase, band_mode, cutoff, want_logy):
    """Same contract as frontend._adaptive_chain: per-ear lists of Y, Q, phase | None, logY | None."""
    rows, T, Fbins = x.shape
    B = rows // ears
    G = len(ctrl_mods)
    N = fc.numel()
    xr = torch.view_as_real(x)
    stack = _ControllerStack(ctrl_mods)
    q0g = q0.view(1, 1, N)
    dqg = dq_vec.view(1, 1, N)
    q = q0.view(1, 1, N).expand(G, B, N)
    h = None
    mem = None
    ys, qs, phs = [], [], []
    for t in range(T):
        q_rows = (q.expand(ears, B, N) if shared else q).reshape(rows, N)
        y, ph = ops.BandFrame.apply(q_rows, xr, t, fc, df, cutoff, want_phase, band_mode)
        ys.append(y)
        qs.append(q.reshape(G * B, N))
        if want_phase:
            phs.append(ph)
        if t == T - 1:
            # The reference runs the controller once more and discards the result (model_torch.py:361-380);
            # that step only consumes dropout RNG and receives zero gradient, so it is skipped.
            break
        yc = torch.log1p(torch.clamp(y, min=0.0)).view(ears, B, N)
        ycd = yc.detach()
        if shared:
            if mem is None:
                mem = torch.zeros_like(ycd)
            feat = torch.cat([yc[0], mem[0], yc[1], mem[1]], dim=-1).unsqueeze(0)      # (1,B,4N)
        else:
            feat = torch.cat([yc, 0.2 * ycd], dim=-1)                                  # (G,B,2N)
        pre, h = stack.step(feat, h, training)
        q_new = _next_q(torch.tanh(pre), q0g, dqg, dq_mode)
        # batch-global non-finite fallback of the reference, decided per controller, on the device
        ok = torch.isfinite(q_new).flatten(1).all(dim=1).view(G, 1, 1)
        # (built from Q0 / zeros directly: NaN * 0 is NaN, and the reference cuts the graph on the fallback branch)
        q = torch.where(ok, torch.nan_to_num(q_new, nan=0.0, posinf=0.0, neginf=0.0), q0g.expand_as(q_new))
        h = torch.where(ok, torch.nan_to_num(h, nan=0.0, posinf=0.0, neginf=0.0), torch.zeros_like(h))
        if shared:
            mem = 0.8 * mem + 0.2 * ycd
    y_all = torch.stack(ys, dim=1)
    q_all = torch.stack(qs, dim=1)
    ph_all = torch.stack(phs, dim=1) if want_phase else None
    lx_all = _log_energy(y_all) if want_logy else None
    split = lambda t, k: [t[i * B:(i + 1) * B] for i in range(k)] if t is not None else None
    return split(y_all, ears), split(q_all, G), split(ph_all, ears), split(lx_all, ears)


def install():
    frontend.CHAIN_ENGINE = run
