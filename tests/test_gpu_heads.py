"""GPU parity of the fused sector heads (csrc/heads.cu) against the torch.nn modules they replace (the same SubHead
parameters through model_torch.py:869-906's formulation): outputs and every parameter / body gradient, ragged batches,
dropout consistency, and the whole active model with native heads against the reference's golden logits."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(batch, seed=0, n_sectors=8, n_cls=5):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from biear_b200 import model_torch as mt
    torch.manual_seed(seed)
    heads = torch.nn.ModuleList([mt.SubHead(200, n_dist_class=n_cls) for _ in range(n_sectors)]).to(DEV)
    body = torch.randn(batch, 200, device=DEV).clamp_min(0.0)     # the body ends in ReLU (+ dropout)
    return heads, body


def _torch_heads(heads, body):
    outs = [h(body) for h in heads]
    return (torch.cat([o[0] for o in outs], dim=1), torch.cat([o[1] for o in outs], dim=1), torch.stack([o[2] for o in outs], dim=1))


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.mark.parametrize("batch,n_sectors,n_cls", [(256, 8, 5), (33, 8, 5), (1, 3, 2), (70, 2, 8)])
def test_heads_forward_backward_against_torch(batch, n_sectors, n_cls):
    from biear_b200 import ops, _lib
    heads, body = _setup(batch, 1, n_sectors, n_cls)
    heads.eval()
    g = torch.Generator().manual_seed(3)
    ups = [torch.randn(s, generator=g).to(DEV) for s in ((batch, n_sectors), (batch, n_sectors), (batch, n_sectors, n_cls))]
    res = {}
    for native in (False, True):
        b = body.clone().requires_grad_(True)
        for p in heads.parameters():
            p.grad = None
        _lib.reset_launch_count()
        outs = ops.sector_heads(b, list(heads), False) if native else _torch_heads(heads, b)
        sum((u * o).sum() for u, o in zip(ups, outs)).backward()
        torch.cuda.synchronize()
        if native:
            assert _lib.launch_count() == 4        # forward, backward, two fixed-order reductions
        res[native] = ([o.detach().clone() for o in outs], b.grad.clone(), {n: p.grad.clone() for n, p in heads.named_parameters()})
    for a, b_ in zip(res[True][0], res[False][0]):
        assert _rel(a, b_) <= 2e-5
    assert _rel(res[True][1], res[False][1]) <= 2e-5
    worst = max(_rel(res[True][2][n], res[False][2][n]) for n in res[False][2])
    print(f"[heads B={batch} S={n_sectors}] worst parameter-gradient difference vs torch.nn {worst:.2e}")
    assert worst <= 5e-5


def test_heads_dropout_masks_are_regenerated_by_the_backward():
    """Train mode: finite difference of the seed-pinned loss along a random direction (weights and body)."""
    from biear_b200 import ops
    heads, body = _setup(19, 2)
    heads.train()
    up = torch.randn(19, 8, generator=torch.Generator().manual_seed(1)).to(DEV)
    params = list(heads.parameters())

    def loss_fn(b):
        torch.manual_seed(77)
        s, a, d = ops.sector_heads(b, list(heads), True)
        return ((up * s).sum() + (up * a).sum() + (up.unsqueeze(-1) * d).sum()).double()

    b = body.clone().requires_grad_(True)
    loss_fn(b).backward()
    g = [p.grad.clone() for p in params] + [b.grad.clone()]
    torch.manual_seed(5)
    dirs = [torch.randn_like(p) * p.detach().abs().mean().clamp_min(1e-3) for p in params] + [torch.randn_like(body) * 0.1]
    eps = 1e-3
    with torch.no_grad():
        for p, d in zip(params, dirs):
            p.add_(eps * d)
        lp = loss_fn(body + eps * dirs[-1])
        for p, d in zip(params, dirs):
            p.sub_(2 * eps * d)
        lm = loss_fn(body - eps * dirs[-1])
        for p, d in zip(params, dirs):
            p.add_(eps * d)
    fd = float((lp - lm) / (2 * eps))
    an = float(sum((gi.double() * di.double()).sum() for gi, di in zip(g, dirs)))
    assert abs(fd - an) <= 0.03 * abs(an) + 1e-3, (fd, an)
    # about 20 % of the shared activations are dropped: train and eval outputs differ, two seeds differ
    torch.manual_seed(1)
    a1 = ops.sector_heads(body, list(heads), True)[0]
    torch.manual_seed(2)
    a2 = ops.sector_heads(body, list(heads), True)[0]
    assert not torch.equal(a1, a2)


def test_full_active_model_native_heads_match_torch_heads():
    """The drop-in active model with the native heads against the same model with native_heads = False (eval mode)."""
    from biear_b200 import model_torch as mt
    from oracle import biear_oracle as orc
    torch.manual_seed(0)
    m = mt.build_model_active(use_cc=True, fb_alpha=0.0, deltaQ_base=1.0, deltaQ_low_factor=0.3, deltaQ_high_factor=5.0,
                              deltaQ_mode="relative").to(DEV).eval()
    for mod in m.modules():                       # cuDNN's RNN backward insists on training mode (no dropout inside nn.GRU here)
        if isinstance(mod, torch.nn.GRU):
            mod.train()
    wl, wr = orc.synth_binaural(12, seed=9)
    tl, tr = torch.from_numpy(wl).to(DEV), torch.from_numpy(wr).to(DEV)
    x3 = torch.randn(12, 100, device=DEV)
    res = {}
    for native in (True, False):
        m.native_heads = native
        for p in m.parameters():
            p.grad = None
        s, a, d = m(tl, tr, x3)
        (s.sum() + (a * a).sum() + (d * d).sum()).backward()
        res[native] = ((s.detach().clone(), a.detach().clone(), d.detach().clone()),
                       {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    for x, y in zip(res[True][0], res[False][0]):
        assert _rel(x, y) <= 2e-5
    assert res[True][1].keys() == res[False][1].keys()
    for n in res[False][1]:
        assert _rel(res[True][1][n], res[False][1][n]) <= 2e-4, (n, _rel(res[True][1][n], res[False][1][n]))
