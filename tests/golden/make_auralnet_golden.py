"""Golden vectors for the AuralNet-style comparison model from the UNMODIFIED reference:
    python tests/golden/make_auralnet_golden.py       # writes tests/golden/auralnet_golden.npz

reference.build_model_auralnet_active() under torch.manual_seed(0), eval mode, 3 synthetic clips (one scaled beyond +-1 so
that the input clamp of model_torch.py:1196-1197 acts), x3 random.  Stores the three outputs and a subsample of the
gradients of a fixed scalar loss, in fp32 and (same weights, same graph) in fp64."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import biear_oracle as orc  # noqa: E402
from tests.golden.make_golden import import_reference, sub  # noqa: E402
from tests.golden.make_model_golden import loss_weights  # noqa: E402

GRAD_KEYS = ("attn_L.proj.weight", "attn_R.encoder.layers.0.self_attn.in_proj_weight",
             "attn_diff.encoder.layers.1.linear2.weight", "cc_proj.weight", "body.0.weight", "subheads.5.dist.4.weight")


def inputs(batch=3):
    wl, wr = orc.synth_binaural(batch, seed=4321)
    wl[1] *= 3.0                                                     # beyond +-1: clamped by the model
    x3 = np.random.RandomState(3).standard_normal((batch, 100)).astype(np.float32)
    return wl, wr, x3


def main():
    ref_model, _ = import_reference()
    torch.set_num_threads(8)
    wl, wr, x3 = inputs()
    torch.manual_seed(0)
    m = ref_model.build_model_auralnet_active().eval()
    ws, wa, wd = loss_weights(3)
    out = {}
    for tag, dtype in (("a32", torch.float32), ("a64", torch.float64)):
        mm = copy.deepcopy(m).to(dtype)
        tl, tr, t3 = (torch.from_numpy(a).to(dtype) for a in (wl, wr, x3))
        if dtype == torch.float32:
            sound, aoa, dist = mm(tl, tr, t3)
        else:                                                        # forward() casts to fp32: the same graph by hand
            tl, tr = torch.clamp(tl, -1.0, 1.0), torch.clamp(tr, -1.0, 1.0)
            xl = torch.clamp(torch.log(mm.fb_L(tl) + 1e-8), -12.0, 12.0)
            xr = torch.clamp(torch.log(mm.fb_R(tr) + 1e-8), -12.0, 12.0)
            feats = [mm.attn_L(xl).mean(1), mm.attn_R(xr).mean(1), mm.attn_diff(xl - xr).mean(1), mm.cc_proj(t3)]
            body = mm.body(torch.cat(feats, -1))
            outs = [h(body) for h in mm.subheads]
            sound = torch.cat([o[0] for o in outs], 1)
            aoa = torch.cat([o[1] for o in outs], 1)
            dist = torch.stack([o[2] for o in outs], 1)
        loss = (torch.from_numpy(ws).to(dtype) * sound).sum() + (torch.from_numpy(wa).to(dtype) * aoa).sum() \
            + (torch.from_numpy(wd).to(dtype) * dist).sum()
        loss.backward()
        out[f"{tag}.sound"], out[f"{tag}.aoa"], out[f"{tag}.dist"] = (t.detach().numpy() for t in (sound, aoa, dist))
        params = dict(mm.named_parameters())
        for k in GRAD_KEYS:
            out[f"{tag}.grad.{k}"] = sub(params[k].grad.numpy())
    path = os.path.join(HERE, "auralnet_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e3:.0f} kB", len(out), "arrays")
    for k in ("sound", "aoa", "dist"):
        a, b = out[f"a32.{k}"], out[f"a64.{k}"]
        print(k, "fp32 vs fp64:", float(np.max(np.abs(a - b)) / np.max(np.abs(b))))


if __name__ == "__main__":
    main()
