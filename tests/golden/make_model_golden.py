"""Golden vectors for the FULL active model (front-end + back-end) from the UNMODIFIED reference:
    python tests/golden/make_model_golden.py          # writes tests/golden/model_golden.npz

reference.build_model_active(conf/config.yaml settings) under torch.manual_seed(0), the two Q controllers replaced by
oracle.synth_controller(11/12) so that Q actually moves, eval mode, 3 synthetic clips, x3 from the reference's own CC.
Stores logits / predictions and a subsample of the gradients of a fixed scalar loss, in fp32 and in fp64 (the fp64 run
measures how far the reference's fp32 is from exact: the phase path is ill conditioned)."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import biear_oracle as orc  # noqa: E402
from tests.golden.make_golden import CONFIG_YAML, import_reference, load_ctrl, sub  # noqa: E402

GRAD_KEYS = ("bifb.fb_L.q_out.8.weight", "bifb.fb_R.q_rnn.weight_ih_l0", "encoder_ild.gru1.weight_ih_l0",
             "encoder_ipd.gru2.weight_hh_l0", "cc_proj.weight", "body.0.weight", "subheads.3.aoa.0.weight")


def loss_weights(batch):
    rs = np.random.RandomState(17)
    return (rs.standard_normal((batch, 8)).astype(np.float32), rs.standard_normal((batch, 8)).astype(np.float32),
            rs.standard_normal((batch, 8, 5)).astype(np.float32))


def main():
    ref_model, ref_utils = import_reference()
    torch.set_num_threads(8)
    batch = 3
    wl, wr = orc.synth_binaural(batch, seed=1234)
    x3 = np.stack([ref_utils.compute_cross_correlation_feature(a, b, 16000, 100, 3.0) for a, b in zip(wl, wr)])
    torch.manual_seed(0)
    m = ref_model.build_model_active(use_cc=True, fb_alpha=0.0, fixed_frontend_q=False, **CONFIG_YAML)
    load_ctrl(m.bifb.fb_L, orc.synth_controller(11))
    load_ctrl(m.bifb.fb_R, orc.synth_controller(12))
    m.eval()
    out = {"x3": x3.astype(np.float32)}
    ws, wa, wd = loss_weights(batch)
    for tag, dtype in (("model32", torch.float32), ("model64", torch.float64)):
        mm = copy.deepcopy(m).to(dtype)
        if dtype == torch.float64:     # forward() casts its inputs to fp32: run the same graph in fp64 by hand
            mm.forward = None
        tl, tr = torch.from_numpy(wl).to(dtype), torch.from_numpy(wr).to(dtype)
        t3 = torch.from_numpy(x3).to(dtype)
        if dtype == torch.float32:
            sound, aoa, dist = mm(tl, tr, t3)
        else:
            YL, YR, QL, QR, XL, XR = mm.bifb(tl, tr)
            x1 = torch.clamp(torch.log(YL + 1e-8), -12.0, 12.0)
            x2 = torch.clamp(torch.log(YR + 1e-8), -12.0, 12.0)
            pl = mm._subband_phase_from_X(XL, QL, mm.bifb.f_fft, mm.bifb.fc)
            pr = mm._subband_phase_from_X(XR, QR, mm.bifb.f_fft, mm.bifb.fc)
            feats = torch.cat([mm.encoder_ild(x1, x2), mm.encoder_ipd(pl, pr), mm.cc_proj(t3)], dim=-1)
            body = mm.body(feats)
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
    path = os.path.join(HERE, "model_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e3:.0f} kB", len(out), "arrays")
    for k in ("sound", "aoa", "dist"):
        a, b = out[f"model32.{k}"], out[f"model64.{k}"]
        print(k, "fp32 vs fp64:", float(np.max(np.abs(a - b)) / np.max(np.abs(b))))


if __name__ == "__main__":
    main()
