"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py          # writes tests/golden/*.npz

Imports /root/reference/model_torch.py and /root/reference/utils.py (the latter with
empty stand-ins for its absent, unused top-level imports `librosa` / `gammatone`),
feeds them the deterministic inputs/weights from oracle.biear_oracle.synth_* and stores
only the OUTPUTS; tests regenerate the inputs from the same seeds.  The reference does
not exist on the GPU box, so nothing in tests/ or bench.py imports it -- only this script.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import biear_oracle as orc  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    for name in ("librosa", "gammatone", "gammatone.gtgram"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["gammatone"].gtgram = sys.modules["gammatone.gtgram"]
    sys.modules["gammatone.gtgram"].gtgram = None
    import model_torch as ref_model  # noqa
    import utils as ref_utils  # noqa
    sys.path.pop(0)
    return ref_model, ref_utils


def load_ctrl(module, weights):
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    missing = module.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing


SUB_STEP = 11
SUB_MIN = 2000


def sub(a, step=SUB_STEP):
    """Strided subsample of big arrays (flattened, every 11th element) to keep fixtures small."""
    a = np.asarray(a)
    return a.reshape(-1)[::step].copy() if a.size > SUB_MIN else a.copy()


CONFIG_YAML = dict(deltaQ_base=1.0, deltaQ_low_factor=0.3, deltaQ_high_factor=5.0, deltaQ_mode="relative")
CONFIG_SINGLE = dict(deltaQ_base=2.0, deltaQ_low_factor=0.5, deltaQ_high_factor=5.0, deltaQ_mode="absolute")


def upstream(batch, seed=3, t=19, n=100):
    rs = np.random.RandomState(seed)
    return {k: rs.standard_normal((batch, t, n)).astype(np.float32)
            for k in ("gYL", "gYR", "gPL", "gPR", "gQL", "gQR")}


def run_dual(ref_model, dtype, batch, wl, wr, w_l, w_r, kw, out, tag):
    torch.manual_seed(0)
    bifb = ref_model.BinauralAdaptiveGammatoneFB(alpha=0.0, fixed_frontend_q=False, **kw)
    load_ctrl(bifb.fb_L, w_l)
    load_ctrl(bifb.fb_R, w_r)
    bifb = bifb.to(dtype).eval()
    host = ref_model.DeepEarActiveWaveform(fixed_frontend_q=True)  # only for _subband_phase_from_X
    taps = {"L": [], "R": []}

    def mk(side):
        def hook(_m, _i, o):
            o.retain_grad()
            taps[side].append(o)
        return hook
    bifb.fb_L.q_out.register_forward_hook(mk("L"))
    bifb.fb_R.q_out.register_forward_hook(mk("R"))

    tl = torch.from_numpy(wl).to(dtype)
    tr = torch.from_numpy(wr).to(dtype)
    up = {k: torch.from_numpy(v).to(dtype) for k, v in upstream(batch).items()}

    def fwd():
        taps["L"].clear(); taps["R"].clear()
        yl, yr, ql, qr, xl, xr = bifb(tl, tr)
        pl = host._subband_phase_from_X(xl, ql, bifb.f_fft, bifb.fc)
        pr = host._subband_phase_from_X(xr, qr, bifb.f_fft, bifb.fc)
        return yl, yr, ql, qr, xl, xr, pl, pr

    def grads(loss):
        bifb.zero_grad()
        loss.backward()
        g = {}
        for side, fb in (("L", bifb.fb_L), ("R", bifb.fb_R)):
            for name, prm in fb.named_parameters():
                g[f"{side}.{name}"] = prm.grad.detach().cpu().numpy().copy()
            g[f"{side}.tap"] = torch.stack([o.grad if o.grad is not None else torch.zeros_like(o) for o in taps[side]], 1).cpu().numpy().copy()
        return g

    yl, yr, ql, qr, xl, xr, pl, pr = fwd()
    np_ = lambda x: x.detach().cpu().numpy()
    out[f"{tag}.YL"], out[f"{tag}.YR"] = np_(yl), np_(yr)
    out[f"{tag}.QL"], out[f"{tag}.QR"] = np_(ql), np_(qr)
    out[f"{tag}.PL"], out[f"{tag}.PR"] = np_(pl), np_(pr)
    if dtype == torch.float32:
        out[f"{tag}.XL"] = np_(xl[:, ::3])      # frames 0,3,...,18
        out[f"{tag}.XR0"] = np_(xr[:1, ::6])
    # loss A: through Y and the Q regulariser path only (well conditioned)
    loss_a = (up["gYL"] * torch.log(yl + 1e-8)).sum() + (up["gYR"] * torch.log(yr + 1e-8)).sum() \
        + (up["gQL"] * ql).sum() + (up["gQR"] * qr).sum()
    for k, v in grads(loss_a).items():
        out[f"{tag}.gradA.{k}"] = sub(v)
    # loss B: through phase only (ill conditioned in fp32; compare against the fp64 record)
    yl, yr, ql, qr, xl, xr, pl, pr = fwd()
    loss_b = (up["gPL"] * pl).sum() + (up["gPR"] * pr).sum()
    for k, v in grads(loss_b).items():
        out[f"{tag}.gradB.{k}"] = sub(v)


def main():
    ref_model, ref_utils = import_reference()
    torch.set_num_threads(8)
    out = {}

    # ---- constants --------------------------------------------------------------------
    fb = ref_model.BinauralAdaptiveGammatoneFB(**CONFIG_YAML)
    out["const.fc"] = fb.fc.numpy()
    out["const.Q0"] = fb.Q0.numpy()
    out["const.f_fft"] = fb.f_fft.numpy()
    out["const.deltaQ_yaml"] = fb.fb_L.deltaQ_vec.numpy()
    out["const.win_fn"] = fb.fb_L.win_fn.numpy()
    fb2 = ref_model.BinauralAdaptiveGammatoneFB(Nbands=64, **CONFIG_SINGLE)
    out["const.fc64"] = fb2.fc.numpy()
    out["const.Q064"] = fb2.Q0.numpy()
    out["const.deltaQ_single64"] = fb2.fb_L.deltaQ_vec.numpy()

    # ---- dual adaptive, conf/config.yaml settings ---------------------------------------
    batch = 3
    wl, wr = orc.synth_binaural(batch, seed=1234)
    w_l = orc.synth_controller(11)
    w_r = orc.synth_controller(12)
    run_dual(ref_model, torch.float32, batch, wl, wr, w_l, w_r, CONFIG_YAML, out, "dual32")
    run_dual(ref_model, torch.float64, batch, wl, wr, w_l, w_r, CONFIG_YAML, out, "dual64")
    # strong controller: many Q entries driven onto the clamp bounds (dense near-uniform rows)
    w_l2 = orc.synth_controller(21, out_std=0.3)
    w_r2 = orc.synth_controller(22, out_std=0.3)
    run_dual(ref_model, torch.float32, 2, wl[:2], wr[:2], w_l2, w_r2, CONFIG_YAML, out, "clamp32")
    run_dual(ref_model, torch.float64, 2, wl[:2], wr[:2], w_l2, w_r2, CONFIG_YAML, out, "clamp64")
    # absolute mode
    run_dual(ref_model, torch.float32, 2, wl[:2], wr[:2], w_l, w_r, CONFIG_SINGLE, out, "abs32")

    # ---- ragged input lengths: short (padded) and long (truncated) ---------------------------
    fixed = ref_model.BinauralAdaptiveGammatoneFB(fixed_frontend_q=True).eval()
    with torch.no_grad():
        tl, tr = torch.from_numpy(wl), torch.from_numpy(wr)
        yl, yr, ql, qr, xl, xr = fixed(tl, tr)
        out["fixed.YL"], out["fixed.YR"] = yl.numpy(), yr.numpy()
        out["fixed.QL"] = ql.numpy()
        ys, _, _, _, _, _ = fixed(tl[:, :9000], tr[:, :9000])
        out["fixed.YL_short9000"] = ys.numpy()
        yl2, _, _, _, _, _ = fixed(torch.cat([tl, tl], 1), torch.cat([tr, tr], 1))
        out["fixed.YL_long32000"] = yl2.numpy()
        host = ref_model.DeepEarActiveWaveform(fixed_frontend_q=True)
        out["fixed.PL"] = host._subband_phase_from_X(xl, ql, fixed.f_fft, fixed.fc).numpy()
        aur = ref_model.AuralNetGammatoneFB().eval()
        out["auralnet.YL"] = aur(tl).numpy()
        # band-count sweep (BASELINE config 5)
        f64 = ref_model.BinauralAdaptiveGammatoneFB(Nbands=64, fixed_frontend_q=True).eval()
        out["fixed64.YL"] = f64(tl, tr)[0].numpy()

    # ---- single controller (config_single_ctrl.yaml) -----------------------------------------
    torch.manual_seed(0)
    sc = ref_model.BinauralAdaptiveGammatoneFB_SingleController(**CONFIG_SINGLE)
    w_s = orc.synth_controller(31, in_mult=4)
    load_ctrl(sc, w_s)
    sc.eval()
    tl2 = torch.from_numpy(wl[:2]); tr2 = torch.from_numpy(wr[:2])
    yl, yr, q, _, _, _ = sc(tl2, tr2)
    up = {k: torch.from_numpy(v) for k, v in upstream(2).items()}
    loss = (up["gYL"] * torch.log(yl + 1e-8)).sum() + (up["gYR"] * torch.log(yr + 1e-8)).sum() + (up["gQL"] * q).sum()
    loss.backward()
    out["single.YL"], out["single.YR"], out["single.Q"] = yl.detach().numpy(), yr.detach().numpy(), q.detach().numpy()
    for name, prm in sc.named_parameters():
        out[f"single.gradA.{name}"] = sub(prm.grad.numpy())
    # float64 twin of the same run: how far the reference's own fp32 result is from the truth (conditioning record)
    sc64 = ref_model.BinauralAdaptiveGammatoneFB_SingleController(**CONFIG_SINGLE)
    load_ctrl(sc64, w_s)
    sc64 = sc64.double().eval()
    yl, yr, q, _, _, _ = sc64(tl2.double(), tr2.double())
    loss = (up["gYL"].double() * torch.log(yl + 1e-8)).sum() + (up["gYR"].double() * torch.log(yr + 1e-8)).sum() \
        + (up["gQL"].double() * q).sum()
    loss.backward()
    out["single64.YL"], out["single64.Q"] = yl.detach().numpy(), q.detach().numpy()
    for name, prm in sc64.named_parameters():
        out[f"single64.gradA.{name}"] = sub(prm.grad.numpy())

    # ---- CC feature ---------------------------------------------------------------------------
    cw_l, cw_r = orc.synth_binaural(6, seed=77)
    out["cc.default"] = np.stack([ref_utils.compute_cross_correlation_feature(a, b, 16000) for a, b in zip(cw_l, cw_r)])
    out["cc.lags64_1ms"] = np.stack([ref_utils.compute_cross_correlation_feature(a, b, 16000, 64, 1.0) for a, b in zip(cw_l[:2], cw_r[:2])])
    out["cc.lags128_5ms"] = np.stack([ref_utils.compute_cross_correlation_feature(a, b, 16000, 128, 5.0) for a, b in zip(cw_l[:2], cw_r[:2])])
    q16 = lambda x: (np.round(x * 32767) / 32768).astype(np.float32)   # int16-quantised audio
    out["cc.int16"] = np.stack([ref_utils.compute_cross_correlation_feature(q16(a), q16(b), 16000) for a, b in zip(cw_l[:2], cw_r[:2])])
    z = np.zeros(16000, np.float32)
    out["cc.silence"] = ref_utils.compute_cross_correlation_feature(z, z, 16000)
    dc = np.full(16000, 0.25, np.float32)
    out["cc.dc_vs_noise"] = ref_utils.compute_cross_correlation_feature(dc + cw_l[0] * 0.1, cw_r[0], 16000)

    path = os.path.join(HERE, "frontend_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path)/1e6:.2f} MB", len(out), "arrays")
    print("torch", torch.__version__, "numpy", np.__version__)


if __name__ == "__main__":
    main()
