"""Measurement arms that ride along with bench.py's main line (all measurement code, never a product path):

  reference_cpu(...)        the reference ITSELF (oracle/_ref staging of model_torch.py / utils.py) on the host cores:
                            bench.py --impl reference and the cpu_baseline object (kind "reference"; falls back to the
                            oracle port, kind "port", only when oracle/_ref is not staged)
  gpu_eager_reference(...)  the same reference code on the same B200 in PyTorch eager (the real same-box competitor)
  fixed_q(...)              BASELINE config 3: fixed-Q front-end + phase + CC, forward only, with its own HBM roofline
  full_step(...)            BASELINE config 4: full active training step (front-end + back-end + losses + clips + Adam)
  variants(...)             BASELINE config 5: single controller, band-count / lag sweeps, 10 s clips, AuralNet filterbank
  compute_roofline(...)     what actually bounds the adaptive path: issue slots / FMA pipe / MUFU / SURVEY 8(d) ceilings
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS, T, NBANDS, NBINS = 16000, 19, 100, 513
CONFIG_YAML = dict(deltaQ_base=1.0, deltaQ_low_factor=0.3, deltaQ_high_factor=5.0, deltaQ_mode="relative")
REG_Q_W = REG_SMOOTH_W = 1e-3
UNIT = "audio-s/s"
A_FIXED = 128000 + 15200 + 15200 + 400     # SURVEY 8(d) A_fixed = 158 800 B per clip: both ears' waveforms in; Y, phase (both ears each), CC out


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# the reference itself
# ------------------------------------------------------------------------------------------------
class ReferenceStep:
    """One front-end training step of the workload through the reference's own code: BinauralAdaptiveGammatoneFB forward
    (model_torch.py:492-573), DeepEarActiveWaveform._subband_phase_from_X for both ears (:1039-1063), the log-energy
    features (:1080-1083), the CC feature (utils.py:390-420; on the CPU in a process pool over the cores, the way
    create_h5_data/data_save.py:213-221 runs it), the bench's loss and loss.backward().  Train-mode dropout."""

    def __init__(self, batch, device="cpu", with_cc=True, cc_workers=None):
        from oracle import stage_ref
        self.ref_model, self.ref_utils = stage_ref.load_reference()
        import bench
        self.device = torch.device(device)
        torch.manual_seed(0)
        m = self.ref_model.BinauralAdaptiveGammatoneFB(alpha=0.0, fixed_frontend_q=False, **CONFIG_YAML)
        with torch.no_grad():
            for fb in (m.fb_L, m.fb_R):
                torch.nn.init.normal_(fb.q_out[-1].weight, std=0.02)
        self.m = m.to(self.device).train()
        wl, wr = bench.synth_binaural(batch, seed=1234)
        self.wl_np, self.wr_np = wl, wr
        self.wl, self.wr = torch.from_numpy(wl).to(self.device), torch.from_numpy(wr).to(self.device)
        rs = np.random.RandomState(3)
        self.up = {k: torch.from_numpy(rs.standard_normal((batch, T, NBANDS)).astype(np.float32)).to(self.device)
                   for k in ("gYL", "gYR", "gPL", "gPR")}
        self.gc = torch.from_numpy(rs.standard_normal((batch, NBANDS)).astype(np.float32)).to(self.device)
        self.log_q0 = torch.log(self.m.Q0 + 1e-8).view(1, 1, -1)
        self.with_cc = with_cc
        self.pool = None
        if with_cc:
            n = cc_workers or max(1, len(os.sched_getaffinity(0)))
            self.pool = stage_ref.CcPool(min(n, batch))
        self.cc_ms = 0.0
        self.phase_fn = self.ref_model.DeepEarActiveWaveform._subband_phase_from_X

    def __call__(self):
        m = self.m
        for p in m.parameters():
            p.grad = None
        fut = None
        if self.with_cc:      # the CC of this batch runs in the worker pool next to the filterbank
            fut = self.pool.submit(self.wl_np, self.wr_np, FS, NBANDS, 3.0)
        yl, yr, ql, qr, xl, xr = m(self.wl, self.wr)
        phl = self.phase_fn(None, xl, ql, m.f_fft, m.fc)
        phr = self.phase_fn(None, xr, qr, m.f_fft, m.fc)
        lx = lambda y: torch.clamp(torch.log(y + 1e-8), -12.0, 12.0)
        lq = torch.log(0.5 * (ql + qr) + 1e-8)
        loss = (self.up["gYL"] * lx(yl)).mean() + (self.up["gYR"] * lx(yr)).mean() \
            + (self.up["gPL"] * phl).mean() + (self.up["gPR"] * phr).mean() \
            + REG_Q_W * ((lq - self.log_q0) ** 2).mean() + REG_SMOOTH_W * ((lq[..., 1:] - lq[..., :-1]) ** 2).mean()
        if fut is not None:
            cc = torch.from_numpy(fut()).to(self.device)
            loss = loss + (self.gc * cc).mean()
        loss.backward()
        return float(loss.detach())

    def close(self):
        if self.pool is not None:
            self.pool.close()


def reference_cpu(sample_batch=16, steps=None, warmup=3, budget_s=25.0, min_steps=10):
    """Timed on the host cores this process may use.  Returns the cpu_baseline object."""
    from oracle import stage_ref
    try:     # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1)
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        pass
    if not stage_ref.available():
        import bench
        d = bench.cpu_baseline_port(sample_batch=sample_batch, steps=steps, warmup=max(1, warmup), budget_s=budget_s)
        d["note"] = "oracle/_ref not staged on this machine: the oracle PORT was timed instead of the reference itself"
        return d
    step = ReferenceStep(sample_batch, "cpu", with_cc=True)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and n >= min_steps and (el >= budget_s or n >= 60)):
            break
    step.close()
    return {"value": sample_batch * n / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"{n} steps (after {warmup} warm-up) of batch {sample_batch} of the same workload (same clip / weight "
                      f"recipe, train-mode dropout) through the UNMODIFIED reference code staged in oracle/_ref "
                      f"(model_torch.BinauralAdaptiveGammatoneFB + _subband_phase_from_X, fwd+bwd, torch CPU fp32; "
                      f"utils.compute_cross_correlation_feature in a process pool over the cores), {el:.1f} s, "
                      f"os.cpu_count()={os.cpu_count()}",
            "ms_per_step": el / n * 1e3, "steps": n}


def reference_cpu_full_step(sample_batch=16, warmup=2, budget_s=10.0, min_steps=5):
    """one FULL training step of the unmodified reference (oracle/_ref:
    model_torch.build_model_active with the conf/config.yaml settings, train mode) on the host cores -- CC in the worker
    pool, model forward, the losses and Q regularisers of train_biear.py:417-431, 476-490, backward, the two global-norm
    clips (:523-525) and Adam with the two parameter groups (:617-621).  The CPU counterpart of `full_step`."""
    from oracle import stage_ref
    if not stage_ref.available():
        return {"unavailable": "oracle/_ref not staged on this machine"}
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        pass
    import bench
    import torch.nn.functional as F
    ref_model, _ = stage_ref.load_reference()
    torch.manual_seed(0)
    model = ref_model.build_model_active(use_cc=True, fb_alpha=0.0, **CONFIG_YAML)
    with torch.no_grad():
        for fb in (model.bifb.fb_L, model.bifb.fb_R):
            torch.nn.init.normal_(fb.q_out[-1].weight, std=0.02)
    model.train()
    fb_params = list(model.bifb.parameters())
    be_params = [p for n, p in model.named_parameters() if not n.startswith("bifb.")]
    opt = torch.optim.Adam([{"params": fb_params, "lr": 5e-5}, {"params": be_params, "lr": 1e-4}], weight_decay=1e-5, eps=1e-7)
    B = sample_batch
    wl_np, wr_np = bench.synth_binaural(B, 1234)
    wl, wr = torch.from_numpy(wl_np), torch.from_numpy(wr_np)
    rs = np.random.RandomState(5)
    pres = torch.from_numpy((rs.uniform(size=(B, 8)) < 0.25).astype(np.float32))
    ang = torch.from_numpy(rs.uniform(size=(B, 8)).astype(np.float32))
    dist_cls = torch.from_numpy(rs.randint(0, 5, size=(B, 8)))
    log_q0 = torch.log(model.bifb.Q0 + 1e-8).view(1, 1, -1)
    pos_w = torch.tensor(3.0)
    pool = stage_ref.CcPool(min(max(1, len(os.sched_getaffinity(0))), B))

    def step():
        fut = pool.submit(wl_np, wr_np, FS, NBANDS, 3.0)
        x3 = torch.from_numpy(fut())
        opt.zero_grad(set_to_none=True)
        sound, aoa, dl = model(wl, wr, x3)
        l_sound = F.binary_cross_entropy_with_logits(sound, pres, pos_weight=pos_w)
        l_aoa = (F.smooth_l1_loss(aoa, ang, beta=0.02, reduction="none") * pres).sum() / pres.sum().clamp_min(1.0)
        l_dist = (F.cross_entropy(dl.reshape(-1, 5), dist_cls.reshape(-1), reduction="none") * pres.reshape(-1)).sum() \
            / pres.sum().clamp_min(1.0)
        lq = torch.log(model.last_Q + 1e-8)
        reg = REG_Q_W * ((lq - log_q0) ** 2).mean() + REG_SMOOTH_W * ((lq[..., 1:] - lq[..., :-1]) ** 2).mean()
        loss = 0.2 * l_sound + 0.45 * l_aoa + 0.35 * l_dist + reg
        loss.backward()
        torch.nn.utils.clip_grad_norm_(fb_params, 0.2)
        torch.nn.utils.clip_grad_norm_(be_params, 3.0)
        opt.step()
        return float(loss.detach())

    try:
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        n = 0
        while True:
            loss = step()
            n += 1
            el = time.perf_counter() - t0
            if n >= min_steps and (el >= budget_s or n >= 30):
                break
    finally:
        pool.close()
    return {"value": B * n / el, "unit": UNIT, "ms_per_step": el / n * 1e3, "steps": n, "batch": B, "loss": loss,
            "cores": torch.get_num_threads(), "kind": "reference",
            "what": f"{n} full training steps (after {warmup} warm-up) of batch {B} through the UNMODIFIED reference staged in "
                    "oracle/_ref (build_model_active, conf/config.yaml settings, train mode; CC in a process pool; losses, Q "
                    "regularisers, two global-norm clips, Adam with two groups) on the host cores, torch CPU fp32"}


def gpu_eager_reference(batch=256, steps=5, warmup=2, device="cuda:0"):
    """The reference's own front-end code in PyTorch eager on this GPU: same batch, same loss, fwd+bwd, train mode.
    CC is left out (the reference computes it offline with numpy on the CPU; there is no GPU path for it there)."""
    from oracle import stage_ref
    if not stage_ref.available():
        return {"unavailable": "oracle/_ref not staged"}
    torch.cuda.reset_peak_memory_stats()
    step = ReferenceStep(batch, device, with_cc=False)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    mem = torch.cuda.max_memory_allocated() / 1e9
    del step
    torch.cuda.empty_cache()
    return {"value": batch / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "batch": batch,
            "peak_memory_gb": round(mem, 2),
            "what": "UNMODIFIED reference front-end (oracle/_ref model_torch.py: bifb + _subband_phase_from_X) fwd+bwd, "
                    "PyTorch eager CUDA on this GPU, train mode, same loss; CC excluded (CPU-only in the reference)"}


# ------------------------------------------------------------------------------------------------
# BASELINE config 3: fixed-Q front-end + phase + CC, forward only
# ------------------------------------------------------------------------------------------------
def fixed_q(batch=4096, device="cuda:0", reps=None):
    import biear_b200
    dev = torch.device(device)
    fb = biear_b200.BinauralAdaptiveGammatoneFB(fixed_frontend_q=True).to(dev).eval()
    fb.graph_replay = False     # captured explicitly below
    g = torch.Generator(device="cpu").manual_seed(batch)
    n_in = 4
    ins = [(torch.rand((batch, FS), generator=g) * 2 - 1).to(dev) for _ in range(n_in)]

    def run(i):
        o = fb.forward_features(ins[i % n_in], ins[(i + 1) % n_in], want_phase=True, want_cc=True, want_logenergy=True)
        return o["logYL"], o["logYR"], o["phaseL"], o["phaseR"], o["cc"]

    with torch.no_grad():
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            run(0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        for i in range(n_in):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep = run(i)
            graphs.append((gr, keep))
        for gr, _ in graphs:
            gr.replay()
        torch.cuda.synchronize()
        reps = reps or max(8, 16384 // batch)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            graphs[r % n_in][0].replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    cps = batch / (us * 1e-6)
    peak, _, src = _peaks()
    del graphs, ins
    torch.cuda.empty_cache()
    return {"value": cps, "unit": "clips/s", "batch": batch, "us_per_batch": us,
            "workload": f"fixed-Q front-end (both ears) + sub-band phase + log-energy + CC, forward only, batch {batch}, "
                        f"inputs rotate over {n_in} resident batches ({n_in * batch * FS * 4 / 1e6:.0f} MB > L2)",
            "roofline": {"bound": "hbm", "achieved": cps * A_FIXED / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": cps * A_FIXED / 1e9 / peak, "bytes_per_clip": A_FIXED, "peak_source": src}}


# ------------------------------------------------------------------------------------------------
# BASELINE config 4: the full active training step
# ------------------------------------------------------------------------------------------------
def build_full_step(batch, dev, world=1, rank=0):
    """(full_step callable, fwd+bwd GraphedStep, model, flat-grad numel): forward + losses + backward as one CUDA graph,
    [all-reduce,] the two global-norm clips + Adam (train_biear.py:523-525, 617-621) as a second one."""
    import torch.nn.functional as F
    import bench
    from biear_b200 import GraphedStep, model_torch as mt, ops
    from biear_b200.dist import FlatGradAllReducer
    torch.manual_seed(0)
    model = mt.build_model_active(use_cc=True, fb_alpha=0.0, **CONFIG_YAML)
    with torch.no_grad():
        for fb in (model.bifb.fb_L, model.bifb.fb_R):
            torch.nn.init.normal_(fb.q_out[-1].weight, std=0.02)
    model = model.to(dev).train()
    model.bifb.graph_replay = False     # the whole step is captured as one graph below
    fb_params = list(model.bifb.parameters())
    be_params = [p for n, p in model.named_parameters() if not n.startswith("bifb.")]
    params = fb_params + be_params
    opt = torch.optim.Adam([{"params": fb_params, "lr": 5e-5}, {"params": be_params, "lr": 1e-4}], weight_decay=1e-5,
                           eps=1e-7, capturable=True, fused=True)
    wl, wr = bench.synth_binaural(batch, 1234 + rank)
    wl, wr = torch.from_numpy(wl).to(dev), torch.from_numpy(wr).to(dev)
    rs = np.random.RandomState(5)
    y = torch.zeros(batch, 8, 7, device=dev)
    y[..., 0] = torch.from_numpy((rs.uniform(size=(batch, 8)) < 0.25).astype(np.float32)).to(dev)
    y[..., 1] = torch.from_numpy(rs.uniform(size=(batch, 8)).astype(np.float32)).to(dev)
    dist_cls = torch.from_numpy(rs.randint(0, 5, size=(batch, 8))).to(dev)
    pos_w = torch.tensor(3.0, device=dev)

    def loss_fn(a, b):
        x3 = ops.cc_feature(a, b)
        sound, aoa, dl = model(a, b, x3)
        pres = y[..., 0]
        l_sound = F.binary_cross_entropy_with_logits(sound, pres, pos_weight=pos_w)
        l_aoa = (F.smooth_l1_loss(aoa, y[..., 1], beta=0.02, reduction="none") * pres).sum() / pres.sum().clamp_min(1.0)
        l_dist = (F.cross_entropy(dl.reshape(-1, 5), dist_cls.reshape(-1), reduction="none") * pres.reshape(-1)).sum() \
            / pres.sum().clamp_min(1.0)
        reg = ops.q_regularizers(model.last_QL, model.last_QR, model.bifb.Q0, REG_Q_W, REG_SMOOTH_W)[0]   # train_biear.py:476-490
        return 0.2 * l_sound + 0.45 * l_aoa + 0.35 * l_dist + reg

    step = GraphedStep(loss_fn, (wl, wr), params, warmup=3, flat_grads=world > 1)
    red = FlatGradAllReducer(params, flat=step.flat) if world > 1 else None

    def update():
        torch.nn.utils.clip_grad_norm_(fb_params, 0.2, foreach=True)
        torch.nn.utils.clip_grad_norm_(be_params, 3.0, foreach=True)
        opt.step()

    state = {"graph": None}

    def full():
        loss = step(wl, wr)
        if red is not None:
            red()
        if state["graph"] is None:
            update()
        else:
            state["graph"].replay()
        return loss

    for _ in range(3):
        full()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        update()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, pool=step.pool()):
        update()
    state["graph"] = g
    return full, step, model, sum(p.numel() for p in params)


def full_step(batch=256, steps=30, device="cuda:0", world=1, rank=0, dist=None):
    """world > 1: every rank calls this (batch clips each); the gradients of the whole model (one flat bucket of 1 634 780
    floats, written by the step's graph) are all-reduced over NCCL between the backward and the clips; time = max over ranks."""
    dev = torch.device(device)
    full, step, model, numel = build_full_step(batch, dev, world=world, rank=rank)
    for _ in range(3):
        full()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = full()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    out = {"value": batch * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "batch_per_gpu": batch,
           "n_gpus": world, "loss": float(loss), "launches_of_ours_per_step": step.launches_per_replay, "parameters": numel,
           "allreduce_floats": numel if world > 1 else 0,
           "workload": f"BASELINE config 4 on {world} GPU(s): full active training step (front-end + ILD/IPD encoders with "
                       "native GRU recurrences + body + 8 native sector heads, the reference's losses + Q regularisers, "
                       + ("NCCL all-reduce of the whole model's gradient bucket, " if world > 1 else "")
                       + "two global-norm clips, Adam with two groups), forward + backward as one CUDA graph, clips + Adam "
                       "as a second; resident inputs"}
    del full, step, model
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# BASELINE config 5: the other front-end variants (single controller, band-count / lag sweeps, 10 s clips, AuralNet FB)
# ------------------------------------------------------------------------------------------------
def _time_calls(fn, warmup=4, reps=20):
    for _ in range(warmup):          # (the drop-in's transparent CUDA-graph replay records itself on the third call)
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def variants(batch=256, device="cuda:0"):
    """Throughput of the variants BASELINE.json config 5 names, each fwd+bwd in train mode at `batch` clips through the
    module's own API (forward_features + autograd, transparent graph replay on), random-init controllers with a non-zero
    last layer.  Values in audio-s/s (1 clip = 1 s x 2 ears)."""
    import biear_b200
    from biear_b200 import ops
    dev = torch.device(device)
    g = torch.Generator(device="cpu").manual_seed(99)
    out = {"batch": batch, "unit": UNIT,
           "workload": "fwd+bwd (Y, phase, log-energy features; loss = fixed random functionals), train mode, resident inputs, "
                       "module API with transparent CUDA-graph replay"}

    def frontend_case(mod, n_bands, b):
        mod = mod.to(dev).train()
        with torch.no_grad():
            for name, prm in mod.named_parameters():
                if name.endswith("q_out.8.weight"):
                    torch.nn.init.normal_(prm, std=0.02)
        wl = (torch.rand((b, FS), generator=g) * 2 - 1).to(dev)
        wr = (torch.rand((b, FS), generator=g) * 2 - 1).to(dev)
        ups = [torch.randn((b, T, n_bands), generator=g).to(dev) for _ in range(4)]
        has_grad = any(p.requires_grad for p in mod.parameters())

        def call():
            with torch.set_grad_enabled(has_grad):
                o = mod.forward_features(wl, wr, want_phase=True, want_logenergy=True)
                if has_grad:
                    for p_ in mod.parameters():
                        p_.grad = None
                    ((ups[0] * o["logYL"]).sum() + (ups[1] * o["logYR"]).sum() + (ups[2] * o["phaseL"]).sum()
                     + (ups[3] * o["phaseR"]).sum()).backward()
        ms = _time_calls(call)
        del mod
        return {"ms_per_step": ms, "value": b / (ms * 1e-3)}

    single_kw = dict(deltaQ_base=2.0, deltaQ_low_factor=0.5, deltaQ_high_factor=5.0, deltaQ_mode="absolute")   # config_single_ctrl.yaml
    cases = [("single_controller", lambda: frontend_case(biear_b200.BinauralAdaptiveGammatoneFB_SingleController(**single_kw), NBANDS, batch))]
    for nb in (32, 64, 128):
        cases.append((f"dual_bands_{nb}", lambda nb=nb: frontend_case(
            biear_b200.BinauralAdaptiveGammatoneFB(Nbands=nb, alpha=0.0, **CONFIG_YAML), nb, batch)))
        cases.append((f"single_controller_bands_{nb}", lambda nb=nb: frontend_case(
            biear_b200.BinauralAdaptiveGammatoneFB_SingleController(Nbands=nb, **single_kw), nb, batch)))
    # 10 s clips = 10 x 1 s segments folded into the batch (the reference itself truncates to the first second)
    cases.append(("dual_10s_clips", lambda: frontend_case(biear_b200.BinauralAdaptiveGammatoneFB(alpha=0.0, **CONFIG_YAML), NBANDS, 10 * batch)))
    for key, fn in cases:
        try:
            out[key] = fn()
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": repr(e)[:200]}
        torch.cuda.empty_cache()
    # AuralNet filterbank (fixed Q0, GEMM form; model_torch.py:70-195), forward only, both ears
    try:
        fb = biear_b200.AuralNetGammatoneFB().to(dev).eval()
        wav = (torch.rand((2 * batch, FS), generator=g) * 2 - 1).to(dev)
        with torch.no_grad():
            ms = _time_calls(lambda: fb(wav))
        out["auralnet_fb_forward"] = {"ms_per_step": ms, "value": batch / (ms * 1e-3)}
    except Exception as e:  # noqa: BLE001
        out["auralnet_fb_forward"] = {"error": repr(e)[:200]}
    # CC lag-range sweep (utils.py:390-420), num_lags == Nbands == 100
    wl = (torch.rand((batch, FS), generator=g) * 2 - 1).to(dev)
    wr = (torch.rand((batch, FS), generator=g) * 2 - 1).to(dev)
    for ms_lag in (1.0, 3.0, 5.0):
        try:
            ms = _time_calls(lambda: ops.cc_feature(wl, wr, max_lag_ms=ms_lag))
            out[f"cc_max_lag_{ms_lag:g}ms"] = {"ms_per_step": ms, "value": batch / (ms * 1e-3)}
        except Exception as e:  # noqa: BLE001
            out[f"cc_max_lag_{ms_lag:g}ms"] = {"error": repr(e)[:200]}
    return out


# ------------------------------------------------------------------------------------------------
# the compute-side roofline of the adaptive path
# ------------------------------------------------------------------------------------------------
def gaussian_window_counts(q, fc, df=15.625, cutoff=6.0, n_bins=NBINS):
    """Bins inside the |u| <= cutoff window of every (row, frame, band), as csrc/band_dev.cuh band_params() bounds it."""
    bw = fc.view(1, 1, -1) / (q + 1e-8) + 1e-8
    half = cutoff * bw
    k_lo = torch.clamp(torch.floor((fc.view(1, 1, -1) - half) / df), min=0)
    k_hi = torch.clamp(torch.ceil((fc.view(1, 1, -1) + half) / df), max=n_bins - 1)
    return (k_hi - k_lo + 1).clamp_min(0)


def compute_roofline(model, dev_in, value_per_gpu, fwd_us, ncu_summary_path):
    """What bounds the adaptive path ("report the HBM fraction AND the fp32/MUFU fraction"):
      * issue-slot and FMA-pipe utilisation of the two recurrence kernels, from the committed ncu capture;
      * Gaussian evaluations per second of the forward recurrence (live: window sizes from this run's Q, the kernel's
        CUDA-event time) against the MUFU limit 148 SM x 16 ex2/clk x SM clock;
      * the step's throughput against SURVEY.md 8(d)'s compute ceilings (1.6 M dense / 5 M truncated audio-s/s)."""
    _, sm_mhz, src = _peaks()
    wl, wr = dev_in[0]
    B = wl.shape[0]
    with torch.no_grad():
        was = model.training
        model.eval()
        o = model.forward_features(wl, wr, want_phase=False)
        model.train(was)
        fb = model.fb_L
        useful = float(gaussian_window_counts(o["QL"], fb.fc).sum() + gaussian_window_counts(o["QR"], fb.fc).sum())
    dense = 2.0 * B * T * NBANDS * NBINS
    mufu_peak = 148 * 16 * sm_mhz * 1e6
    out = {
        "bound": "fp32 issue / FMA pipe + 19-step dependency chain (not HBM, not tensor cores: W(Q) is private per (clip, frame))",
        "gaussian_evals_per_launch": {"useful_window": useful, "dense": dense, "window_fraction": useful / dense},
        "gaussian_evals_per_s": useful / (fwd_us * 1e-6),
        "mufu_peak_per_s": mufu_peak,
        "mufu_frac": useful / (fwd_us * 1e-6) / mufu_peak,
        "mufu_peak_source": f"148 SM x 16 ex2/clk x {sm_mhz:.0f} MHz ({src})",
        "survey_ceiling_audio_s_per_s": {"dense": 1.6e6, "truncated": 5.0e6},
        "value_frac_of_ceiling": {"dense": value_per_gpu / 1.6e6, "truncated": value_per_gpu / 5.0e6},
    }
    try:
        with open(ncu_summary_path) as f:
            k = json.load(f)["kernels"]
        for short, prefix in (("seq_fwd", "seq_fwd2_kernel"), ("seq_bwd", "seq_bwd_kernel")):
            name = next(n for n in k if n.startswith(prefix) and k[n]["duration_us"] > 50)
            d = k[name]
            rec = {"issue_slot_pct": d.get("issue_active_pct"), "warp_instructions": d.get("warp_instructions"),
                   "duration_us_under_ncu": d.get("duration_us")}
            for key in ("pipe_fma_cycles_pct", "pipe_fma_inst_pct", "pipe_alu_pct", "pipe_xu_pct", "pipe_lsu_pct",
                        "smem_wavefronts_pct"):
                if d.get(key) is not None:
                    rec[key] = d[key]
            if d.get("pipe_fma_cycles_pct") is not None:     # the fp32 roofline fraction of the kernel: share of the
                rec["fma_pipe_frac"] = d["pipe_fma_cycles_pct"] / 100.0      # cycles in which the FMA pipe is busy
            out[short] = rec
        out["ncu_source"] = os.path.relpath(ncu_summary_path, ROOT)
    except Exception as e:  # noqa: BLE001
        out["ncu_source"] = f"unavailable ({e!r})"
    return out
