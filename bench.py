#!/usr/bin/env python
"""bench.py -- binaural audio-seconds/second of the BiEAR active-mode front-end, forward + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

Workload (BASELINE.json configs[1]): batch 256 synthetic binaural clips (1 s, 2 ears, 16 kHz, fp32) per GPU,
adaptive Q (dual controllers, conf/config.yaml settings), CC on, train mode.  One step = STFT of all frames,
the 19-frame Q recurrence (band energies + sub-band phase + controller), the CC feature, a loss over
log-energies / phase / Q regularisers and the backward pass into the controller weights (dQ closed form);
with N > 1 ranks the batch is sharded (weak scaling, 256 clips per rank) and the controller gradients are
all-reduced over NCCL every step.

Prints ONE JSON line (rank 0).  `value` is measured with the inputs resident in HBM, `e2e` through the
same public call with pinned HOST buffers (H2D of both waveforms and D2H of the loss inside the timed region).
`--impl reference` times the reference's own code (oracle/_ref staging of model_torch.py / utils.py; the oracle port only
if that staging is absent) on the host cores instead.  The line also carries `roofline` (HBM, as the contract asks, plus the
compute-side bound under `roofline.compute`), `cpu_baseline`, and -- at N = 1 -- the sub-records `gpu_eager_reference`
(the reference code in PyTorch eager on the same GPU), `fixed_q` (BASELINE config 3) and `full_step` (config 4).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS, T, NBANDS, NBINS = 16000, 19, 100, 513
CONFIG_YAML = dict(deltaQ_base=1.0, deltaQ_low_factor=0.3, deltaQ_high_factor=5.0, deltaQ_mode="relative")
REG_Q_W = REG_SMOOTH_W = 1e-3                      # conf/config.yaml
# algorithmic bytes per audio-second, module-boundary form (X materialised), fwd+bwd
A_FULL = 128000 + 15200 + 15200 + 155952 + 15200 + 15200
METRIC = "binaural audio-sec/sec front-end fwd+bwd"
UNIT = "audio-s/s"
N_ROTATE = 8                                       # resident input batches cycled through (> L2 in total)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synth_binaural(batch, seed, n=FS):
    """AR(1)-tilted noise, per-clip integer ITD in [-12,12] and ILD gain in [0.5,1], joint peak-normalised
    (SURVEY.md 8(d)); numpy RandomState so every rank/run sees the same clips for a given seed."""
    rs = np.random.RandomState(seed)
    src = rs.standard_normal((batch, n + 64)).astype(np.float32)
    s = np.empty_like(src)
    acc = np.zeros(batch, np.float32)
    for i in range(src.shape[1]):
        acc = 0.9 * acc + src[:, i]
        s[:, i] = acc
    itd = rs.randint(-12, 13, size=batch)
    ild = rs.uniform(0.5, 1.0, size=batch).astype(np.float32)
    idx = (32 - itd)[:, None] + np.arange(n)[None, :]
    wl = s[:, 32:32 + n]
    wr = ild[:, None] * np.take_along_axis(s, idx, axis=1) + 0.01 * rs.standard_normal((batch, n)).astype(np.float32)
    peak = np.maximum(np.abs(wl).max(1), np.abs(wr).max(1))[:, None]
    return np.ascontiguousarray(wl / peak, np.float32), np.ascontiguousarray(wr / peak, np.float32)


class ClockSampler:
    """In-process NVML poller for the timed region (the fields of B200_PROFILING.md's clocks line).  NVML is
    initialised BEFORE the timed region: spawning nvidia-smi next to a running launch loop was measured to stall
    kernel launches for ~1 s while it initialises (round-1 log: 51 ms/step vs 8 ms/step for the same code)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index, period_s=0.01):
        self.period = period_s
        self.rows = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.handle = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._sample()                      # first call pays NVML's lazy setup, outside the timed region
            self.rows.clear()
        except Exception as e:  # noqa: BLE001
            self.handle = None
            self.err = repr(e)

    def _sample(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        try:
            power = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:  # noqa: BLE001
            power = None
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:  # noqa: BLE001
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.rows.append((sm, power, mask))

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                self._sample()
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(self.period)

    def start(self):
        if self.handle is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable: {self.err}"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        sm = [r[0] for r in self.rows]
        power = [r[1] for r in self.rows if r[1] is not None]
        reasons = sorted({name for _, _, m in self.rows for name, bit in self.REASONS if m & bit})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": reasons}


class SmiSampler:
    """Fallback when NVML cannot be loaded in-process: an nvidia-smi poller, started (and given 2 s to finish
    its own initialisation) BEFORE the timed region so that it cannot stall the launch loop."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        import subprocess
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.rows.append([c.strip() for c in ln.split(",")])
                                             for ln in self.proc.stdout], daemon=True).start()
            time.sleep(2.0)
        except Exception:  # noqa: BLE001
            self.proc = None

    def start(self):
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0])), smax.append(float(r[1])), power.append(float(r[2]))
                reasons.update(n for n, v in zip(names, r[3:7]) if v.lower().startswith("active"))
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_sampler(gpu_index):
    s = ClockSampler(gpu_index)
    return s if s.handle is not None else SmiSampler(gpu_index)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_frontend(device):
    import biear_b200
    torch.manual_seed(0)
    m = biear_b200.BinauralAdaptiveGammatoneFB(alpha=0.0, fixed_frontend_q=False, **CONFIG_YAML)
    with torch.no_grad():
        for fb in (m.fb_L, m.fb_R):   # default zero-init pins Q == Q0; make the controller actually act
            torch.nn.init.normal_(fb.q_out[-1].weight, std=0.02)
    m.graph_replay = False      # the whole step is captured by GraphedStep below; no per-module graphs inside it
    return m.to(device).train()


def make_loss(model, up):
    """The public-API call a user makes: waveforms in, scalar loss out (backward leaves the gradients on the
    parameters)."""
    from biear_b200 import ops
    params = [p for p in model.parameters() if p.requires_grad]
    # a random linear functional of every feature stands in for the back-end (it makes all upstream gradients dense and
    # non-trivial): mean(up * x) per feature, written as ONE dot product with the 1/numel folded into the fixed weights
    wts = {k: (v / v.numel()).reshape(-1) for k, v in up.items()}

    branches = {}

    def loss_fn(wl, wr):
        # one call for everything the back-end consumes: log band energies (model_torch.py:1080-1083, fused into the
        # band stage), sub-band phases, Q, and the CC feature (forked stream)
        o = model.forward_features(wl, wr, want_phase=True, want_cc=True, want_logenergy=True)
        # The five feature terms are independent of each other (like the per-ear encoders of the real back-end): each
        # runs on its own forked stream, so the few small kernels of a term -- and, in the backward, of its gradient --
        # overlap with the other terms' instead of queueing behind them.
        cur = torch.cuda.current_stream(wl.device)
        terms = []
        for key, x in (("gYL", o["logYL"]), ("gYR", o["logYR"]), ("gPL", o["phaseL"]), ("gPR", o["phaseR"]), ("gC", o["cc"])):
            st = branches.setdefault(key, torch.cuda.Stream(device=wl.device, priority=-1))
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                terms.append(torch.dot(wts[key], x.reshape(-1)))
        # ... plus the reference's Q regularisers on (QL + QR) / 2 (train_biear.py:476-490): value and gradient from one
        # kernel (biear_q_regularizers)
        st = branches.setdefault("reg", torch.cuda.Stream(device=wl.device, priority=-1))
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            terms.append(ops.q_regularizers(o["QL"], o["QR"], model.Q0, REG_Q_W, REG_SMOOTH_W)[0])
        for key, t in zip(("gYL", "gYR", "gPL", "gPR", "gC", "reg"), terms):
            cur.wait_stream(branches[key])
            t.record_stream(cur)
        return torch.stack(terms).sum()

    return loss_fn, params


def make_step(model, up):
    """Eager step (profiling tools): zero the gradients, forward, backward."""
    loss_fn, params = make_loss(model, up)

    def step(wl, wr):
        for p in params:
            p.grad = None
        loss = loss_fn(wl, wr)
        loss.backward()
        return loss

    return step, params


def bind_to_gpu_numa_node(gpu_index):
    """Run this rank's host threads on the CPUs next to its GPU (NVML's ideal affinity), so that the pinned staging
    buffers of the end-to-end path are allocated on the GPU's own NUMA node: with 8 ranks each pushing 32.8 MB per step
    through the host, cross-socket traffic is what limits the H2D streams.  Best effort: containers may forbid it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:  # noqa: BLE001
        pass


def run_ours(args):
    from biear_b200 import GraphedStep, _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # NCCL prints its banner on stdout: keep stdout = the JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    bind_to_gpu_numa_node(local)
    B = args.batch
    model = build_frontend(dev)
    rs = np.random.RandomState(3)
    up = {k: torch.from_numpy(rs.standard_normal((B, T, NBANDS)).astype(np.float32)).to(dev)
          for k in ("gYL", "gYR", "gPL", "gPR")}
    up["gC"] = torch.from_numpy(rs.standard_normal((B, NBANDS)).astype(np.float32)).to(dev)
    loss_fn, params = make_loss(model, up)
    flat_numel = sum(p.numel() for p in params)

    from biear_b200.dist import FlatGradAllReducer, captured_average
    host_reduce = dist is not None and (args.eager or not args.graph_allreduce)
    current = {"reducer": FlatGradAllReducer(params) if (dist is not None and args.eager) else None}
    # Data-parallel gradient exchange: ONE flat-bucket NCCL all-reduce (sum, x 1/world) per step, issued by the host right
    # after the step's graph replay (the bucket is written by the graph itself: no copies).  --graph-allreduce records it
    # INSIDE the step's CUDA graph instead; measured at 2 GPUs that is slower (0.885 vs 0.866 ms per step: NCCL's
    # graph-captured launch costs more than the two host-issued launches it saves), so it is not the default.
    grad_sync = captured_average(world) if (dist is not None and not host_reduce) else None
    if dist is not None:
        warm = torch.zeros(flat_numel, device=dev)
        dist.all_reduce(warm)                      # communicator set-up outside any capture
        torch.cuda.synchronize()

    def allreduce_grads():
        if current["reducer"] is not None:
            current["reducer"]()          # one flat-bucket NCCL all-reduce (sum, x 1/world)

    # resident inputs: N_ROTATE distinct batches, cycled, so a step never finds its inputs in L2
    host = [synth_binaural(B, seed=1234 + 17 * rank + i) for i in range(2)]
    dev_in = []
    for i in range(N_ROTATE):
        wl, wr = host[i % 2]
        sh = (i // 2) * 7
        dev_in.append((torch.from_numpy(np.roll(wl, sh, axis=1)).to(dev), torch.from_numpy(np.roll(wr, sh, axis=1)).to(dev)))
    if args.e2e_f32:
        pinned = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in host]
    else:   # end-to-end input = 16-bit PCM, the wire format of audio (the reference's harness expects int16-range input too,
            # train_biear.py:463-467): half the host->device bytes; biear_pcm16_to_f32 converts on the copy stream
        to_pcm = lambda x: torch.from_numpy(np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16)).pin_memory()
        pinned = [(to_pcm(a), to_pcm(b)) for a, b in host]
    e2e_bytes_per_sample = 4 if args.e2e_f32 else 2

    # The step (forward + loss + backward) is captured once per resident batch as a CUDA graph (biear_b200.GraphedStep)
    # and replayed: ~60 launches per step issued from Python are host-bound, the replay is not.  --eager times the
    # same calls issued one by one.
    if args.eager:
        def eager(wl, wr):
            for p in params:
                p.grad = None
            loss = loss_fn(wl, wr)
            loss.backward()
            return loss
        steps_res = [lambda i=i: eager(*dev_in[i]) for i in range(N_ROTATE)]
        e2e_graph = None
        launches_per_step = None
    else:
        graphs, pool = [], None
        for i in range(N_ROTATE):
            gs = GraphedStep(loss_fn, dev_in[i], params, warmup=2 if i == 0 else 1, pool=pool, copy_inputs=False,
                             flat_grads=dist is not None, grad_sync=grad_sync)
            pool = gs.pool()
            graphs.append(gs)
        # end-to-end: two graphs with their own static inputs, so the H2D of step i+1 (copy stream) overlaps step i
        e2e_graphs = [GraphedStep(loss_fn, dev_in[0], params, warmup=1, pool=pool, copy_inputs=True,
                                  flat_grads=dist is not None, grad_sync=grad_sync) for _ in range(2)]
        e2e_graph = e2e_graphs[0]
        copy_stream = torch.cuda.Stream(device=dev)
        launches_per_step = graphs[0].launches_per_replay
        if host_reduce:        # every graph writes its gradients into its own flat bucket: all-reduce that, no copies
            for gs in graphs + e2e_graphs:
                gs.reducer = FlatGradAllReducer(params, flat=gs.flat)

        def replay(gs, *inputs):
            loss = gs(*inputs)
            current["reducer"] = getattr(gs, "reducer", None)
            return loss
        steps_res = [lambda gs=gs: replay(gs) for gs in graphs]

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident-input timing ----------------------------------------------------------------
    for i in range(args.warmup):
        steps_res[i % N_ROTATE]()
        allreduce_grads()
    sync_all()
    sampler = make_sampler(local) if rank == 0 else None
    sync_all()
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        steps_res[i % N_ROTATE]()
        allreduce_grads()
    e1.record()
    sync_all()
    launches = _lib.launch_count() if launches_per_step is None else launches_per_step * args.steps
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host waveforms in, loss out ------------------------------------------
    # Every step copies ITS inputs from pinned host memory (H2D, copy stream) and reads its loss back (D2H); the copy of
    # step i+1 is issued before step i's replay, so transfer and compute overlap like in any prefetching input pipeline.
    staged = {}

    def stage(i):
        gs = e2e_graphs[i % 2]
        copy_stream.wait_stream(torch.cuda.current_stream(dev))   # the buffer's previous replay has been enqueued
        staged[i] = gs.load(*pinned[i % 2], stream=copy_stream)

    def e2e_step(i):
        if e2e_graph is None:
            loss = steps_eager_e2e(*pinned[i % 2])
        else:
            if i not in staged:
                stage(i)
            torch.cuda.current_stream(dev).wait_event(staged.pop(i))
            stage(i + 1)                      # issued BEFORE this step's replay: waits only for step i-1 (done)
            loss = replay(e2e_graphs[i % 2])
        allreduce_grads()
        # D2H read of the step's result, every step: the scalar goes to pinned host memory behind the step, and the host
        # picks it up after it has issued the NEXT step (an input pipeline's usual one-step run-ahead), so the GPU does
        # not idle through the host's turnaround; the last step's value is read before the timed region ends
        slot = i % 2
        host_loss[slot].copy_(loss.detach().reshape(()), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        prev = pending.pop("p", None)
        pending["p"] = (slot, ev)
        if prev is not None:
            prev[1].synchronize()
            losses.append(float(host_loss[prev[0]]))

    def e2e_drain():
        prev = pending.pop("p", None)
        if prev is not None:
            prev[1].synchronize()
            losses.append(float(host_loss[prev[0]]))

    def steps_eager_e2e(a, b):
        from biear_b200 import ops
        wl = a.to(dev, non_blocking=True)
        wr = b.to(dev, non_blocking=True)
        if wl.dtype == torch.int16:
            wl, wr = ops.pcm16_to_f32(wl), ops.pcm16_to_f32(wr)
        for p in params:
            p.grad = None
        loss = loss_fn(wl, wr)
        loss.backward()
        return loss

    host_loss = torch.zeros(2, dtype=torch.float32).pin_memory()
    pending, losses = {}, []
    n_warm = max(2, args.warmup // 2)
    for i in range(n_warm):
        e2e_step(i)
    e2e_drain()
    staged.clear()                            # the timed region stages its own first batch
    losses.clear()
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    for i in range(n_warm, n_warm + args.steps):
        e2e_step(i)
    e2e_drain()                               # the last step's loss is on the host before the clock stops
    e1.record()
    sync_all()
    assert len(losses) == args.steps and all(np.isfinite(losses)), "every step's loss must have been read back"
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    wall_e2e = (time.perf_counter() - t0) * 1e3

    # ---- dominant kernel alone, for the roofline line ----------------------------------------------
    roof = kernel_roofline(model, dev_in, dev, B)

    # BASELINE config 4 at N > 1 (every rank takes part: its all-reduce is the 1 634 780-float bucket of the whole model)
    full_multi = None
    if dist is not None and not args.no_extra and not args.eager:
        import bench_extra
        try:
            full_multi = bench_extra.full_step(B, steps=30, device=str(dev), world=world, rank=rank, dist=dist)
        except Exception as e:  # noqa: BLE001
            full_multi = {"error": repr(e)[:300]}

    times = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(times[0]), float(times[1])
    if rank != 0:
        if dist is not None:
            if not args.eager:
                del graphs, e2e_graphs, steps_res
                import gc
                gc.collect()
                torch.cuda.synchronize()
            dist.destroy_process_group()
        return
    clips = B * world * args.steps
    value = clips / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    step_frac = value / world * A_FULL / 1e9 / peak
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BiEAR active front-end fwd+bwd, adaptive Q (dual) + phase + CC, batch {B} x 1 s "
                               f"binaural clips @16 kHz per GPU, conf/config.yaml settings, train mode",
                   "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                   "l2": f"inputs rotate over {N_ROTATE} resident batches "
                         f"({N_ROTATE * B * 2 * FS * 4 / 1e6:.0f} MB > 126 MB L2)",
                   "issue": "eager launches" if args.eager else "one CUDA graph per step (biear_b200.GraphedStep)",
                   "allreduce_floats": flat_numel if world > 1 else 0,
                   "allreduce": ("none (1 GPU)" if world == 1 else "NCCL, issued by the host after each replay" if host_reduce
                                 else "NCCL, recorded inside the step's CUDA graph")},
        "e2e": {"value": clips / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * B * FS * e2e_bytes_per_sample,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall_e2e / args.steps,
                "input": "float32 waveforms in pinned host memory" if args.e2e_f32 else
                         "16-bit PCM waveforms in pinned host memory, converted to float32 on the device (biear_pcm16_to_f32)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "roofline_step": {"bound": "hbm", "achieved": value / world * A_FULL / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": step_frac, "bytes_per_audio_s": A_FULL, "peak_source": peak_src,
                          "note": "whole step per GPU; the adaptive path is bound by the 19-step dependency chain "
                                  "and fp32/MUFU work, not HBM (SURVEY.md 8(d))"},
    }
    import bench_extra
    roof["compute"] = bench_extra.compute_roofline(model, dev_in, value / world, roof["us_per_launch"], NCU_SUMMARY)
    if world == 1 and not args.no_extra:
        # sub-records (1 GPU only; each bounded to a few seconds): the same-box competitor and BASELINE configs 3 / 4
        for key, fn in (("gpu_eager_reference", lambda: bench_extra.gpu_eager_reference(B, steps=5, warmup=2, device=str(dev))),
                        ("fixed_q", lambda: bench_extra.fixed_q(4096, device=str(dev))),
                        ("full_step", lambda: bench_extra.full_step(B, steps=30, device=str(dev))),
                        ("variants", lambda: bench_extra.variants(B, device=str(dev)))):
            try:
                out[key] = fn()
            except Exception as e:  # noqa: BLE001  (a failing side measurement must not lose the main line)
                out[key] = {"error": repr(e)[:300]}
    if full_multi is not None:
        out["full_step"] = full_multi
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = bench_extra.reference_cpu(sample_batch=16, budget_s=20.0)
        if isinstance(out.get("full_step"), dict) and "error" not in out["full_step"]:
            try:      # config 1 (ii): the reference's own full training step on the host cores, next to `full_step`
                out["full_step"]["cpu_reference"] = bench_extra.reference_cpu_full_step(sample_batch=16, budget_s=10.0)
            except Exception as e:  # noqa: BLE001
                out["full_step"]["cpu_reference"] = {"error": repr(e)[:300]}
    print(json.dumps(out), flush=True)
    if dist is not None:
        if not args.eager:          # graphs that recorded NCCL work must be gone before the communicator is
            del graphs, e2e_graphs, steps_res
            import gc
            gc.collect()
            torch.cuda.synchronize()
        dist.destroy_process_group()


NCU_SUMMARY = os.path.join(ROOT, "profiles", "r2_ncu_summary.json")
# floats the forward recurrence saves per (row, step) for the backward: H 128 + gates 512 + LN in/out 4 x 128 + log1p(Y)
# 100 + rstd 2
SAVED_PER_ROW_STEP = 128 + 512 + 4 * 128 + NBANDS + 2


def kernel_roofline(model, dev_in, dev, B):
    """The dominant kernel alone: seq_fwd2_kernel, the persistent forward recurrence (band stage + controller, all 19
    frames, both ears), timed with CUDA events around graph replays of just the recurrence call on spectra of rotating
    input batches.  The call's two tiny companion launches (weight packing, ~5 us, and the early-exit replay check,
    ~3 us) are inside the window: < 2 % of it.  `traffic` comes from the committed ncu capture of the same kernel."""
    from biear_b200 import ops
    fb = model.fb_L
    mods = [model.fb_L, model.fb_R]
    with torch.no_grad():
        xs = [torch.view_as_real(fb._spectra([a, b])) for a, b in dev_in]
        st = lambda f: [f(m).detach() for m in mods]
        w = {"w_ih": st(lambda m: m.q_rnn.weight_ih_l0), "w_hh": st(lambda m: m.q_rnn.weight_hh_l0),
             "b_ih": st(lambda m: m.q_rnn.bias_ih_l0), "b_hh": st(lambda m: m.q_rnn.bias_hh_l0),
             "w1": st(lambda m: m.q_out[0].weight), "b1": st(lambda m: m.q_out[0].bias),
             "ln1_g": st(lambda m: m.q_out[1].weight), "ln1_b": st(lambda m: m.q_out[1].bias),
             "w2": st(lambda m: m.q_out[4].weight), "b2": st(lambda m: m.q_out[4].bias),
             "ln2_g": st(lambda m: m.q_out[5].weight), "ln2_b": st(lambda m: m.q_out[5].bias),
             "w3": st(lambda m: m.q_out[8].weight), "b3": st(lambda m: m.q_out[8].bias)}

        def run_all():
            keep = []
            for xr in xs:
                keep.append(ops.adaptive_sequence(xr, fb.fc, fb.Q0, fb.deltaQ_vec, w, fb.deltaQ_mode == "relative", True,
                                                  True, fb.cutoff, fb.df, seed=1))
            return keep
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            run_all()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run_all()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(xs))
    rows = 2 * B
    alg = rows * (T * NBINS * 8 + 6 * T * NBANDS * 4 + (T - 1) * SAVED_PER_ROW_STEP * 4)
    peak, peak_src = measured_peaks()
    ach = alg / (us * 1e-6) / 1e9
    traffic, traffic_src = None, None
    try:
        with open(NCU_SUMMARY) as f:
            k = json.load(f)["kernels"]
        name = next(n for n in k if n.startswith("seq_fwd2_kernel") and k[n]["duration_us"] > 50)
        if B == 256:                       # the capture was taken at the benchmark batch
            traffic, traffic_src = k[name]["dram_bytes"], "profiles/r2_ncu_summary.json (dram__bytes_read+write per launch)"
    except Exception:
        pass
    return {"kernel": "seq_fwd2_kernel (persistent forward recurrence: band stage + Q controller, 19 frames, both ears)",
            "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "traffic_source": traffic_src, "us_per_launch": us, "algorithmic_bytes_per_launch": alg,
            "peak_source": peak_src,
            "note": "HBM is not what bounds this kernel: ~47 MFLOP of dependent fp32/MUFU work per audio-second on a "
                    "19-step dependency chain (SURVEY.md 8(d), DESIGN.md 3.1); see profiles/ for issue-slot utilisation"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference, timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_factory(sample_batch):
    from oracle import biear_oracle as orc          # checker / CPU baseline only
    cfg = orc.FrontEndConfig(deltaq_base=1.0, deltaq_low=0.3, deltaq_high=5.0, deltaq_mode="relative")
    c = orc.constants(cfg)
    wl, wr = orc.synth_binaural(sample_batch, seed=1234)
    tl, tr = torch.from_numpy(wl), torch.from_numpy(wr)
    pl = orc.to_torch(orc.synth_controller(11), requires_grad=True)
    pr = orc.to_torch(orc.synth_controller(12), requires_grad=True)
    rs = np.random.RandomState(3)
    up = {k: torch.from_numpy(rs.standard_normal((sample_batch, T, NBANDS)).astype(np.float32))
          for k in ("gYL", "gYR", "gPL", "gPR")}
    gc = torch.from_numpy(rs.standard_normal((sample_batch, NBANDS)).astype(np.float32))
    log_q0 = torch.log(c["Q0"] + 1e-8).view(1, 1, -1)

    def step():
        for p in list(pl.values()) + list(pr.values()):
            p.grad = None
        yl, yr, ql, qr, xl, xr = orc.binaural_forward(tl, tr, pl, pr, cfg, c=c)
        phl = orc.subband_phase(xl, ql, c["f_fft"], c["fc"])
        phr = orc.subband_phase(xr, qr, c["f_fft"], c["fc"])
        cc = torch.from_numpy(orc.cc_feature_batch(wl, wr))
        lq = torch.log(0.5 * (ql + qr) + 1e-8)
        loss = (up["gYL"] * orc.log_energy(yl)).mean() + (up["gYR"] * orc.log_energy(yr)).mean() \
            + (up["gPL"] * phl).mean() + (up["gPR"] * phr).mean() + (gc * cc).mean() \
            + REG_Q_W * ((lq - log_q0) ** 2).mean() + REG_SMOOTH_W * ((lq[..., 1:] - lq[..., :-1]) ** 2).mean()
        loss.backward()
        return float(loss)

    return step


def cpu_baseline_port(sample_batch=16, budget_s=20.0, steps=None, warmup=1):
    """The oracle PORT timed on the host cores (only used when oracle/_ref -- the reference itself -- is not staged)."""
    try:     # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1)
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:  # noqa: BLE001
        pass
    step = cpu_step_factory(sample_batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    while True:
        step()
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and (el >= budget_s or n >= 50)):
            break
    return {"value": sample_batch * n / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "steps": n,
            "sample": f"{n} steps of batch {sample_batch} (same clips/weights recipe), oracle/biear_oracle.py "
                      f"(torch CPU fp32 restatement of model_torch.py + utils.py CC), {el:.1f} s, "
                      f"os.cpu_count()={os.cpu_count()}",
            "ms_per_step": el / n * 1e3}


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (oracle/_ref) on the box's host cores, all the
    threads it can use, on our arm's config / metric / unit; each step a bounded sample (16 clips) of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import bench_extra
    sample = 16
    warm = max(3, min(args.warmup, 5))
    steps = max(10, min(args.steps, 40))
    cb = bench_extra.reference_cpu(sample_batch=sample, steps=steps, warmup=warm)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": cb.get("steps", steps),
        "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BiEAR active front-end fwd+bwd, adaptive Q (dual) + phase + CC, batch {args.batch} x 1 s "
                               f"binaural clips @16 kHz per GPU, conf/config.yaml settings, train mode",
                   "note": f"CPU arm ({cb['kind']}): each step is a bounded sample of {sample} clips of that workload, train-mode dropout"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the gpu_eager_reference / fixed_q / full_step sub-records")
    ap.add_argument("--eager", action="store_true", help="issue every launch from Python instead of replaying CUDA graphs")
    ap.add_argument("--graph-allreduce", action="store_true", help="record the gradient all-reduce inside the step's CUDA graph")
    ap.add_argument("--e2e-f32", action="store_true", help="end-to-end arm with float32 host waveforms instead of 16-bit PCM")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
