"""Optional instrumentation of the front-end calls (off by default; the drop-in namespace switches it on with
BIEAR_TIMING=1 so that a run of the reference's unchanged scripts reports the front-end's share of a step).

A span records a CUDA event pair on the current stream (device time the front-end's work occupied, including any gaps in
which the GPU waited for the host to issue the next launch) and the host wall time spent inside the call."""
from __future__ import annotations

import time
from contextlib import contextmanager

import torch

enabled = False
_spans = {}          # name -> list of (host seconds, event0, event1, batch)


@contextmanager
def span(name: str, batch: int = 0):
    if not enabled or not torch.cuda.is_available() or torch.cuda.is_current_stream_capturing():
        yield
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    try:
        yield
    finally:
        e1.record()
        _spans.setdefault(name, []).append((time.perf_counter() - t0, e0, e1, batch))


def time_backward_node(name: str, batch: int, node):
    """Record a span around the execution of one autograd node (the graphed front-end's backward replay)."""
    if not enabled or node is None:
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = [0.0]

    def pre(grad_outputs):
        t[0] = time.perf_counter()
        e0.record()

    def post(grad_inputs, grad_outputs):
        e1.record()
        _spans.setdefault(name, []).append((time.perf_counter() - t[0], e0, e1, batch))

    node.register_prehook(pre)
    node.register_hook(post)


def reset():
    _spans.clear()


def summary(skip: int = 4):
    """name -> {calls, batch (mode), device_ms, host_ms}: means over the calls after the first `skip` (warm-up) of the most
    frequent batch size."""
    torch.cuda.synchronize()
    out = {}
    for name, rows in _spans.items():
        sizes = [r[3] for r in rows]
        mode = max(set(sizes), key=sizes.count) if sizes else 0
        sel = [r for r in rows if r[3] == mode][skip:] or rows      # (the first calls: lazy set-up and graph capture)
        out[name] = {"calls": len(rows), "batch": mode,
                     "device_ms": sum(r[1].elapsed_time(r[2]) for r in sel) / len(sel),
                     "host_ms": 1e3 * sum(r[0] for r in sel) / len(sel)}
    return out


def report(prefix: str = "[biear_b200 timing]"):
    s = summary()
    for name, d in s.items():
        print(f"{prefix} {name}: {d['calls']} calls, batch {d['batch']}: {d['device_ms']:.3f} ms device, "
              f"{d['host_ms']:.3f} ms host per call", flush=True)
    if "frontend.forward[train]" in s and "frontend.backward" in s:
        f, b = s["frontend.forward[train]"], s["frontend.backward"]
        print(f"{prefix} front-end per training step (forward + backward): {f['device_ms'] + b['device_ms']:.3f} ms device, "
              f"{f['host_ms'] + b['host_ms']:.3f} ms host", flush=True)
    return s
