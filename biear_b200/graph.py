"""CUDA-graph capture of a whole front-end training step.

One step of the hot path is ~60 launches (STFT, the persistent recurrence kernels, the weight-gradient GEMMs, the
CC kernel and the PyTorch glue of the loss); issued eagerly from Python they are host-bound on a busy box (measured:
15 ms/step of wall time around 2 ms of kernel time).  Capturing the step once and replaying it removes the host
from the loop.  The reference has no counterpart (it issues ~2000 eager launches and ~50 host syncs per step,
SURVEY.md 3.1-3.2).

    step = GraphedStep(fn, example_inputs, params)      # fn(*inputs) -> scalar loss; captured with its backward
                                                        # (flat_grads=True: the gradients are slices of step.flat)
    loss = step(wavL, wavR)                             # copies the inputs into the static buffers, replays
    # params[i].grad now hold this step's gradients; loss is a device scalar (static tensor)

Dropout stays random across replays: while capturing, the front-end reads its Philox seed from a device counter that
the graph itself advances (ops._captured_seed).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence

import torch


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class GraphedStep:
    def __init__(self, fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor],
                 params: Iterable[torch.nn.Parameter], warmup: int = 3, pool=None, copy_inputs: bool = True,
                 flat_grads: bool = False, grad_sync: Optional[Callable[[torch.Tensor], None]] = None):
        """grad_sync (needs flat_grads): called on the flat gradient bucket INSIDE the capture, right after the backward --
        e.g. an NCCL all-reduce + scaling for data-parallel training, which then replays as part of the step's graph
        instead of being issued by the host after it."""
        self.params = [p for p in params if p.requires_grad]
        self._pcm_stage = {}
        assert grad_sync is None or flat_grads, "grad_sync works on the flat gradient bucket (flat_grads=True)"
        self.copy_inputs = copy_inputs
        self.static_inputs = [x.clone() for x in example_inputs] if copy_inputs else list(example_inputs)
        dev = self.static_inputs[0].device
        # warm-up and capture run on a HIGH-priority stream: the kernels the front-end forks onto its own (lowest-priority)
        # side stream -- the CC feature -- then yield the SMs to the persistent recurrence kernels whenever both are
        # eligible, and fill the SMs those leave idle
        side = torch.cuda.Stream(device=dev, priority=-1)
        side.wait_stream(torch.cuda.current_stream(dev))
        # Gradients are taken with torch.autograd.grad (no AccumulateGrad nodes): those nodes remember the stream
        # they were created on, and one left over from an earlier eager backward on the legacy default stream would
        # make the capture illegal.
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):          # lazy one-time setup (tables, smem opt-in, allocator warm-up)
                torch.autograd.grad(fn(*self.static_inputs), self.params, allow_unused=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool, stream=side, capture_error_mode="thread_local"):
            self.loss = fn(*self.static_inputs)
            grads = torch.autograd.grad(self.loss, self.params, allow_unused=True)
            if flat_grads:
                # one contiguous bucket (written by the graph itself) whose slices become the .grad tensors: a
                # data-parallel step then needs a single all-reduce on `flat` and no per-parameter copies
                self.flat = torch.cat([(g if g is not None else torch.zeros_like(p)).reshape(-1)
                                       for g, p in zip(grads, self.params)])
                views, off = [], 0
                for p in self.params:
                    views.append(self.flat[off:off + p.numel()].view_as(p))
                    off += p.numel()
                grads = views
                if grad_sync is not None:
                    grad_sync(self.flat)
        self.launches_per_replay = _lib.launch_count() - n0   # kernels of libbiear_b200.so recorded in the graph
        self.loss = self.loss.detach()
        self.grads = list(grads)                      # static tensors (graph pool) rewritten by every replay
        for p, g in zip(self.params, self.grads):
            p.grad = g

    def pool(self):
        return self.graph.pool()

    def load(self, *inputs: torch.Tensor, stream: Optional[torch.cuda.Stream] = None):
        """Copy (host or device) inputs into the graph's static input buffers, optionally on another stream so that
        the H2D transfer of the next step overlaps this step's replay; returns an event to wait on before replay()."""
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            for i, (dst, src) in enumerate(zip(self.static_inputs, inputs)):
                if src.dtype == torch.int16 and dst.dtype == torch.float32:
                    # 16-bit PCM wire format: half the host->device bytes; converted next to the copy (biear_pcm16_to_f32)
                    from . import ops
                    stage = self._pcm_stage.get(i)
                    if stage is None:
                        stage = self._pcm_stage[i] = torch.empty(dst.shape, dtype=torch.int16, device=dst.device)
                    stage.copy_(src, non_blocking=True)
                    ops.pcm16_to_f32(stage, dst)
                else:
                    dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        return ev

    def replay(self) -> torch.Tensor:
        return self.__call__()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if self.copy_inputs and inputs:
            for dst, src in zip(self.static_inputs, inputs):
                dst.copy_(src, non_blocking=True)
        for p, g in zip(self.params, self.grads):
            if p.grad is not g:                       # e.g. after optimizer.zero_grad(set_to_none=True)
                p.grad = g
        self.graph.replay()
        return self.loss
