"""Batch data-parallel plumbing for the front-end: one process per GPU, contiguous equal batch
shards, weights replicated, and ONE flat-bucket all-reduce of the (tiny) gradients per step.

The reference has no distributed code (single device, train_biear.py:120); the only exchange the path needs is
the sum of the per-shard gradients before the two global-norm clips (train_biear.py:523-525).  Clips are
independent, so there is no data-path collective: 1 634 780 fp32 gradients (6.5 MB) over NCCL/NVLink per step.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank; shards differ by at most one item, earlier ranks get the extras."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class FlatGradAllReducer:
    """Averages the .grad of `params` across the process group through one persistent flat fp32 bucket.

    Loss terms are batch means, so with equal shards the average of the per-rank gradients equals the full-batch
    gradient (LayerNorm only, no BatchNorm anywhere on the path).  Parameters whose .grad is None on a rank
    contribute zeros.  `weight` lets unequal shards be combined exactly: pass shard_size / global_batch and the
    result is the full-batch-mean gradient.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 flat: Optional[torch.Tensor] = None):
        """flat: an existing bucket that already holds the gradients, parameter after parameter (e.g.
        GraphedStep(..., flat_grads=True).flat, whose slices ARE the .grad tensors): no copies are made then."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.aliased = flat is not None
        if flat is not None:
            assert flat.numel() == self.numel and flat.dtype == torch.float32
        self.flat = flat if flat is not None else torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @torch.no_grad()
    def __call__(self, weight: Optional[float] = None) -> torch.Tensor:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if not self.aliased:
            for p, v in zip(self.params, self.views):
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
        if world > 1:
            if weight is not None:
                self.flat.mul_(weight)
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            elif dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)    # the 1 / world scaling rides in the collective
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / world)
        if not self.aliased:
            for p, v in zip(self.params, self.views):
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    p.grad.copy_(v)
        return self.flat


def captured_average(world: int, group: Optional[dist.ProcessGroup] = None):
    """grad_sync callback for GraphedStep(flat_grads=True, grad_sync=...): sum the flat gradient bucket over the ranks and
    scale by 1 / world, both recorded inside the step's CUDA graph (NCCL collectives are capturable), so that the exchange
    replays with the step -- behind the kernel that produces the bucket -- instead of being issued by the host per step."""
    def sync(flat: torch.Tensor):
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.mul_(1.0 / world)
    return sync
