"""World-size-2 gloo test of the data-parallel plumbing: shard the batch, all-reduce the flat
gradient bucket, and require the result to equal the full-batch gradient.  Runs on CPU, so the module under the
all-reduce is the plain-PyTorch back-end of the drop-in namespace (the front-end itself needs a GPU; its
multi-GPU run is bench.py under torchrun)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from biear_b200 import model_torch as mt
from biear_b200.dist import FlatGradAllReducer, captured_average, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(n):
    rs = np.random.RandomState(5)
    f = lambda *s: torch.from_numpy(rs.standard_normal(s).astype(np.float32))
    return f(n, 19, 100), f(n, 19, 100), f(n, 100), f(n, 19, 100), f(n, 19, 100), f(n, 8)


def _loss(model, batch, denom):
    x1, x2, x3, x4, x5, tgt = batch
    sound, aoa, dist_logits = model(x1, x2, x3, x4, x5)
    return ((sound - tgt) ** 2).sum() / denom + (aoa ** 2).sum() / denom + (dist_logits ** 2).sum() / denom


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = mt.build_model(use_cc=True).eval()          # eval: dropout off, so shards are comparable
    lo, hi = shard_bounds(n, rank, world)
    batch = tuple(t[lo:hi] for t in _inputs(n))
    _loss(model, batch, hi - lo).backward()             # per-shard mean loss
    red = FlatGradAllReducer(model.parameters())
    flat = red(weight=(hi - lo) / n)                    # exact for unequal shards
    assert red.numel == 1288468                         # fixed-Q / passive parameter count
    if rank == 0:
        np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
    # the capturable form (what GraphedStep records into the step's graph) gives the plain average of the buckets
    mine = torch.full((5,), float(rank + 1))
    captured_average(world)(mine)
    assert torch.allclose(mine, torch.full((5,), sum(range(1, world + 1)) / world))
    # every rank must hold the same reduced gradients
    chk = torch.tensor([float(flat.double().sum())], dtype=torch.float64)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    assert all(torch.equal(b, both[0]) for b in both)
    dist.destroy_process_group()


def test_shard_bounds():
    for n in (1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flat_allreduce_equals_full_batch_gradient(tmp_path):
    n, world = 7, 2                                      # unequal shards (4 + 3)
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    flat = np.load(tmp_path / "flat.npy")
    torch.manual_seed(0)
    model = mt.build_model(use_cc=True).eval()
    _loss(model, _inputs(n), n).backward()
    full = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).numpy()
    err = np.max(np.abs(flat - full)) / np.max(np.abs(full))
    assert err < 1e-5, err
