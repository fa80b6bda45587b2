"""Host-side helpers for the data-parallel path (one process per GPU).

The loss shards by batch item with no data-path collective: every rank runs the fused step on
its own images.  The only cross-rank operations are (i) the scalar loss mean for logging and
(ii) timing reductions for the benchmark; the networks' gradient all-reduce is DDP's job."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_range(n_items: int, r: int, w: int):
    """Contiguous, equal shards (the reference loader uses drop_last, loader.py:61)."""
    if n_items % w != 0:
        raise ValueError(f"batch {n_items} is not divisible by world size {w}")
    per = n_items // w
    return r * per, (r + 1) * per


def shard_batch(tensors, r: int, w: int):
    """Slice dim 0 of every tensor (or list of tensors) to this rank's shard."""
    def cut(t):
        lo, hi = shard_range(t.shape[0], r, w)
        return t[lo:hi].contiguous()
    return {k: ([cut(t) for t in v] if isinstance(v, (list, tuple)) else (cut(v) if torch.is_tensor(v) else v))
            for k, v in tensors.items()}


def mean_over_ranks(x: torch.Tensor) -> torch.Tensor:
    """Global-batch loss from equal-size per-rank means (processor.py:212 takes the batch mean)."""
    if world() == 1:
        return x
    y = x.detach().clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM)
    return y / world()


def max_over_ranks(value: float, device=None) -> float:
    if world() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
