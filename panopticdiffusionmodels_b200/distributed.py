"""Multi-GPU plumbing of the sampling path: one process per GPU, the sample batch sharded by rank, weights
replicated, no collective inside the denoising loop, ONE all-gather of the finished ``(z, pred_mask)`` per batch.

Mirrors ``utils.sample2dir`` of the reference (``utils.py:561-640``: ``batch_size = mini_batch_size * num_processes``,
``amortize`` loop, ``accelerator.gather`` at ``:585-588``) and its per-rank seeding
(``train_t2i_discrete.py:237`` ``set_seed(seed, device_specific=True)``)."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist

from .utils import amortize


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def rank_seed(seed: int) -> int:
    """device-specific seed: ``seed + rank`` (accelerate's ``set_seed(..., device_specific=True)``)."""
    return seed + world()[0]


def gather_samples(z: torch.Tensor, pred_mask: Optional[torch.Tensor]):
    """All-gather along dim 0, rank-major (what ``accelerator.gather`` returns)."""
    rank, n = world()
    if n == 1:
        return z, pred_mask

    def gather(t):
        t = t.contiguous()
        out = torch.empty((n * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(out, t)
        else:
            dist.all_gather(list(out.chunk(n, dim=0)), t)
        return out

    return gather(z), (None if pred_mask is None else gather(pred_mask))


def sample_all(sample_fn: Callable[[int], Tuple[torch.Tensor, Optional[torch.Tensor]]], n_samples: int,
               mini_batch_size: int) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Draw ``n_samples`` joint samples over all ranks.  ``sample_fn(b)`` returns this rank's ``(z, pred_mask)`` for a
    mini-batch of ``b``; every global batch is ``mini_batch_size * world`` samples followed by one all-gather; the
    tail is trimmed to ``n_samples`` exactly like ``utils.sample2dir`` (``utils.py:564, 575, 599-601``)."""
    rank, n = world()
    zs: List[torch.Tensor] = []
    ms: List[torch.Tensor] = []
    done = 0
    for _ in amortize(n_samples, mini_batch_size * n):
        z, pm = sample_fn(mini_batch_size)
        z, pm = gather_samples(z, pm)
        take = min(z.shape[0], n_samples - done)
        zs.append(z[:take])
        if pm is not None:
            ms.append(pm[:take])
        done += take
    return torch.cat(zs, 0), (torch.cat(ms, 0) if ms else None)
