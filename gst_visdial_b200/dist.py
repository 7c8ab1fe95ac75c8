"""Sharding of independent dialogs over ranks and the single final gather.

Images are independent units, so there is no collective inside a dialog, a round or a step: rank r of n processes the
contiguous block of global image indices given by ``shard_range`` and the generated token ids / perplexities are
gathered once at the end (NCCL over NVLink on GPUs, gloo in the CPU tests).  This replaces nn.DataParallel's
per-forward parameter broadcast + output gather (generate.py:67,77).
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of rank ``rank``; blocks differ by at most one element."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the default group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def gather_results(tensors: List[torch.Tensor], counts: List[int]) -> List[torch.Tensor]:
    """All-gathers per-rank result tensors whose first dimension is that rank's shard size (``counts[r]``) and returns
    them concatenated in global image order.  Ragged shards are padded to the largest one for the collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tensors
    world = dist.get_world_size()
    mx = max(counts)
    out = []
    for t in tensors:
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        out.append(torch.cat([bufs[r][: counts[r]] for r in range(world)], 0))
    return out
