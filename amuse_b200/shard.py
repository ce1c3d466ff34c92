"""Clip sharding across GPUs: every clip is independent through
AST -> denoise -> decode, so the batch is cut into contiguous per-rank slices, weights are
broadcast once at init and poses are gathered once at the end.  No collective inside the loop.
Works on any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of ``range(n_global)`` owned by ``rank``; the first ``n_global % world``
    ranks get one extra clip.  Concatenating the slices in rank order restores the batch."""
    base, rem = divmod(n_global, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_state_dict(sd: Dict[str, torch.Tensor], src: int = 0, device: Optional[torch.device] = None):
    """In-place broadcast of every tensor of ``sd`` from ``src`` (all ranks hold same-shaped tensors)."""
    for k in sd:
        t = sd[k].to(device) if device is not None else sd[k]
        t = t.contiguous()
        dist.broadcast(t, src=src)
        sd[k] = t
    return sd


def gather_clips(local: torch.Tensor, n_global: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank clip tensors ``[n_local, ...]`` (ragged over ranks) to ``dst`` in clip order."""
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [shard_range(n_global, r, world) for r in range(world)]
    n_max = max(b - a for a, b in sizes)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs: Optional[List[torch.Tensor]] = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
