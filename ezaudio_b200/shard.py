"""Multi-GPU mode of the hot path: prompts are independent units (no cross-sample op anywhere: LayerNorm is per token,
the CFG-rescale std is per sample, src/inference.py:17-18), so a batch of prompts is split contiguously across ranks
(CFG pair kept on the same rank), weights replicated, and NO collective runs inside the sampling loop.
torch.distributed (NCCL on GPUs, gloo in CPU tests) is used only at the edges: optional gather of the waveforms."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first n_items % world ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_prompts(prompts: Sequence[str], world: int, rank: int) -> List[str]:
    a, b = shard_range(len(prompts), world, rank)
    return list(prompts[a:b])


def gather_waveforms(local: torch.Tensor, n_total: int, dist=None) -> torch.Tensor:
    """local: (n_local, T) waveforms of this rank's shard (device tensor for NCCL, CPU tensor for gloo).
    Returns (n_total, T) on every rank, in prompt order.  Shards may be ragged by one row: pad to the max, then trim."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = [shard_range(n_total, world, r) for r in range(world)]
    mx = max(b - a for a, b in per)
    pad = torch.zeros(mx, local.shape[1], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, per)], 0)
