"""Utterance-batch sharding across GPUs. Utterances are independent in eval (BatchNorm uses
running statistics, attention never crosses the batch axis), so rank r simply owns a contiguous slice of the batch and
the only exchange is one all-gather of the (U_local, labels) fp32 logits. torch.distributed is the plumbing (NCCL over
NVLink on GPUs; the same code runs under gloo on CPU for the host-logic tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_utterances: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: rank r gets [lo, hi); the first (n % world) ranks take one extra utterance."""
    base, extra = divmod(n_utterances, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: tuple, rank: int, world: int) -> tuple:
    """Slice every per-utterance tensor/list of the reference batch tuple (utils/dataset.py:291-292)."""
    n = len(batch[9]) if not torch.is_tensor(batch[9]) else batch[9].shape[0]
    lo, hi = shard_range(n, rank, world)
    return tuple(x[lo:hi] for x in batch)


def gather_logits(local_logits: torch.Tensor, n_utterances: int) -> torch.Tensor:
    """All-gather the per-rank logits back into batch order. Ragged shards are padded to the largest one."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_logits
    world = dist.get_world_size()
    sizes = [shard_range(n_utterances, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    labels = local_logits.shape[1]
    padded = local_logits.new_zeros(mx, labels)
    padded[: local_logits.shape[0]] = local_logits
    out = local_logits.new_empty(world * mx, labels)
    dist.all_gather_into_tensor(out, padded.contiguous())
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
