"""Batch sharding of utterances across ranks (inference shards naturally, no data-path collective).

Rank r of G takes the contiguous slice ``[r*N/G, (r+1)*N/G)`` of the utterance list; weights are replicated.
The only communication is a barrier and (optionally) a gather of results, done by the caller with
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def shard_slice(n_items: int, rank: int, world_size: int) -> slice:
    """Contiguous, balanced split: the first ``n_items % world_size`` ranks get one extra item."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def enhance_sharded(enhance: Callable[[torch.Tensor], torch.Tensor], waves: torch.Tensor, rank: int, world_size: int,
                    micro_batch: int = 64, gather: bool = False, group=None) -> Optional[torch.Tensor]:
    """Run ``enhance`` over this rank's slice of ``waves`` (N, L) in micro-batches.

    Returns this rank's outputs, or -- with ``gather=True`` -- the full (N, L) result on every rank
    (all_gather of padded shards; only used by tests / small jobs, never on the timed path)."""
    sl = shard_slice(waves.shape[0], rank, world_size)
    mine = waves[sl]
    outs = [enhance(mine[i:i + micro_batch]) for i in range(0, mine.shape[0], micro_batch)]
    local = torch.cat(outs, 0) if outs else waves.new_zeros((0, waves.shape[1]))
    if not gather:
        return local
    import torch.distributed as dist
    per = -(-waves.shape[0] // world_size)
    padded = local.new_zeros((per, waves.shape[1]))
    padded[:local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(bufs, padded, group=group)
    parts = [bufs[r][:shard_slice(waves.shape[0], r, world_size).stop - shard_slice(waves.shape[0], r, world_size).start]
             for r in range(world_size)]
    return torch.cat(parts, 0)
