"""Batch-of-videos sharding across the GPUs of one box (SURVEY.md §8 e).

Every (video, position) token update is independent given that video's logits, so the batch is
partitioned contiguously across ranks, each rank runs the whole reverse chain on its videos with no
communication (its Philox stream keyed by GLOBAL row index, so the gathered result equals the
single-GPU run bit for bit), and the only collective is one all-gather of the int64 `[B_local, N]`
tokens at the end (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous `[begin, end)` slice of the video batch owned by `rank` (remainder spread over the
    first ranks)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world {world_size}")
    base, extra = divmod(global_batch, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_tokens(local_tokens: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather int64 `[B_local, N]` token grids into `[global_batch, N]` on every rank.

    Equal shards use a single `all_gather_into_tensor`; ragged shards are padded to the largest
    shard and trimmed after the gather.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert local_tokens.shape[0] == global_batch
        return local_tokens
    world, rank = dist.get_world_size(), dist.get_rank()
    N = local_tokens.shape[1]
    sizes = [shard_range(global_batch, world, r) for r in range(world)]
    widest = max(e - b for b, e in sizes)
    assert local_tokens.shape[0] == sizes[rank][1] - sizes[rank][0]
    send = local_tokens.contiguous()
    if send.shape[0] != widest:
        pad = torch.zeros(widest - send.shape[0], N, dtype=send.dtype, device=send.device)
        send = torch.cat([send, pad], 0)
    recv = torch.empty(world * widest, N, dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send)
    if all(e - b == widest for b, e in sizes):
        return recv
    parts = [recv[r * widest: r * widest + (e - b)] for r, (b, e) in enumerate(sizes)]
    return torch.cat(parts, 0)
