"""Utterance sharding across the GPUs of one box.

Every op of the decoder is per utterance, so ranks never exchange activations: rank r of W
decodes utterances [r*B/W, (r+1)*B/W) with replicated weights, and the only collective is the
gather of the waveforms to rank 0 (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_utterances: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) slice of the utterance list for `rank` (first ranks get the remainder)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, rem = divmod(n_utterances, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_waveforms(wave: torch.Tensor, counts: Optional[List[int]] = None, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank waveforms [b_r,1,S] to `dst`; returns the concatenated [sum b_r,1,S] there, None elsewhere.
    `counts` = utterances per rank (needed when shards are ragged; equal shards by default)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return wave
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = counts or [wave.shape[0]] * world
    bmax = max(counts)
    send = wave
    if wave.shape[0] < bmax:                      # pad ragged shards to a common shape for the collective
        send = torch.zeros((bmax,) + tuple(wave.shape[1:]), dtype=wave.dtype, device=wave.device)
        send[: wave.shape[0]] = wave
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send.contiguous(), bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
