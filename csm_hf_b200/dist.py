"""Batch-sharded generation across the GPUs of one node.

Sequences are independent everywhere on the path; the only cross-sequence operation of the
reference is the stop rule `torch.all(new_frame == 0)` over the whole batch
(modeling_csm.py:662).  So: one process per GPU, a full weight replica each, rank r owns a
contiguous slice of the batch, no collective on the data path, and ONE all-gather of the
emitted frame tokens ([B_local, n, 32] int64, 256 B per sequence-frame) at the end.  The
global stop rule is evaluated on the gathered tensor, which gives exactly the frames the
single-GPU reference would have kept (generation is deterministic, so running a shard past
the global stop point cannot change earlier frames).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first batch % world ranks get one extra sequence."""
    q, r = divmod(batch, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def truncate_at_global_stop(frames: torch.Tensor) -> torch.Tensor:
    """Reference stop rule on the gathered batch: drop the first all-zero frame and everything after."""
    if frames.shape[1] == 0:
        return frames
    allzero = (frames == 0).all(dim=2).all(dim=0)          # [n]
    idx = torch.nonzero(allzero)
    n = int(idx[0]) if idx.numel() else frames.shape[1]
    return frames[:, :n]


def all_gather_frames(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """[B_local, n, 32] on every rank -> [B, n, 32] on every rank (rank order == batch order)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (batch + world - 1) // world
    n = local.shape[1]
    pad = torch.zeros(per, n, local.shape[2], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(batch, r, world)
        parts.append(out[r][: hi - lo])
    del rank
    return torch.cat(parts, dim=0)


def generate_sharded(model, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], max_new_frames: int = 100,
                     temperature: float = 1.0, topk: int = 50, use_cache: bool = True, stop_on_all_zeros: bool = True,
                     group=None) -> torch.Tensor:
    """`CSMModel.generate` for a batch split over the ranks of `group`.  Every rank passes the
    full batch and receives the full result."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B = input_ids.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    if hi > lo:
        model.seq_base = lo      # sampling noise is keyed by the global sequence index: sharding does not change a sequence's draw
        mask = attention_mask[lo:hi] if attention_mask is not None else None
        local = model.generate(input_ids[lo:hi], mask, max_new_frames=max_new_frames, temperature=temperature,
                               topk=topk, use_cache=use_cache, stop_on_all_zeros=False)
    else:
        local = torch.zeros(0, max_new_frames, 32, dtype=torch.long, device=input_ids.device)
    frames = all_gather_frames(local, B, group)
    return truncate_at_global_stop(frames) if stop_on_all_zeros else frames
