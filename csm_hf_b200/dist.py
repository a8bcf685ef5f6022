"""Batch-sharded generation across the GPUs of one node.

Sequences are independent everywhere on the path; the only cross-sequence operation of the
reference is the stop rule `torch.all(new_frame == 0)` over the whole batch
(modeling_csm.py:662).  So: one process per GPU, a full weight replica each, rank r owns a
contiguous slice of the batch, no collective on the data path, and all-gathers of the emitted
frame tokens ([B_local, n, 32] int64, 256 B per sequence-frame):

  * stop_on_all_zeros=False: ONE all-gather at the end;
  * stop_on_all_zeros=True: every rank generates `stop_check_every` frames at a time
    (CSMModel.generate_more continues the same decode loop), the chunk is all-gathered and the
    reference's stop rule is evaluated on the gathered frames -- the gather IS the stop-flag
    exchange -- so a batch that ends early stops within one chunk instead of burning the whole
    frame budget.  The result is exactly what the single-GPU loop returns (generation is
    deterministic, running a shard a few frames past the stop point cannot change earlier frames).
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
    return torch.cat(parts, dim=0)


def _comm_device(model, like: torch.Tensor) -> torch.device:
    """Tensors handed to the collective live where the backend needs them: on the model's GPU under NCCL."""
    if dist.get_backend() == "nccl":
        return torch.device(model.device)
    return like.device


def generate_sharded(model, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], max_new_frames: int = 100,
                     temperature: float = 1.0, topk: int = 50, use_cache: bool = True, stop_on_all_zeros: bool = True,
                     group=None, stop_check_every: int = 16) -> torch.Tensor:
    """`CSMModel.generate` for a batch split over the ranks of `group`.  Every rank passes the
    full batch and receives the full result (on the device of `input_ids`)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B = input_ids.shape[0]
    lo, hi = shard_bounds(B, rank, world)
    nb = hi - lo
    out_dev = input_ids.device
    comm = _comm_device(model, input_ids)
    chunk = max_new_frames if not stop_on_all_zeros else max(1, min(stop_check_every, max_new_frames))
    old_base = getattr(model, "seq_base", 0)
    model.seq_base = lo          # sampling noise is keyed by the global sequence index: sharding does not change a draw
    try:
        if nb > 0:
            mask = attention_mask[lo:hi] if attention_mask is not None else None
            local = model.generate(input_ids[lo:hi], mask, max_new_frames=chunk, temperature=temperature, topk=topk,
                                   use_cache=use_cache, stop_on_all_zeros=False, reserve_frames=max_new_frames)
        else:
            # an empty shard generates nothing but keeps the per-call sampling counter (part of the noise seed)
            # in step with the other ranks
            if not (temperature == 0 or topk == 1):
                model._sample_calls = getattr(model, "_sample_calls", 0) + 1
            local = torch.zeros(0, chunk, 32, dtype=torch.long, device=comm)
        parts, done = [], 0
        while True:
            got = all_gather_frames(local.to(comm), B, group)
            done += got.shape[1]
            if stop_on_all_zeros:
                cut = truncate_at_global_stop(got)
                parts.append(cut)
                if cut.shape[1] < got.shape[1]:
                    break
            else:
                parts.append(got)
            if done >= max_new_frames:
                break
            n = min(chunk, max_new_frames - done)
            local = model.generate_more(nb, n, stop_on_all_zeros=False) if nb > 0 else \
                torch.zeros(0, n, 32, dtype=torch.long, device=comm)
        return torch.cat(parts, dim=1).to(out_dev)
    finally:
        model.seq_base = old_base


def allreduce_gradients(model, group=None, bucket_bytes: int = 256 << 20) -> int:
    """Data-parallel training (plain DP, one replica per GPU): average the parameter gradients
    over the ranks after loss.backward().  Gradients are packed into flat buckets of ~`bucket_bytes` (3.1 GB of bf16
    gradients for csm-1b -> a dozen NCCL all-reduces over NVLink instead of 187), summed, scaled by 1/world and
    unpacked in place.  Works with any backend (gloo on CPU in the tests).  -> number of buckets reduced."""
    if getattr(model, "_grads_reduced", False):      # training.enable_data_parallel averaged them during the backward
        model._grads_reduced = False
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for _, p in sorted(model.named_parameters()) if p.grad is not None]
    if world == 1 or not grads:
        return 0
    flat = getattr(model, "_grad_flat", None)
    if flat is not None and len(grads) == len(list(model.parameters())) and \
            all(g.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for g in grads):
        # gradients of csm_train_step are views into one flat buffer (training.py): one collective, no copies
        # (the few padding elements between the views are reduced along; nothing reads them)
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)      # (the average inside the collective)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(world)
        return 1
    buckets, cur, cur_bytes = [], [], 0
    for g in grads:
        nbytes = g.numel() * g.element_size()
        if cur and (cur_bytes + nbytes > bucket_bytes or g.dtype != cur[0].dtype):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(g)
        cur_bytes += nbytes
    if cur:
        buckets.append(cur)
    for b in buckets:
        flat = torch.cat([g.reshape(-1) for g in b])
        if flat.dtype == torch.bfloat16 and flat.device.type == "cpu":
            flat = flat.float()                       # (gloo has no bf16 sum)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in b:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    return len(buckets)
