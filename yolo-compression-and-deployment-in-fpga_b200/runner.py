"""Batch sharding over the GPUs of one box.

Frames are independent at inference (trackers frozen, slim_yolo_v2.py:215,28-29), so the batch is split contiguously
by frame index, weights are replicated at load, and there is NO collective on the hot path.  Only the final
per-frame detection lists are gathered (fixed-capacity [frames][max_det] records + counts), in frame order.
One process per GPU; torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of the frames rank `rank` owns; the first n % world ranks get one extra frame."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(dets: torch.Tensor, counts: torch.Tensor, n_frames: int, group=None):
    """dets: [local_frames, max_det, 8] int32 view of yolo_b200_det records, counts: [local_frames] int32.
    Returns (all_dets [n_frames, md, 8], all_counts [n_frames]) on every rank, in global frame order, where md is the
    largest detection count of any frame (the lists are trimmed to their filled part before they travel).
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dets, counts
    rank = dist.get_rank(group)
    max_local = shard_range(n_frames, 0, world)[1]            # rank 0 always has the largest shard
    # only the filled part of the fixed-capacity lists travels: trim to the largest count of any frame on any rank
    mc = counts.max().to(torch.int64).reshape(1) if counts.numel() else torch.zeros(1, dtype=torch.int64, device=counts.device)
    dist.all_reduce(mc, op=dist.ReduceOp.MAX, group=group)
    md = max(1, min(int(mc.item()), dets.shape[1]))
    pad_d = torch.zeros((max_local, md, 8), dtype=dets.dtype, device=dets.device)
    pad_c = torch.zeros((max_local,), dtype=counts.dtype, device=counts.device)
    lo, hi = shard_range(n_frames, rank, world)
    pad_d[:hi - lo] = dets[:, :md]
    pad_c[:hi - lo] = counts
    out_d = [torch.empty_like(pad_d) for _ in range(world)]
    out_c = [torch.empty_like(pad_c) for _ in range(world)]
    dist.all_gather(out_d, pad_d, group=group)
    dist.all_gather(out_c, pad_c, group=group)
    ds: List[torch.Tensor] = []
    cs: List[torch.Tensor] = []
    for r in range(world):
        l, h = shard_range(n_frames, r, world)
        ds.append(out_d[r][:h - l])
        cs.append(out_c[r][:h - l])
    return torch.cat(ds), torch.cat(cs)
