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


class DetectionGatherer:
    """Asynchronous, double-buffered all-gather of the per-rank detection lists (NCCL over NVLink on GPUs).

    The gather of step i runs on the collective's own stream while step i+1 computes, so it costs the hot path nothing;
    `finish()` (or the next use of the same buffer) waits for it.  No host synchronisation: the lists travel at a fixed
    capacity `cap` <= max_det (callers pass the number of anchors per frame, which bounds any count), so no count has to
    reach the host first.  Usage per step:  buf = g.buffers(i); <write detections into buf.dets / buf.counts>; g.launch(i)."""

    class _Buf:
        def __init__(self, local_frames, max_det, cap, world, device):
            self.dets = torch.zeros((local_frames, max_det, 8), dtype=torch.int32, device=device)
            self.counts = torch.zeros((local_frames,), dtype=torch.int32, device=device)
            self.send = torch.zeros((local_frames, cap, 8), dtype=torch.int32, device=device)
            self.all_dets = torch.zeros((world * local_frames, cap, 8), dtype=torch.int32, device=device)
            self.all_counts = torch.zeros((world * local_frames,), dtype=torch.int32, device=device)
            self.work = []

    def __init__(self, local_frames: int, max_det: int, cap: int, device, group=None, depth: int = 2):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cap = min(cap, max_det)
        self.bufs = [self._Buf(local_frames, max_det, self.cap, self.world, device) for _ in range(depth)]

    def buffers(self, step: int):
        b = self.bufs[step % len(self.bufs)]
        for w in b.work:            # the gather that last used this buffer must have read it (stream-side wait, host does not block)
            w.wait()
        b.work = []
        return b

    def launch(self, step: int):
        b = self.bufs[step % len(self.bufs)]
        if self.world == 1:
            return b
        b.send.copy_(b.dets[:, :self.cap])
        b.work = [dist.all_gather_into_tensor(b.all_dets, b.send, group=self.group, async_op=True),
                  dist.all_gather_into_tensor(b.all_counts, b.counts, group=self.group, async_op=True)]
        return b

    def finish(self):
        for b in self.bufs:
            for w in b.work:
                w.wait()
            b.work = []
