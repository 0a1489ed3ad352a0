"""Batch sharding over the GPUs of one box.

Frames are independent at inference (trackers frozen, slim_yolo_v2.py:215,28-29), so the batch is split contiguously
by frame index, weights are replicated at load, and there is NO collective on the hot path.  Only the final
per-frame detection lists are gathered (fixed-capacity [frames][max_det] records + counts), in frame order.
One process per GPU; torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).

On GPUs the lists are COLLECTED ON ONE RANK by copy-engine peer writes (PeerCollector): every rank squeezes its lists to their
filled part on the device and copies exactly those bytes into its slot of a buffer that lives on the collecting GPU
(CUDA IPC, NVLink).  No collective kernel runs, so nothing competes with the persistent convolution CTAs for SMs, and
nobody receives lists it did not ask for.  gather_detections (a padded all-gather) remains for CPU / gloo."""
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


class _DevView:
    """A raw device pointer as something torch.as_tensor understands (__cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2, "strides": None}


class PeerCollector:
    """Collects every rank's detection lists on rank 0 by copy-engine peer writes; see the module docstring.

    Per step:  buf = pc.buffers(i);  <forward writes buf.dets / buf.counts>;  pc.launch(i).   At the end: pc.finish().
    launch(i) packs step i on the compute stream and, WITHOUT stalling the host on step i, ships step i - 1 (whose record
    count has reached the host by then): the copy of step i - 1 overlaps the kernels of step i.  Rank 0 reads
    pc.collected(step) after finish() + a barrier: a list of (offsets int32 [frames + 1], records int32 [total, 8]) per rank."""

    class _Buf:
        def __init__(self, frames, max_det, cap, device):
            self.dets = torch.zeros((frames, max_det, 8), dtype=torch.int32, device=device)
            self.counts = torch.zeros((frames,), dtype=torch.int32, device=device)
            self.packed = torch.zeros((frames * cap, 8), dtype=torch.int32, device=device)
            self.offsets = torch.zeros((frames + 1,), dtype=torch.int32, device=device)
            self.total = torch.zeros((1,), dtype=torch.int32).pin_memory()
            self.ev_packed = torch.cuda.Event()
            self.ev_total = torch.cuda.Event()
            self.ev_sent = torch.cuda.Event()
            self.step = -1
            self.pending = False

    def __init__(self, ctx, frames: int, max_det: int, cap: int, device, group=None, depth: int = 3):
        self.ctx, self.frames, self.group, self.depth = ctx, frames, group, depth
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cap = min(cap, max_det)
        rec_bytes = frames * self.cap * 32
        self.off_bytes = (frames + 1) * 4
        self.slot_bytes = (rec_bytes + self.off_bytes + 255) // 256 * 256
        self.rec_bytes = rec_bytes
        import ctypes as C
        self._C = C
        total_bytes = depth * self.world * self.slot_bytes
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        if self.rank == 0:
            ctx._check(ctx.L.yolo_b200_ipc_alloc(ctx._h, total_bytes, C.byref(ptr), handle))
        box = [bytes(handle.raw)]
        if self.world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        if self.rank != 0:
            ctx._check(ctx.L.yolo_b200_ipc_open(ctx._h, box[0], C.byref(ptr)))
        self.base = int(ptr.value)
        self.stream = torch.cuda.Stream(device=device)
        self.bufs = [self._Buf(frames, max_det, self.cap, device) for _ in range(depth)]
        self.sent_bytes = 0

    def _slot(self, step, rank):
        return self.base + ((step % self.depth) * self.world + rank) * self.slot_bytes

    def buffers(self, step: int):
        b = self.bufs[step % self.depth]
        if b.pending:
            self._ship(b)
        b.ev_sent.synchronize()                # its previous contents have left
        return b

    def launch(self, step: int):
        b = self.bufs[step % self.depth]
        cur = torch.cuda.current_stream()
        self.ctx.set_stream(cur.cuda_stream)
        self.ctx._check(self.ctx.L.yolo_b200_pack_detections(self.ctx._h, b.dets.data_ptr(), b.counts.data_ptr(), self.frames,
                                                             b.packed.data_ptr(), b.offsets.data_ptr()))
        b.ev_packed.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(b.ev_packed)
            b.total.copy_(b.offsets[self.frames:], non_blocking=True)
            b.ev_total.record(self.stream)
        b.step, b.pending = step, True
        prev = self.bufs[(step - 1) % self.depth]
        if step > 0 and prev.pending and prev.step == step - 1:
            self._ship(prev)
        return b

    def _ship(self, b):
        b.ev_total.synchronize()
        total = int(b.total[0])
        L, h, st = self.ctx.L, self.ctx._h, self.stream.cuda_stream
        slot = self._slot(b.step, self.rank)
        self.ctx._check(L.yolo_b200_copy_async(h, slot, b.packed.data_ptr(), total * 32, st))
        self.ctx._check(L.yolo_b200_copy_async(h, slot + self.rec_bytes, b.offsets.data_ptr(), self.off_bytes, st))
        b.ev_sent.record(self.stream)
        b.pending = False
        self.sent_bytes += total * 32 + self.off_bytes

    def finish(self):
        for b in sorted(self.bufs, key=lambda x: x.step):
            if b.pending:
                self._ship(b)
        self.stream.synchronize()

    def collected(self, step: int):
        """Rank 0, after finish() and a barrier: [(offsets, records)] per rank for `step` (views of the collection buffer)."""
        assert self.rank == 0
        out = []
        for r in range(self.world):
            slot = self._slot(step, r)
            off = torch.as_tensor(_DevView(slot + self.rec_bytes, self.off_bytes), device="cuda").view(torch.int32)
            total = int(off[self.frames])
            rec = torch.as_tensor(_DevView(slot, max(total, 1) * 32), device="cuda").view(torch.int32).reshape(-1, 8)[:total]
            out.append((off, rec))
        return out

    def close(self):
        if self.base:
            self.stream.synchronize()
            self.ctx.L.yolo_b200_ipc_close(self.ctx._h, self.base, 0 if self.rank == 0 else 1)
            self.base = 0
