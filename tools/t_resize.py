#!/usr/bin/env python
"""Times yolo_b200_resize_u8bgr alone (CUDA events) on B frames of SHxSW -> HxW; under ncu it is the capture target:
    ncu --set full --clock-control none -k regex:resize -s 3 -c 1 -f -o gpurun_out/resize python tools/t_resize.py 64
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yolo_b200  # noqa: E402,F401
from yolo_b200 import lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
SH, SW, H, W = 480, 640, 416, 416
ctx = lib.Context(0)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ctx.set_stream(s.cuda_stream)                       # events below are recorded on the stream the kernel runs on
src = torch.randint(0, 256, (B, SH, SW, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty((B, H, W, 3), dtype=torch.uint8, device="cuda")
for _ in range(4):
    ctx.resize_u8bgr(src, B, SH, SW, dst, H, W)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ctx.resize_u8bgr(src, B, SH, SW, dst, H, W)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
nbytes = B * (SH * SW * 3 + H * W * 3)
print("resize %d x %dx%d -> %dx%d: %.4f ms, %.0f GB/s algorithmic" % (B, SH, SW, H, W, ms, nbytes / ms / 1e6))
ctx.close()
