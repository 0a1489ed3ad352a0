#!/usr/bin/env python
"""Minimal driver for ncu: N forward passes of the bench workload (batch x 416x416 int8), nothing else."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo_b200  # noqa
from yolo_b200 import export as ex, lib

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
H = W = 416
qnet = ex.random_quantnet(seed=0, calib_hw=(H, W), calib_frames=2)
ctx = lib.Context(0)
ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
rng = np.random.default_rng(0)
img = rng.integers(0, 256, (batch, H, W, 3), dtype=np.uint8).astype(np.float32)
img /= 255.
img -= np.array((0.406, 0.456, 0.485), dtype=np.float32)
img /= np.array((0.225, 0.224, 0.229), dtype=np.float32)
q = np.clip(np.rint(img[..., ::-1] * (2.0 ** qnet.sa[0])), -128, 127).astype(np.int8)
x = torch.zeros((batch, H, W, 4), dtype=torch.int8)
x[..., :3] = torch.from_numpy(q)
d = x.cuda()
dets = torch.zeros((batch, 4096, 8), dtype=torch.int32, device="cuda")
counts = torch.zeros((batch,), dtype=torch.int32, device="cuda")
for _ in range(steps):
    ctx.forward_int8_dev(d, batch, H, W, dets, counts)
ctx.sync()
print("done", counts[:4].tolist())
