#!/usr/bin/env python
"""Minimal driver for ncu: N forward passes of the bench workload (batch x 416x416 RGB444 frames), nothing else.
usage: profile_step.py [steps] [batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload definition shared with the bench)
import yolo_b200  # noqa
from yolo_b200 import export as ex, lib

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
H = W = 416
qnet = bench.make_qnet()
ctx = lib.Context(0)
ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=bench.CONF, nms_thresh=bench.NMS, max_det=4096)
d = torch.from_numpy(ex.synthetic_frames_rgb444(batch, H, W, seed=0).view(np.int16)).cuda()
dets = torch.zeros((batch, 4096, 8), dtype=torch.int32, device="cuda")
counts = torch.zeros((batch,), dtype=torch.int32, device="cuda")
for _ in range(steps):
    ctx.forward_rgb444_dev(d, batch, H, W, dets, counts)
ctx.sync()
print("done", counts[:4].tolist())
