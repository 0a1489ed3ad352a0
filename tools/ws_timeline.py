#!/usr/bin/env python
"""Debug: per-tile clock64 timeline of CTA 0 of every conv_ws launch (needs the -DYB_WS_TIMELINE build in
yolo-compression-and-deployment-in-fpga_b200/build_dbg/libyolo_b200_dbg.so; see the nvcc line in DESIGN.md)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import yolo_b200  # noqa
from yolo_b200 import export as ex, lib

lib._lib = lib.load_library(os.path.join(ROOT, "yolo-compression-and-deployment-in-fpga_b200", "build_dbg", "libyolo_b200_dbg%s.so" % os.environ.get("YB_RP_EXP", "")))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = W = 416
qnet = bench.make_qnet()
ctx = lib.Context(0)
ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=bench.CONF, nms_thresh=bench.NMS, max_det=4096)
d = torch.from_numpy(ex.synthetic_frames_rgb444(batch, H, W, seed=0).view(np.int16)).cuda()
dets = torch.zeros((batch, 4096, 8), dtype=torch.int32, device="cuda")
counts = torch.zeros((batch,), dtype=torch.int32, device="cuda")
for i in range(2):
    print("==== pass", i, flush=True)
    ctx.forward_rgb444_dev(d, batch, H, W, dets, counts)
    ctx.sync()
