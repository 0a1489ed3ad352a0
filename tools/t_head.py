#!/usr/bin/env python
"""Times decode + NMS alone (CUDA events) on the bench's dense prediction maps, for several batch sizes; with
YOLO_B200_DBG=1 loads the -DYB_NMS_TIMELINE build (tools/dbg_build.sh) and prints the per-phase cycle counts of frames 0/1."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import yolo_b200  # noqa
from yolo_b200 import export as ex, lib

if os.environ.get("YOLO_B200_LIB"):
    lib._lib = lib.load_library(os.environ["YOLO_B200_LIB"])
elif os.environ.get("YOLO_B200_DBG"):
    lib._lib = lib.load_library(os.path.join(ROOT, "yolo-compression-and-deployment-in-fpga_b200", "build_dbg", "libyolo_b200_dbg.so"))
H = W = 416
sparse = float(os.environ.get("HEAD_BIAS", "0"))
qnet = ex.random_quantnet(seed=0, calib_hw=(H, W), calib_frames=2, calib_input="rgb444", head_bias_shift=sparse)
ctx = lib.Context(0)
ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=bench.CONF, nms_thresh=bench.NMS, max_det=4096)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
for B in [int(a) for a in sys.argv[1:]] or [256, 148, 64, 1]:
    d = torch.from_numpy(ex.synthetic_frames_rgb444(B, H, W, seed=0).view(np.int16)).cuda()
    dets = torch.zeros((B, 4096, 8), dtype=torch.int32, device="cuda")
    counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    ctx.forward_rgb444_dev(d, B, H, W, dets, counts)
    ctx.sync()
    pred = ctx.L  # noqa
    # the prediction map of the last forward stays in the context: run the head on it repeatedly
    import ctypes as C
    pred_ptr = C.c_void_p(); gh = C.c_int(); gw = C.c_int()
    x8 = torch.empty((B, H, W, 4), dtype=torch.int8, device="cuda")
    ctx.quantize_rgb444(d, B, H, W, x8)
    p, g1, g2 = ctx.backbone(x8, B, H, W)
    reps = 1 if os.environ.get("YOLO_B200_DBG") else 10
    for i in range(2 if reps > 1 else 0):
        ctx.detect(p, B, g1, g2, H, W, dets, counts)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(reps):
        ctx.detect(p, B, g1, g2, H, W, dets, counts)
    e1.record(stream)
    torch.cuda.synchronize()
    print("batch %d: decode + NMS %.4f ms, mean detections %.1f" % (B, e0.elapsed_time(e1) / reps, counts.float().mean().item()), flush=True)
