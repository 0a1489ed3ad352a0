"""Host-side cost of queueing one forward pass (CPU time per call, GPU kept idle-bound with a tiny batch)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, yolo_b200
from yolo_b200 import export as ex, lib
H = W = 416
ctx = lib.Context(0)
ctx.load_quantnet(bench.make_qnet(), contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
for B in (1, 64):
    d = torch.from_numpy(ex.synthetic_frames_rgb444(B, H, W, seed=0).view(np.int16)).cuda()
    dets = torch.zeros((B, 4096, 8), dtype=torch.int32, device="cuda"); counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    for i in range(5): ctx.forward_rgb444_dev(d, B, H, W, dets, counts)
    ctx.sync()
    t0 = time.perf_counter()
    for i in range(50): ctx.forward_rgb444_dev(d, B, H, W, dets, counts)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    print("batch %d: %.1f us of host time per queued forward pass (12 launches); %.1f us per pass until the GPU is done" % (B, (t1 - t0) / 50 * 1e6, (t2 - t0) / 50 * 1e6), flush=True)
