"""Per-kernel CUDA-event times of the device-resident forward pass at several batch sizes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, yolo_b200
from yolo_b200 import export as ex, lib
H = W = 416
ctx = lib.Context(0)
ctx.load_quantnet(bench.make_qnet(), contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
names = ["conv1", "conv2", "conv3_1", "conv3_2", "conv4_1", "conv4_2", "conv5", "conv6", "conv7", "pred", "head"]
for B in [int(a) for a in sys.argv[1:]] or [256, 128, 64, 32]:
    d = torch.from_numpy(ex.synthetic_frames_rgb444(B, H, W, seed=0).view(np.int16)).cuda()
    dets = torch.zeros((B, 4096, 8), dtype=torch.int32, device="cuda"); counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    for i in range(3): ctx.forward_rgb444_dev(d, B, H, W, dets, counts)
    ctx.enable_timing(True)
    acc = None
    for i in range(10):
        ctx.forward_rgb444_dev(d, B, H, W, dets, counts)
        t = np.array(ctx.layer_times_ms())
        acc = t if acc is None else acc + t
    ctx.enable_timing(False)
    acc /= 10
    print("batch", B, "total %.3f ms" % acc.sum(), "us/frame %.2f" % (acc.sum() / B * 1e3), {n: round(float(v), 4) for n, v in zip(names, acc)}, flush=True)
