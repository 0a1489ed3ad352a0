import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench, yolo_b200
from yolo_b200 import export as ex, lib
B, H, W = 256, 416, 416
qnet = bench.make_qnet()
ctx = lib.Context(0)
ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
d = torch.from_numpy(ex.synthetic_frames_rgb444(B, H, W, seed=0).view(np.int16)).cuda()
dq = torch.empty((B, H, W, 4), dtype=torch.int8, device="cuda")
dets = torch.zeros((B, 4096, 8), dtype=torch.int32, device="cuda"); counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
ctx.quantize_rgb444(d, B, H, W, dq)
ctx.enable_timing(True)
for name, fn in (("int8", lambda: ctx.forward_int8_dev(dq, B, H, W, dets, counts)), ("rgb444 fused", lambda: ctx.forward_rgb444_dev(d, B, H, W, dets, counts))):
    acc = []
    for i in range(6):
        fn(); t = ctx.layer_times_ms(); acc.append(t[0])
    print(name, "conv1 ms:", np.round(acc, 4))
