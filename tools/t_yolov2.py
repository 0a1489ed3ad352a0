"""Per-layer CUDA-event times of the yolo_v2 graph (batch 64 at 416x416)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, yolo_b200
from yolo_b200 import export as ex, lib
H = W = 416
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
import faulthandler, time; faulthandler.dump_traceback_later(100, exit=True)
t0 = time.time()
q = ex.random_quantnet_yolo_v2(seed=0, calib_hw=(H, W), calib_frames=1)
print("quantnet %.1f s" % (time.time() - t0), flush=True)
ctx = lib.Context(0)
ctx.load_quantnet(q, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=1024)
x = torch.from_numpy(bench.ol_quantize(ex.synthetic_frames_f32(B, H, W, seed=13).numpy(), q.sa[0])).cuda()
dets = torch.zeros((B, 1024, 8), dtype=torch.int32, device="cuda"); counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
for i in range(3): ctx.forward_int8_dev(x, B, H, W, dets, counts)
ctx.sync(); print("warm-up done %.1f s" % (time.time() - t0), flush=True)
ctx.enable_timing(True)
acc = None
for i in range(5):
    ctx.forward_int8_dev(x, B, H, W, dets, counts)
    t = np.array(ctx.layer_times_ms())
    acc = t if acc is None else acc + t
acc /= 5
dims = bench.yolo_v2_maps(q, H, W)
tot = 0
for l, ((ci, co, a, p), g, (h, w)) in enumerate(zip(q.layers, q.graph, dims)):
    mac = h * w * g["ksize"] ** 2 * ci * co * B
    print("layer %2d %4d->%4d k%d %3dx%-3d pool %d: %.4f ms  %.0f TOPS" % (l, ci, co, g["ksize"], h, w, p, acc[l], 2 * mac / (acc[l] * 1e-3) / 1e12))
    tot += acc[l]
print("head %.4f ms; total %.3f ms" % (acc[len(q.layers)], acc.sum()))
