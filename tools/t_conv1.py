"""First-layer kernel alone (int8 NHWC4 input through the single-layer entry point): time per call at several batch sizes.
usage: t_conv1.py [slim|yolov2] batch..."""
import os, sys, time, faulthandler
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, yolo_b200
from yolo_b200 import export as ex, lib
faulthandler.dump_traceback_later(50, exit=True)
if os.environ.get("YB_WATCH"):
    lib._lib = lib.load_library(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "yolo-compression-and-deployment-in-fpga_b200", "build_dbg", "libyolo_b200_watch.so"))
H = W = 416
which = sys.argv[1] if len(sys.argv) > 1 else "slim"
q = ex.random_quantnet_yolo_v2(seed=0, calib_hw=(H, W), calib_frames=1) if which == "yolov2" else bench.make_qnet()
co = q.layers[0][1]
ctx = lib.Context(0)
ctx.load_quantnet(q, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=1024)
for B in [int(a) for a in sys.argv[2:]] or [64]:
    x = torch.randint(-128, 128, (B, H, W, 4), dtype=torch.int8, device="cuda"); x[..., 3] = 0
    out = torch.zeros((B, H // 2, W // 2, co), dtype=torch.int8, device="cuda")
    print(which, "batch", B, "...", flush=True)
    for i in range(2): ctx.conv_layer(0, x, B, H, W, out)
    ctx.sync()
    ctx.sync(); t0 = time.perf_counter()
    for i in range(20): ctx.conv_layer(0, x, B, H, W, out)
    ctx.sync()
    print(which, "batch", B, "%.4f ms per call (host clock around 20 calls on the library's stream)" % ((time.perf_counter() - t0) / 20 * 1e3), flush=True)
