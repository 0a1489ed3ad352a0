"""Timeline of one yolo_b200_forward_rgb444 call (YOLO_B200_TRACE_HOST=1 makes the library print it)."""
import os, sys
os.environ["YOLO_B200_TRACE_HOST"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, yolo_b200
from yolo_b200 import export as ex, lib
B, H, W = 256, 416, 416
ctx = lib.Context(0)
ctx.load_quantnet(bench.make_qnet(), contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
hs = torch.from_numpy(ex.synthetic_frames_rgb444(B, H, W, seed=0).view(np.int16)).pin_memory()
hd = torch.zeros((B, 4096, 8), dtype=torch.int32).pin_memory(); hc = torch.zeros((B,), dtype=torch.int32).pin_memory()
if len(sys.argv) > 1: ctx.set_host_chunk(int(sys.argv[1]))
for i in range(4):
    print("call", i, file=sys.stderr, flush=True)
    assert ctx.L.yolo_b200_forward_rgb444(ctx._h, hs.data_ptr(), B, H, W, hd.data_ptr(), hc.data_ptr()) == 0
