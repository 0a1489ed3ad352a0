"""Where does the host spend its time in the streamed host path?  (submit = queueing one call, wait = blocking for its results)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, yolo_b200
from yolo_b200 import export as ex, lib
B, H, W = 256, 416, 416
rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(rank)
ctx = lib.Context(rank)
ctx.load_quantnet(bench.make_qnet(), contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
hs = [torch.from_numpy(ex.synthetic_frames_rgb444(B, H, W, seed=s).view(np.int16)).pin_memory() for s in range(3)]
outs = [(torch.zeros((B, 4096, 8), dtype=torch.int32).pin_memory(), torch.zeros((B,), dtype=torch.int32).pin_memory()) for _ in range(2)]
L = ctx.L
for chunk in (64, 128, 0):
    ctx.set_host_chunk(chunk)
    for rep in range(2):
        tickets, ts, tw = [], 0.0, 0.0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(12):
            if len(tickets) == 2:
                a = time.perf_counter(); ctx.wait(tickets.pop(0)); tw += time.perf_counter() - a
            a = time.perf_counter()
            t = L.yolo_b200_submit_rgb444(ctx._h, hs[i % 3].data_ptr(), B, H, W, outs[i % 2][0].data_ptr(), outs[i % 2][1].data_ptr())
            ts += time.perf_counter() - a
            assert t >= 0
            tickets.append(t)
        while tickets:
            a = time.perf_counter(); ctx.wait(tickets.pop(0)); tw += time.perf_counter() - a
        el = time.perf_counter() - t0
    print("rank %d chunk %3d: %.3f ms per step (%d frames/s); host time per step: submit %.3f ms, wait %.3f ms" % (rank, chunk, el / 12 * 1e3, int(B * 12 / el), ts / 12 * 1e3, tw / 12 * 1e3), flush=True)
