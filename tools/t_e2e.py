"""Host path (pinned host frames in, detections out): blocking call and the submit / wait stream, by chunk size."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench, yolo_b200
from yolo_b200 import export as ex, lib
B, H, W = 256, 416, 416
qnet = bench.make_qnet()
ctx = lib.Context(0)
ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
hs = [torch.from_numpy(ex.synthetic_frames_rgb444(B, H, W, seed=s).view(np.int16)).pin_memory() for s in range(3)]
outs = [(torch.zeros((B, 4096, 8), dtype=torch.int32).pin_memory(), torch.zeros((B,), dtype=torch.int32).pin_memory()) for _ in range(2)]
L = ctx.L
def step(i):
    rc = L.yolo_b200_forward_rgb444(ctx._h, hs[i % 3].data_ptr(), B, H, W, outs[0][0].data_ptr(), outs[0][1].data_ptr())
    assert rc == 0
def stream(k):
    tickets = []
    for i in range(k):
        if len(tickets) == 2:
            ctx.wait(tickets.pop(0))
        t = L.yolo_b200_submit_rgb444(ctx._h, hs[i % 3].data_ptr(), B, H, W, outs[i % 2][0].data_ptr(), outs[i % 2][1].data_ptr())
        assert t >= 0
        tickets.append(t)
    while tickets:
        ctx.wait(tickets.pop(0))
for chunk in (0, 32, 64, 128, 64, 128, 0):
    ctx.set_host_chunk(chunk)
    for i in range(3): step(i)
    ts, tp = [], []
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        for i in range(10): step(i)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t) / 10 * 1e3)
    stream(4)
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        stream(10)
        torch.cuda.synchronize(); tp.append((time.perf_counter() - t) / 10 * 1e3)
    print("chunk", chunk, "blocking ms/step", np.round(ts, 3), "fps", int(B / min(ts) * 1e3), "| submit/wait ms/step", np.round(tp, 3), "fps", int(B / min(tp) * 1e3), flush=True)
