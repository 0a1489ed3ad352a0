"""e2e time of yolo_b200_forward_rgb444 for explicit chunk schedules (YOLO_B200_CHUNKS experiment hook)."""
import os, sys, time, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench, yolo_b200
    from yolo_b200 import export as ex, lib
    B, H, W = 256, 416, 416
    ctx = lib.Context(0)
    ctx.load_quantnet(bench.make_qnet(), contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
    hs = [torch.from_numpy(ex.synthetic_frames_rgb444(B, H, W, seed=s).view(np.int16)).pin_memory() for s in range(3)]
    hd = torch.zeros((B, 4096, 8), dtype=torch.int32).pin_memory(); hc = torch.zeros((B,), dtype=torch.int32).pin_memory()
    def step(i): assert ctx.L.yolo_b200_forward_rgb444(ctx._h, hs[i % 3].data_ptr(), B, H, W, hd.data_ptr(), hc.data_ptr()) == 0
    for i in range(3): step(i)
    ts = []
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        for i in range(10): step(i)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t) / 10 * 1e3)
    print(os.environ.get("YOLO_B200_CHUNKS"), "ms/step %.3f" % min(ts), "fps", int(B / min(ts) * 1e3), flush=True)
else:
    for sched in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, YOLO_B200_CHUNKS=sched))
