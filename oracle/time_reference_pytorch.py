#!/usr/bin/env python
"""oracle/time_reference_pytorch.py — ORACLE tooling (test / measurement infrastructure).  Times the UNMODIFIED reference
module `models/slim_yolo_v2.py::SlimYOLOv2_quantize_bnfuse.forward(x, quantization=True)` on this machine's CPU cores
(SURVEY.md 8d "CPU baselines": batch 1 — its head only handles batch element 0, :348-350 —, 3 warm-up + 10 timed, 416x416 and
240x320, 1 thread and all threads) and writes profiles/ref_pytorch_cpu_r2.json, which bench.py carries in its JSON line as the
labelled `cpu_baseline_pytorch` entry.  The reference cannot travel to the GPU box (/root/reference does not exist there), so
this number is taken in the build container; cores, CPU model and torch version are recorded beside it.

    PYTHONDONTWRITEBYTECODE=1 python oracle/time_reference_pytorch.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as gg  # noqa: E402  (import shims for pycocotools / np.int, nothing of the reference is edited)

ROOT = gg.ROOT


def main():
    refmod, quantize_tensor, quantize_tensor_b = gg.import_reference()
    ex = gg.load_pkg().export
    out = {"what": "reference SlimYOLOv2_quantize_bnfuse.forward(x, quantization=True), random-init weights quantised by the "
                   "reference rule, batch 1, conf 0.1 / nms 0.5 (dense detections: NumPy NMS included), 3 warm-up + 10 timed",
           "where": "build container (the reference does not exist on the GPU box)", "torch": torch.__version__,
           "cpu": next((l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?"),
           "logical_cpus": os.cpu_count(), "runs": []}
    for (H, W) in ((416, 416), (240, 320)):
        torch.manual_seed(0)
        net = refmod.SlimYOLOv2_quantize_bnfuse("cpu", input_size=[H, W], num_classes=2, trainable=False, conf_thresh=0.1,
                                                 nms_thresh=0.5, anchor_size=ex.ANCHOR_SIZE_MASK).eval()
        with torch.no_grad():
            for c in [net.conv1.convs[0], net.conv2.convs[0], net.conv3_1.convs[0], net.conv3_2.convs[0], net.conv4_1.convs[0],
                      net.conv4_2.convs[0], net.conv5.convs[0], net.conv6.convs[0], net.conv7.convs[0], net.pred]:
                qw, s_w = quantize_tensor(c.weight.detach().clone(), 8, False)
                qb, s_b = quantize_tensor_b(c.bias.detach().clone(), 8, False)
                c.weight[...] = qw / s_w
                c.bias[...] = qb / s_b
            net(ex.synthetic_frames_f32(2, H, W, seed=1000), quantization=True)          # calibration call
        x = ex.synthetic_frames_f32(1, H, W, seed=2000)
        for threads in (1, os.cpu_count()):
            torch.set_num_threads(threads)
            with torch.no_grad():
                for _ in range(3):
                    net(x, quantization=True)
                t0 = time.perf_counter()
                for _ in range(10):
                    b, s, c = net(x, quantization=True)
                dt = (time.perf_counter() - t0) / 10
            out["runs"].append({"h": H, "w": W, "threads": threads, "ms_per_frame": dt * 1e3, "frames_per_s": 1.0 / dt, "detections": int(len(s))})
            print(out["runs"][-1], flush=True)
    with open(os.path.join(ROOT, "profiles", "ref_pytorch_cpu_r2.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
