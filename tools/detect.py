#!/usr/bin/env python
"""Image detection CLI for the fixed-point slim_yolo_v2 path on the GPU: the counterpart of the reference's test.py /
demo.py image mode for `-v slim_yolo_v2_q_bf` (test.py:13-99,165-181; demo.py:99-120), with the same flags where they
apply.  Every image goes through the whole hot path on the device: the uint8 BGR image (resized on the GPU with cv2.resize semantics, as
BaseTransform does, data/__init__.py:36) is normalised, quantised, convolved, decoded and NMS-ed by libyolo_b200.so.

    python tools/detect.py --trained_model slim_yolo_v2_retune_quantize1.pth --images dir/ -size 416 --out output/
    python tools/detect.py --trained_model random --images synthetic:4          (no checkpoint / no images at hand)
    python tools/detect.py --trained_model ckpt.pth --video clip.avi --batch 32  (demo.py's video / camera modes, batched)
"""
import argparse
import glob
import os
import sys
import time

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolo_b200  # noqa: E402,F401
from yolo_b200 import export as ex, lib  # noqa: E402

MASK_CLASSES = ("face", "face_mask")                         # data/voc_mask.py:22-24
CLASS_COLORS = [(0, 0, 255), (0, 255, 0)]


def vis(img, bboxes, scores, cls_inds, thresh):
    """test.py:34-69 (dataset 'mask' branch)."""
    for i, box in enumerate(bboxes):
        if scores[i] > thresh:
            xmin, ymin, xmax, ymax = [int(v) for v in box]
            c = CLASS_COLORS[int(cls_inds[i]) % len(CLASS_COLORS)]
            cv2.rectangle(img, (xmin, ymin), (xmax, ymax), c, 1)
            cv2.rectangle(img, (xmin, abs(ymin) - 20), (xmax, ymin), c, -1)
            cv2.putText(img, MASK_CLASSES[int(cls_inds[i]) % len(MASK_CLASSES)], (xmin, ymin - 5), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 0), 1)
    return img


def run_video(ctx, args, size):
    """demo.py video / camera modes (:60-96, :123-158) with the frames batched: `--batch` frames are read, go through
    yolo_b200_forward_u8bgr_resize in one call (resize, normalise, quantise, network, decode, NMS on the GPU), are annotated
    and appended to <out>/detections.avi (MJPG, the source's frame size and rate)."""
    cap = cv2.VideoCapture(int(args.video) if args.video.isdigit() else args.video)
    if not cap.isOpened():
        raise SystemExit("cannot open video source %s" % args.video)
    os.makedirs(args.out, exist_ok=True)
    fps = cap.get(cv2.CAP_PROP_FPS) or 25.0
    writer, frames, t_dev, done = None, 0, 0.0, False
    while not done:
        batch = []
        while len(batch) < args.batch:
            ok, frame = cap.read()
            if not ok or (args.max_frames and frames + len(batch) >= args.max_frames):
                done = True
                break
            batch.append(frame)
        if not batch:
            break
        t0 = time.time()
        dets, counts = ctx.forward_u8bgr_resize(np.stack(batch), size)
        t_dev += time.time() - t0
        for k, im in enumerate(batch):
            boxes, scores, cls, _ = lib.dets_to_arrays(dets[k], int(min(counts[k], 1024)))
            h, w = im.shape[:2]
            out = vis(im, boxes * np.array([[w, h, w, h]], dtype=np.float32), scores, cls, args.visual_threshold)   # demo.py:143-149
            if writer is None:
                writer = cv2.VideoWriter(os.path.join(args.out, "detections.avi"), cv2.VideoWriter_fourcc(*"MJPG"), fps, (w, h))
            writer.write(out)
        frames += len(batch)
    cap.release()
    if writer is not None:
        writer.release()
    print("%d frames, %.1f ms in yolo_b200_forward_u8bgr_resize (%.0f frames/s incl. copies)" % (frames, 1e3 * t_dev, frames / max(t_dev, 1e-9)))


def main():
    ap = argparse.ArgumentParser(description="slim_yolo_v2 fixed-point detection on B200")
    ap.add_argument("-v", "--version", default="slim_yolo_v2_q_bf", help="only slim_yolo_v2_q_bf runs on this path")
    ap.add_argument("-size", "--input_size", default=416, type=int)
    ap.add_argument("--trained_model", default="random", help="q_bf state_dict (.pth) or 'random'")
    ap.add_argument("--conf_thresh", default=0.1, type=float)
    ap.add_argument("--nms_thresh", default=0.50, type=float)
    ap.add_argument("--visual_threshold", default=0.3, type=float)
    ap.add_argument("--cuda", action="store_true", default=True, help="kept for compatibility: this path always runs on CUDA")
    ap.add_argument("--images", default="synthetic:2", help="directory / glob of images, or synthetic:N")
    ap.add_argument("--video", default=None, help="video file (demo.py --mode video, :123-158) or camera index (--mode camera, :60-96): frames are batched, detected on the GPU and written, annotated, to <out>/detections.avi")
    ap.add_argument("--max_frames", default=0, type=int, help="stop a --video run after this many frames (0 = to the end)")
    ap.add_argument("--out", default="output")
    ap.add_argument("--batch", default=64, type=int)
    args = ap.parse_args()
    if args.version != "slim_yolo_v2_q_bf":
        raise SystemExit("only -v slim_yolo_v2_q_bf is implemented (the fixed-point hot path)")
    size = (args.input_size, args.input_size)
    if args.trained_model == "random":
        qnet = ex.random_quantnet(seed=0, calib_hw=size, calib_frames=2, head_bias_shift=-3.0)
    else:
        sd = torch.load(args.trained_model, map_location="cpu")
        qnet = ex.quantnet_from_state_dict(sd, anchors=ex.ANCHOR_SIZE_MASK, num_classes=2)     # test.py:165-172
    ctx = lib.Context(0)
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_P, head_mode=lib.HEAD_PYTHON, conf_thresh=args.conf_thresh,
                      nms_thresh=args.nms_thresh, max_det=1024)
    if args.video is not None:
        run_video(ctx, args, size)
        ctx.close()
        return
    if args.images.startswith("synthetic:"):
        rng = np.random.default_rng(0)
        imgs = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(int(args.images.split(":")[1]))]
        names = ["synthetic_%d" % i for i in range(len(imgs))]
    else:
        files = sorted(glob.glob(os.path.join(args.images, "*")) if os.path.isdir(args.images) else glob.glob(args.images))
        imgs = [cv2.imread(f, cv2.IMREAD_COLOR) for f in files]
        names = [os.path.splitext(os.path.basename(f))[0] for f in files]
        imgs, names = zip(*[(i, n) for i, n in zip(imgs, names) if i is not None]) if files else ((), ())
    os.makedirs(args.out, exist_ok=True)
    t_dev = 0.0
    for b0 in range(0, len(imgs), args.batch):
        chunk = imgs[b0:b0 + args.batch]
        # data/__init__.py:36 (cv2.resize, bilinear) runs on the GPU too: images of one size go in one call
        t0 = time.time()
        dets = np.zeros((len(chunk), 1024), dtype=lib.DET_DTYPE)
        counts = np.zeros(len(chunk), dtype=np.int32)
        shapes = sorted({im.shape[:2] for im in chunk})
        for shp in shapes:
            idx = [k for k, im in enumerate(chunk) if im.shape[:2] == shp]
            d, c_ = ctx.forward_u8bgr_resize(np.stack([chunk[k] for k in idx]), size)
            dets[idx] = d
            counts[idx] = c_
        t_dev += time.time() - t0
        for k, im in enumerate(chunk):
            boxes, scores, cls, _ = lib.dets_to_arrays(dets[k], int(min(counts[k], 1024)))
            h, w = im.shape[:2]
            boxes = boxes * np.array([[w, h, w, h]], dtype=np.float32)          # test.py:88-90
            out = vis(im.copy(), boxes, scores, cls, args.visual_threshold)
            cv2.imwrite(os.path.join(args.out, names[b0 + k] + ".jpg"), out)
            print("%s: %d detections (%d above %.2f)" % (names[b0 + k], len(scores), int((scores > args.visual_threshold).sum()), args.visual_threshold))
    print("%d images, %.1f ms in yolo_b200_forward_u8bgr_resize" % (len(imgs), 1e3 * t_dev))
    ctx.close()


if __name__ == "__main__":
    main()
