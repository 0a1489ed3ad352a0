"""Drop-in for the reference's `models/slim_yolo_v2.py::SlimYOLOv2_quantize_bnfuse` (`-v slim_yolo_v2_q_bf`,
test.py:165-172, demo.py:192-197) whose fixed-point inference runs on the hand-written CUDA kernels behind
include/yolo_b200.h.

Same constructor, same `forward(x, target=None, quantization=False, find=False)` -> `(bboxes, scores, cls_inds)`,
same attributes (`input_size`, `stride`, `trainable`, `set_grid`) and the same 42 state_dict keys
(`conv*.convs.0.{weight,bias}`, `pred.{weight,bias}`, `a_tracker_*.{scale,first_a}`), so
`net.load_state_dict(torch.load(...))` works on reference checkpoints.

Which path runs:
  * `quantization=True`, inference  -> the fixed-point hot path (contract P) on the GPU through the C-ABI, with or without
    `find=True` (the overflow probe + the /2**k rescaling of retune_bias_quantize_findbest.py:364 run on the GPU too).
    This is what `utils/vocapi_evaluator_mask.py:69` calls.  Requires CUDA; there is no CPU fallback.  Fresh trackers
    (first_a == 0) are calibrated on the first batch by the GPU tracker pass, as the reference's first call does.
  * `quantization=False`, inference -> activations are NOT quantised in the reference (slim_yolo_v2.py:18-19), i.e.
    a plain float network: stock PyTorch ops + the same head, outside the fixed-point path.
  * `trainable=True` -> training is out of scope; raises.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import export as ex


class AveragedRangeTracker(nn.Module):
    """Buffers of slim_yolo_v2.py:9-15; the arithmetic lives in the kernels' epilogue."""

    def __init__(self, momentum=0.1):
        super().__init__()
        self.momentum = momentum
        self.register_buffer("scale", torch.zeros(1))
        self.register_buffer("first_a", torch.zeros(1))


class Conv2d_fuse(nn.Module):
    """conv(bias) + LeakyReLU(0.125), utils/modules.py:20-29 (kept for its state_dict key layout)."""

    def __init__(self, in_channels, out_channels, ksize, padding=0, stride=1, dilation=1, leakyReLU=False):
        super().__init__()
        self.convs = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, ksize, stride=stride, padding=padding, dilation=dilation),
            nn.LeakyReLU(0.125, inplace=True) if leakyReLU else nn.ReLU(inplace=True))

    def forward(self, x):
        return self.convs(x)


class SlimYOLOv2_quantize_bnfuse(nn.Module):
    def __init__(self, device, input_size=None, num_classes=20, trainable=False, conf_thresh=0.01, nms_thresh=0.5,
                 anchor_size=None, hr=False):
        super().__init__()
        self.device = device
        self.input_size = input_size
        self.num_classes = num_classes
        self.trainable = trainable
        self.conf_thresh = conf_thresh
        self.nms_thresh = nms_thresh
        self.anchor_size = torch.tensor(anchor_size)
        self.anchor_number = len(anchor_size)
        self.stride = 16
        self.set_grid(input_size)

        self.a_tracker_in = AveragedRangeTracker()
        self.conv1 = Conv2d_fuse(3, 16, 3, 1, leakyReLU=True)
        self.a_tracker1 = AveragedRangeTracker()
        self.pool1 = nn.MaxPool2d(2, 2)
        self.conv2 = Conv2d_fuse(16, 32, 3, 1, leakyReLU=True)
        self.a_tracker2 = AveragedRangeTracker()
        self.pool2 = nn.MaxPool2d(2, 2)
        self.conv3_1 = Conv2d_fuse(32, 64, 3, 1, leakyReLU=True)
        self.a_tracker3_1 = AveragedRangeTracker()
        self.conv3_2 = Conv2d_fuse(64, 64, 3, 1, leakyReLU=True)
        self.a_tracker3_2 = AveragedRangeTracker()
        self.pool3 = nn.MaxPool2d(2, 2)
        self.conv4_1 = Conv2d_fuse(64, 128, 3, 1, leakyReLU=True)
        self.a_tracker4_1 = AveragedRangeTracker()
        self.conv4_2 = Conv2d_fuse(128, 128, 3, 1, leakyReLU=True)
        self.a_tracker4_2 = AveragedRangeTracker()
        self.pool4 = nn.MaxPool2d(2, 2)
        self.conv5 = Conv2d_fuse(128, 256, 3, 1, leakyReLU=True)
        self.a_tracker5 = AveragedRangeTracker()
        self.conv6 = Conv2d_fuse(256, 256, 3, 1, leakyReLU=True)
        self.a_tracker6 = AveragedRangeTracker()
        self.conv7 = Conv2d_fuse(256, 256, 3, 1, leakyReLU=True)
        self.a_tracker7 = AveragedRangeTracker()
        self.pred = nn.Conv2d(256, self.anchor_number * (1 + 4 + self.num_classes), 3, 1, padding=1)
        self.a_tracker_pred = AveragedRangeTracker()

        self._ctx = None
        self._ctx_key = None
        self.last_overflow = 0

    # -- reference API ---------------------------------------------------------------------------------
    def set_grid(self, input_size):
        """slim_yolo_v2.py:105-109: only the input size matters here; grid/anchor tensors live in the head kernel."""
        self.input_size = input_size
        self.scale = np.array([[[input_size[1], input_size[0], input_size[1], input_size[0]]]])

    def _convs(self):
        return [self.conv1.convs[0], self.conv2.convs[0], self.conv3_1.convs[0], self.conv3_2.convs[0],
                self.conv4_1.convs[0], self.conv4_2.convs[0], self.conv5.convs[0], self.conv6.convs[0],
                self.conv7.convs[0], self.pred]

    def _trackers(self):
        return [getattr(self, k) for k in ex.SLIM_TRACKER_KEYS]

    def quantnet(self, calib_frames=None) -> ex.QuantNet:
        """Export the current parameters to the fixed-point form (reference rule, retune_bias_quantize.py:73-119)."""
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        return ex.quantnet_from_state_dict(sd, calib_frames=calib_frames, anchors=self.anchor_size.tolist(),
                                           num_classes=self.num_classes)

    # the /2**k constants the reference hard-codes in its find branch, slim_yolo_v2.py:227,240,...,327
    FIND_SHIFTS = (11, 10, 10, 11, 11, 10, 11, 11, 11, 10)

    def _load(self, ctx, find):
        """(Re)program the context from the module's parameters.  find=True: the reference divides every layer's output by
        2**k before its tracker (slim_yolo_v2.py:222-227 ... :327); y / 2**k with y = acc * 2^-(sa_i + sw) + b * 2^-sb is the
        same layer with BOTH exponents raised by k (the leaky-ReLU is positively homogeneous), so the find branch is the
        fake-quant contract with shifted weight / bias exponent tables: no separate kernel path."""
        from . import lib
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        fresh = any(float(t.first_a) == 0 for t in self._trackers())
        if fresh:                                    # exponents are about to be calibrated on the GPU: placeholders
            for k in ex.SLIM_TRACKER_KEYS:
                sd[k + ".scale"] = torch.ones(1); sd[k + ".first_a"] = torch.ones(1)
        qnet = ex.quantnet_from_state_dict(sd, anchors=self.anchor_size.tolist(), num_classes=self.num_classes)
        if find:
            qnet.sw = [e + k for e, k in zip(qnet.sw, self.FIND_SHIFTS)]
            qnet.sb = [e + k for e, k in zip(qnet.sb, self.FIND_SHIFTS)]
        ctx.load_quantnet(qnet, contract=lib.CONTRACT_P, head_mode=lib.HEAD_PYTHON,
                          conf_thresh=float(self.conf_thresh), nms_thresh=float(self.nms_thresh), max_det=4096)

    def update_trackers(self, x, find=False, momentum=None):
        """One calibration batch on the GPU (yolo_b200_update_trackers_f32): fresh trackers take scale = 127 / max|a| (the
        reference's first call, slim_yolo_v2.py:25-27); trackers that have been called before move by the exponential
        average of :31 (what the reference does while `trainable`).  The `scale` / `first_a` buffers end up with the values
        the reference's buffers would hold."""
        from . import lib
        if not (torch.cuda.is_available() and x.is_cuda):
            raise lib.YoloB200Error("the trackers are calibrated by CUDA kernels; input is on %s and there is no CPU fallback" % x.device)
        if self._ctx is None:
            self._ctx = lib.Context(x.device.index or 0)
        self._load(self._ctx, find)
        trackers = self._trackers()
        scales = np.array([float(t.scale) if float(t.first_a) != 0 else 0.0 for t in trackers], dtype=np.float32)
        n, c, h, w = x.shape
        self._ctx.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
        self._ctx.update_trackers_f32(x.contiguous().float(), n, h, w, self.a_tracker_in.momentum if momentum is None else momentum, scales)
        for t, sc in zip(trackers, scales):
            t.scale.fill_(float(sc))
            t.first_a.fill_(1)
        self._ctx_key = None

    def _context(self, x, find=False):
        from . import lib
        if not (torch.cuda.is_available() and x.is_cuda):
            raise lib.YoloB200Error("quantization=True runs the fixed-point path on CUDA kernels; input is on %s and "
                                    "there is no CPU fallback" % x.device)
        trackers = self._trackers()
        if any(float(t.first_a) == 0 for t in trackers):
            # first call with fresh trackers: the reference calibrates on this batch (slim_yolo_v2.py:25-27)
            self.update_trackers(x, find)
        key = (tuple(int(p._version) for p in self.parameters()), tuple(float(t.scale) for t in trackers),
               float(self.conf_thresh), float(self.nms_thresh), x.device.index or 0, bool(find))
        if self._ctx is None or self._ctx_key != key:
            if self._ctx is None:
                self._ctx = lib.Context(x.device.index or 0)
            self._load(self._ctx, find)
            self._ctx_key = key
        return self._ctx

    def forward_batch(self, x, find=False, frames=None):
        """Fixed-point inference for a whole batch: list of (bboxes, scores, cls_inds) per frame.
        (The reference's head only looks at batch element 0, slim_yolo_v2.py:348-350.)"""
        from . import lib
        ctx = self._context(x, find)
        n, c, h, w = x.shape
        x = x.contiguous().float()
        ctx.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
        if find:
            # the overflow probe of the find branch (slim_yolo_v2.py:222-226 ... :322-326): every layer's output, BEFORE the
            # division by 2**k, must stay below the 16-bit accumulator range
            mx = ctx.measure_f32(x, n, h, w)
            for l, k in enumerate(self.FIND_SHIFTS):
                if mx[l + 1] * 2.0 ** k >= 2 ** (16 - 1):
                    print("too high!!!")
                    print(mx[l + 1] * 2.0 ** k)
                    raise AssertionError("layer %d: max |output| %g exceeds the 16-bit accumulator (slim_yolo_v2.py:222-226)"
                                         % (l, mx[l + 1] * 2.0 ** k))
        if frames is not None:                       # trackers and the probe saw the whole batch; decode only these frames
            x = x[:frames].contiguous()
            n = x.shape[0]
        dets = torch.empty((n, ctx.params.max_det, 8), dtype=torch.int32, device=x.device)
        counts = torch.empty((n,), dtype=torch.int32, device=x.device)
        ctx.forward_f32_dev(x, n, h, w, dets, counts)
        counts_h = counts.cpu().numpy()
        dets_h = dets.cpu().numpy().view(lib.DET_DTYPE).reshape(n, ctx.params.max_det)
        self.last_overflow = ctx.overflow_count()
        out = []
        for i in range(n):
            b, s, cl, _ = lib.dets_to_arrays(dets_h[i], int(min(counts_h[i], ctx.params.max_det)))
            out.append((b, s, cl))
        return out

    def forward(self, x, target=None, quantization=False, find=False):
        if self.trainable:
            raise NotImplementedError("training (slim_yolo_v2.py:360-382) is outside this library's scope")
        if quantization:
            # (the reference's head only decodes batch element 0, slim_yolo_v2.py:348-350, but its trackers see the whole batch)
            return self.forward_batch(x, find, frames=1)[0]
        return self._forward_float(x, find)

    # -- float path (quantization=False): stock PyTorch, outside the fixed-point hot path -------------------
    def _forward_float(self, x, find):
        retune = ex.SHIPPED_RETUNE     # the /2**k constants hard-coded at slim_yolo_v2.py:227...327
        with torch.no_grad():
            y = x
            for l, (conv, (cin, cout, activ, pool)) in enumerate(zip(self._convs(), ex.SLIM_YOLO_V2_LAYERS)):
                y = conv(y)
                if activ:
                    y = F.leaky_relu(y, 0.125)
                if find:
                    if y.abs().max() >= 2 ** 15:
                        raise AssertionError("layer %d exceeds the 16-bit accumulator (slim_yolo_v2.py:222-226)" % l)
                    y = y / 2 ** retune[l]
                if pool:
                    y = F.max_pool2d(y, 2, 2)
            B, abC, H, W = y.shape
            A, Cn = self.anchor_number, self.num_classes
            p = y.permute(0, 2, 3, 1).reshape(B, H * W, abC)[0]
            obj = torch.sigmoid(p[:, :A].reshape(-1, 1))
            cls = torch.softmax(p[:, A:(1 + Cn) * A].reshape(-1, Cn), 1) * obj
            t = p[:, (1 + Cn) * A:].reshape(H * W, A, 4)
            gy, gx = torch.meshgrid(torch.arange(H, device=y.device), torch.arange(W, device=y.device), indexing="ij")
            grid = torch.stack([gx, gy], -1).float().reshape(H * W, 1, 2)
            xy = (torch.sigmoid(t[..., :2]) + grid) * self.stride
            wh = torch.exp(t[..., 2:]) * self.anchor_size.to(y.device).float() * self.stride
            box = torch.cat([xy - wh / 2, xy + wh / 2], -1).reshape(-1, 4)
            box = torch.clamp(box / torch.tensor(self.scale[0, 0], device=y.device).float(), 0., 1.)
            return postprocess(box.cpu().numpy(), cls.cpu().numpy(), self.conf_thresh, self.nms_thresh, Cn)


def postprocess(boxes, probs, conf_thresh, nms_thresh, num_classes):
    """Host restatement of postprocess/nms (slim_yolo_v2.py:145-210) for the float path; ties in the score sort go to
    the higher anchor index (the rule the CUDA head uses)."""
    cls = np.argmax(probs, axis=1)
    scores = probs[np.arange(len(cls)), cls]
    sel = np.where(scores >= np.float32(conf_thresh))[0]
    boxes, scores, cls = boxes[sel], scores[sel], cls[sel]
    keep = np.zeros(len(sel), dtype=bool)
    for c in range(num_classes):
        inds = np.where(cls == c)[0]
        order = inds[np.argsort(scores[inds], kind="stable")[::-1]]
        while order.size:
            i, rest = order[0], order[1:]
            keep[i] = True
            w = np.maximum(1e-28, np.minimum(boxes[i, 2], boxes[rest, 2]) - np.maximum(boxes[i, 0], boxes[rest, 0]))
            h = np.maximum(1e-28, np.minimum(boxes[i, 3], boxes[rest, 3]) - np.maximum(boxes[i, 1], boxes[rest, 1]))
            inter = w * h
            area = lambda b: (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
            ovr = inter / (area(boxes[i]) + area(boxes[rest]) - inter)
            order = rest[ovr <= np.float32(nms_thresh)]
    return boxes[keep], scores[keep], cls[keep]
