"""Checkpoint -> fixed-point network exporter (host-side logic, no GPU needed).

The reference has NO exporter: `c_embedding/weight.h` is produced by an unpublished step
(`.MISSING_LARGE_BLOBS:1`).  This module is that missing link, following the reference's own rules:

* BatchNorm folding ................. conv+bn2conv.py:126-150 (`fuse_conv_and_bn`), generalised to nested modules
  (`fold_bn_state_dict`; the reference's loop :317-326 only visits top-level children).
* weight / bias quantisation ........ retune_bias_quantize.py:73-97 (`quantize_tensor`, `quantize_tensor_b`):
  per-tensor ``s = 2**floor(log2(127 / max|t|))``, ``q = round(s * t)`` (torch.round = half-to-even).
* activation scale calibration ...... models/slim_yolo_v2.py:16-38 (`AveragedRangeTracker`, first-call rule).
* accumulator scale ``retune`` ...... models/slim_yolo_v2.py:222-227 (largest r with max|y| * 2**r < 2**15).
* layer list ........................ models/slim_yolo_v2.py:58-87, c_embedding/yolo_forward.c:1202-1262.
* weight.h burst order .............. c_embedding/yolo_forward.c:165-173,696-701 (inferred; header is missing).
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# (cin, cout, activ, pool) — yolo_forward.c:1202-1262 / slim_yolo_v2.py:58-87
SLIM_YOLO_V2_LAYERS = [
    (3, 16, 1, 1), (16, 32, 1, 1), (32, 64, 1, 0), (64, 64, 1, 1), (64, 128, 1, 0),
    (128, 128, 1, 1), (128, 256, 1, 0), (256, 256, 1, 0), (256, 256, 1, 0), (256, 35, 0, 0),
]
# state_dict key prefixes of SlimYOLOv2_quantize_bnfuse in layer order (slim_yolo_v2.py:58-87)
SLIM_CONV_KEYS = ["conv1.convs.0", "conv2.convs.0", "conv3_1.convs.0", "conv3_2.convs.0", "conv4_1.convs.0",
                  "conv4_2.convs.0", "conv5.convs.0", "conv6.convs.0", "conv7.convs.0", "pred"]
SLIM_TRACKER_KEYS = ["a_tracker_in", "a_tracker1", "a_tracker2", "a_tracker3_1", "a_tracker3_2", "a_tracker4_1",
                     "a_tracker4_2", "a_tracker5", "a_tracker6", "a_tracker7", "a_tracker_pred"]

# data/config.py:10-14
ANCHOR_SIZE = [[1.19, 1.98], [2.79, 4.59], [4.53, 8.92], [8.06, 5.29], [10.32, 10.65]]
ANCHOR_SIZE_MASK = [[0.27894, 0.49337], [0.8669, 1.37835], [1.82727, 2.8404], [3.4131, 5.05744], [5.8903, 7.6757]]
ANCHOR_SIZE_COCO = [[0.53, 0.79], [1.71, 2.36], [2.89, 6.44], [6.33, 3.79], [9.03, 9.74]]

# tables exactly as shipped in yolo_forward.c:32-35 (scale_a[0]: `65536` stored in a const char -> 0)
SHIPPED_SCALE_W = [6, 8, 8, 9, 9, 9, 10, 10, 10, 9]
SHIPPED_SCALE_B = [7, 6, 5, 5, 5, 6, 5, 5, 5, 10]
SHIPPED_SCALE_A = [0, 4, 8, 8, 8, 8, 8, 8, 8, 16, 4]
SHIPPED_RETUNE = [11, 10, 10, 11, 11, 10, 11, 11, 11, 10]


def cstride(c: int) -> int:
    """Channel stride of an int8 NHWC feature map (matches yolo_b200_cstride)."""
    return 4 if c <= 4 else (c + 15) // 16 * 16


def pow2_scale_exponent(t: torch.Tensor, bitwidth: int = 8) -> int:
    """log2 of the reference's power-of-two scale for tensor t (retune_bias_quantize.py:79-84)."""
    _max = t.abs().max()
    scale = (2 ** (bitwidth - 1) - 1) / _max
    return int(torch.floor(torch.log2(scale)).item())


def quantize_pow2(t: torch.Tensor, bitwidth: int = 8):
    """Returns (int tensor q, exponent e) with q = round(t * 2**e)  (retune_bias_quantize.py:73-97)."""
    e = pow2_scale_exponent(t, bitwidth)
    q = torch.round((2.0 ** e) * t)
    return q, e


@dataclass
class QuantNet:
    """A BN-fused fixed-point network: int8 weights/biases plus the exponent tables of yolo_forward.c:32-35."""
    layers: List[tuple]
    w: List[np.ndarray]            # int8 [cout][3][3][cin]  (OHWI)
    b: List[np.ndarray]            # int8 [cout]
    sw: List[int]
    sb: List[int]
    sa: List[int]                  # len(layers)+1, sa[0] = input
    retune: Optional[List[int]]   # None = never derived from real activations: contract F / weight.h export refused
    anchors: List[List[float]] = field(default_factory=lambda: [list(a) for a in ANCHOR_SIZE_MASK])
    num_classes: int = 2
    stride: int = 16
    graph: Optional[List[dict]] = None   # per layer: ksize / in_from / reorg / concat_with (include/yolo_b200.h, ABI 2); None = chain of 3x3

    def sha256(self) -> str:
        h = hashlib.sha256()
        for w, b in zip(self.w, self.b):
            h.update(np.ascontiguousarray(w).tobytes())
            h.update(np.ascontiguousarray(b).tobytes())
        h.update(np.asarray(self.sw + self.sb + self.sa + (self.retune or []), dtype=np.int32).tobytes())
        return h.hexdigest()

    def require_retune(self, what: str) -> List[int]:
        if self.retune is None:
            raise ValueError("%s needs retune[] (yolo_forward.c:35), which this network never derived from real activations: "
                             "pass calib_frames= or retune= to quantnet_from_state_dict" % what)
        return self.retune

    def dequantized_state_dict(self) -> Dict[str, torch.Tensor]:
        """The 42-key state_dict SlimYOLOv2_quantize_bnfuse.load_state_dict expects (weights as q/2**e,
        trackers frozen at 2**sa) — retune_bias_quantize.py:411-415 stores exactly this form."""
        sd: Dict[str, torch.Tensor] = {}
        for l, key in enumerate(SLIM_CONV_KEYS):
            w = torch.from_numpy(self.w[l].astype(np.float32)).permute(0, 3, 1, 2).contiguous()
            sd[key + ".weight"] = w / (2.0 ** self.sw[l])
            sd[key + ".bias"] = torch.from_numpy(self.b[l].astype(np.float32)) / (2.0 ** self.sb[l])
        for l, key in enumerate(SLIM_TRACKER_KEYS):
            sd[key + ".scale"] = torch.tensor([2.0 ** self.sa[l]])
            sd[key + ".first_a"] = torch.ones(1)
        return sd


def fold_bn(conv_w: torch.Tensor, conv_b: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
            mean: torch.Tensor, var: torch.Tensor, eps: float = 1e-5):
    """Conv + BatchNorm2d (eval statistics) -> one conv: `fuse_conv_and_bn`, conv+bn2conv.py:126-150, in the same float32
    operation order (W' = diag(gamma / sqrt(eps + var)) W; b' = b + (beta - gamma * mean / sqrt(var + eps))).  The
    reference multiplies by the diagonal matrix with torch.mm; every other term of those dot products is an exact zero,
    so the row-wise product below is bit-identical (tests/test_host.py checks it against the reference's own output)."""
    with torch.no_grad():
        cout = conv_w.shape[0]
        scale = gamma.div(torch.sqrt(eps + var))
        w = (scale.view(cout, 1) * conv_w.reshape(cout, -1)).view(conv_w.shape)
        b_conv = conv_b if conv_b is not None else torch.zeros(cout, dtype=conv_w.dtype)
        b_bn = beta - gamma.mul(mean).div(torch.sqrt(var + eps))
        return w, b_conv + b_bn


def fold_bn_state_dict(sd: Dict[str, torch.Tensor], eps: float = 1e-5) -> Dict[str, torch.Tensor]:
    """state_dict of an un-fused network (SlimYOLOv2, slim_yolo_v2.py:385: blocks `X.convs = Sequential(conv, bn, act)`,
    utils/modules.py:6-18) -> state_dict of its BN-fused twin (SlimYOLOv2_quantize_bnfuse / Conv2d_fuse: `X.convs.0` with
    a bias, no `X.convs.1`).  Generalises the loop of conv+bn2conv.py:317-326, which only visits top-level children: every
    BatchNorm in the dict (any nesting depth, recognised by its `running_var`) is folded into the module one index before
    it in the same Sequential, its keys are dropped, and later indices of that Sequential move down by one — the layout
    the reference's `nn.Sequential(fused, *rest)` produces.  Keys of modules without a BatchNorm pass through."""
    def order(bp):      # BatchNorms of one Sequential from the LAST to the first: a fold only renumbers members behind it,
        parent, _, idx = bp.rpartition(".")        # so the indices of the folds still to come stay valid
        return (parent, -int(idx) if idx.isdigit() else 0)
    bn_prefixes = sorted((k[:-len(".running_var")] for k in sd if k.endswith(".running_var")), key=order)
    out = dict(sd)
    for bp in bn_prefixes:
        parent, _, idx = bp.rpartition(".")
        if not idx.isdigit() or int(idx) == 0:
            raise ValueError("BatchNorm %s does not follow a convolution inside a Sequential" % bp)
        cp = "%s.%d" % (parent, int(idx) - 1)
        if cp + ".weight" not in out or out[cp + ".weight"].dim() != 4:
            raise ValueError("no convolution at %s for BatchNorm %s" % (cp, bp))
        w, b = fold_bn(out[cp + ".weight"].float(), out[cp + ".bias"].float() if cp + ".bias" in out else None,
                       out[bp + ".weight"].float(), out[bp + ".bias"].float(), out[bp + ".running_mean"].float(),
                       out[bp + ".running_var"].float(), eps)
        out[cp + ".weight"], out[cp + ".bias"] = w, b
        for k in [k for k in out if k.startswith(bp + ".")]:
            del out[k]
        # later members of the Sequential move down by one index
        moved = {}
        for k in [k for k in out if k.startswith(parent + ".")]:
            head, _, rest = k[len(parent) + 1:].partition(".")
            if head.isdigit() and int(head) > int(idx):
                moved["%s.%d.%s" % (parent, int(head) - 1, rest)] = out.pop(k)
        out.update(moved)
    return out


def _float_convs_from_state_dict(sd: Dict[str, torch.Tensor]):
    ws, bs = [], []
    for key in SLIM_CONV_KEYS:
        ws.append(sd[key + ".weight"].detach().float().cpu())
        bs.append(sd[key + ".bias"].detach().float().cpu())
    return ws, bs


def random_float_convs(seed: int = 0, layers: Sequence[tuple] = SLIM_YOLO_V2_LAYERS):
    """Random-init weights of the named architecture, drawn exactly as constructing the reference module
    after torch.manual_seed(seed) would (nn.Conv2d default init, construction order of slim_yolo_v2.py:58-87)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    ws, bs = [], []
    for cin, cout, _, _ in layers:
        conv = torch.nn.Conv2d(cin, cout, 3, 1, padding=1)
        ws.append(conv.weight.detach().clone())
        bs.append(conv.bias.detach().clone())
    torch.random.set_rng_state(g)
    return ws, bs


def calibrate(ws: List[torch.Tensor], bs: List[torch.Tensor], frames: torch.Tensor,
              layers: Sequence[tuple] = SLIM_YOLO_V2_LAYERS):
    """One calibration call of the fake-quant forward (slim_yolo_v2.py:212-328 with quantization=True on
    fresh trackers): every tracker takes scale = 127/max|a| on its first call (:22-27).  Also derives retune
    per layer from the same activations (:222-227).  ws/bs are the DE-quantised (q/s) float tensors.
    Returns (sa exponents [L+1], retune [L])."""
    sa: List[int] = []
    retune: List[int] = []
    with torch.no_grad():
        x = frames.float()
        e = pow2_scale_exponent(x)
        sa.append(e)
        x = torch.round(x * 2.0 ** e) / 2.0 ** e
        for l, (cin, cout, activ, pool) in enumerate(layers):
            y = F.conv2d(x, ws[l], bs[l], stride=1, padding=1)
            if activ:
                y = F.leaky_relu(y, 0.125)
            m = float(y.abs().max())
            # largest r with m * 2**r < 2**15
            r = int(math.floor(math.log2((2.0 ** 15) / m))) if m > 0 else 15
            while m * 2.0 ** r >= 2.0 ** 15:
                r -= 1
            retune.append(r)
            e = pow2_scale_exponent(y)
            sa.append(e)
            y = torch.round(y * 2.0 ** e) / 2.0 ** e
            if pool:
                y = F.max_pool2d(y, 2, 2)
            x = y
    return sa, retune


def quantize_convs(ws: List[torch.Tensor], bs: List[torch.Tensor]):
    """Reference weight/bias quantisation (retune_bias_quantize.py:111-119): returns int8 OHWI weights,
    int8 biases, exponents, and the de-quantised float tensors the fake-quant model runs with."""
    qw, qb, sw, sb, dw, db = [], [], [], [], [], []
    for w, b in zip(ws, bs):
        q, e = quantize_pow2(w)
        qw.append(q.permute(0, 2, 3, 1).contiguous().numpy().astype(np.int8))
        sw.append(e)
        dw.append(q / 2.0 ** e)
        q2, e2 = quantize_pow2(b)
        qb.append(q2.numpy().astype(np.int8))
        sb.append(e2)
        db.append(q2 / 2.0 ** e2)
    return qw, qb, sw, sb, dw, db


def build_quantnet(ws: List[torch.Tensor], bs: List[torch.Tensor], calib_frames: torch.Tensor,
                   layers: Sequence[tuple] = SLIM_YOLO_V2_LAYERS, anchors=None, num_classes: int = 2) -> QuantNet:
    qw, qb, sw, sb, dw, db = quantize_convs(ws, bs)
    sa, retune = calibrate(dw, db, calib_frames, layers)
    return QuantNet(list(layers), qw, qb, sw, sb, sa, retune,
                    [list(a) for a in (anchors or ANCHOR_SIZE_MASK)], num_classes)


def synthetic_frames_f32(n: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """Synthetic camera frames through BaseTransform's arithmetic (data/__init__.py:30-56, test.py:79-80):
    uint8 BGR image -> /255 -> -mean -> /std -> RGB, CHW.  No resize (image generated at the target size)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    x = img.astype(np.float32)
    x /= 255.
    x -= np.array((0.406, 0.456, 0.485), dtype=np.float32)
    x /= np.array((0.225, 0.224, 0.229), dtype=np.float32)
    x = x[..., (2, 1, 0)]
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))


def synthetic_image_u8(seed: int, h: int = 480, w: int = 640, kind: str = "noise") -> np.ndarray:
    """Seeded uint8 BGR camera image [h][w][3].  "noise": SURVEY.md 8(d) config 2 (uniform noise: after ten layers a
    random-init network answers almost the same value in every cell).  "scene": a smooth random field, 40 random
    half-transparent rectangles and a little noise, so that the prediction map (and hence the scores) varies from cell to cell."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    # smooth random field (a coarse random grid interpolated bilinearly) + 40 random filled rectangles + a little noise
    gy, gx = 13, 17
    coarse = rng.integers(0, 256, size=(gy, gx, 3)).astype(np.float64)
    yy = np.linspace(0, gy - 1, h); xx = np.linspace(0, gx - 1, w)
    y0 = np.minimum(yy.astype(int), gy - 2); x0 = np.minimum(xx.astype(int), gx - 2)
    fy = (yy - y0)[:, None, None]; fx = (xx - x0)[None, :, None]
    img = ((1 - fy) * ((1 - fx) * coarse[y0][:, x0] + fx * coarse[y0][:, x0 + 1]) +
           fy * ((1 - fx) * coarse[y0 + 1][:, x0] + fx * coarse[y0 + 1][:, x0 + 1]))
    for _ in range(40):
        ry, rx = int(rng.integers(0, h)), int(rng.integers(0, w))
        hh, ww = int(rng.integers(8, h // 2)), int(rng.integers(8, w // 2))
        img[ry:ry + hh, rx:rx + ww] = 0.5 * img[ry:ry + hh, rx:rx + ww] + 0.5 * rng.integers(0, 256, size=3)
    img = np.rint(img).astype(np.int32) + rng.integers(-12, 13, size=(h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def rgb444_lut(sa: int = 0) -> np.ndarray:
    """pixel_norm_quantize (c_embedding/yolo_forward.c:57-85) for all 4096 RGB444 codes -> int8 [4096][4] (R,G,B,0):
    mask WITHOUT shifting (R in 0..15, G in {0,16..240}, B in {0,256..3840}), /255., -mean, /std in the reference's
    float/double mix, * 2^sa, truncation toward zero, low byte kept.  Host replica used to build calibration frames;
    the device table is built by the C library (yolo_b200_rgb444_lut) and tests compare the two."""
    codes = np.arange(4096, dtype=np.int64)
    out = np.zeros((4096, 4), dtype=np.int8)
    for ch, (mask, mean, sd) in enumerate(((0x000f, 0.485, 0.229), (0x00f0, 0.456, 0.224), (0x0f00, 0.406, 0.225))):
        v = (codes & mask).astype(np.float32)
        v = (v.astype(np.float64) / 255.).astype(np.float32)
        v = (v.astype(np.float64) - mean).astype(np.float32)
        v = (v.astype(np.float64) / sd).astype(np.float32)
        t = np.trunc(v.astype(np.float64) * 2.0 ** sa).astype(np.int64)
        out[:, ch] = (t & 0xff).astype(np.uint8).view(np.int8)
    return out


def synthetic_frames_rgb444(n: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """Synthetic camera frames in the C path's input format: uint16 [n][h][w], 0x0BGR (ov7670.h:203-230)."""
    return np.random.default_rng(seed).integers(0, 4096, size=(n, h, w), dtype=np.uint16)


def random_quantnet(seed: int = 0, calib_hw=(416, 416), calib_frames: int = 2, head_bias_shift: float = 0.0,
                    anchors=None, calib_input: str = "f32", head_gain: float = 1.0, weight_gain: float = 1.0) -> QuantNet:
    """Random-init slim_yolo_v2 of the named architecture, quantised and calibrated by the reference's rules
    (SURVEY.md 8d 'Weights for all configs').  head_bias_shift is added to the 5 objectness biases of `pred`
    BEFORE quantisation: random init otherwise puts most anchors above the threshold (dense NMS worst case);
    a negative shift gives the sparse detections of a trained network."""
    ws, bs = random_float_convs(seed)
    if weight_gain != 1.0:
        # nn.Conv2d's default init shrinks the signal layer after layer until the biases dominate and every cell of the
        # prediction map holds nearly the same values; a He-like gain on every layer's weights keeps the input's structure
        # alive through the ten layers (used by the detection-parity fixtures, which need distinct scores)
        ws = [w * weight_gain for w in ws]
    if head_gain != 1.0:
        # a random-init `pred` answers with almost the same value everywhere (a few hundred distinct scores per frame):
        # scaling its weights spreads the prediction map over the int8 range like a trained head does
        ws[-1] = ws[-1] * head_gain
    if head_bias_shift:
        bs[-1] = bs[-1].clone()
        bs[-1][:5] += head_bias_shift
    if calib_input == "rgb444":
        # calibrate on what the RGB444 front end delivers at scale_a[0] = 0 (the shipped table): integers in [-2, 65]
        u16 = synthetic_frames_rgb444(calib_frames, calib_hw[0], calib_hw[1], seed=1000 + seed)
        q = rgb444_lut(0)[u16.astype(np.int64)][..., :3].astype(np.float32)
        frames = torch.from_numpy(np.ascontiguousarray(q.transpose(0, 3, 1, 2)))
    else:
        frames = synthetic_frames_f32(calib_frames, calib_hw[0], calib_hw[1], seed=1000 + seed)
    return build_quantnet(ws, bs, frames, anchors=anchors)


def quantnet_from_state_dict(sd: Dict[str, torch.Tensor], calib_frames: Optional[torch.Tensor] = None,
                             anchors=None, num_classes: int = 2, retune: Optional[Sequence[int]] = None) -> QuantNet:
    """q_bf state_dict (float or already-quantised `*_retune_quantize*.pth`, retune_bias_quantize.py:411-415)
    -> QuantNet.  The layer list comes from the state_dict itself (the head's width is `pred.weight.shape[0]`, which must
    equal len(anchors) * (5 + num_classes), slim_yolo_v2.py:87).  Activation exponents come from the checkpoint's trackers
    when they were calibrated (first_a != 0), else from `calib_frames`.

    `retune[]` (the 16-bit accumulator scale of contract F / weight.h, yolo_forward.c:35) only means something when it was
    derived from real activations: it is taken from `calib_frames` (the 2^15 rule of slim_yolo_v2.py:222-227) or from an
    explicit table (`retune=`, e.g. SHIPPED_RETUNE); with neither it is None and the QuantNet refuses contract F and the
    weight.h export (contract P, the PyTorch fake-quant arithmetic, never reads it)."""
    ws, bs = _float_convs_from_state_dict(sd)
    anchors = [list(a) for a in (anchors or ANCHOR_SIZE_MASK)]
    layers = [(int(w.shape[1]), int(w.shape[0]), a, p) for w, (_, _, a, p) in zip(ws, SLIM_YOLO_V2_LAYERS)]
    for l in range(1, len(layers)):
        if layers[l][0] != layers[l - 1][1]:
            raise ValueError("state_dict: layer %d takes %d channels but layer %d produces %d" % (l, layers[l][0], l - 1, layers[l - 1][1]))
    want = len(anchors) * (5 + num_classes)
    if layers[-1][1] != want:
        raise ValueError("pred.weight has %d output channels; %d anchors x (5 + %d classes) needs %d"
                         % (layers[-1][1], len(anchors), num_classes, want))
    qw, qb, sw, sb, dw, db = quantize_convs(ws, bs)
    have_trackers = all((k + ".scale") in sd and float(sd[k + ".first_a"]) != 0 for k in SLIM_TRACKER_KEYS)
    if calib_frames is None and not have_trackers:
        raise ValueError("state_dict has uncalibrated activation trackers: pass calib_frames")
    sa_c = rt_c = None
    if calib_frames is not None:
        sa_c, rt_c = calibrate(dw, db, calib_frames, layers)
    sa = [int(math.floor(math.log2(float(sd[k + ".scale"])))) for k in SLIM_TRACKER_KEYS] if have_trackers else sa_c
    if retune is not None:
        if len(retune) != len(layers):
            raise ValueError("retune table needs %d entries" % len(layers))
        rt = [int(r) for r in retune]
    else:
        rt = rt_c          # None when the checkpoint brought its own trackers and no calibration frames were given
    return QuantNet(layers, qw, qb, sw, sb, sa, rt, anchors, num_classes)


# ---- yolo_v2 / darknet19 (BASELINE configs[4]) --------------------------------------------------------------------

def yolo_v2_layers(num_classes: int = 20, num_anchors: int = 5):
    """The BN-folded yolo_v2 graph as (cin, cout, activ, pool) tuples + the ABI-2 graph fields of include/yolo_b200.h:
    darknet19 (backbone/darknet.py:40-108: conv_1 .. conv_6, maxpool_4 / maxpool_5 written as the pool of the layer in front
    of them), convsets_1, route_layer (1x1 on C_5, the map in front of maxpool_5) + reorg, the concat, convsets_2 and the
    1x1 pred (models/yolo_v2.py:29-40,165-177).  Every leaky-ReLU is the 1/8 shift (the head's slope is 0.125,
    utils/modules.py:25; the backbone's 0.1, darknet.py:18, is not a shift: SURVEY.md 8d config 5 [DECISION])."""
    L, G = [], []

    def add(cin, cout, ks, activ=1, pool=0, **g):
        L.append((cin, cout, activ, pool))
        G.append(dict(ksize=ks, **g))
    add(3, 32, 3, pool=1)
    add(32, 64, 3, pool=1)
    add(64, 128, 3); add(128, 64, 1); add(64, 128, 3, pool=1)
    add(128, 256, 3); add(256, 128, 1); add(128, 256, 3, pool=1)                                  # conv_4, maxpool_4
    add(256, 512, 3); add(512, 256, 1); add(256, 512, 3); add(512, 256, 1); add(256, 512, 3, pool=1)    # conv_5 (C_5 = un-pooled output of layer 12), maxpool_5
    add(512, 1024, 3); add(1024, 512, 1); add(512, 1024, 3); add(1024, 512, 1); add(512, 1024, 3)       # conv_6
    add(1024, 1024, 3); add(1024, 1024, 3)                                                        # convsets_1 (layers 18, 19)
    add(512, 64, 1, in_from=13, reorg=1)                                                          # route_layer on C_5 + reorg (layer 20)
    add(1280, 1024, 3, in_from=20, concat_with=21)                                                # cat([reorg(route), convsets_1]) -> convsets_2
    add(1024, num_anchors * (5 + num_classes), 1, activ=0)                                        # pred
    return L, G


def input_exponent(layers, graph, sa, l):
    """Activation exponent of layer l's input: its source's output exponent; a concat brings both parts to the smaller one."""
    if l == 0:
        return sa[0]
    g = (graph or [{}] * len(layers))[l]
    src = g.get("in_from", 0) - 1 if g.get("in_from", 0) else l - 1
    e = sa[src + 1]
    if g.get("concat_with", 0):
        e = min(e, sa[g["concat_with"]])
    return e


def reorg_nchw(x: torch.Tensor) -> torch.Tensor:
    """reorg_layer(stride=2), utils/modules.py:43-57: out[b, (2*dy+dx)*C + c, y, x] = in[b, c, 2y+dy, 2x+dx]."""
    b, c, h, w = x.shape
    x = x.view(b, c, h // 2, 2, w // 2, 2).permute(0, 3, 5, 1, 2, 4).contiguous()
    return x.view(b, 4 * c, h // 2, w // 2)


def calibrate_graph(ws, bs, frames, layers, graph):
    """calibrate() for a graph network: one fake-quant forward with fresh trackers over the yolo_v2 graph; the parts of a concat
    are re-rounded to the smaller of their two exponents, as the integer path does.  Returns (sa [L+1], retune [L])."""
    sa, retune = [], []
    pre, post = {}, {}                  # un-pooled / pooled (consumer-visible) fake-quant outputs per layer
    with torch.no_grad():
        x = frames.float()
        e = pow2_scale_exponent(x)
        sa.append(e)
        x0 = torch.round(x * 2.0 ** e) / 2.0 ** e
        for l, (cin, cout, activ, pool) in enumerate(layers):
            g = graph[l]
            if l == 0:
                xin = x0
            else:
                src = g.get("in_from", 0) - 1 if g.get("in_from", 0) else l - 1
                xin = pre[src] if g.get("in_from", 0) else post[src]
                if g.get("concat_with", 0):
                    k = g["concat_with"] - 1
                    a = reorg_nchw(post[k]) if graph[k].get("reorg", 0) else post[k]
                    ec = min(sa[k + 1], sa[src + 1])
                    xin = torch.cat([torch.round(a * 2.0 ** ec) / 2.0 ** ec, torch.round(xin * 2.0 ** ec) / 2.0 ** ec], 1)
            ks = g.get("ksize", 3) or 3
            y = F.conv2d(xin, ws[l], bs[l], stride=1, padding=ks // 2)
            if activ:
                y = F.leaky_relu(y, 0.125)
            m = float(y.abs().max())
            r = int(math.floor(math.log2((2.0 ** 15) / m))) if m > 0 else 15
            while m * 2.0 ** r >= 2.0 ** 15:
                r -= 1
            retune.append(r)
            e = pow2_scale_exponent(y)
            sa.append(e)
            y = torch.round(y * 2.0 ** e) / 2.0 ** e
            pre[l] = y
            post[l] = F.max_pool2d(y, 2, 2) if pool else y
    return sa, retune


def random_quantnet_yolo_v2(seed: int = 0, calib_hw=(416, 416), calib_frames: int = 1, num_classes: int = 20, anchors=None,
                            weight_gain: float = 2.0) -> QuantNet:
    """Random-init, BN-folded yolo_v2 (darknet19 backbone) quantised with the reference's power-of-two rule and calibrated by
    the first-call tracker rule.  nn.Conv2d's default init in construction order; weight_gain keeps the
    signal alive through 23 layers (see random_quantnet)."""
    layers, graph = yolo_v2_layers(num_classes, len(anchors or ANCHOR_SIZE))
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    ws, bs = [], []
    for (cin, cout, _, _), gg in zip(layers, graph):
        conv = torch.nn.Conv2d(cin, cout, gg["ksize"], 1, padding=gg["ksize"] // 2)
        ws.append(conv.weight.detach().clone() * weight_gain)
        bs.append(conv.bias.detach().clone())
    torch.random.set_rng_state(g)
    qw, qb, sw, sb, dw, db = quantize_convs(ws, bs)
    frames = synthetic_frames_f32(calib_frames, calib_hw[0], calib_hw[1], seed=1000 + seed)
    sa, retune = calibrate_graph(dw, db, frames, layers, graph)
    return QuantNet(layers, qw, qb, sw, sb, sa, retune, [list(a) for a in (anchors or ANCHOR_SIZE)], num_classes, stride=32, graph=graph)


# ---- weight.h ---------------------------------------------------------------------------------------

TM, TN = 32, 16   # accelerator tile: Tm output x Tn input channels (main.c:44, yolo_forward.c:1181)


def pack_weight_h_order(w_ohwi: np.ndarray) -> np.ndarray:
    """[cout][3][3][cin] -> flat int8 in the burst order load_weight reads (yolo_forward.c:165-173):
    tap-major with tap stride cin*cout; inside a tap [cout/Tm][cin/Tn][Tm][Tn] (callers advance by Tn*Tm
    per (out-group, in-group) pass, :697).  Layers thinner than the tile use their own size as the group.
    The order INSIDE a Tn*Tm block is not determinable from the reference; out-major is this repo's choice."""
    cout, _, _, cin = w_ohwi.shape
    tm, tn = min(TM, cout), min(TN, cin)
    gco, gci = -(-cout // tm), -(-cin // tn)
    buf = np.zeros((3, 3, gco, gci, tm, tn), dtype=np.int8)
    for go in range(gco):
        for gi in range(gci):
            blk = w_ohwi[go * tm:(go + 1) * tm, :, :, gi * tn:(gi + 1) * tn]   # [<=tm][3][3][<=tn]
            buf[:, :, go, gi, :blk.shape[0], :blk.shape[3]] = blk.transpose(1, 2, 0, 3)
    return buf.reshape(-1)


def unpack_weight_h_order(flat: np.ndarray, cout: int, cin: int) -> np.ndarray:
    tm, tn = min(TM, cout), min(TN, cin)
    gco, gci = -(-cout // tm), -(-cin // tn)
    buf = np.asarray(flat, dtype=np.int8).reshape(3, 3, gco, gci, tm, tn)
    w = np.zeros((gco * tm, 3, 3, gci * tn), dtype=np.int8)
    for go in range(gco):
        for gi in range(gci):
            w[go * tm:(go + 1) * tm, :, :, gi * tn:(gi + 1) * tn] = buf[:, :, go, gi].transpose(2, 0, 1, 3)
    return w[:cout, :, :, :cin]


def write_weight_h(net: QuantNet, path: str) -> None:
    """Emit a weight.h with the symbols yolo_forward.c uses (w_conv0..9 / b_conv0..9, :1204-1260) plus the
    exponent tables of :32-35 as comments."""
    net.require_retune("weight.h (the C driver's tables)")
    with open(path, "w") as f:
        f.write("/* generated by yolo_b200 export.write_weight_h — layout: see pack_weight_h_order */\n")
        f.write("/* scale_w = %s\n   scale_b = %s\n   scale_a = %s\n   retune  = %s */\n"
                % (net.sw, net.sb, net.sa, net.retune))
        for l, (w, b) in enumerate(zip(net.w, net.b)):
            flat = pack_weight_h_order(w)
            f.write("const char w_conv%d[%d] = {%s};\n" % (l, flat.size, ",".join(str(int(v)) for v in flat)))
            f.write("const char b_conv%d[%d] = {%s};\n" % (l, b.size, ",".join(str(int(v)) for v in b)))


def read_weight_h(path: str, layers: Sequence[tuple] = SLIM_YOLO_V2_LAYERS):
    """Parse a weight.h written in the layout above -> (weights OHWI, biases)."""
    import re
    txt = open(path).read()
    ws, bs = [], []
    for l, (cin, cout, _, _) in enumerate(layers):
        mw = re.search(r"w_conv%d\[\d*\]\s*=\s*\{([^}]*)\}" % l, txt)
        mb = re.search(r"b_conv%d\[\d*\]\s*=\s*\{([^}]*)\}" % l, txt)
        if not mw or not mb:
            raise ValueError("weight.h lacks w_conv%d / b_conv%d" % (l, l))
        flat = np.array([int(v) for v in mw.group(1).split(",")], dtype=np.int8)
        ws.append(unpack_weight_h_order(flat, cout, cin))
        bs.append(np.array([int(v) for v in mb.group(1).split(",")], dtype=np.int8)[:cout])
    return ws, bs
