"""ctypes access to the CPU oracle (oracle/ref_int8.c, oracle/_ref/libtierA.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None
_tierA = None

I8P = C.POINTER(C.c_int8)


class OracleDet(C.Structure):
    _fields_ = [("x1", C.c_float), ("y1", C.c_float), ("x2", C.c_float), ("y2", C.c_float), ("score", C.c_float),
                ("cls", C.c_int32), ("anchor_index", C.c_int32), ("pad_", C.c_int32)]


class OracleNet(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("cin", C.c_int * 32), ("cout", C.c_int * 32), ("activ", C.c_int * 32),
                ("pool", C.c_int * 32), ("sw", C.c_int * 32), ("sb", C.c_int * 32), ("sa", C.c_int * 33),
                ("retune", C.c_int * 32), ("contract", C.c_int), ("round_mode", C.c_int)]


def build(force=False):
    so = os.path.join(ORACLE_DIR, "_ref", "liboracle.so")
    src = os.path.join(ORACLE_DIR, "ref_int8.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_conv_layer.restype = C.c_int
        _lib.oracle_requant.restype = C.c_int
        _lib.oracle_requant.argtypes = [C.c_int64] + [C.c_int] * 9
        _lib.oracle_nms_python.restype = C.c_int
        _lib.oracle_head_c.restype = C.c_int
        _lib.oracle_backbone.restype = C.c_int
        _lib.oracle_set_threads.restype = C.c_int
    return _lib


def set_threads(n=None):
    """OpenMP threads of the oracle (default: every logical CPU); returns the count in force."""
    return lib().oracle_set_threads(int(n or (os.cpu_count() or 1)))


def tierA():
    """The reference's own C functions (None when oracle/_ref/libtierA.so was not built/shipped)."""
    global _tierA
    if _tierA is None:
        build()
        so = os.path.join(ORACLE_DIR, "_ref", "libtierA.so")
        if not os.path.exists(so):
            return None
        _tierA = C.CDLL(so)
        _tierA.tierA_sigmoid.restype = C.c_float
        _tierA.tierA_sigmoid.argtypes = [C.c_float]
        _tierA.tierA_dequantize.restype = C.c_float
        _tierA.tierA_box_iou.restype = C.c_float
        _tierA.tierA_tables.restype = C.POINTER(C.c_int8)
        _tierA.tierA_anchors.restype = C.POINTER(C.c_float)
        _tierA.tierA_decode_txtytwth.argtypes = [C.c_float] * 4 + [C.c_int] * 3 + [C.POINTER(C.c_int)]
        _tierA.tierA_sort_nms.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_float,
                                          C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _tierA


def _p8(a):
    return a.ctypes.data_as(I8P)


def cs(c):
    return 4 if c <= 4 else (c + 15) // 16 * 16


def conv_layer(x, w_ohwi, b, cin, cout, sa_i, sw, sb, retune, sa_o, activ, pool, contract, round_mode=0):
    """x: int8 [n][h][w][cs_in] -> (int8 [n][h'][w'][cs(cout)], overflow count)."""
    x = np.ascontiguousarray(x, dtype=np.int8)
    n, h, w_, cs_in = x.shape
    oh, ow = (h // 2, w_ // 2) if pool else (h, w_)
    out = np.zeros((n, oh, ow, cs(cout)), dtype=np.int8)
    ovf = C.c_int64(0)
    w_ohwi = np.ascontiguousarray(w_ohwi, dtype=np.int8)
    b = np.ascontiguousarray(b, dtype=np.int8)
    rc = lib().oracle_conv_layer(_p8(x), n, h, w_, cs_in, cin, _p8(w_ohwi), _p8(b), cout, cs(cout),
                                 int(sa_i), int(sw), int(sb), int(retune), int(sa_o), int(activ), int(pool),
                                 int(contract), int(round_mode), _p8(out), C.byref(ovf))
    assert rc == 0
    return out, ovf.value


def conv_acc(x, w_ohwi, cin, cout):
    x = np.ascontiguousarray(x, dtype=np.int8)
    n, h, w_, cs_in = x.shape
    acc = np.zeros((n, h, w_, cout), dtype=np.int32)
    w_ohwi = np.ascontiguousarray(w_ohwi, dtype=np.int8)
    lib().oracle_conv_acc(_p8(x), n, h, w_, cs_in, cin, _p8(w_ohwi), cout, acc.ctypes.data_as(C.POINTER(C.c_int32)))
    return acc


def backbone(qnet, x_nhwc4, contract, round_mode=0, keep_layers=True):
    """Runs every layer; returns (list of per-layer int8 maps, overflow count)."""
    cur = np.ascontiguousarray(x_nhwc4, dtype=np.int8)
    outs, total = [], 0
    for l, (cin, cout, activ, pool) in enumerate(qnet.layers):
        cur, ovf = conv_layer(cur, qnet.w[l], qnet.b[l], cin, cout, qnet.sa[l], qnet.sw[l], qnet.sb[l],
                              qnet.retune[l], qnet.sa[l + 1], activ, pool, contract, round_mode)
        total += ovf
        outs.append(cur)
    return outs, total


def _shr_rne(v, s):
    """numpy int32 round-half-even right shift (the alignment of a concat's parts, include/yolo_b200.h: concat_with)."""
    if s <= 0:
        return v
    fl = v >> s
    rem = v - (fl << s)
    half = 1 << (s - 1)
    return fl + ((rem > half) | ((rem == half) & ((fl & 1) == 1)))


def backbone_graph(qnet, x_nhwc4, contract, round_mode=0):
    """A graph network (yolo_v2 / darknet19: 1x1 layers, a route that reads an un-pooled map, reorg, concat) on the CPU
    oracle: every convolution is oracle_conv_layer (a 1x1 kernel = the centre tap of a 3x3 one) WITHOUT its pool; pooling,
    reorg (utils/modules.py:43-57), concat (yolo_v2.py:171-174) and the exponent alignment are restated in numpy.
    Returns (list of consumer-visible per-layer outputs, i.e. after the pool, overflow count)."""
    import yolo_b200  # noqa: F401
    from yolo_b200 import export as ex
    graph = qnet.graph or [{}] * len(qnet.layers)
    pre, post, total = {}, {}, 0
    x0 = np.ascontiguousarray(x_nhwc4, dtype=np.int8)
    for l, (cin, cout, activ, pool) in enumerate(qnet.layers):
        g = graph[l]
        if l == 0:
            xin = x0
        else:
            src = g.get("in_from", 0) - 1 if g.get("in_from", 0) else l - 1
            xin = (pre[src] if g.get("in_from", 0) else post[src])[..., :qnet.layers[src][1]]
            if g.get("concat_with", 0):
                k = g["concat_with"] - 1
                a = post[k][..., :qnet.layers[k][1]]
                if graph[k].get("reorg", 0):
                    n, h, w, c = a.shape
                    a = a.reshape(n, h // 2, 2, w // 2, 2, c).transpose(0, 1, 3, 2, 4, 5).reshape(n, h // 2, w // 2, 4 * c)
                ea, eb = qnet.sa[k + 1], qnet.sa[src + 1]
                e = min(ea, eb)
                xin = np.concatenate([np.clip(_shr_rne(a.astype(np.int32), ea - e), -128, 127),
                                      np.clip(_shr_rne(xin.astype(np.int32), eb - e), -128, 127)], axis=-1).astype(np.int8)
            pad = cs(cin) - xin.shape[-1]
            if pad:
                xin = np.concatenate([xin, np.zeros(xin.shape[:-1] + (pad,), np.int8)], axis=-1)
        w = qnet.w[l]
        if w.shape[1] == 1:                                   # 1x1 -> centre tap
            w3 = np.zeros((w.shape[0], 3, 3, w.shape[3]), np.int8)
            w3[:, 1, 1, :] = w[:, 0, 0, :]
            w = w3
        sa_i = ex.input_exponent(qnet.layers, graph, qnet.sa, l)
        y, ovf = conv_layer(xin, w, qnet.b[l], cin, cout, sa_i, qnet.sw[l], qnet.sb[l], qnet.retune[l], qnet.sa[l + 1], activ, 0,
                            contract, round_mode)
        total += ovf
        pre[l] = y
        if pool:
            n, h, w_, c = y.shape
            y = y[:, :h // 2 * 2, :w_ // 2 * 2].reshape(n, h // 2, 2, w_ // 2, 2, c).max(axis=(2, 4))
        post[l] = y
    return [post[l] for l in range(len(qnet.layers))], total


def quantize_f32(x_nchw, sa):
    x = np.ascontiguousarray(x_nchw, dtype=np.float32)
    n, c, h, w = x.shape
    assert c == 3
    out = np.zeros((n, h, w, 4), dtype=np.int8)
    ovf = C.c_int64(0)
    lib().oracle_quantize_f32(x.ctypes.data_as(C.POINTER(C.c_float)), n, h, w, int(sa), _p8(out), C.byref(ovf))
    return out, ovf.value


def quantize_rgb444(frames_u16, sa):
    f = np.ascontiguousarray(frames_u16, dtype=np.uint16)
    out = np.zeros(f.shape + (4,), dtype=np.int8)
    lib().oracle_quantize_rgb444(f.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_size_t(f.size), int(sa), _p8(out))
    return out


def rgb444_lut(sa):
    codes = np.arange(4096, dtype=np.uint16)
    return quantize_rgb444(codes, sa)


def decode_python(pred, A, Cn, sa_pred, anchors, stride, in_h, in_w):
    """pred: int8 [gh][gw][cs] -> boxes [N,4], scores [N], cls [N]."""
    pred = np.ascontiguousarray(pred, dtype=np.int8)
    gh, gw, csz = pred.shape
    N = gh * gw * A
    boxes = np.zeros((N, 4), np.float32); scores = np.zeros(N, np.float32); cls = np.zeros(N, np.int32)
    anc = np.ascontiguousarray(anchors, dtype=np.float32)
    lib().oracle_decode_python(_p8(pred), gh, gw, csz, A, Cn, int(sa_pred), anc.ctypes.data_as(C.POINTER(C.c_float)),
                               int(stride), int(in_h), int(in_w), boxes.ctypes.data_as(C.POINTER(C.c_float)),
                               scores.ctypes.data_as(C.POINTER(C.c_float)), cls.ctypes.data_as(C.POINTER(C.c_int32)))
    return boxes, scores, cls


def nms_python(boxes, scores, cls, Cn, conf_thresh, nms_thresh, max_det=4096):
    N = len(scores)
    dets = (OracleDet * max_det)()
    cnt = lib().oracle_nms_python(boxes.ctypes.data_as(C.POINTER(C.c_float)), scores.ctypes.data_as(C.POINTER(C.c_float)),
                                  cls.ctypes.data_as(C.POINTER(C.c_int32)), N, Cn, C.c_float(conf_thresh),
                                  C.c_float(nms_thresh), dets, max_det)
    return dets_to_arrays(dets, min(cnt, max_det)), cnt


def head_python(pred, A, Cn, sa_pred, anchors, stride, in_h, in_w, conf_thresh, nms_thresh, max_det=4096):
    b, s, c = decode_python(pred, A, Cn, sa_pred, anchors, stride, in_h, in_w)
    return nms_python(b, s, c, Cn, conf_thresh, nms_thresh, max_det)


def head_c(pred, A, sa_pred, anchors, stride, conf_thresh, nms_thresh, max_det=4096):
    pred = np.ascontiguousarray(pred, dtype=np.int8)
    gh, gw, csz = pred.shape
    anc = np.ascontiguousarray(anchors, dtype=np.float32)
    dets = (OracleDet * max_det)()
    cnt = lib().oracle_head_c(_p8(pred), gh, gw, csz, A, int(sa_pred), anc.ctypes.data_as(C.POINTER(C.c_float)),
                              int(stride), C.c_float(conf_thresh), C.c_float(nms_thresh), dets, max_det)
    return dets_to_arrays(dets, min(cnt, max_det)), cnt


def dets_to_arrays(dets, cnt):
    a = np.zeros((cnt, 4), np.float32); s = np.zeros(cnt, np.float32); c = np.zeros(cnt, np.int64); idx = np.zeros(cnt, np.int64)
    for i in range(cnt):
        d = dets[i]
        a[i] = (d.x1, d.y1, d.x2, d.y2); s[i] = d.score; c[i] = d.cls; idx[i] = d.anchor_index
    return a, s, c, idx
