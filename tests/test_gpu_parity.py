"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI of include/yolo_b200.h, against
(1) the golden vectors of the unmodified reference module (tests/golden) and (2) the CPU oracle on seeded inputs.
Integer feature maps must be bit-exact; the float head is held to 1e-5 with identical kept sets."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import golden_util as gu
import oracle_lib as ol
import yolo_b200  # noqa: F401
from yolo_b200 import export as ex
from yolo_b200 import lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    c = lib.Context(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def layer_shapes(qnet, h, w):
    out = []
    for (cin, cout, activ, pool) in qnet.layers:
        if pool:
            h, w = h // 2, w // 2
        out.append((h, w))
    return out


def run_backbone(ctx, x8):
    """x8: int8 [n,h,w,4] numpy -> list of per-layer outputs (numpy)"""
    n, h, w, _ = x8.shape
    d = dev(x8)
    pred, gh, gw = ctx.backbone(d, n, h, w)
    ctx.sync()
    qshapes = []
    hh, ww = h, w
    for l in range(ctx.params.num_layers):
        if ctx.params.layers[l].pool:
            hh, ww = hh // 2, ww // 2
        qshapes.append((hh, ww))
    return [ctx.layer_output(l, n, *qshapes[l]) for l in range(ctx.params.num_layers)], (pred, gh, gw)


def det_arrays(ctx, pred_ptr, n, gh, gw, in_h, in_w):
    md = ctx.params.max_det
    d_dets = torch.zeros((n, md, 8), dtype=torch.int32, device="cuda")
    d_counts = torch.zeros((n,), dtype=torch.int32, device="cuda")
    ctx.detect(pred_ptr, n, gh, gw, in_h, in_w, d_dets, d_counts)
    ctx.sync()
    dets = d_dets.cpu().numpy().view(lib.DET_DTYPE).reshape(n, md)
    return dets, d_counts.cpu().numpy()


# ---- (1) golden vectors from the reference module: contract P ------------------------------------------

@pytest.mark.parametrize("name", gu.FIXTURES)
def test_contract_p_against_reference_golden(ctx, name):
    g, qnet, frames = gu.load(name)
    H, W = int(g["H"]), int(g["W"])
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_P, conf_thresh=float(g["conf_thresh"]), nms_thresh=float(g["nms_thresh"]))
    for i in range(int(g["n_frames"])):
        x = frames[i:i + 1].contiguous().cuda()
        x8 = torch.empty((1, H, W, 4), dtype=torch.int8, device="cuda")
        ctx.quantize_f32(x, 1, H, W, x8)
        ctx.sync()
        assert gu.sha(x8.cpu().numpy()[0]) == str(g["f%d_map0_sha256" % i])
        outs, (pred, gh, gw) = run_backbone(ctx, x8.cpu().numpy())
        assert ctx.overflow_count() == 0
        for l, o in enumerate(outs):
            assert gu.sha(o[0]) == str(g["f%d_map%d_sha256" % (i, l + 1)]), "layer %d differs from the reference" % l
        dets, counts = det_arrays(ctx, pred, 1, gh, gw, H, W)
        b, s, c, idx = lib.dets_to_arrays(dets[0], int(counts[0]))
        assert np.all(np.diff(idx) > 0)
        if int(g["f%d_tie_robust" % i]):
            assert counts[0] == len(g["f%d_scores" % i])
            np.testing.assert_array_equal(c, g["f%d_cls" % i])
            np.testing.assert_allclose(s, g["f%d_scores" % i], atol=1e-5, rtol=0)
            np.testing.assert_allclose(b, g["f%d_bboxes" % i], atol=1e-5, rtol=0)
        else:
            # the reference's own kept set hinges on NumPy's unstable argsort (tied scores that overlap): both lists must
            # be greedy outcomes of the reference's candidates under some order of the ties, and agree elsewhere
            rk, rb_, rs_, rc_ = gu.reference_kept_indices(g, i)
            conf, nt = float(g["conf_thresh"]), float(g["nms_thresh"])
            assert gu.greedy_consistent(rb_, rs_, rc_, rk, conf, nt) == []
            assert gu.greedy_consistent(rb_, rs_, rc_, idx, conf, nt) == []
            differ = sorted(set(rk.tolist()) ^ set(idx.tolist()))
            assert len([a for a in differ if not gu.tie_affected(rb_, rs_, rc_, a, nt)]) <= len(differ) // 2
            both = sorted(set(rk.tolist()) & set(idx.tolist()))
            sel = np.isin(idx, both)
            np.testing.assert_allclose(s[sel], rs_[both], atol=1e-5, rtol=0)
            np.testing.assert_allclose(b[sel], rb_[both], atol=1e-5, rtol=0)
        # every frame: identical kept set to the oracle (same deterministic tie rule)
        (ob, os_, oc, oidx), ocnt = ol.head_python(outs[-1][0], 5, 2, qnet.sa[10], qnet.anchors, 16, H, W,
                                                   float(g["conf_thresh"]), float(g["nms_thresh"]))
        assert counts[0] == ocnt
        np.testing.assert_array_equal(idx, oidx)
        np.testing.assert_array_equal(c, oc)
        np.testing.assert_allclose(s, os_, atol=1e-5, rtol=0)
        np.testing.assert_allclose(b, ob, atol=1e-5, rtol=0)


def test_forward_f32_host_api_matches_golden(ctx):
    """The whole call a user makes (host buffers in, detections out)."""
    g, qnet, frames = gu.load("ref_p_64x96")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_P, conf_thresh=float(g["conf_thresh"]), nms_thresh=float(g["nms_thresh"]))
    dets, counts = ctx.forward_f32(frames.numpy())
    for i in range(int(g["n_frames"])):
        b, s, c, idx = lib.dets_to_arrays(dets[i], int(counts[i]))
        assert counts[i] == len(g["f%d_scores" % i])
        np.testing.assert_array_equal(c, g["f%d_cls" % i])
        np.testing.assert_allclose(s, g["f%d_scores" % i], atol=1e-5, rtol=0)
        np.testing.assert_allclose(b, g["f%d_bboxes" % i], atol=1e-5, rtol=0)


def test_dropin_module_matches_golden():
    """SlimYOLOv2_quantize_bnfuse drop-in: load_state_dict of a reference-format checkpoint, forward(quantization=True)."""
    from yolo_b200 import model
    g, qnet, frames = gu.load("ref_p_64x96")
    net = model.SlimYOLOv2_quantize_bnfuse(torch.device("cuda"), input_size=[64, 96], num_classes=2, trainable=False,
                                           conf_thresh=float(g["conf_thresh"]), nms_thresh=float(g["nms_thresh"]),
                                           anchor_size=qnet.anchors)
    missing = net.load_state_dict(qnet.dequantized_state_dict())
    assert not missing.missing_keys and not missing.unexpected_keys
    net = net.cuda().eval()
    for i in range(int(g["n_frames"])):
        b, s, c = net(frames[i:i + 1].cuda(), quantization=True)
        assert net.last_overflow == 0
        np.testing.assert_array_equal(c, g["f%d_cls" % i])
        np.testing.assert_allclose(s, g["f%d_scores" % i], atol=1e-5, rtol=0)
        np.testing.assert_allclose(b, g["f%d_bboxes" % i], atol=1e-5, rtol=0)
        assert b.dtype == np.float32 and c.dtype == np.int64


def test_dropin_find_branch_and_tracker_averaging_match_the_reference():
    """forward(x, quantization=True, find=True) — what retune_bias_quantize_findbest.py:364 and the evaluator issue — on the
    GPU: fresh trackers are calibrated in the find branch on the first batch (first-call rule), two more un-frozen batches move
    them by the exponential average (slim_yolo_v2.py:31), and the detections of an unseen frame equal the reference's.
    The tracker `scale` buffers must hold the reference's float32 values."""
    from yolo_b200 import model
    g, qnet, frames = gu.load("ref_p_64x96_find")
    H, W, seed = int(g["H"]), int(g["W"]), int(g["seed"])
    net = model.SlimYOLOv2_quantize_bnfuse(torch.device("cuda"), input_size=[H, W], num_classes=2, trainable=False,
                                           conf_thresh=float(g["conf_thresh"]), nms_thresh=float(g["nms_thresh"]),
                                           anchor_size=qnet.anchors)
    sd = qnet.dequantized_state_dict()
    for l, key in enumerate(ex.SLIM_CONV_KEYS):            # the checkpoint holds the plain exponents: undo the find shift
        sd[key + ".weight"] = sd[key + ".weight"] * 2.0 ** net.FIND_SHIFTS[l]
        sd[key + ".bias"] = sd[key + ".bias"] * 2.0 ** net.FIND_SHIFTS[l]
    for k in ex.SLIM_TRACKER_KEYS:                          # fresh trackers
        sd[k + ".scale"] = torch.zeros(1); sd[k + ".first_a"] = torch.zeros(1)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    calib = ex.synthetic_frames_f32(2, H, W, seed=1000 + seed).cuda()
    net(calib, quantization=True, find=True)                # first call: calibrates (and decodes frame 0, discarded)
    # (in the find branch an activation is acc * 2^-46 + b * 2^-23: more than 24 significant bits, so the reference's own
    # float32 convolution output - and with it 127 / max|a| - carries a rounding that depends on MKL-DNN's summation order;
    # the integer pass here is exact.  Hence a few-ulp tolerance on the float buffers; the exponents floor(log2(scale)) and
    # every quantised map are compared exactly elsewhere.)
    got = np.array([float(t.scale) for t in net._trackers()], np.float32)
    np.testing.assert_allclose(got, g["tracker_scales_first"], rtol=1e-6, atol=0)
    for bi in range(g["tracker_scales_ema"].shape[0]):
        batch = (ex.synthetic_frames_f32(2, H, W, seed=3000 + seed + bi) * (1.0 + 0.5 * bi)).cuda()
        net.update_trackers(batch, find=True)
        got = np.array([float(t.scale) for t in net._trackers()], np.float32)
        np.testing.assert_allclose(got, g["tracker_scales_ema"][bi], rtol=1e-6, atol=0)
    assert [int(np.floor(np.log2(float(t.scale)))) for t in net._trackers()] == qnet.sa
    b, s, c = net(frames[:1].cuda(), quantization=True, find=True)
    assert net.last_overflow == 0
    rk, rb_, rs_, rc_ = gu.reference_kept_indices(g, 0)
    if int(g["f0_tie_robust"]):
        np.testing.assert_array_equal(c, g["f0_cls"])
        np.testing.assert_allclose(s, g["f0_scores"], atol=1e-5, rtol=0)
        np.testing.assert_allclose(b, g["f0_bboxes"], atol=1e-5, rtol=0)
    else:
        assert abs(len(s) - len(g["f0_scores"])) <= max(2, len(g["f0_scores"]) // 10)
    # the overflow probe (slim_yolo_v2.py:222-226): max|y| per tracker, exact; un-pooled layers can be checked against the
    # reference's own integer maps: max|q| = RNE(max|y| * 2^sa) because rounding is monotone and odd
    x = frames[:1].cuda().contiguous()
    mx = net._ctx.measure_f32(x, 1, H, W)
    assert mx[0] == float(frames[:1].abs().max())
    for l, (cin, cout, activ, pool) in enumerate(qnet.layers):
        assert mx[l + 1] * 2.0 ** net.FIND_SHIFTS[l] < 2 ** 15
        if not pool:
            assert int(np.abs(g["f0_map%d" % (l + 1)].astype(np.int32)).max()) == int(np.rint(mx[l + 1] * 2.0 ** qnet.sa[l + 1]))
    # ... and it fires when a layer leaves the 16-bit accumulator range
    net.FIND_SHIFTS = (30,) + tuple(net.FIND_SHIFTS[1:])
    net._ctx_key = None
    with pytest.raises((AssertionError, lib.YoloB200Error)):
        net(frames[:1].cuda(), quantization=True, find=True)


# ---- (2) oracle on seeded inputs: contract F, all rounding modes, random tables -----------------------

def random_tables(rng, qnet):
    """Random but valid exponent tables (bit-exactness must hold for ANY table)."""
    L = len(qnet.layers)
    sa = [int(rng.integers(0, 9)) for _ in range(L + 1)]
    sw = [int(rng.integers(4, 12)) for _ in range(L)]
    sb = [int(rng.integers(2, 12)) for _ in range(L)]
    rt = [int(rng.integers(6, 14)) for _ in range(L)]
    return sa, sw, sb, rt


@pytest.mark.parametrize("mode", [lib.ROUND_RNE, lib.ROUND_FLOOR, lib.ROUND_HALF_UP])
@pytest.mark.parametrize("tables", ["calibrated", "shipped", "random0", "random1"])
def test_contract_f_against_oracle(ctx, mode, tables):
    g, qnet, frames = gu.load("ref_p_64x96")
    import copy
    q = copy.deepcopy(qnet)
    import zlib
    rng = np.random.default_rng(zlib.crc32(('%d-%s' % (mode, tables)).encode()))
    if tables == "shipped":
        q.sa, q.sw, q.sb, q.retune = list(ex.SHIPPED_SCALE_A), list(ex.SHIPPED_SCALE_W), list(ex.SHIPPED_SCALE_B), list(ex.SHIPPED_RETUNE)
    elif tables.startswith("random"):
        q.sa, q.sw, q.sb, q.retune = random_tables(rng, q)
    ctx.load_quantnet(q, contract=lib.CONTRACT_F, round_mode=mode)
    x8 = rng.integers(-128, 128, (2, 48, 80, 4), dtype=np.int8)
    x8[..., 3] = 0
    outs, _ = run_backbone(ctx, x8)
    ref, _ = ol.backbone(q, x8, contract=0, round_mode=mode)
    for l, (a, b) in enumerate(zip(outs, ref)):
        np.testing.assert_array_equal(a, b, err_msg="layer %d" % l)


@pytest.mark.parametrize("hw", [(26, 26), (13, 13), (15, 20), (30, 40), (2, 2), (1, 1), (17, 33)])
@pytest.mark.parametrize("layer", [0, 1, 2, 3, 4, 5, 6, 8, 9])
def test_single_layer_ragged_shapes(ctx, hw, layer):
    """Per-layer entry point (first_conv...conv_last replacement) on odd / tiny maps, both contracts."""
    g, qnet, frames = gu.load("ref_p_64x96")
    cin, cout, activ, pool = qnet.layers[layer]
    h, w = hw
    if pool and (h < 2 or w < 2):
        pytest.skip("pooled layer needs a 2x2 map")
    rng = np.random.default_rng(layer * 100 + h)
    cs_in = ex.cstride(cin)
    x = np.zeros((3, h, w, cs_in), dtype=np.int8)
    x[..., :cin] = rng.integers(-128, 128, (3, h, w, cin), dtype=np.int8)
    for contract in (lib.CONTRACT_F, lib.CONTRACT_P):
        ctx.load_quantnet(qnet, contract=contract)
        oh, ow = (h // 2, w // 2) if pool else (h, w)
        d_out = torch.full((3, oh, ow, ex.cstride(cout)), 77, dtype=torch.int8, device="cuda")
        ctx.conv_layer(layer, dev(x), 3, h, w, d_out)
        ctx.sync()
        ref, _ = ol.conv_layer(x, qnet.w[layer], qnet.b[layer], cin, cout, qnet.sa[layer], qnet.sw[layer], qnet.sb[layer],
                               qnet.retune[layer], qnet.sa[layer + 1], activ, pool, contract)
        np.testing.assert_array_equal(d_out.cpu().numpy(), ref)


@pytest.mark.parametrize("nhw", [(3, 32, 36), (2, 48, 100), (2, 34, 38), (5, 208, 208), (1, 416, 416)])
def test_chain_kernels_on_odd_geometries(ctx, nhw):
    """The whole chain (conv1 writing x-split rows for the row-pair kernel where its output width is even, plain rows where it is
    odd: 34x38 -> 17x19; the CTA-pair kernel on 13x13 / 26x26 maps with odd and even tile counts) against the oracle on all maps."""
    g, qnet, frames = gu.load("ref_p_64x96")
    n, h, w = nhw
    rng = np.random.default_rng(n * 1000 + h + w)
    x8 = rng.integers(-128, 128, (n, h, w, 4), dtype=np.int8)
    x8[..., 3] = 0
    for contract in (lib.CONTRACT_F, lib.CONTRACT_P):
        ctx.load_quantnet(qnet, contract=contract)
        outs, _ = run_backbone(ctx, x8)
        ref, _ = ol.backbone(qnet, x8, contract=contract)
        for l, (a, b) in enumerate(zip(outs, ref)):
            np.testing.assert_array_equal(a, b, err_msg="layer %d" % l)


@pytest.mark.parametrize("n", [1, 4, 5, 7])
@pytest.mark.parametrize("hw", [(26, 26), (13, 13), (15, 20)])
@pytest.mark.parametrize("layer", [6, 7])
def test_cta_pair_kernel_tile_counts(ctx, layer, hw, n):
    """conv5 / conv6 shapes on the CTA-pair kernel (tcgen05.mma.cta_group::2) for batch sizes that give odd tile counts (the pair's
    second tile is then a dummy) and fewer pairs than SMs."""
    g, qnet, frames = gu.load("ref_p_64x96")
    cin, cout, activ, pool = qnet.layers[layer]
    h, w = hw
    rng = np.random.default_rng(layer * 100 + h + n)
    x = rng.integers(-128, 128, (n, h, w, cin), dtype=np.int8)
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_F)
    d_out = torch.full((n, h, w, ex.cstride(cout)), 77, dtype=torch.int8, device="cuda")
    ctx.conv_layer(layer, dev(x), n, h, w, d_out)
    ctx.sync()
    ref, _ = ol.conv_layer(x, qnet.w[layer], qnet.b[layer], cin, cout, qnet.sa[layer], qnet.sw[layer], qnet.sb[layer],
                           qnet.retune[layer], qnet.sa[layer + 1], activ, pool, lib.CONTRACT_F)
    np.testing.assert_array_equal(d_out.cpu().numpy(), ref)


def test_saturation_extremes(ctx):
    """All-max / all-min inputs drive the 16-bit and 8-bit saturation points of contract F."""
    g, qnet, frames = gu.load("ref_p_64x96")
    import copy
    q = copy.deepcopy(qnet)
    q.sa, q.sw, q.sb, q.retune = list(ex.SHIPPED_SCALE_A), list(ex.SHIPPED_SCALE_W), list(ex.SHIPPED_SCALE_B), list(ex.SHIPPED_RETUNE)
    ctx.load_quantnet(q, contract=lib.CONTRACT_F)
    for fill in (127, -128):
        x8 = np.full((1, 32, 32, 4), fill, dtype=np.int8)
        x8[..., 3] = 0
        outs, _ = run_backbone(ctx, x8)
        ref, _ = ol.backbone(q, x8, contract=0)
        for l, (a, b) in enumerate(zip(outs, ref)):
            np.testing.assert_array_equal(a, b, err_msg="layer %d" % l)


def test_contract_p_overflow_counter(ctx):
    """An input hotter than the calibration saturates int8; the library counts it (the reference never clamps)."""
    g, qnet, frames = gu.load("ref_p_64x96")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_P)
    x8 = np.full((1, 32, 32, 4), 127, dtype=np.int8)
    outs, _ = run_backbone(ctx, x8)
    ref, ovf = ol.backbone(qnet, x8, contract=1)
    for a, b in zip(outs, ref):
        np.testing.assert_array_equal(a, b)
    assert ctx.overflow_count() == ovf and ovf > 0
    assert ctx.overflow_count() == 0      # resets on read


# ---- RGB444 front end and the C path configuration (BASELINE configs[0]) -----------------------------

def test_rgb444_frontend_and_shipped_tables_320x240(ctx):
    """One synthetic camera frame (uint16 [240][320], 0x0BGR) through the as-shipped tables, contract F, against the
    oracle: LUT, every layer's int8 map, and the detection list."""
    g, qnet, frames = gu.load("ref_p_64x96")
    import copy
    q = copy.deepcopy(qnet)
    q.sa, q.sw, q.sb, q.retune = list(ex.SHIPPED_SCALE_A), list(ex.SHIPPED_SCALE_W), list(ex.SHIPPED_SCALE_B), list(ex.SHIPPED_RETUNE)
    q.anchors = [list(a) for a in ex.ANCHOR_SIZE_COCO]
    ctx.load_quantnet(q, contract=lib.CONTRACT_F, conf_thresh=0.01, nms_thresh=0.5)
    np.testing.assert_array_equal(ctx.rgb444_lut(), ol.rgb444_lut(q.sa[0]))
    rng = np.random.default_rng(0)
    cases = [rng.integers(0, 4096, (240, 320), dtype=np.uint16), np.zeros((240, 320), np.uint16),
             np.full((240, 320), 0x0fff, np.uint16), (np.arange(320, dtype=np.uint16)[None, :] * 12 % 4096).repeat(240, 0)]
    fr = np.stack(cases)
    d8 = torch.empty((4, 240, 320, 4), dtype=torch.int8, device="cuda")
    ctx.quantize_rgb444(dev(fr), 4, 240, 320, d8)
    ctx.sync()
    x8 = d8.cpu().numpy()
    np.testing.assert_array_equal(x8, ol.quantize_rgb444(fr, q.sa[0]))
    outs, (pred, gh, gw) = run_backbone(ctx, x8)
    assert (gh, gw) == (15, 20)
    ref, _ = ol.backbone(q, x8, contract=0)
    for l, (a, b) in enumerate(zip(outs, ref)):
        np.testing.assert_array_equal(a, b, err_msg="layer %d" % l)
    dets, counts = ctx.forward_rgb444(fr)
    for i in range(4):
        (ob, os_, oc, oidx), ocnt = ol.head_python(ref[-1][i], 5, 2, q.sa[10], q.anchors, 16, 240, 320, 0.01, 0.5)
        b, s, c, idx = lib.dets_to_arrays(dets[i], int(counts[i]))
        assert counts[i] == ocnt
        np.testing.assert_array_equal(idx, oidx)
        np.testing.assert_allclose(s, os_, atol=1e-5, rtol=0)
        np.testing.assert_allclose(b, ob, atol=1e-5, rtol=0)


def test_c_head_mode_against_oracle(ctx):
    """HEAD_C: decode, strict threshold, conf_sort's order INCLUDING the order its swap history leaves among equal scores
    (yolo_forward.c:1114-1126; pinned on the reference's compiled function by
    tests/test_oracle.py::test_c_sort_nms_tie_order_matches_reference_c), class-agnostic NMS: identical lists, no tolerance.
    Narrow value ranges make most scores collide; the wide range exercises the tie-free fast path."""
    g, qnet, frames = gu.load("ref_p_64x96")
    import copy
    q = copy.deepcopy(qnet)
    q.anchors = [list(a) for a in ex.ANCHOR_SIZE_COCO]
    rng = np.random.default_rng(5)
    saw_ties = 0
    for trial, (thresh, span) in enumerate(((0.01, 60), (0.2, 60), (0.3, 60), (0.01, 3), (0.05, 8), (0.1, 20), (0.0, 1))):
        ctx.load_quantnet(q, contract=lib.CONTRACT_F, head_mode=lib.HEAD_C, conf_thresh=thresh, nms_thresh=0.5, max_det=4096)
        for (n, gh, gw) in ((2, 15, 20), (1, 26, 26)):
            pred = np.zeros((n, gh, gw, 48), dtype=np.int8)
            pred[..., :35] = rng.integers(-span, span, (n, gh, gw, 35), dtype=np.int8)
            dets, counts = det_arrays(ctx, dev(pred), n, gh, gw, gh * 16, gw * 16)
            for i in range(n):
                (ob, os_, oc, oidx), ocnt = ol.head_c(pred[i], 5, q.sa[10], q.anchors, 16, thresh, 0.5)
                b, s, c, idx = lib.dets_to_arrays(dets[i], int(counts[i]))
                saw_ties += len(set(os_.tolist())) != len(os_)
                assert counts[i] == ocnt
                np.testing.assert_array_equal(idx, oidx)
                np.testing.assert_array_equal(c, oc)
                np.testing.assert_allclose(s, os_, atol=1e-6, rtol=0)
                np.testing.assert_array_equal(b, ob)
    assert saw_ties >= 6


@pytest.mark.parametrize("sa_pred", [2, 4, 6])
@pytest.mark.parametrize("nms_thresh", [0.3, 0.5, 0.7])
def test_python_head_random_prediction_maps_against_oracle(ctx, sa_pred, nms_thresh):
    """Decode + per-class greedy NMS on random prediction maps (dense, heavily overlapping boxes of every size, both the
    spatially indexed and the exhaustive suppression paths) against the oracle: identical kept sets, classes, order."""
    import copy
    g, qnet, frames = gu.load("ref_p_64x96")
    q = copy.deepcopy(qnet)
    q.sa = list(q.sa)
    q.sa[10] = sa_pred
    rng = np.random.default_rng(1000 * sa_pred + int(nms_thresh * 10))
    ctx.load_quantnet(q, contract=lib.CONTRACT_F, head_mode=lib.HEAD_PYTHON, conf_thresh=0.05, nms_thresh=nms_thresh, max_det=4096)
    for (n, gh, gw) in ((2, 26, 26), (3, 13, 13), (2, 15, 20)):
        pred = np.zeros((n, gh, gw, 48), dtype=np.int8)
        pred[..., :35] = rng.integers(-128, 128, (n, gh, gw, 35), dtype=np.int8)
        pred[0, : gh // 2, :, 25:35:4] = -128          # a region of tiny boxes
        dets, counts = det_arrays(ctx, dev(pred), n, gh, gw, gh * 16, gw * 16)
        for i in range(n):
            (ob, os_, oc, oidx), ocnt = ol.head_python(pred[i], 5, 2, sa_pred, q.anchors, 16, gh * 16, gw * 16, 0.05, nms_thresh)
            b, s_, c, idx = lib.dets_to_arrays(dets[i], int(counts[i]))
            assert counts[i] == ocnt, "frame %d of %s: %d kept, oracle %d" % (i, (n, gh, gw), counts[i], ocnt)
            np.testing.assert_array_equal(idx, oidx)
            np.testing.assert_array_equal(c, oc)
            np.testing.assert_allclose(s_, os_, atol=1e-6, rtol=0)
            np.testing.assert_allclose(b, ob, atol=1e-6, rtol=0)


# ---- batch semantics / edge cases --------------------------------------------------------------------

def test_empty_batch_and_errors(ctx):
    g, qnet, frames = gu.load("ref_p_64x96")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_P)
    dets, counts = ctx.forward_int8(np.zeros((0, 64, 96, 4), np.int8))
    assert dets.shape[0] == 0 and counts.shape[0] == 0
    L = ctx.L
    assert L.yolo_b200_conv_layer(ctx._h, 99, None, 1, 8, 8, None) < 0
    assert b"out of range" in L.yolo_b200_last_error()
    assert L.yolo_b200_forward_int8(ctx._h, None, 1, 64, 96, None, None) < 0
    # an input too large for the per-frame candidate buffer is refused, not truncated
    big = np.zeros((1, 16 * 40, 16 * 40, 4), np.int8)
    with pytest.raises(lib.YoloB200Error):
        ctx.forward_int8(big)
    fresh = lib.Context(0)
    try:
        out = np.zeros(1, np.int32)
        rc = fresh.L.yolo_b200_forward_int8(fresh._h, np.zeros((1, 32, 32, 4), np.int8).ctypes.data, 1, 32, 32,
                                            out.ctypes.data, out.ctypes.data)
        assert rc < 0 and b"no network loaded" in fresh.L.yolo_b200_last_error()
        # malformed tables are rejected at load
        bad = lib.make_params(qnet)
        bad.layers[3].cin = 7
        with pytest.raises(lib.YoloB200Error):
            fresh.load(qnet.w, qnet.b, bad)
    finally:
        fresh.close()


@pytest.mark.parametrize("contract", [lib.CONTRACT_F, lib.CONTRACT_P])
def test_bench_workload_batch256_against_oracle(ctx, contract):
    """The configuration bench.py times (BASELINE configs[2]): 256 RGB444 camera frames of 416x416, the bench's own network
    and first input set, through the device entry point AND the host entry point yolo_b200_forward_rgb444.  Frames 0, 1,
    127, 254, 255 (first / last of the canvas, a middle one) are compared with the oracle: every layer's map and the
    detection list; all 256 counts and lists must agree between the two entry points."""
    B, H, W = 256, 416, 416
    check = [0, 1, 127, 254, 255]
    qnet = ex.random_quantnet(seed=0, calib_hw=(H, W), calib_frames=2, calib_input="rgb444")      # bench.make_qnet()
    ctx.load_quantnet(qnet, contract=contract, round_mode=lib.ROUND_RNE, conf_thresh=0.1, nms_thresh=0.5, max_det=4096)
    frames = ex.synthetic_frames_rgb444(B, H, W, seed=0)                                          # bench: seed 100 * rank + set
    d = torch.from_numpy(frames.view(np.int16)).cuda()
    d_dets = torch.zeros((B, 4096, 8), dtype=torch.int32, device="cuda")
    d_counts = torch.zeros((B,), dtype=torch.int32, device="cuda")
    ctx.forward_rgb444_dev(d, B, H, W, d_dets, d_counts)
    ctx.sync()
    counts = d_counts.cpu().numpy()
    dets = d_dets.cpu().numpy().view(lib.DET_DTYPE).reshape(B, 4096)
    shapes = layer_shapes(qnet, H, W)
    x8 = ol.quantize_rgb444(frames[check], qnet.sa[0])
    ref, _ = ol.backbone(qnet, x8, contract=contract)
    for l, (oh, ow) in enumerate(shapes):
        got = ctx.layer_output(l, B, oh, ow)
        for k, f in enumerate(check):
            assert np.array_equal(got[f], ref[l][k]), "layer %d, frame %d of the 256-frame batch differs from the oracle" % (l, f)
        del got
    for k, f in enumerate(check):
        (ob, os_, oc, oidx), ocnt = ol.head_python(ref[-1][k], 5, 2, qnet.sa[10], qnet.anchors, 16, H, W, 0.1, 0.5)
        b, s_, c, idx = lib.dets_to_arrays(dets[f], int(counts[f]))
        assert counts[f] == ocnt
        np.testing.assert_array_equal(idx, oidx)
        np.testing.assert_array_equal(c, oc)
        np.testing.assert_allclose(s_, os_, atol=1e-5, rtol=0)
        np.testing.assert_allclose(b, ob, atol=1e-5, rtol=0)
    # the host entry point (chunked copies, head on a second stream) must return the same lists for all 256 frames
    hdets, hcounts = ctx.forward_rgb444(frames)
    np.testing.assert_array_equal(hcounts, counts)
    for f in range(B):
        n = int(counts[f])
        assert hdets[f][:n].tobytes() == dets[f][:n].tobytes(), "host and device entry points differ on frame %d" % f


def test_frames_are_independent_at_full_size(ctx):
    """BASELINE-size property check (416x416): a frame's result does not depend on its batch neighbours or position,
    and repeated runs are bit-identical."""
    qnet = ex.random_quantnet(seed=0, calib_hw=(416, 416), calib_frames=1)
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.3)
    rng = np.random.default_rng(9)
    x8 = rng.integers(-100, 100, (6, 416, 416, 4), dtype=np.int8)
    x8[..., 3] = 0
    outs, _ = run_backbone(ctx, x8)
    pred_all = outs[-1].copy()
    perm = np.array([3, 0, 5, 1, 4, 2])
    outs2, _ = run_backbone(ctx, x8[perm])
    np.testing.assert_array_equal(outs2[-1], pred_all[perm])
    outs3, _ = run_backbone(ctx, x8[2:3])
    np.testing.assert_array_equal(outs3[-1][0], pred_all[2])
    d1, c1 = ctx.forward_int8(x8)
    d2, c2 = ctx.forward_int8(x8)
    np.testing.assert_array_equal(c1, c2)
    assert d1.tobytes() == d2.tobytes()
    # frame 0 of the 416 golden fixture is also checked bit-for-bit above; here: oracle on the last layer of one frame
    ref, _ = ol.conv_layer(outs[-2][2:3], qnet.w[9], qnet.b[9], 256, 35, qnet.sa[9], qnet.sw[9], qnet.sb[9], qnet.retune[9],
                           qnet.sa[10], 0, 0, 0)
    np.testing.assert_array_equal(pred_all[2:3], ref)


def test_pipelined_host_entry_points_do_not_depend_on_the_chunk_size(ctx):
    """yolo_b200_forward_rgb444 / _int8 (host buffers; copies overlapped with compute chunk by chunk) return the same
    detections for every chunk size, and the same as the device-buffer entry point."""
    qnet = ex.random_quantnet(seed=3, calib_hw=(64, 96), calib_frames=2, calib_input="rgb444")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=512)
    n, h, w = 7, 64, 96
    frames = ex.synthetic_frames_rgb444(n, h, w, seed=9)
    d_dets = torch.zeros((n, 512, 8), dtype=torch.int32, device="cuda")
    d_counts = torch.zeros((n,), dtype=torch.int32, device="cuda")
    ctx.forward_rgb444_dev(dev(frames.view(np.int16)), n, h, w, d_dets, d_counts)
    ctx.sync()
    ref_d, ref_c = d_dets.cpu().numpy(), d_counts.cpu().numpy()
    assert ref_c.sum() > 0
    x8 = ol.quantize_rgb444(frames, qnet.sa[0])
    try:
        for chunk in (0, 1, 2, 3, 128):
            ctx.set_host_chunk(chunk)
            dets, counts = ctx.forward_rgb444(frames)
            np.testing.assert_array_equal(counts, ref_c)
            dets8, counts8 = ctx.forward_int8(x8)
            np.testing.assert_array_equal(counts8, ref_c)
            for i in range(n):
                k = int(ref_c[i])
                np.testing.assert_array_equal(dets[i][:k].view(np.int32).reshape(k, 8), ref_d[i, :k])
                np.testing.assert_array_equal(dets8[i][:k].view(np.int32).reshape(k, 8), ref_d[i, :k])
    finally:
        ctx.set_host_chunk(64)


def test_submit_wait_pairs_in_flight_match_the_blocking_call(ctx):
    """yolo_b200_submit_rgb444 / yolo_b200_wait with two calls in flight (the tail of one call overlaps the copy of the next)
    return exactly what the blocking entry point returns for each batch; a third submission is refused."""
    qnet = ex.random_quantnet(seed=3, calib_hw=(64, 96), calib_frames=2, calib_input="rgb444")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=512)
    n, h, w = 9, 64, 96
    batches = [torch.from_numpy(ex.synthetic_frames_rgb444(n, h, w, seed=20 + b).view(np.int16)).pin_memory() for b in range(5)]
    want = [ctx.forward_rgb444(b.numpy().view(np.uint16)) for b in batches]
    assert sum(int(c.sum()) for _, c in want) > 0
    ctx.set_host_chunk(4)
    try:
        outs = [(torch.zeros((n, 512, 8), dtype=torch.int32).pin_memory(), torch.zeros((n,), dtype=torch.int32).pin_memory()) for _ in batches]
        tickets = []
        for b, (dets, counts) in zip(batches, outs):
            if len(tickets) == 2:
                ctx.wait(tickets.pop(0))
            tickets.append(ctx.submit_rgb444(b.numpy().view(np.uint16), dets.numpy().view(lib.DET_DTYPE).reshape(n, 512), counts.numpy()))
            if len(tickets) == 2:        # both slots busy: a third call (blocking or not) is refused, nothing is disturbed
                with pytest.raises(RuntimeError):
                    ctx.forward_rgb444(batches[0].numpy().view(np.uint16))
        while tickets:
            ctx.wait(tickets.pop(0))
        with pytest.raises(RuntimeError):
            ctx.wait(0)                  # nothing in flight
        for (dets, counts), (wd, wc) in zip(outs, want):
            np.testing.assert_array_equal(counts.numpy(), wc)
            for i in range(n):
                k = int(wc[i])
                np.testing.assert_array_equal(dets.numpy()[i, :k], wd[i][:k].view(np.int32).reshape(k, 8))
    finally:
        ctx.set_host_chunk(64)


def test_fused_rgb444_front_end_edge_cases(ctx):
    """The first layer with the camera quantiser fused in: upper nibble of the pixels ignored (camera_to_inpBuf masks
    it, yolo_forward.c:87-123), frame widths that are / are not a multiple of 4 and a frame pointer that is only 2-byte
    aligned (both fall back to quantiser + generic first layer): always the layer-1 map of the oracle."""
    qnet = ex.random_quantnet(seed=5, calib_hw=(64, 96), calib_frames=2, calib_input="rgb444")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=512)
    rng = np.random.default_rng(11)
    for (n, h, w) in ((3, 64, 96), (2, 34, 52), (2, 40, 98), (1, 33, 47)):
        clean = rng.integers(0, 4096, (n, h, w), dtype=np.uint16)
        dirty = clean | (rng.integers(0, 16, (n, h, w), dtype=np.uint16) << 12)
        x8 = ol.quantize_rgb444(clean, qnet.sa[0])
        ref, _ = ol.backbone(qnet, x8, contract=0)
        for frames, shift in ((clean, 0), (dirty, 0), (dirty, 1)):
            buf = torch.zeros((n * h * w + 4,), dtype=torch.int16, device="cuda")
            view = buf[shift:shift + n * h * w]
            view.copy_(torch.from_numpy(frames.reshape(-1).view(np.int16)))
            d_dets = torch.zeros((n, 512, 8), dtype=torch.int32, device="cuda")
            d_counts = torch.zeros((n,), dtype=torch.int32, device="cuda")
            ctx.forward_rgb444_dev(view, n, h, w, d_dets, d_counts)
            ctx.sync()
            for l in (0, 1, 9):
                shp = ref[l].shape
                np.testing.assert_array_equal(ctx.layer_output(l, shp[0], shp[1], shp[2]), ref[l],
                                              err_msg="shape %s shift %d layer %d" % ((n, h, w), shift, l))


def test_uint8_image_front_end_matches_basetransform_and_f32_path(ctx):
    """yolo_b200_forward_u8bgr: BaseTransform (without resize) + BGR->RGB + tracker quantiser as a fused table lookup.
    The table equals the reference's float32 arithmetic evaluated in numpy; the quantised map equals the f32 front end
    on the transformed image; detections equal the f32 entry point's."""
    g, qnet, frames = gu.load("ref_p_64x96")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_P, conf_thresh=0.1, nms_thresh=0.5, max_det=512)
    rng = np.random.default_rng(21)
    n, h, w = 3, 64, 96
    img = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)                 # BGR, as cv2.imread delivers
    # data/__init__.py:44-52 in numpy float32, then test.py:79 (BGR -> RGB, HWC -> CHW)
    x = img.astype(np.float32)
    x /= 255.
    x -= np.array((0.406, 0.456, 0.485), dtype=np.float32)
    x /= np.array((0.225, 0.224, 0.229), dtype=np.float32)
    x = np.ascontiguousarray(x[..., (2, 1, 0)].transpose(0, 3, 1, 2))
    ref8, ovf = ol.quantize_f32(x, qnet.sa[0])
    assert ovf == 0
    lut = ctx.u8bgr_lut()
    v = np.arange(256, dtype=np.float32)
    for ch, (mean, sd) in enumerate(((0.485, 0.229), (0.456, 0.224), (0.406, 0.225))):
        t = v / np.float32(255.)
        t = (t - np.float32(mean)) / np.float32(sd)
        np.testing.assert_array_equal(lut[ch], np.rint(t * np.float32(2.0 ** qnet.sa[0])).astype(np.int8))
    d_q = torch.zeros((n, h, w, 4), dtype=torch.int8, device="cuda")
    ctx.quantize_u8bgr(dev(img), n, h, w, d_q)
    ctx.sync()
    np.testing.assert_array_equal(d_q.cpu().numpy(), ref8)
    dets_u8, counts_u8 = ctx.forward_u8bgr(img)                              # fused into the first layer
    dets_f, counts_f = ctx.forward_f32(x)
    np.testing.assert_array_equal(counts_u8, counts_f)
    assert counts_f.sum() > 0
    for i in range(n):
        k = int(counts_f[i])
        np.testing.assert_array_equal(dets_u8[i][:k].view(np.int32), dets_f[i][:k].view(np.int32))
    ctx.set_conv_backend(1)                                                  # un-fused path (stand-alone quantiser + dp4a)
    try:
        dets_b, counts_b = ctx.forward_u8bgr(img)
    finally:
        ctx.set_conv_backend(0)
    np.testing.assert_array_equal(counts_b, counts_f)


@pytest.mark.parametrize("contract", [lib.CONTRACT_P, lib.CONTRACT_F])
def test_gpu_calibration_matches_the_reference_rule(ctx, contract):
    """yolo_b200_calibrate_f32 (tracker first-call rule + overflow guard, on the GPU) derives the same exponent tables as
    the exporter's float32 restatement of the reference's calibration call, from deliberately wrong starting tables,
    and the re-programmed context then reproduces the exporter-calibrated network bit for bit."""
    import copy
    ws, bs = ex.random_float_convs(seed=4)
    frames = ex.synthetic_frames_f32(2, 64, 96, seed=77)
    ref = ex.build_quantnet(ws, bs, frames)                       # CPU: quantise weights, calibrate on `frames`
    wrong = copy.deepcopy(ref)
    wrong.sa = [max(0, e - 1) for e in ref.sa]
    wrong.retune = [r - 1 for r in ref.retune]
    ctx.load_quantnet(wrong, contract=contract, conf_thresh=0.1, nms_thresh=0.5)
    sa, rt = ctx.calibrate_f32(frames.cuda(), 2, 64, 96)
    assert sa == list(ref.sa), (sa, ref.sa)
    assert rt == list(ref.retune), (rt, ref.retune)
    test = ex.synthetic_frames_f32(2, 64, 96, seed=78)
    x8, _ = ol.quantize_f32(test.numpy(), ref.sa[0])
    outs, _ = run_backbone(ctx, x8)
    want, _ = ol.backbone(ref, x8, contract=contract)
    for l, (a_, b_) in enumerate(zip(outs, want)):
        np.testing.assert_array_equal(a_, b_, err_msg="layer %d after GPU calibration" % l)


def test_legacy_yolo_forward_symbol(ctx):
    """yolo_forward(18,22,16,20,32,16, camera, vga) as main.c:44-49 calls it: detections are drawn into the camera
    buffer and the frame is copied to the VGA buffer (yolo_forward.c:1280-1281)."""
    g, qnet, frames = gu.load("ref_p_64x96")
    import copy
    q = copy.deepcopy(qnet)
    q.sa, q.sw, q.sb, q.retune = list(ex.SHIPPED_SCALE_A), list(ex.SHIPPED_SCALE_W), list(ex.SHIPPED_SCALE_B), list(ex.SHIPPED_RETUNE)
    q.anchors = [list(a) for a in ex.ANCHOR_SIZE_COCO]
    ctx.load_quantnet(q, contract=lib.CONTRACT_F, conf_thresh=0.3, nms_thresh=0.5, max_det=64)
    ctx.set_default()
    cam = np.random.default_rng(1).integers(0, 4096, (240, 320), dtype=np.uint16)
    orig = cam.copy()
    vga = np.zeros((240, 320), dtype=np.uint16)
    ctx.L.yolo_forward(bytes([18]), bytes([22]), bytes([16]), bytes([20]), bytes([32]), bytes([16]), cam.ctypes.data, vga.ctypes.data)
    np.testing.assert_array_equal(vga, cam)
    dets, counts = ctx.forward_rgb444(orig[None])
    expect = orig.copy()
    n = int(min(counts[0], 64))
    rc = ctx.L.yolo_b200_draw_rectangles(expect.ctypes.data, 240, 320, dets[0].ctypes.data, n, 1)
    assert rc == 0
    np.testing.assert_array_equal(cam, expect)
    changed = cam != orig
    assert n == 0 or changed.any()
    assert set(np.unique(cam[changed])) <= {0x000f, 0x00f0}


# ---- the two convolution back ends (tcgen05 implicit GEMM vs integer dot product) ---------------------------

def test_tensor_core_and_dot_product_kernels_agree_at_full_size(ctx):
    """416x416, batch 3: every layer's map from the tcgen05 path equals the dp4a path bit for bit."""
    qnet = ex.random_quantnet(seed=0, calib_hw=(416, 416), calib_frames=1)
    rng = np.random.default_rng(11)
    x8 = rng.integers(-100, 100, (3, 416, 416, 4), dtype=np.int8)
    x8[..., 3] = 0
    for contract in (lib.CONTRACT_F, lib.CONTRACT_P):
        ctx.load_quantnet(qnet, contract=contract)
        ctx.set_conv_backend(1)
        ref, _ = run_backbone(ctx, x8)
        ctx.set_conv_backend(0)
        got, _ = run_backbone(ctx, x8)
        for l, (a, b) in enumerate(zip(got, ref)):
            np.testing.assert_array_equal(a, b, err_msg="layer %d" % l)
    ctx.set_conv_backend(0)


WS_LAYERS = (1, 2, 3, 4, 5, 6, 7, 8, 9)     # conv_ws.cu: weights resident in shared memory (1-5, 9) or streamed tap by tap (6-8)


@pytest.mark.parametrize("backend", [2, 4, 5])
@pytest.mark.parametrize("layer", [1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_tensor_core_layer_against_oracle(ctx, layer, backend):
    """Each tensor-core layer alone, one tcgen05 kernel forced (2 = streaming, 4/5 = weight-stationary with the fp32 /
    integer epilogue), against the CPU oracle (odd sizes, several images)."""
    if backend >= 4 and layer not in WS_LAYERS:
        pytest.skip("weights do not fit the weight-stationary kernel")
    g, qnet, frames = gu.load("ref_p_64x96")
    cin, cout, activ, pool = qnet.layers[layer]
    rng = np.random.default_rng(layer)
    for (n, h, w) in ((3, 14, 18), (5, 13, 13), (2, 40, 21)):
        x = np.zeros((n, h, w, ex.cstride(cin)), dtype=np.int8)
        x[..., :cin] = rng.integers(-128, 128, (n, h, w, cin), dtype=np.int8)
        for contract in (lib.CONTRACT_F, lib.CONTRACT_P):
            ctx.load_quantnet(qnet, contract=contract)
            ctx.set_conv_backend(backend)
            try:
                oh, ow = (h // 2, w // 2) if pool else (h, w)
                d_out = torch.full((n, oh, ow, ex.cstride(cout)), 77, dtype=torch.int8, device="cuda")
                ctx.conv_layer(layer, dev(x), n, h, w, d_out)
                ctx.sync()
            finally:
                ctx.set_conv_backend(0)
            ref, _ = ol.conv_layer(x, qnet.w[layer], qnet.b[layer], cin, cout, qnet.sa[layer], qnet.sw[layer], qnet.sb[layer],
                                   qnet.retune[layer], qnet.sa[layer + 1], activ, pool, contract)
            np.testing.assert_array_equal(d_out.cpu().numpy(), ref, err_msg="shape %s contract %d" % ((n, h, w), contract))


@pytest.mark.parametrize("pool", [1, 0])
def test_first_layer_kernel_against_oracle(ctx, pool):
    """conv1 on the warp-level integer-MMA kernel (NHWC4 input), pooled and un-pooled, odd sizes, both contracts."""
    import copy
    g, qnet0, frames = gu.load("ref_p_64x96")
    qnet = copy.deepcopy(qnet0)
    cin, cout, activ, _ = qnet.layers[0]
    qnet.layers[0] = (cin, cout, activ, pool)
    rng = np.random.default_rng(7 + pool)
    # widths that are not a multiple of 4 fall back to the dp4a kernel (the mma kernel fetches aligned pixel quads)
    for (n, h, w) in ((2, 33, 47), (3, 16, 32), (1, 50, 70), (2, 33, 44), (1, 50, 68), (2, 5, 4), (1, 18, 100)):
        x = np.zeros((n, h, w, 4), dtype=np.int8)
        x[..., :3] = rng.integers(-128, 128, (n, h, w, 3), dtype=np.int8)
        for contract in (lib.CONTRACT_F, lib.CONTRACT_P):
            ctx.load_quantnet(qnet, contract=contract)
            oh, ow = (h // 2, w // 2) if pool else (h, w)
            outs = []
            for backend in (0, 1):
                ctx.set_conv_backend(backend)
                try:
                    d_out = torch.full((n, oh, ow, 16), 77, dtype=torch.int8, device="cuda")
                    ctx.conv_layer(0, dev(x), n, h, w, d_out)
                    ctx.sync()
                finally:
                    ctx.set_conv_backend(0)
                outs.append(d_out.cpu().numpy())
            ref, _ = ol.conv_layer(x, qnet.w[0], qnet.b[0], cin, cout, qnet.sa[0], qnet.sw[0], qnet.sb[0],
                                   qnet.retune[0], qnet.sa[1], activ, pool, contract)
            np.testing.assert_array_equal(outs[0], ref, err_msg="mma kernel, shape %s contract %d" % ((n, h, w), contract))
            np.testing.assert_array_equal(outs[1], ref, err_msg="dp4a kernel, shape %s contract %d" % ((n, h, w), contract))


@pytest.mark.parametrize("pool_shape", [(1, 2, 256), (2, 6, 264), (3, 50, 272), (1, 33, 320), (2, 130, 416), (1, 20, 640), (1, 416, 416)])
def test_first_layer_tcgen05_kernel_against_oracle(ctx, pool_shape):
    """conv1 on the tcgen05 slot kernel (conv_fs.cu: pooled, 16 channels, pooled width >= 128): int8 NHWC4 input through the
    single-layer entry point under both contracts, and RGB444 frames through the fused front end (x-split hand-off to conv2
    where the chain takes it): layer-1 and layer-2 maps of the oracle.  Shapes: one tile, tiles that cross rows and frames,
    odd heights, a ragged last tile, a ring of 8 rows (640 wide)."""
    g, qnet0, frames = gu.load("ref_p_64x96")
    n, h, w = pool_shape
    rng = np.random.default_rng(100 + h + w)
    cin, cout, activ, pool = qnet0.layers[0]
    assert pool == 1 and cout == 16
    x = np.zeros((n, h, w, 4), dtype=np.int8)
    x[..., :3] = rng.integers(-128, 128, (n, h, w, 3), dtype=np.int8)
    for contract in (lib.CONTRACT_F, lib.CONTRACT_P):
        ctx.load_quantnet(qnet0, contract=contract)
        d_out = torch.full((n, h // 2, w // 2, 16), 77, dtype=torch.int8, device="cuda")
        ctx.conv_layer(0, dev(x), n, h, w, d_out)
        ctx.sync()
        ref, _ = ol.conv_layer(x, qnet0.w[0], qnet0.b[0], cin, cout, qnet0.sa[0], qnet0.sw[0], qnet0.sb[0],
                               qnet0.retune[0], qnet0.sa[1], activ, pool, contract)
        np.testing.assert_array_equal(d_out.cpu().numpy(), ref, err_msg="shape %s contract %d" % ((n, h, w), contract))
    if h < 32:
        return
    qnet = ex.random_quantnet(seed=5, calib_hw=(64, 96), calib_frames=2, calib_input="rgb444")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.1, nms_thresh=0.5, max_det=512)
    u16 = rng.integers(0, 65536, (n, h, w), dtype=np.uint16)
    x8 = ol.quantize_rgb444(u16 & 0x0fff, qnet.sa[0])
    ref, _ = ol.backbone(qnet, x8, contract=0)
    d = torch.from_numpy(u16.reshape(-1).view(np.int16)).cuda()
    d_dets = torch.zeros((n, 512, 8), dtype=torch.int32, device="cuda")
    d_counts = torch.zeros((n,), dtype=torch.int32, device="cuda")
    ctx.forward_rgb444_dev(d, n, h, w, d_dets, d_counts)
    ctx.sync()
    for l in (0, 1):
        shp = ref[l].shape
        np.testing.assert_array_equal(ctx.layer_output(l, shp[0], shp[1], shp[2]), ref[l], err_msg="rgb444 shape %s layer %d" % ((n, h, w), l))
    if w % 16:
        return
    # uint8 BGR images (BaseTransform without resize + tracker quantiser as three byte tables, fused into the same kernel)
    ctx.load_quantnet(qnet0, contract=lib.CONTRACT_P, conf_thresh=0.1, nms_thresh=0.5, max_det=512)
    img = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    lut = ctx.u8bgr_lut()                                                     # [R, G, B][256] int8
    x8 = np.zeros((n, h, w, 4), dtype=np.int8)
    for c in range(3):
        x8[..., c] = lut[c][img[..., 2 - c]]
    ref, _ = ol.backbone(qnet0, x8, contract=1)
    ctx.forward_u8bgr(img)
    for l in (0, 1):
        shp = ref[l].shape
        np.testing.assert_array_equal(ctx.layer_output(l, shp[0], shp[1], shp[2]), ref[l], err_msg="u8 bgr shape %s layer %d" % ((n, h, w), l))


def test_first_layer_tcgen05_kernel_32_channels_many_tiles_per_cta(ctx):
    """darknet19's first layer (3 -> 32 + pool) on the tcgen05 slot kernel with 40 frames of 416x416: ~90 tiles per CTA over four
    TMEM buffers (the configuration in which issuer warps that shared buffers once ran two uses ahead of the drain and hung);
    frames 0 / 17 / 39 against the oracle, both contracts."""
    qnet = ex.random_quantnet_yolo_v2(seed=0, calib_hw=(64, 96), calib_frames=1)
    cin, cout, activ, pool = qnet.layers[0]
    assert (cout, pool) == (32, 1)
    n, h, w = 40, 416, 416
    rng = np.random.default_rng(5)
    x = np.zeros((n, h, w, 4), dtype=np.int8)
    x[..., :3] = rng.integers(-128, 128, (n, h, w, 3), dtype=np.int8)
    for contract in (lib.CONTRACT_F, lib.CONTRACT_P):
        ctx.load_quantnet(qnet, contract=contract)
        d_out = torch.full((n, h // 2, w // 2, 32), 77, dtype=torch.int8, device="cuda")
        for _ in range(3):
            ctx.conv_layer(0, dev(x), n, h, w, d_out)
        ctx.sync()
        got = d_out.cpu().numpy()
        for f in (0, 17, 39):
            ref, _ = ol.conv_layer(x[f:f + 1], qnet.w[0], qnet.b[0], cin, cout, qnet.sa[0], qnet.sw[0], qnet.sb[0],
                                   qnet.retune[0], qnet.sa[1], activ, pool, contract)
            np.testing.assert_array_equal(got[f:f + 1], ref, err_msg="frame %d contract %d" % (f, contract))


def test_weight_stationary_layers_with_cp_async_producers_in_a_subprocess():
    """YOLO_B200_WS_TMA=0 (read once per process) keeps the cp.async producers for conv3_1 / conv4_1 / conv4_2: same
    results as the oracle (the TMA-fed default is what every other test runs)."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np, torch
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import yolo_b200
        from yolo_b200 import lib, export as ex
        import golden_util as gu, oracle_lib as ol
        g, qnet, frames = gu.load("ref_p_64x96")
        ctx = lib.Context(0)
        rng = np.random.default_rng(3)
        for layer in (2, 4, 5):
            cin, cout, activ, pool = qnet.layers[layer]
            for (n, h, w) in ((3, 14, 18), (2, 40, 20)):
                x = np.zeros((n, h, w, ex.cstride(cin)), dtype=np.int8)
                x[..., :cin] = rng.integers(-128, 128, (n, h, w, cin), dtype=np.int8)
                ctx.load_quantnet(qnet, contract=lib.CONTRACT_F)
                oh, ow = (h // 2, w // 2) if pool else (h, w)
                d_out = torch.zeros((n, oh, ow, ex.cstride(cout)), dtype=torch.int8, device="cuda")
                ctx.conv_layer(layer, torch.from_numpy(x).cuda(), n, h, w, d_out); ctx.sync()
                ref, _ = ol.conv_layer(x, qnet.w[layer], qnet.b[layer], cin, cout, qnet.sa[layer], qnet.sw[layer], qnet.sb[layer],
                                       qnet.retune[layer], qnet.sa[layer + 1], activ, pool, lib.CONTRACT_F)
                assert (d_out.cpu().numpy() == ref).all(), (layer, n, h, w)
        print("OK")
    """) % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, YOLO_B200_WS_TMA="0"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


# ---- the epilogue arithmetic alone: exact-fp32 fast paths and the integer path vs the oracle -----------------

def _edge_accumulators(rng, n=200000):
    """Accumulators that stress rounding ties, both saturation points and the fp32 exactness boundary."""
    parts = [rng.integers(-2 ** 25, 2 ** 25, n // 4), rng.integers(-2 ** 17, 2 ** 17, n // 4), rng.integers(-4096, 4096, n // 4)]
    edges = []
    for e in range(0, 26):
        for d in (-3, -2, -1, 0, 1, 2, 3):
            edges += [2 ** e + d, -(2 ** e) + d, 3 * 2 ** e // 2 + d, -(3 * 2 ** e // 2) + d]
    parts.append(np.array(edges * 8, dtype=np.int64))
    a = np.concatenate(parts).astype(np.int64)
    return np.clip(a, -(2 ** 25), 2 ** 25).astype(np.int32)


@pytest.mark.parametrize("contract", [lib.CONTRACT_F, lib.CONTRACT_P])
@pytest.mark.parametrize("tables", ["calibrated", "shipped", "random0", "random1", "random2", "random3"])
def test_epilogue_arithmetic_against_oracle(ctx, contract, tables):
    g, qnet, frames = gu.load("ref_p_64x96")
    import copy
    import zlib
    q = copy.deepcopy(qnet)
    rng = np.random.default_rng(zlib.crc32(('%d-%s' % (contract, tables)).encode()))
    if tables == "shipped":
        q.sa, q.sw, q.sb, q.retune = list(ex.SHIPPED_SCALE_A), list(ex.SHIPPED_SCALE_W), list(ex.SHIPPED_SCALE_B), list(ex.SHIPPED_RETUNE)
    elif tables.startswith("random"):
        q.sa, q.sw, q.sb, q.retune = random_tables(rng, q)
    try:
        ctx.load_quantnet(q, contract=contract)
    except lib.YoloB200Error:
        pytest.skip("table outside the supported exponent range")
    acc = _edge_accumulators(rng)
    d_acc = dev(acc)
    L = ol.lib()
    fast_seen = 0
    for layer in range(len(q.layers)):
        cin, cout, activ, pool = q.layers[layer]
        ref = np.array([L.oracle_requant(int(a), int(q.b[layer][i % cout]), q.sa[layer], q.sw[layer], q.sb[layer], q.retune[layer],
                                         q.sa[layer + 1], activ, contract, 0) for i, a in enumerate(acc[:20000])], dtype=np.int8)
        for force in (False, True):
            d_out = torch.zeros(acc.shape, dtype=torch.int8, device="cuda")
            epi = ctx.debug_requant(layer, d_acc, acc.size, d_out, force_generic=force)
            ctx.sync()
            out = d_out.cpu().numpy()
            fast_seen += epi > 0
            np.testing.assert_array_equal(out[:20000], ref, err_msg="layer %d epi %d" % (layer, epi))
            if not force:
                fast = out
            else:
                np.testing.assert_array_equal(fast, out, err_msg="fp32 and integer epilogues differ, layer %d" % layer)
    if tables in ("calibrated",):
        assert fast_seen >= 9, "the exact-fp32 epilogue should apply to calibrated tables"


def test_detect_cli_on_synthetic_images(tmp_path):
    """tools/detect.py (the test.py / demo.py counterpart) runs end to end and writes one annotated image per input."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "detect.py"), "--trained_model", "random", "--images", "synthetic:3",
                        "-size", "160", "--out", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".jpg")]) == 3
    assert "3 images" in r.stdout


def test_graft_entry_smoke_runs():
    """__graft_entry__.smoke() (what the driver runs on the GPU box): host entry point + layer read-back + oracle check."""
    import importlib, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    g = importlib.import_module("__graft_entry__")
    g.smoke()


# ---- image resize in front of the path (cv2.resize of base_transform, data/__init__.py:36) ---------------------------

def _resize_oracle():
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle_resize_u8", os.path.join(ROOT, "oracle", "resize_u8.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _gpu_resize(ctx, imgs, dh, dw, misalign=0):
    n, sh, sw, _ = imgs.shape
    buf = torch.full((n * dh * dw * 3 + 8,), 0xAB, dtype=torch.uint8, device="cuda")
    out = buf[misalign:misalign + n * dh * dw * 3]
    ctx.resize_u8bgr(dev(imgs), n, sh, sw, out, dh, dw)
    ctx.sync()
    assert int(buf[misalign + n * dh * dw * 3]) == 0xAB                      # nothing written past the end
    return out.cpu().numpy().reshape(n, dh, dw, 3)


@pytest.mark.gpu
def test_resize_kernel_matches_cv2_golden_and_oracle(ctx):
    """yolo_b200_resize_u8bgr == cv2.resize(image, (w, h)) of the reference's base_transform: bit-exact against cv2's own
    outputs (tests/golden/resize_cv2.npz) and against the oracle on geometries that take both kernel variants
    (4 pixels per thread / 1 pixel per thread: odd widths, unaligned destination), batches, up- and down-scaling, the
    exact-2x case where OpenCV switches to the 2x2 mean, single-pixel sources and the reference's 480x640 -> 416x416."""
    ro = _resize_oracle()
    g = np.load(os.path.join(ROOT, "tests", "golden", "resize_cv2.npz"))
    for seed, sh, sw, dh, dw in g["cases"].tolist():
        img = np.random.default_rng(seed).integers(0, 256, (1, sh, sw, 3), dtype=np.uint8)
        np.testing.assert_array_equal(_gpu_resize(ctx, img, dh, dw)[0], g["out_%d" % seed], err_msg="golden %dx%d -> %dx%d" % (sh, sw, dh, dw))
    rng = np.random.default_rng(17)
    cases = [(3, 48, 64, 64, 96, 0), (2, 37, 53, 33, 47, 0), (2, 37, 53, 32, 48, 1), (1, 64, 96, 32, 48, 3), (4, 5, 7, 21, 19, 0),
             (1, 1, 1, 4, 8, 0), (2, 90, 31, 8, 100, 2), (1, 3, 200, 64, 4, 0), (2, 480, 640, 416, 416, 0), (1, 1080, 1920, 416, 416, 0),
             (1, 832, 832, 416, 416, 0), (1, 240, 320, 416, 416, 0)]
    for n, sh, sw, dh, dw, mis in cases:
        imgs = rng.integers(0, 256, (n, sh, sw, 3), dtype=np.uint8)
        got = _gpu_resize(ctx, imgs, dh, dw, misalign=mis)
        np.testing.assert_array_equal(got, ro.resize_batch(imgs, dh, dw), err_msg="%d x %dx%d -> %dx%d (+%d)" % (n, sh, sw, dh, dw, mis))
    # constant images stay constant (weights sum to 2048 everywhere), n = 0 is a no-op, bad shapes are errors
    flat = np.full((1, 23, 29, 3), 255, dtype=np.uint8)
    assert (_gpu_resize(ctx, flat, 40, 44) == 255).all()
    ctx.resize_u8bgr(dev(flat), 0, 23, 29, torch.zeros(4, dtype=torch.uint8, device="cuda"), 4, 4)
    with pytest.raises(Exception):
        ctx.resize_u8bgr(dev(flat), 1, 23, 29, torch.zeros(4, dtype=torch.uint8, device="cuda"), 0, 4)


@pytest.mark.gpu
def test_forward_u8bgr_resize_equals_basetransform_then_forward(ctx):
    """yolo_b200_forward_u8bgr_resize (host images of camera size -> detections) == the reference's order of operations:
    cv2.resize on the host (oracle), then the at-network-size entry point.  Chunked host path (ragged last chunk), the
    device entry point and the identity geometry are covered; detections must be bit-identical."""
    ro = _resize_oracle()
    g, qnet, frames = gu.load("ref_p_64x96")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_P, conf_thresh=0.1, nms_thresh=0.5, max_det=512)
    rng = np.random.default_rng(23)
    n, sh, sw, h, w = 5, 48, 64, 64, 96
    imgs = rng.integers(0, 256, (n, sh, sw, 3), dtype=np.uint8)
    small = ro.resize_batch(imgs, h, w)
    dets_ref, counts_ref = ctx.forward_u8bgr(small)
    assert counts_ref.sum() > 0

    def same(dets, counts):
        np.testing.assert_array_equal(counts, counts_ref)
        for i in range(n):
            k = int(counts_ref[i])
            np.testing.assert_array_equal(dets[i][:k].view(np.int32), dets_ref[i][:k].view(np.int32))

    for chunk in (64, 2, 1):
        ctx.set_host_chunk(chunk)
        try:
            same(*ctx.forward_u8bgr_resize(imgs, (h, w)))
        finally:
            ctx.set_host_chunk(64)
    md = ctx.params.max_det
    d_dets = torch.zeros((n, md, 8), dtype=torch.int32, device="cuda")
    d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    ctx.forward_u8bgr_resize_dev(dev(imgs), n, sh, sw, h, w, d_dets, d_counts)
    ctx.sync()
    np.testing.assert_array_equal(d_counts.cpu().numpy(), counts_ref)
    for i in range(n):
        k = int(counts_ref[i])
        np.testing.assert_array_equal(d_dets[i, :k].cpu().numpy().reshape(-1), dets_ref[i][:k].view(np.int32).reshape(-1))
    same(*ctx.forward_u8bgr_resize(small, (h, w)))                           # identity geometry: no resize
    # a second geometry on the same context rebuilds the tap tables
    imgs2 = rng.integers(0, 256, (2, 100, 80, 3), dtype=np.uint8)
    d2, c2 = ctx.forward_u8bgr_resize(imgs2, (h, w))
    d2r, c2r = ctx.forward_u8bgr(ro.resize_batch(imgs2, h, w))
    np.testing.assert_array_equal(c2, c2r)
    for i in range(2):
        np.testing.assert_array_equal(d2[i][:int(c2r[i])].view(np.int32), d2r[i][:int(c2r[i])].view(np.int32))


@pytest.mark.gpu
def test_detect_cli_video_mode(tmp_path):
    """tools/detect.py --video (demo.py's video mode, :123-158): a synthetic MJPG clip goes through the GPU front end in
    batches (ragged last batch) and comes back as an annotated clip with the same number of frames and frame size."""
    import subprocess
    import sys
    cv2 = pytest.importorskip("cv2")
    src = str(tmp_path / "clip.avi")
    w = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"MJPG"), 10.0, (96, 72))
    if not w.isOpened():
        pytest.skip("this OpenCV build cannot write MJPG")
    rng = np.random.default_rng(2)
    for _ in range(7):
        w.write(rng.integers(0, 256, (72, 96, 3), dtype=np.uint8))
    w.release()
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "detect.py"), "--trained_model", "random", "--video", src,
                        "-size", "64", "--batch", "3", "--out", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "7 frames" in r.stdout
    cap = cv2.VideoCapture(str(out / "detections.avi"))
    n = 0
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        assert frame.shape == (72, 96, 3)
        n += 1
    assert n == 7


@pytest.mark.gpu
def test_peer_collector_packs_and_collects_the_filled_lists(ctx):
    """Multi-GPU collection path on one GPU (world 1: rank 0 writes its own slot through the same IPC buffer and copy
    stream): the packed records and offsets that arrive equal the filled parts of the fixed-capacity lists, for several
    pipelined steps including empty frames."""
    from yolo_b200 import runner
    g, qnet, frames = gu.load("ref_p_64x96")
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_P, conf_thresh=0.1, nms_thresh=0.5, max_det=512)
    n, N = 5, 4 * 6 * 5
    pc = runner.PeerCollector(ctx, n, 512, N, torch.device("cuda", 0), depth=3)
    rng = np.random.default_rng(11)
    want = {}
    for step in range(5):
        buf = pc.buffers(step)
        x8 = rng.integers(-100, 100, (n, 64, 96, 4), dtype=np.int8)
        x8[..., 3] = 0
        if step == 2:
            x8[1] = 0
        ctx.forward_int8_dev(dev(x8), n, 64, 96, buf.dets, buf.counts)
        pc.launch(step)
        ctx.sync()
        want[step] = (buf.counts.cpu().numpy().copy(), buf.dets.cpu().numpy().copy())
    pc.finish()
    for step in (2, 3, 4):                       # the ring holds the last three steps
        (off, rec), = pc.collected(step)
        counts, dets = want[step]
        off = off.cpu().numpy(); rec = rec.cpu().numpy()
        assert off[0] == 0 and np.array_equal(np.diff(off), np.minimum(counts, 512))
        for f in range(n):
            np.testing.assert_array_equal(rec[off[f]:off[f + 1]], dets[f, :counts[f]])
    assert pc.sent_bytes > 0
    pc.close()


# ---- yolo_v2 / darknet19 (BASELINE configs[4]): 1x1 layers, 512 / 1024 / 1280 channels, route + reorg + concat ---------------

def _run_graph(ctx, qnet, x8):
    n, h, w, _ = x8.shape
    pred, gh, gw = ctx.backbone(dev(x8), n, h, w)
    ctx.sync()
    return pred, gh, gw


@pytest.mark.gpu
@pytest.mark.parametrize("contract", [lib.CONTRACT_P, lib.CONTRACT_F])
@pytest.mark.parametrize("hw", [(64, 96), (96, 64)])
def test_yolo_v2_graph_against_oracle(ctx, contract, hw):
    """Every layer of the BN-folded fixed-point yolo_v2 (23 convolutions: 3x3 and 1x1, up to 1280 -> 1024 channels, the route
    from the un-pooled C_5 through a 1x1 layer and reorg, the exponent-aligned concat) and its 20-class detections against the
    generalised oracle (tests/oracle_lib.py: backbone_graph), bit for bit."""
    H, W = hw
    qnet = ex.random_quantnet_yolo_v2(seed=0, calib_hw=(H, W), calib_frames=1)
    ctx.load_quantnet(qnet, contract=contract, round_mode=lib.ROUND_RNE, conf_thresh=0.02, nms_thresh=0.5, max_det=1024)
    x = ex.synthetic_frames_f32(3, H, W, seed=21).numpy()
    x8, _ = ol.quantize_f32(x, qnet.sa[0])
    ref, _ = ol.backbone_graph(qnet, x8, contract=contract)
    pred, gh, gw = _run_graph(ctx, qnet, x8)
    assert (gh, gw) == (H // 32, W // 32) and ctx.slow_path_count() >= 0
    for l, r in enumerate(ref):
        got = ctx.layer_output(l, 3, r.shape[1], r.shape[2])
        assert np.array_equal(got, r), "yolo_v2 layer %d (%s %s) differs from the oracle" % (l, qnet.layers[l], qnet.graph[l])
    dets, counts = det_arrays(ctx, pred, 3, gh, gw, H, W)
    for i in range(3):
        (ob, os_, oc, oidx), ocnt = ol.head_python(ref[-1][i], 5, 20, qnet.sa[-1], qnet.anchors, 32, H, W, 0.02, 0.5, max_det=1024)
        b, s_, c, idx = lib.dets_to_arrays(dets[i], int(counts[i]))
        assert counts[i] == ocnt
        np.testing.assert_array_equal(idx, oidx)
        np.testing.assert_array_equal(c, oc)
        np.testing.assert_allclose(s_, os_, atol=1e-5, rtol=0)
        np.testing.assert_allclose(b, ob, atol=1e-5, rtol=0)


@pytest.mark.gpu
def test_yolo_v2_at_416_and_backends_agree(ctx):
    """416x416 (13x13 grid, the configuration BASELINE configs[4] names): the prediction map of one frame equals the oracle's,
    the host entry point returns the device entry point's detections, and the dot-product back end agrees layer by layer."""
    H = W = 416
    qnet = ex.random_quantnet_yolo_v2(seed=1, calib_hw=(H, W), calib_frames=1)
    ctx.load_quantnet(qnet, contract=lib.CONTRACT_F, conf_thresh=0.02, nms_thresh=0.5, max_det=1024)
    x = ex.synthetic_frames_f32(2, H, W, seed=3).numpy()
    x8, _ = ol.quantize_f32(x, qnet.sa[0])
    ref, _ = ol.backbone_graph(qnet, x8[:1], contract=0)
    pred, gh, gw = _run_graph(ctx, qnet, x8)
    outs = [ctx.layer_output(l, 2, r.shape[1], r.shape[2]) for l, r in enumerate(ref)]
    for l, r in enumerate(ref):
        assert np.array_equal(outs[l][:1], r), "yolo_v2 @416 layer %d differs from the oracle" % l
    dets, counts = det_arrays(ctx, pred, 2, gh, gw, H, W)
    hd, hc = ctx.forward_int8(x8)
    np.testing.assert_array_equal(hc, counts)
    for f in range(2):
        assert hd[f][:counts[f]].tobytes() == dets[f][:counts[f]].tobytes()
    ctx.set_conv_backend(1)
    try:
        _run_graph(ctx, qnet, x8)
        for l, r in enumerate(ref):
            assert np.array_equal(ctx.layer_output(l, 2, r.shape[1], r.shape[2]), outs[l]), "back ends differ on layer %d" % l
    finally:
        ctx.set_conv_backend(0)
