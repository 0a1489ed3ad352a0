"""CPU tests of the host-side logic: the C-ABI library loads and exports every symbol include/yolo_b200.h declares
(no compute without a GPU), parameter tables, the exporter, and the multi-GPU sharding plumbing over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import yolo_b200  # noqa: F401
from yolo_b200 import export as ex
from yolo_b200 import lib, runner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from yolo_b200 import build
    build.build()
    return lib.load_library()


def test_library_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "yolo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(yolo_b200_\w+|yolo_forward)\s*\(", hdr))
    names -= {"yolo_b200_ctx", "yolo_b200_params", "yolo_b200_det", "yolo_b200_layer"}
    assert len(names) >= 28
    for n in names:
        assert hasattr(L, n), "libyolo_b200.so does not export %s" % n
    assert names == set(lib.EXPORTS)


def test_default_params_are_the_shipped_tables(L):
    p = lib.Params()
    assert L.yolo_b200_default_params(C.byref(p)) == 0
    assert p.num_layers == 10
    assert [(p.layers[l].cin, p.layers[l].cout, p.layers[l].activ, p.layers[l].pool) for l in range(10)] == ex.SLIM_YOLO_V2_LAYERS
    assert list(p.scale_w)[:10] == ex.SHIPPED_SCALE_W and list(p.scale_b)[:10] == ex.SHIPPED_SCALE_B
    assert list(p.scale_a)[:11] == ex.SHIPPED_SCALE_A and list(p.retune)[:10] == ex.SHIPPED_RETUNE
    assert [[round(p.anchors[a][0], 4), round(p.anchors[a][1], 4)] for a in range(5)] == ex.ANCHOR_SIZE_COCO
    assert (p.num_anchors, p.num_classes, p.stride) == (5, 2, 16)
    assert abs(p.conf_thresh - 0.01) < 1e-9 and p.nms_thresh == 0.5


def test_cstride(L):
    assert [L.yolo_b200_cstride(c) for c in (3, 4, 16, 35, 256, 425)] == [4, 4, 16, 48, 256, 432]
    assert [ex.cstride(c) for c in (3, 4, 16, 35, 256, 425)] == [4, 4, 16, 48, 256, 432]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback(L):
    h = C.c_void_p()
    rc = L.yolo_b200_create(C.byref(h), 0)
    assert rc < 0 and b"no CPU fallback" in L.yolo_b200_last_error()
    with pytest.raises(lib.YoloB200Error):
        lib.Context(0)
    # null-argument error behaviour needs no device
    assert L.yolo_b200_default_params(None) < 0
    assert L.yolo_b200_set_stream(None, None) < 0
    # the image front end has no host path either: without a context every entry point is an error, not a CPU resize
    buf = np.zeros(4 * 4 * 3, dtype=np.uint8)
    assert L.yolo_b200_resize_u8bgr(None, buf.ctypes.data, 1, 4, 4, buf.ctypes.data, 4, 4) < 0
    assert L.yolo_b200_forward_u8bgr_resize(None, buf.ctypes.data, 1, 4, 4, 4, 4, None, None) < 0
    assert L.yolo_b200_forward_u8bgr_resize_dev(None, None, 1, 4, 4, 4, 4, None, None) < 0


def test_quantise_rule_is_idempotent_on_checkpoints():
    """A `*_retune_quantize*.pth`-style state_dict (weights stored as q/s, retune_bias_quantize.py:411-415)
    re-exports to the same integers."""
    qnet = ex.random_quantnet(seed=3, calib_hw=(64, 64), calib_frames=1)
    sd = qnet.dequantized_state_dict()
    q2 = ex.quantnet_from_state_dict(sd, anchors=qnet.anchors)
    for a, b in zip(qnet.w + qnet.b, q2.w + q2.b):
        np.testing.assert_array_equal(a, b)
    assert (qnet.sw, qnet.sb, qnet.sa) == (q2.sw, q2.sb, q2.sa)
    assert len(sd) == 42


def test_weight_h_order_roundtrip_and_tap_stride():
    rng = np.random.default_rng(0)
    for cin, cout in ((3, 16), (16, 32), (32, 64), (256, 35)):
        w = rng.integers(-128, 128, (cout, 3, 3, cin), dtype=np.int8)
        flat = ex.pack_weight_h_order(w)
        np.testing.assert_array_equal(ex.unpack_weight_h_order(flat, cout, cin), w)
        # load_weight reads tap t at offset t * (padded cin*cout) (yolo_forward.c:169)
        assert flat.size % 9 == 0
        tm, tn = min(32, cout), min(16, cin)
        t4 = flat.reshape(9, -1)[4].reshape(-(-cout // tm), -(-cin // tn), tm, tn)
        assert t4[0, 0, 1, 2] == w[1, 1, 1, 2]


def test_shard_range_partitions_all_frames():
    for n in (0, 1, 7, 256, 2048):
        for world in (1, 2, 3, 8):
            spans = [runner.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_gather_detections_gloo_world2(tmp_path):
    """world_size-2 gloo run of the only cross-rank step: gathering ragged per-rank detection lists in frame order."""
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import yolo_b200
from yolo_b200 import runner
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 5
lo, hi = runner.shard_range(n, rank, world)
dets = torch.zeros((hi - lo, 4, 8), dtype=torch.int32)
counts = torch.zeros((hi - lo,), dtype=torch.int32)
for i, f in enumerate(range(lo, hi)):
    dets[i] = f + 1
    counts[i] = f %% 4
ad, ac = runner.gather_detections(dets, counts, n)
assert ad.shape == (n, 3, 8) and ac.tolist() == [f %% 4 for f in range(n)]   # trimmed to the largest count (3)
assert all(int(ad[f, 0, 0]) == f + 1 for f in range(n))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
''' % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_async_detection_gatherer_gloo_world2(tmp_path):
    """The double-buffered asynchronous gather the bench uses at N > 1 (here over gloo on CPU, world size 2)."""
    script = tmp_path / "g.py"
    script.write_text('''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import yolo_b200
from yolo_b200 import runner
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = runner.DetectionGatherer(local_frames=3, max_det=6, cap=4, device=torch.device("cpu"))
for step in range(5):
    b = g.buffers(step)
    b.dets[:] = 100 * step + 10 * rank + torch.arange(6, dtype=torch.int32).view(1, 6, 1)
    b.counts[:] = torch.tensor([1, 2, 3], dtype=torch.int32) + rank
    g.launch(step)
g.finish()
b = g.bufs[4 %% 2]
assert b.all_dets.shape == (6, 4, 8) and b.all_counts.tolist() == [1, 2, 3, 2, 3, 4]
for r in range(world):
    assert b.all_dets[3 * r, :, 0].tolist() == [400 + 10 * r + k for k in range(4)]
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
''' % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_resize_taps_match_the_oracle(L):
    """yolo_b200_resize_taps (host arithmetic of the GPU resize: OpenCV's float/double coefficient set-up) against
    oracle/resize_u8.py's axis_table, which is pinned on cv2's own outputs (tests/test_oracle.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle_resize_u8", os.path.join(ROOT, "oracle", "resize_u8.py"))
    ro = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ro)
    rng = np.random.default_rng(3)
    pairs = [(480, 416), (640, 416), (375, 416), (1080, 416), (1920, 416), (832, 416), (1, 7), (7, 1), (416, 416), (100, 608)]
    pairs += [tuple(int(v) for v in rng.integers(1, 3000, 2)) for _ in range(200)]
    for src, dst in pairs:
        for horizontal in (0, 1):
            taps = np.zeros((dst, 4), dtype=np.int32)
            assert L.yolo_b200_resize_taps(src, dst, horizontal, taps.ctypes.data) == 0
            i0, i1, c0, c1 = ro.axis_table(src, dst, bool(horizontal))
            np.testing.assert_array_equal(taps, np.stack([i0, i1, c0, c1], axis=1), err_msg="%d -> %d h=%d" % (src, dst, horizontal))
    assert L.yolo_b200_resize_taps(0, 4, 1, None) < 0


def test_bn_fold_matches_the_reference_function():
    """export.fold_bn / fold_bn_state_dict against the reference's own `fuse_conv_and_bn` (conv+bn2conv.py:126-150) run
    on the reference's un-fused SlimYOLOv2 (tests/golden/bnfold_ref.npz, oracle/gen_golden_bnfold.py): bit-exact weights
    and biases, BatchNorm keys gone, the fused keys are the ones SlimYOLOv2_quantize_bnfuse loads."""
    import hashlib
    g = np.load(os.path.join(ROOT, "tests", "golden", "bnfold_ref.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in/")}
    sd["pred.weight"] = torch.zeros(35, 256, 3, 3)                         # a module without BatchNorm passes through
    fused = ex.fold_bn_state_dict(sd)
    assert not any(".convs.1." in k for k in fused)
    assert "pred.weight" in fused and fused["pred.weight"] is sd["pred.weight"]
    names = sorted({k.split("/")[1] for k in g.files if k.startswith("out/")})
    assert len(names) == 6
    for k in names:
        np.testing.assert_array_equal(fused[k].numpy(), g["out/" + k], err_msg=k)
        assert hashlib.sha256(np.ascontiguousarray(fused[k].numpy()).tobytes()).hexdigest() == str(g["sha/" + k])
        assert k.rsplit(".", 1)[0] in ex.SLIM_CONV_KEYS
    # nested modules (what conv+bn2conv.py:317-326 cannot reach) and members after the BatchNorm moving down one index
    nested = {"backbone.layer1.convs.0.weight": sd["conv1.convs.0.weight"], "backbone.layer1.convs.0.bias": sd["conv1.convs.0.bias"],
              "backbone.layer1.convs.3.weight": torch.ones(2)}
    for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
        nested["backbone.layer1.convs.1." + k] = sd["conv1.convs.1." + k]
    f2 = ex.fold_bn_state_dict(nested)
    assert sorted(f2) == ["backbone.layer1.convs.0.bias", "backbone.layer1.convs.0.weight", "backbone.layer1.convs.2.weight"]
    np.testing.assert_array_equal(f2["backbone.layer1.convs.0.weight"].numpy(), g["out/conv1.convs.0.weight"])
    with pytest.raises(ValueError):
        ex.fold_bn_state_dict({"bn.0.running_var": torch.ones(2)})


def test_quantnet_from_state_dict_takes_the_head_width_from_the_checkpoint():
    """The reference's default is num_classes=20 (VOC, test.py:165-172): 5 anchors x 25 = 125 prediction channels."""
    from yolo_b200 import model
    net = model.SlimYOLOv2_quantize_bnfuse("cpu", input_size=[64, 64], num_classes=20, anchor_size=ex.ANCHOR_SIZE)
    q = net.quantnet(calib_frames=ex.synthetic_frames_f32(1, 64, 64))
    assert q.layers[-1] == (256, 125, 0, 0) and q.w[-1].shape == (125, 3, 3, 256) and q.num_classes == 20
    assert q.retune is not None and len(q.retune) == 10
    with pytest.raises(ValueError):          # head width and (anchors, classes) must agree
        ex.quantnet_from_state_dict(net.state_dict(), calib_frames=ex.synthetic_frames_f32(1, 64, 64), anchors=ex.ANCHOR_SIZE, num_classes=2)


def test_retune_is_never_invented():
    """A checkpoint with calibrated trackers and no calibration frames carries no accumulator scale: contract F and the
    weight.h export refuse it instead of using a table tuned on noise; an explicit table or calibration frames fix it."""
    q0 = ex.random_quantnet(seed=0, calib_hw=(64, 64), calib_frames=1)
    sd = q0.dequantized_state_dict()
    q = ex.quantnet_from_state_dict(sd, anchors=q0.anchors, num_classes=2)
    assert q.retune is None and q.sa == q0.sa
    with pytest.raises(ValueError):
        lib.make_params(q, contract=lib.CONTRACT_F)
    with pytest.raises(ValueError):
        ex.write_weight_h(q, os.devnull)
    lib.make_params(q, contract=lib.CONTRACT_P)                      # the fake-quant contract never reads retune
    q2 = ex.quantnet_from_state_dict(sd, anchors=q0.anchors, num_classes=2, retune=ex.SHIPPED_RETUNE)
    assert q2.retune == ex.SHIPPED_RETUNE
    lib.make_params(q2, contract=lib.CONTRACT_F)
    q3 = ex.quantnet_from_state_dict(sd, anchors=q0.anchors, num_classes=2, calib_frames=ex.synthetic_frames_f32(1, 64, 64, seed=1000))
    assert q3.retune == q0.retune and q3.sa == q0.sa


def test_bn_fold_with_two_batchnorms_in_one_sequential():
    """Sequential(conv, bn, relu, conv, bn, relu): both convolutions are fused and the survivors are renumbered the way
    nn.Sequential(fused, *rest) would (conv+bn2conv.py:317-326 generalised).  (Convolutions in front of a BatchNorm have no
    bias, as in the reference's blocks, utils/modules.py:6-18: fuse_conv_and_bn adds a conv bias UNSCALED, :148.)"""
    torch.manual_seed(3)
    blk = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1, bias=False), torch.nn.BatchNorm2d(4), torch.nn.ReLU(),
                              torch.nn.Conv2d(4, 5, 3, padding=1, bias=False), torch.nn.BatchNorm2d(5), torch.nn.ReLU()).eval()
    for m in blk:
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.uniform_(-1, 1); m.running_var.uniform_(0.5, 2); m.weight.data.uniform_(0.5, 1.5); m.bias.data.uniform_(-1, 1)
    sd = {"blk." + k: v for k, v in blk.state_dict().items()}
    out = ex.fold_bn_state_dict(sd)
    assert sorted(out) == ["blk.0.bias", "blk.0.weight", "blk.2.bias", "blk.2.weight"]
    fused = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 5, 3, padding=1), torch.nn.ReLU())
    fused.load_state_dict({k[4:]: v for k, v in out.items()})
    x = torch.randn(2, 3, 8, 8)
    with torch.no_grad():
        assert torch.allclose(fused(x), blk(x), atol=1e-5)
