"""Shared helpers for tests that use tests/golden/*.npz (made by oracle/gen_golden.py from the reference)."""
import hashlib
import os

import numpy as np

import yolo_b200  # noqa: F401
from yolo_b200 import export as ex

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["ref_p_64x96", "ref_p_80x64_sparse", "ref_p_416x416"]
_cache = {}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    """Returns (npz dict, QuantNet rebuilt WITHOUT the reference, float frames [n,3,H,W])."""
    if name in _cache:
        return _cache[name]
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    H, W, seed = int(g["H"]), int(g["W"]), int(g["seed"])
    qnet = ex.random_quantnet(seed=seed, calib_hw=(H, W), calib_frames=2, head_bias_shift=float(g["head_bias_shift"]),
                              anchors=g["anchors"].tolist())
    assert qnet.sha256() == str(g["net_sha256"]), "rebuilt network differs from the one the reference ran"
    assert qnet.sa == g["sa"].tolist()
    import torch
    frames = torch.cat([ex.synthetic_frames_f32(1, H, W, seed=int(s)) for s in g["frame_seeds"]])
    for i in range(int(g["n_frames"])):
        assert sha(frames[i:i + 1].numpy()) == str(g["f%d_frame_sha256" % i])
    _cache[name] = (g, qnet, frames)
    return _cache[name]
