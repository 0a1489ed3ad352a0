"""Writes tests/golden/bnfold_ref.npz: the reference's own BN folding applied to the reference's own un-fused network.
Run in the build container (needs /root/reference):
    python oracle/gen_golden_bnfold.py
A SlimYOLOv2 (models/slim_yolo_v2.py:385, Conv2d = conv + BatchNorm2d + LeakyReLU, utils/modules.py:6-18) is built with
seeded weights and seeded non-trivial BatchNorm statistics; every Conv2d block is folded by the UNMODIFIED
`fuse_conv_and_bn` of conv+bn2conv.py:126-150 (extracted from the file by name: the script's top level imports the
training stack) in the loop of conv+bn2conv.py:317-326.  Stored for the first three layers: the un-fused state_dict's tensors (so the
test needs no reference), the fused weight / bias and their sha256."""
import ast
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import REF, ROOT, import_reference  # noqa: E402

LAYERS = ("conv1", "conv2", "conv3_1")     # layers kept in the fixture (inputs whole, fused tensors whole + sha256): 95 KB


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


def main():
    mod, _, _ = import_reference()
    src = open(os.path.join(REF, "conv+bn2conv.py")).read()
    ns = {"torch": torch, "nn": nn}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "fuse_conv_and_bn":
            exec(compile(ast.Module([node], []), "conv+bn2conv.py", "exec"), ns)
    fuse_conv_and_bn = ns["fuse_conv_and_bn"]
    import utils.modules as um                                            # the reference's (sys.path has REF first)
    torch.manual_seed(7)
    net = mod.SlimYOLOv2("cpu", input_size=[64, 64], num_classes=2, anchor_size=[[1, 1]] * 5)
    g = torch.Generator().manual_seed(8)
    for m in net.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data = torch.rand(m.weight.shape, generator=g) * 1.5 + 0.25
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.3
            m.running_mean = torch.randn(m.running_mean.shape, generator=g) * 0.5
            m.running_var = torch.rand(m.running_var.shape, generator=g) * 2 + 0.05
    net.eval()
    out = {}
    sd = net.state_dict()
    for k, v in sd.items():
        if k.split(".")[0] in LAYERS:
            out["in/" + k] = v.detach().cpu().numpy()
    for name, a in net.named_children():                                  # conv+bn2conv.py:317-326
        if isinstance(a, um.Conv2d) and name in LAYERS:
            for i, b in enumerate(a.convs):
                if isinstance(b, nn.BatchNorm2d):
                    fused = fuse_conv_and_bn(a.convs[i - 1], b)
                    out["sha/%s.convs.0.weight" % name] = np.array(sha(fused.weight))
                    out["sha/%s.convs.0.bias" % name] = np.array(sha(fused.bias))
                    out["out/%s.convs.0.weight" % name] = fused.weight.detach().numpy()
                    out["out/%s.convs.0.bias" % name] = fused.bias.detach().numpy()
                    break
    path = os.path.join(ROOT, "tests", "golden", "bnfold_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), len(out))


if __name__ == "__main__":
    main()
