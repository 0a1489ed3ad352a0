#!/usr/bin/env python
"""Bring-up check for the tcgen05 convolution kernel: every tensor-core layer against the integer dot-product
kernel (same library, backend switch) and, at small sizes, against the CPU oracle.  Prints where results differ."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import yolo_b200  # noqa
from yolo_b200 import export as ex, lib


def main():
    layers = [int(a) for a in sys.argv[1:]] or list(range(1, 10))
    qnet = ex.random_quantnet(seed=0, calib_hw=(64, 96), calib_frames=1)
    ctx = lib.Context(0)
    rng = np.random.default_rng(0)
    shapes = {1: [(2, 16, 32), (3, 26, 30)], 2: [(2, 8, 16), (2, 13, 13)], 3: [(2, 8, 16), (3, 14, 18)], 4: [(2, 8, 16), (3, 13, 13)],
              5: [(2, 8, 16), (3, 26, 26)], 6: [(2, 8, 16), (5, 13, 13)], 7: [(2, 8, 16), (5, 13, 13)], 8: [(1, 8, 16), (4, 15, 20)],
              9: [(2, 8, 16), (3, 13, 13)]}
    bad = 0
    import copy
    shipped = copy.deepcopy(qnet)
    shipped.sa, shipped.sw, shipped.sb, shipped.retune = list(ex.SHIPPED_SCALE_A), list(ex.SHIPPED_SCALE_W), list(ex.SHIPPED_SCALE_B), list(ex.SHIPPED_RETUNE)
    configs = [("F/RNE calibrated", qnet, lib.CONTRACT_F, lib.ROUND_RNE), ("P calibrated", qnet, lib.CONTRACT_P, 0),
               ("F/RNE shipped", shipped, lib.CONTRACT_F, lib.ROUND_RNE), ("F/FLOOR calibrated", qnet, lib.CONTRACT_F, lib.ROUND_FLOOR),
               ("P shipped", shipped, lib.CONTRACT_P, 0)]
    for (cname, net, contract, mode) in configs:
      ctx.load_quantnet(net, contract=contract, round_mode=mode)
      print("==", cname, flush=True)
      for l in layers:
          cin, cout, activ, pool = qnet.layers[l]
          for (n, h, w) in shapes[l]:
              x = np.zeros((n, h, w, ex.cstride(cin)), np.int8)
              x[..., :cin] = rng.integers(-128, 128, (n, h, w, cin), dtype=np.int8)
              dx = torch.from_numpy(x).cuda()
              oh, ow = (h // 2, w // 2) if pool else (h, w)
              outs = []
              for backend in (1, 2, 3, 4, 5):
                  ctx.set_conv_backend(backend)
                  o = torch.full((n, oh, ow, ex.cstride(cout)), 99, dtype=torch.int8, device="cuda")
                  t0 = time.time()
                  try:
                      ctx.conv_layer(l, dx, n, h, w, o)
                  except lib.YoloB200Error as e:
                      if backend >= 4:          # weights do not fit the weight-stationary kernel: nothing to compare
                          outs.append(outs[0])
                          continue
                      raise
                  ctx.sync()
                  outs.append(o.cpu().numpy())
              a, b, c3, w4, w5 = outs
              if (a != c3).any():
                  bad += 1
                  print("   integer-epilogue tcgen05 path differs from dp4a: %d" % (a != c3).sum())
              for nm, wv in (("ws", w4), ("ws/int-epilogue", w5)):
                  dws = a != wv
                  if dws.any():
                      bad += 1
                      idx = np.argwhere(dws)
                      print("   %s kernel differs from dp4a: %d / %d  first %s direct %s ws %s" % (nm, dws.sum(), dws.size, idx[:4].tolist(), a[dws][:6].tolist(), wv[dws][:6].tolist()))
                      print("      ch%%16:", np.bincount(idx[:, 3] % 16, minlength=16).tolist(), " x:", np.bincount(idx[:, 2], minlength=ow).tolist())
                      print("      y:", np.bincount(idx[:, 1], minlength=oh).tolist(), " n:", np.bincount(idx[:, 0], minlength=n).tolist())
              diff = a != b
              print("layer %d cin %3d cout %3d pool %d shape %s: mismatches %d / %d" % (l, cin, cout, pool, (n, h, w), diff.sum(), diff.size), flush=True)
              if diff.any():
                  bad += 1
                  idx = np.argwhere(diff)
                  print("   first:", idx[:6].tolist(), "direct", a[diff][:6].tolist(), "umma", b[diff][:6].tolist())
                  print("   mismatching channels histogram (mod 16):", np.bincount(idx[:, 3] % 16, minlength=16).tolist())
                  print("   mismatching x histogram:", np.bincount(idx[:, 2], minlength=ow).tolist())
                  print("   mismatching y histogram:", np.bincount(idx[:, 1], minlength=oh).tolist())
                  print("   mismatching n histogram:", np.bincount(idx[:, 0], minlength=n).tolist())
    print("UMMA CHECK:", "FAILED (%d cases)" % bad if bad else "ok")
    ctx.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
