"""Writes tests/golden/resize_cv2.npz: outputs of the reference's resize call, cv2.resize(image, (w, h)) with the default
INTER_LINEAR (data/__init__.py:36), on seeded uint8 images.  Run in the build container (cv2 4.13.0):
    python oracle/gen_golden_resize.py
Sources are regenerated from the seed by the tests (numpy default_rng(seed).integers(0, 256, (sh, sw, 3), uint8)); only
the cv2 outputs are stored."""
import os

import cv2
import numpy as np

CASES = [  # (seed, sh, sw, dh, dw)
    (1, 48, 64, 32, 32),      # the reference's 4:3 camera aspect onto a square input, down
    (2, 24, 32, 64, 96),      # up-scaling, non-square
    (3, 37, 53, 32, 48),      # odd sizes
    (4, 64, 96, 32, 48),      # exact 2x decimation (cv2 switches to the 2x2 area mean)
    (5, 32, 48, 32, 48),      # identity
    (6, 1, 1, 8, 12),         # single source pixel
    (7, 97, 33, 32, 64),      # mixed: down in y, up in x
    (8, 128, 64, 32, 32),     # 4x / 2x decimation (not the area path)
]


def main():
    out = {"cases": np.array(CASES, dtype=np.int32), "cv2_version": np.array(cv2.__version__)}
    for seed, sh, sw, dh, dw in CASES:
        img = np.random.default_rng(seed).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        out["out_%d" % seed] = cv2.resize(img, (dw, dh))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "resize_cv2.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
