"""CPU oracle for the image resize in front of the hot path.  TEST INFRASTRUCTURE ONLY (imported by tests/ and
oracle/gen_golden_resize.py; the product never imports it).

The reference resizes with `cv2.resize(image, (size[1], size[0]))` (data/__init__.py:36, called through
BaseTransform.__call__ :55 from test.py:76 / demo.py:71): default interpolation INTER_LINEAR on a uint8 HxWx3 image.
The algorithm lives in a third-party dependency that is not vendored in the reference: OpenCV (opencv-python 4.13.0 in
this image; the reference pins no version).  This file restates OpenCV's published 8-bit bilinear algorithm
(modules/imgproc/src/resize.cpp: cv::hal::resize coefficient set-up, HResizeLinear<uchar,int,short,2048>,
VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>, and the INTER_LINEAR -> INTER_AREA switch for exact 2x
decimation with ResizeAreaFastVec):

  scale  = 1 / ((double)dst / src)
  f      = (float)((d + 0.5) * scale - 0.5);  s = floor(f);  f -= s
  x only: s < 0 -> (s, f) = (0, 0);  s >= src-1 -> (s, f) = (src-1, 0)
  y only: the two rows s, s+1 are clamped into [0, src-1], f is kept
  alpha/beta = (short)rint((1-f) * 2048), (short)rint(f * 2048)          (round half to even)
  H[y][dx]   = S[y][s]*alpha0 + S[y][s+1]*alpha1                        (int32, scale 2^11)
  D[dy][dx]  = (((beta0 * (H[y0] >> 4)) >> 16) + ((beta1 * (H[y1] >> 4)) >> 16) + 2) >> 2
  src == 2*dst on both axes: D = (a + b + c + d + 2) >> 2 over each 2x2 block

Pinned: tests/golden/resize_cv2.npz holds cv2's own outputs (oracle/gen_golden_resize.py, run in the build container);
tests/test_oracle.py checks this restatement against them bit for bit, and against the live cv2 when it is importable.
"""
import numpy as np

COEF_BITS = 11
ONE = 1 << COEF_BITS


def axis_table(src: int, dst: int, clamp_coeff: bool):
    """(i0, i1, c0, c1) per destination index: source taps and 11-bit weights."""
    scale = np.float64(1.0) / (np.float64(dst) / np.float64(src))
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_coeff:                       # horizontal pass: resize.cpp re-anchors the tap at the border
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    c0 = np.rint((np.float32(1) - f) * np.float32(ONE)).astype(np.int32)
    c1 = np.rint(f * np.float32(ONE)).astype(np.int32)
    i0 = np.clip(s, 0, src - 1).astype(np.int32)
    i1 = np.clip(s + 1, 0, src - 1).astype(np.int32)
    return i0, i1, c0, c1


def resize_bilinear_u8(img: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """img uint8 [sh][sw][c] -> uint8 [dh][dw][c], equal to cv2.resize(img, (dw, dh))."""
    assert img.dtype == np.uint8 and img.ndim == 3
    sh, sw, _ = img.shape
    a = img.astype(np.int32)
    if sh == 2 * dh and sw == 2 * dw:
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, a0, a1 = axis_table(sw, dw, True)
    y0, y1, b0, b1 = axis_table(sh, dh, False)
    hbuf = a[:, x0, :] * a0[None, :, None] + a[:, x1, :] * a1[None, :, None]
    s0 = hbuf[y0] >> 4
    s1 = hbuf[y1] >> 4
    out = (((b0[:, None, None] * s0) >> 16) + ((b1[:, None, None] * s1) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def resize_batch(imgs: np.ndarray, dh: int, dw: int) -> np.ndarray:
    return np.stack([resize_bilinear_u8(im, dh, dw) for im in imgs])
