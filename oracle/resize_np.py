"""ORACLE (test infrastructure only — never imported by the product path).

NumPy restatement of ``cv2.resize(img, (w, h))`` (INTER_LINEAR, uint8, 1-4 channels) as this cv2 build computes it — the
call the reference nodes make on every frame that does not have the configured size
(``ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:152-153``, ``lfn3_adapt_node.py:160-161``).  Follows OpenCV 4.x
``modules/imgproc/src/resize.cpp`` (``resizeGeneric_`` with ``HResizeLinear`` / ``VResizeLinear<uchar, int, short>``):
11-bit fixed-point coefficients ``cvRound(w * 2048)`` of the float source coordinate ``(d + 0.5) * scale - 0.5``;
columns clamp the coordinate (left: weight 1 on column 0; right: weight 1 on the last column), rows do not — they clip
the two row indices instead; vertical pass ``((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2``.
Pinned bit for bit against the wheel in ``tests/test_oracle_resize.py`` (down- and up-scaling, odd sizes, exact 2x).
"""
from __future__ import annotations

import numpy as np


def _coords(dn: int, sn: int):
    scale = sn / dn
    f = ((np.arange(dn, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def _fixed(f: np.ndarray):
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int64)
    return a0, a1


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.asarray(src, np.uint8)
    sh, sw = src.shape[:2]
    s = src.reshape(sh, sw, -1).astype(np.int64)
    xs, fx = _coords(dw, sw)
    lo = xs < 0
    xs[lo] = 0; fx[lo] = 0
    hi = xs >= sw - 1
    xs[hi] = sw - 1; fx[hi] = 0
    ax0, ax1 = _fixed(fx)
    rows = s[:, xs, :] * ax0[None, :, None] + s[:, np.minimum(xs + 1, sw - 1), :] * ax1[None, :, None]
    ys, fy = _coords(dh, sh)
    b0, b1 = _fixed(fy)
    S0 = rows[np.clip(ys, 0, sh - 1)]
    S1 = rows[np.clip(ys + 1, 0, sh - 1)]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    out = out.astype(np.uint8)
    return out[..., 0] if src.ndim == 2 else out
