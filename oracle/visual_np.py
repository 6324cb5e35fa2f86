"""NumPy restatement of the nodes' flow visualisation ``flow_to_color``
(``ros2_ws/src/liteflownet3/liteflownet3/sub_n_pub_lfn3_node.py:132-140``):

    mag, ang = cv2.cartToPolar(u, v); hsv[..., 0] = (ang * 180 / np.pi / 2).astype(np.uint8); hsv[..., 1] = 255
    hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8); cv2.cvtColor(hsv, COLOR_HSV2BGR)

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows OpenCV 4.x ``modules/core/src/mathfuncs_core.simd.hpp``
(``v_atan_f32``: the 7th-order polynomial of fastAtan, magnitude as sqrt(fma(x, x, y*y))), ``norm.cpp`` / ``convert_scale``
(``normalize`` NORM_MINMAX = fma(src, scale, shift) in float32) and ``color_hsv`` (``oracle/prefilter_np.py``); pinned
against the wheel in ``tests/test_oracle_visual.py``.
"""
from __future__ import annotations

import numpy as np

from . import prefilter_np

f32, f64 = np.float32, np.float64


def _fma(a, b, c):
    return (np.asarray(a, f32).astype(f64) * np.asarray(b, f32).astype(f64) + np.asarray(c, f32).astype(f64)).astype(f32)


_K = f32(180 / np.pi)
_P1, _P3 = f32(f32(0.9997878412794807) * _K), f32(f32(-0.3258083974640975) * _K)
_P5, _P7 = f32(f32(0.1555786518463281) * _K), f32(f32(-0.04432655554792128) * _K)
_EPS = f32(2.220446049250313e-16)


def cart_to_polar(x: np.ndarray, y: np.ndarray):
    """== cv2.cartToPolar(x, y) for float32 (angle in radians)."""
    x, y = np.asarray(x, f32), np.asarray(y, f32)
    mag = np.sqrt(_fma(x, x, (y * y).astype(f32))).astype(f32)
    ax, ay = np.abs(x), np.abs(y)
    c = (np.minimum(ax, ay) / (np.maximum(ax, ay) + _EPS).astype(f32)).astype(f32)
    cc = (c * c).astype(f32)
    a = (_fma(_fma(_fma(cc, _P7, _P5), cc, _P3), cc, _P1) * c).astype(f32)
    a = np.where(ax >= ay, a, (f32(90) - a).astype(f32))
    a = np.where(x < 0, (f32(180) - a).astype(f32), a)
    a = np.where(y < 0, (f32(360) - a).astype(f32), a)
    return mag, (a * f32(np.pi / 180.0)).astype(f32)


def normalize_minmax_0_255(src: np.ndarray) -> np.ndarray:
    """== cv2.normalize(src, None, 0, 255, cv2.NORM_MINMAX) for float32."""
    src = np.asarray(src, f32)
    smin, smax = float(src.min()), float(src.max())
    scale = 255.0 * (1.0 / (smax - smin) if smax - smin > 2.220446049250313e-16 else 0.0)
    shift = 0.0 - smin * scale
    return _fma(src, f32(scale), f32(shift))


def flow_to_color(flow_hw2: np.ndarray) -> np.ndarray:
    """The node's flow_to_color on a float32 [H,W,2] field (u, v) -> uint8 [H,W,3] BGR."""
    u, v = flow_hw2[..., 0], flow_hw2[..., 1]
    mag, ang = cart_to_polar(u, v)
    hsv = np.zeros(u.shape + (3,), np.uint8)
    hsv[..., 1] = 255
    hue = (((ang * f32(180)).astype(f32) / f32(np.pi)).astype(f32) / f32(2)).astype(f32)
    hsv[..., 0] = hue.astype(np.uint8)
    hsv[..., 2] = normalize_minmax_0_255(mag).astype(np.uint8)
    return prefilter_np.hsv2rgb_u8(hsv)[..., ::-1].copy()
