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


def flow_to_color_speed(flow_hw2: np.ndarray, dt: float, pixel_to_meter: float, max_speed: float) -> np.ndarray:
    """The sub node's dense view (``lfn3_sub_node.py:244-262``) on a float32 [H,W,2] field -> uint8 [H,W,3] BGR:
    ``mag, ang = cartToPolar(u, v); mag_norm = clip(mag / dt * pixel_to_meter / max_speed, 0, 1)``,
    ``hsv = (uint8(ang * 90 / pi), 255, uint8(mag_norm * 255))`` -> HSV2BGR; float32 operations in NumPy's order."""
    u, v = flow_hw2[..., 0], flow_hw2[..., 1]
    mag, ang = cart_to_polar(u, v)
    t = (((mag / f32(dt)).astype(f32) * f32(pixel_to_meter)).astype(f32) / f32(max_speed)).astype(f32)
    hsv = np.zeros(u.shape + (3,), np.uint8)
    hsv[..., 0] = ((ang * f32(90.0)).astype(f32) / f32(np.pi)).astype(f32).astype(np.uint8)
    hsv[..., 1] = 255
    hsv[..., 2] = (np.clip(t, f32(0), f32(1)) * f32(255)).astype(f32).astype(np.uint8)
    return prefilter_np.hsv2rgb_u8(hsv)[..., ::-1].copy()


def flow_arrows(flow_hw2: np.ndarray, step: int = 20) -> np.ndarray:
    """The sub node's arrow overlay (``lfn3_sub_node.py:225-238``): for every grid point (x, y) with stride ``step`` the
    segment (x, y) -> (x + int(u), y + int(v)) that ``cv2.arrowedLine`` draws; int32 [n, 4] in the node's loop order."""
    h, w = flow_hw2.shape[:2]
    out = [(x, y, x + int(flow_hw2[y, x, 0]), y + int(flow_hw2[y, x, 1])) for y in range(0, h, step) for x in range(0, w, step)]
    return np.asarray(out, np.int32).reshape(-1, 4)
