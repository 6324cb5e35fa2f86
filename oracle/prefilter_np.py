"""ORACLE (test infrastructure only — never imported by the product path).

NumPy restatement of the image pre-filter chain of the reference's "adapt" node
(``ros2_ws/src/liteflownet3/liteflownet3/lfn3_adapt_node.py:164-190``):

  :166      ``cv_hsv = cv2.cvtColor(cv_bgr, cv2.COLOR_BGR2HSV)``
  :170-177  ``contrast = std(v) / (mean(v) + 1e-3)`` -> clip limit, ``self.clahe.setClipLimit(clip)``
  :180      ``v_enhanced = self.clahe.apply(v)``                              (oracle/clahe_np.py)
  :183-184  ``cv_rgb = cv2.cvtColor(cv2.merge((h, s, v_enhanced)), cv2.COLOR_HSV2RGB)``
  :189-191  ``cv_rgb = cv2.bilateralFilter(cv_rgb, d, sigmaColor, sigmaSpace)``

as this cv2 build (4.13, AVX2 dispatch) computes each step on uint8 images:
* BGR2HSV (``color_hsv.simd.hpp``, ``RGB2HSV_b``): 12-bit fixed point with the tables ``sdiv = cvRound((255 << 12) / v)``,
  ``hdiv180 = cvRound((180 << 12) / (6 diff))``;
* HSV2RGB (``HSV2RGB_b``): float32, ``s, v`` scaled by ``1/255.f``, ``h * (6/180)``, sector tables, ``1 - s*f`` and
  ``1 - s*(1-f)`` with ONE rounding (fused negative multiply-add), result ``* 255.f`` TRUNCATED to uint8 in the 32-pixel
  vector steps of a row and ROUNDED in the scalar tail (``width % 32`` pixels);
* bilateralFilter (``bilateral_filter.simd.hpp``): REFLECT_101 border, circular support of radius d/2, weights
  ``(float)exp(-0.5 r^2 / sigma_space^2) * (float)exp(-0.5 c^2 / sigma_color^2)`` with c = |db| + |dg| + |dr|, sums
  accumulated in float in row-major support order, ``cvRound(sum * (1.f / wsum))``.
BGR2HSV, HSV2RGB and the chain without the bilateral filter are pinned bit for bit against the wheel in
``tests/test_oracle_prefilter.py``.  The bilateral restatement is NOT: this wheel routes the 8-bit bilateralFilter
through Intel IPP (``cv2.ipp.useIPP()`` is True; with ``cv2.ipp.setUseIPP(False)`` the mismatches drop from 13 to 3 in
1.9 M values), whose arithmetic is not published; the restatement of OpenCV's own code differs from it at rounding ties
(a few values per 100 000, off by one) — parity UNPINNED for that step: the device filter (``ofb_bilateral_u8c3``) is
bit-exact with this restatement and held to "at most one grey level in at most 1 value per 10 000" against the wheel.
"""
from __future__ import annotations

import math

import numpy as np

from . import clahe_np

_SHIFT = 12
_SDIV = np.zeros(256, np.int64)
_HDIV = np.zeros(256, np.int64)
for _i in range(1, 256):
    _SDIV[_i] = int(np.rint((255 << _SHIFT) / (1.0 * _i)))
    _HDIV[_i] = int(np.rint((180 << _SHIFT) / (6.0 * _i)))
_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])   # (b, g, r) <- tab index


def bgr2hsv_u8(img: np.ndarray, rgb_order: bool = False) -> np.ndarray:
    c0, g, c2 = (img[..., i].astype(np.int64) for i in range(3))
    b, r = (c2, c0) if rgb_order else (c0, c2)
    v = np.maximum(np.maximum(b, g), r)
    diff = v - np.minimum(np.minimum(b, g), r)
    vr = np.where(v == r, -1, 0)
    vg = np.where(v == g, -1, 0)
    s = (diff * _SDIV[v] + (1 << (_SHIFT - 1))) >> _SHIFT
    h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))))
    h = (h * _HDIV[diff] + (1 << (_SHIFT - 1))) >> _SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def _hsv2rgb_core(hsv: np.ndarray, simd: bool) -> np.ndarray:
    f32 = np.float32
    one = f32(1)
    h = hsv[..., 0].astype(f32)
    s = (hsv[..., 1].astype(f32) * f32(1.0 / 255.0)).astype(f32)
    v = (hsv[..., 2].astype(f32) * f32(1.0 / 255.0)).astype(f32)
    h = (h * f32(6.0 / 180.0)).astype(f32)
    sec = np.floor(h).astype(np.int64)
    f = (h - sec.astype(f32)).astype(f32)
    sec %= 6
    t1 = (v * (one - s)).astype(f32)
    # 1 - s*f and 1 - s*(1-f) with ONE rounding (fused negative multiply-add) in the vector body and in the scalar tail
    # alike (the tail is compiled in the same AVX2 + FMA translation unit and gets contracted)
    s64, f64 = s.astype(np.float64), f.astype(np.float64)
    t2 = (v * (1.0 - s64 * f64).astype(f32)).astype(f32)
    t3 = (v * (1.0 - s64 * (one - f).astype(np.float64)).astype(f32)).astype(f32)
    tab = np.stack([v, t1, t2, t3], -1)
    idx = _SECTOR[sec]
    b = np.take_along_axis(tab, idx[..., 0:1], -1)[..., 0]
    g = np.take_along_axis(tab, idx[..., 1:2], -1)[..., 0]
    r = np.take_along_axis(tab, idx[..., 2:3], -1)[..., 0]
    out = (np.stack([r, g, b], -1) * f32(255.0)).astype(f32)
    out = np.floor(out) if simd else np.rint(out)      # the vector body truncates, the scalar tail rounds (cvRound)
    return np.clip(out, 0, 255).astype(np.uint8)


HSV2RGB_VECTOR_PIXELS = 32     # pixels per step of the wheel's AVX2 loop; the last width % 32 pixels of a row take the scalar tail


def hsv2rgb_u8(hsv: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(hsv, COLOR_HSV2RGB) on uint8.  The wheel converts each row in steps of 32 pixels with its vector
    code, which TRUNCATES the result, and the remaining ``width % 32`` pixels with scalar code, which ROUNDS it
    (``saturate_cast``) — both reproduced here, so the result depends on the column as the wheel's does."""
    w = hsv.shape[1]
    body = (w // HSV2RGB_VECTOR_PIXELS) * HSV2RGB_VECTOR_PIXELS
    out = np.empty(hsv.shape, np.uint8)
    if body:
        out[:, :body] = _hsv2rgb_core(hsv[:, :body], True)
    if body < w:
        out[:, body:] = _hsv2rgb_core(hsv[:, body:], False)
    return out


def bilateral_support(d: int, sigma_space: float):
    if sigma_space <= 0:
        sigma_space = 1.0
    gs = -0.5 / (sigma_space * sigma_space)
    radius = int(np.rint(sigma_space * 1.5)) if d <= 0 else d // 2
    radius = max(radius, 1)
    offs, sw = [], []
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r = math.sqrt(float(i * i) + float(j * j))
            if r > radius:
                continue
            offs.append((i, j))
            sw.append(np.float32(math.exp(r * r * gs)))
    return radius, offs, np.array(sw, np.float32)


def bilateral_u8c3(img: np.ndarray, d: int, sigma_color: float, sigma_space: float) -> np.ndarray:
    h, w, _ = img.shape
    if sigma_color <= 0:
        sigma_color = 1.0
    gc = -0.5 / (sigma_color * sigma_color)
    cw = np.array([np.float32(math.exp(i * i * gc)) for i in range(256 * 3)], np.float32)
    radius, offs, sw = bilateral_support(d, sigma_space)
    tmp = np.pad(img, ((radius, radius), (radius, radius), (0, 0)), mode="reflect").astype(np.int64)   # REFLECT_101
    c = tmp[radius:radius + h, radius:radius + w]
    sums = np.zeros((h, w, 3), np.float32)
    wsum = np.zeros((h, w), np.float32)
    for (i, j), swk in zip(offs, sw):
        nb = tmp[radius + i:radius + i + h, radius + j:radius + j + w]
        wgt = (swk * cw[np.abs(nb - c).sum(-1)]).astype(np.float32)
        wsum = (wsum + wgt).astype(np.float32)
        sums = (nb.astype(np.float64) * wgt[..., None].astype(np.float64) + sums.astype(np.float64)).astype(np.float32)   # fma
    inv = (np.float32(1) / wsum).astype(np.float32)
    return np.clip(np.rint((sums * inv[..., None]).astype(np.float32)), 0, 255).astype(np.uint8)


def adaptive_clip(v: np.ndarray, clip_min: float, clip_max: float, c_min: float, c_max: float) -> float:
    contrast = np.std(v) / (np.mean(v) + 1e-3)
    return float(np.clip(clip_min + (contrast - c_min) / (c_max - c_min) * (clip_max - clip_min), clip_min, clip_max))


def adapt_prefilter_np(bgr: np.ndarray, apply_clahe=True, clip=None, clip_range=(1.0, 4.0, 0.1, 0.8), tile_grid=(8, 8),
                       bilateral=None) -> np.ndarray:
    """lfn3_adapt_node.py:164-191: BGR frame -> the RGB frame the node feeds on.  ``clip`` None = adaptive."""
    if apply_clahe:
        hsv = bgr2hsv_u8(bgr)
        v = hsv[..., 2]
        if clip is None:
            clip = adaptive_clip(v, *clip_range)
        hsv = np.stack([hsv[..., 0], hsv[..., 1], clahe_np.clahe_apply(v, clip, tile_grid)], -1)
        rgb = hsv2rgb_u8(hsv)
    else:
        rgb = np.ascontiguousarray(bgr[..., ::-1])
    if bilateral is not None:
        rgb = bilateral_u8c3(rgb, *bilateral)
    return rgb
