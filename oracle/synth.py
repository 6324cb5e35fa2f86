"""Seeded synthetic frame-pair generators.  Test infrastructure.

``synth_pair(h, w, seed, shift)`` is the corpus generator: smooth noise, shifted
by a sub-pixel translation.  cv2 is used for the blur/warp when available (it is
on both the build container and the GPU box); a NumPy fallback keeps the
generator importable without cv2 (used by ``bench.py`` for throughput inputs,
where only the statistics of the texture matter).
"""
from __future__ import annotations

import numpy as np


def _smooth_noise(h: int, w: int, seed: int, sigma: float = 2.5) -> np.ndarray:
    import cv2

    rng = np.random.default_rng(seed)
    a = (rng.random((h + 16, w + 16)) * 255.0).astype(np.float32)
    a = cv2.GaussianBlur(a, (0, 0), sigma)
    a = (a - a.min()) / max(float(a.max() - a.min()), 1e-6) * 255.0
    return a.astype(np.float32)


def synth_pair(h: int, w: int, seed: int, shift=(1.7, -0.9)):
    """Return (prev, next) uint8 [h, w]; next is prev translated by ``shift`` px
    (true flow = shift everywhere)."""
    import cv2

    a = _smooth_noise(h, w, seed)
    sx, sy = float(shift[0]), float(shift[1])
    m = np.array([[1, 0, sx], [0, 1, sy]], dtype=np.float64)
    b = cv2.warpAffine(a, m, (a.shape[1], a.shape[0]), flags=cv2.INTER_CUBIC,
                       borderMode=cv2.BORDER_REFLECT)
    a8 = np.clip(np.round(a[8:8 + h, 8:8 + w]), 0, 255).astype(np.uint8)
    b8 = np.clip(np.round(b[8:8 + h, 8:8 + w]), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(a8), np.ascontiguousarray(b8)


def synth_warp_pair(h: int, w: int, seed: int, angle_deg=1.5, zoom=1.02):
    """Rotation + zoom about the centre (non-constant flow)."""
    import cv2

    a = _smooth_noise(h, w, seed)
    c = (a.shape[1] / 2.0, a.shape[0] / 2.0)
    m = cv2.getRotationMatrix2D(c, angle_deg, zoom)
    b = cv2.warpAffine(a, m, (a.shape[1], a.shape[0]), flags=cv2.INTER_CUBIC,
                       borderMode=cv2.BORDER_REFLECT)
    a8 = np.clip(np.round(a[8:8 + h, 8:8 + w]), 0, 255).astype(np.uint8)
    b8 = np.clip(np.round(b[8:8 + h, 8:8 + w]), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(a8), np.ascontiguousarray(b8)


def low_texture_pair(h: int, w: int, seed: int, roll=(2, 3)):
    """Constant 128 + U{0,1,2} noise + one bright square, rolled by (dy, dx)."""
    rng = np.random.default_rng(seed)
    a = (128 + rng.integers(0, 3, size=(h, w))).astype(np.uint8)
    y0, x0 = h // 3, w // 3
    a[y0:y0 + h // 6, x0:x0 + w // 6] = 250
    b = np.roll(a, roll, axis=(0, 1))
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def panning_sequence(h: int, w: int, n_frames: int, seed: int = 100, max_shift: float = 8.0):
    """``n_frames`` uint8 frames of one texture panning with per-frame shifts
    ~U(-max_shift, max_shift) (corpus C2).  Pure NumPy + cv2.warpAffine."""
    import cv2

    rng = np.random.default_rng(seed)
    pad = int(np.ceil(max_shift)) * n_frames + 16
    pad = min(pad, 256)
    base = _smooth_noise(h + 2 * pad - 16, w + 2 * pad - 16, seed)
    frames = []
    ox = oy = 0.0
    for _ in range(n_frames):
        m = np.array([[1, 0, -pad + ox], [0, 1, -pad + oy]], dtype=np.float64)
        f = cv2.warpAffine(base, m, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                           borderMode=cv2.BORDER_REFLECT)
        frames.append(np.clip(np.round(f), 0, 255).astype(np.uint8))
        dx, dy = rng.uniform(-max_shift, max_shift, size=2)
        ox = float(np.clip(ox + dx, -pad + 8, pad - 8))
        oy = float(np.clip(oy + dy, -pad + 8, pad - 8))
    return frames


def cheap_texture(h: int, w: int, seed: int) -> np.ndarray:
    """Fast NumPy-only textured uint8 frame for throughput runs (no cv2)."""
    rng = np.random.default_rng(seed)
    small = rng.random((h // 8 + 3, w // 8 + 3)).astype(np.float32)
    big = np.kron(small, np.ones((8, 8), dtype=np.float32))[:h + 8, :w + 8]
    # 3 box passes ~ gaussian
    for _ in range(3):
        big = (big[:-1, :-1] + big[1:, :-1] + big[:-1, 1:] + big[1:, 1:]) * 0.25
    big = big[:h, :w]
    big = (big - big.min()) / max(float(big.max() - big.min()), 1e-6) * 255.0
    return np.ascontiguousarray(np.round(big).astype(np.uint8))


def subpixel_shift(frame: np.ndarray, sx: float, sy: float) -> np.ndarray:
    """``frame`` translated by a real-valued (sx, sy) with bilinear weights and wrap-around
    (NumPy only; for throughput inputs where the content moves by a non-integer amount, as the
    panning corpus C2 of SURVEY.md §8d does)."""
    ix, iy = int(np.floor(sx)), int(np.floor(sy))
    ax, ay = float(sx - ix), float(sy - iy)
    f = frame.astype(np.float32)
    r00 = np.roll(f, (iy, ix), axis=(0, 1))
    r01 = np.roll(f, (iy, ix + 1), axis=(0, 1))
    r10 = np.roll(f, (iy + 1, ix), axis=(0, 1))
    r11 = np.roll(f, (iy + 1, ix + 1), axis=(0, 1))
    out = (1 - ay) * ((1 - ax) * r00 + ax * r01) + ay * ((1 - ax) * r10 + ax * r11)
    return np.ascontiguousarray(np.clip(np.round(out), 0, 255).astype(np.uint8))


def high_contrast_pair(h: int, w: int, seed: int, roll=(2, -3)):
    """Dark frame with random bright/dark rectangles + U{0,1,2} noise, rolled by (dy, dx): strong
    edges next to flat areas — the stress case for the accumulation precision of the box blur."""
    rng = np.random.default_rng(seed)
    a = np.full((h, w), 20, np.uint8)
    for _ in range(40):
        y, x = int(rng.integers(0, h - 40)), int(rng.integers(0, w - 40))
        a[y:y + int(rng.integers(8, 40)), x:x + int(rng.integers(8, 40))] = int(rng.integers(0, 2)) * 235 + 10
    a = (a + rng.integers(0, 3, size=a.shape)).astype(np.uint8)
    b = np.roll(a, roll, axis=(0, 1))
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def synth_net(h: int, w: int, seed: int, pitch: float = 17.0, line: int = 3, bgr: bool = True):
    """A fishing-net-like frame for the junction detector (junction_detector.cpp): dark, slightly wavy grid lines of
    width `line` every `pitch` pixels on a brighter noisy background -> cells of about (pitch - line)^2 pixels."""
    import cv2
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    ang = rng.uniform(-0.15, 0.15)
    u = xx * np.cos(ang) + yy * np.sin(ang) + 2.0 * np.sin(yy / 23.0 + rng.uniform(0, 6))
    v = -xx * np.sin(ang) + yy * np.cos(ang) + 2.0 * np.sin(xx / 19.0 + rng.uniform(0, 6))
    du = np.abs((u + rng.uniform(0, pitch)) % pitch - pitch / 2)
    dv = np.abs((v + rng.uniform(0, pitch)) % pitch - pitch / 2)
    net = (np.minimum(du, dv) < line / 2.0)
    base = 150 + 40 * np.sin(xx / 90.0) * np.cos(yy / 70.0) + rng.normal(0, 6, (h, w))
    img = np.where(net, 40 + rng.normal(0, 5, (h, w)), base)
    img = cv2.GaussianBlur(img.astype(np.float32), (0, 0), 0.8)
    g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if not bgr:
        return g
    out = np.stack([np.clip(g.astype(np.int32) + rng.integers(-8, 9, (h, w)), 0, 255) for _ in range(3)], -1)
    return out.astype(np.uint8)
