"""NumPy restatement of OpenCV's pyramidal Lucas-Kanade path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows upstream OpenCV 4.x
``modules/imgproc/src/pyramids.cpp`` (``pyrDown``), ``modules/video/src/lkpyramid.cpp``
(``calcScharrDeriv``, ``buildOpticalFlowPyramid``, ``LKTrackerInvoker::operator()``,
``SparsePyrLKOpticalFlowImpl::calc``) — OpenCV is the reference's un-vendored dependency
(``ros2_ws/src/nueflow/setup.py:29``).  Integer stages are bit-exact; the tracker accumulates in
float32 in plain raster order (cv2 accumulates in SIMD lanes), so positions agree to ~1e-4 px.
Pinned against the cv2 4.13.0 wheel in ``tests/test_oracle_sparse.py``.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
W_BITS = 14


def _reflect101(idx, n):
    if n == 1:
        return np.zeros_like(idx)
    p = 2 * (n - 1)
    idx = np.mod(idx, p)
    return np.where(idx >= n, p - idx, idx)


def pyr_down(img: np.ndarray) -> np.ndarray:
    """== cv2.pyrDown for uint8: [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int32)
    src = img.astype(np.int32)
    xi = _reflect101(2 * np.arange(ow)[:, None] + np.arange(-2, 3)[None, :], w)      # [ow,5]
    rows = (src[:, xi] * k[None, None, :]).sum(-1)                                     # [h, ow]
    yi = _reflect101(2 * np.arange(oh)[:, None] + np.arange(-2, 3)[None, :], h)      # [oh,5]
    out = (rows[yi, :] * k[None, :, None]).sum(1)                                      # [oh, ow]
    return ((out + 128) >> 8).astype(np.uint8)


def scharr_deriv(img: np.ndarray) -> np.ndarray:
    """== calcScharrDeriv: int16 [h,w,2] = (dx, dy), BORDER_REFLECT_101."""
    h, w = img.shape
    s = img.astype(np.int32)
    yi0 = _reflect101(np.arange(h) - 1, h)
    yi1 = _reflect101(np.arange(h) + 1, h)
    t0 = (s[yi0] + s[yi1]) * 3 + s * 10
    t1 = s[yi1] - s[yi0]
    xi0 = _reflect101(np.arange(w) - 1, w)
    xi1 = _reflect101(np.arange(w) + 1, w)
    dx = t0[:, xi1] - t0[:, xi0]
    dy = (t1[:, xi1] + t1[:, xi0]) * 3 + t1 * 10
    return np.stack([dx, dy], -1).astype(np.int16)


def build_pyramid(img: np.ndarray, win=(21, 21), max_level=3):
    """Levels of buildOpticalFlowPyramid (without its border): stop when the next level is <= win."""
    levels = [img]
    for _ in range(max_level):
        h, w = levels[-1].shape
        nh, nw = (h + 1) // 2, (w + 1) // 2
        if nw <= win[0] or nh <= win[1]:
            break
        levels.append(pyr_down(levels[-1]))
    return levels


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _cv_round(v):
    return int(np.rint(np.float64(v)))


def _patch(img_pad, x0, y0, win, iw, shift, border):
    """Fixed-point bilinear window [win_h, win_w] starting at integer (x0, y0) of the padded image."""
    ww, wh = win
    ys, xs = y0 + border, x0 + border
    a = img_pad[ys:ys + wh + 1, xs:xs + ww + 1].astype(np.int64)
    v = a[:-1, :-1] * iw[0] + a[:-1, 1:] * iw[1] + a[1:, :-1] * iw[2] + a[1:, 1:] * iw[3]
    return _descale(v, shift)


def _weights(a, b):
    a, b = f32(a), f32(b)
    one = f32(1.0)
    s = f32(1 << W_BITS)
    iw00 = _cv_round(f32(f32((one - a) * (one - b)) * s))
    iw01 = _cv_round(f32(f32(a * (one - b)) * s))
    iw10 = _cv_round(f32(f32((one - a) * b) * s))
    return iw00, iw01, iw10, (1 << W_BITS) - iw00 - iw01 - iw10


def calc_pyrlk(prev, nxt, prev_pts, next_pts=None, win=(21, 21), max_level=3, max_count=30, epsilon=0.01,
               flags=0, min_eig_threshold=1e-4):
    """Restatement of cv2.calcOpticalFlowPyrLK → (nextPts [N,1,2], status [N,1], err [N,1])."""
    ww, wh = win
    pts = np.asarray(prev_pts, f32).reshape(-1, 2)
    n = len(pts)
    P = build_pyramid(prev, win, max_level)
    Q = build_pyramid(nxt, win, max_level)
    L = len(P) - 1
    max_count = min(max(int(max_count), 0), 100)
    eps2 = f32(min(max(float(epsilon), 0.0), 10.0) ** 2)
    status = np.ones(n, np.uint8)
    err = np.zeros(n, f32)
    out = np.zeros((n, 2), f32)
    use_init = bool(flags & 4)
    if use_init:
        init = np.asarray(next_pts, f32).reshape(-1, 2)
    half = np.array([(ww - 1) * 0.5, (wh - 1) * 0.5], f32)
    FLT_SCALE = f32(1.0 / (1 << 20))
    border = max(ww, wh) + 2
    for level in range(L, -1, -1):
        I, J = P[level], Q[level]
        rows, cols = I.shape
        Ip = np.pad(I, border, mode="reflect")
        Jp = np.pad(J, border, mode="reflect")
        D = scharr_deriv(I)
        Dx = np.pad(D[..., 0], border, mode="constant")
        Dy = np.pad(D[..., 1], border, mode="constant")
        for i in range(n):
            prev_pt = (pts[i] * f32(1.0 / (1 << level))).astype(f32)
            if level == L:
                next_pt = (init[i] * f32(1.0 / (1 << level))).astype(f32) if use_init else prev_pt.copy()
            else:
                next_pt = (out[i] * f32(2.0)).astype(f32)
            out[i] = next_pt
            prev_pt = (prev_pt - half).astype(f32)
            ix, iy = int(np.floor(prev_pt[0])), int(np.floor(prev_pt[1]))
            if ix < -ww or ix >= cols or iy < -wh or iy >= rows:
                if level == 0:
                    status[i] = 0
                    err[i] = 0
                continue
            iw = _weights(prev_pt[0] - f32(ix), prev_pt[1] - f32(iy))
            Iw = _patch(Ip, ix, iy, win, iw, W_BITS - 5, border)
            dIx = _patch(Dx, ix, iy, win, iw, W_BITS, border)
            dIy = _patch(Dy, ix, iy, win, iw, W_BITS, border)
            A11 = f32((dIx * dIx).astype(f32).sum(dtype=f32)) * FLT_SCALE
            A12 = f32((dIx * dIy).astype(f32).sum(dtype=f32)) * FLT_SCALE
            A22 = f32((dIy * dIy).astype(f32).sum(dtype=f32)) * FLT_SCALE
            Dt = f32(f32(A11 * A22) - f32(A12 * A12))
            min_eig = f32(f32(A22 + A11) - np.sqrt(f32(f32((A11 - A22) * (A11 - A22)) + f32(f32(4.0) * A12 * A12)))) \
                / f32(2 * ww * wh)
            if flags & 8:
                err[i] = min_eig
            if min_eig < min_eig_threshold or Dt < np.finfo(f32).eps:
                if level == 0:
                    status[i] = 0
                continue
            Dt = f32(1.0) / Dt
            next_pt = (next_pt - half).astype(f32)
            prev_delta = np.zeros(2, f32)
            for j in range(max_count):
                jx, jy = int(np.floor(next_pt[0])), int(np.floor(next_pt[1]))
                if jx < -ww or jx >= cols or jy < -wh or jy >= rows:
                    if level == 0:
                        status[i] = 0
                    break
                jw = _weights(next_pt[0] - f32(jx), next_pt[1] - f32(jy))
                diff = _patch(Jp, jx, jy, win, jw, W_BITS - 5, border) - Iw
                b1 = f32((diff * dIx).astype(f32).sum(dtype=f32)) * FLT_SCALE
                b2 = f32((diff * dIy).astype(f32).sum(dtype=f32)) * FLT_SCALE
                delta = np.array([f32(f32(f32(A12 * b2) - f32(A22 * b1)) * Dt),
                                  f32(f32(f32(A12 * b1) - f32(A11 * b2)) * Dt)], f32)
                next_pt = (next_pt + delta).astype(f32)
                out[i] = (next_pt + half).astype(f32)
                if f32(delta[0] * delta[0] + delta[1] * delta[1]) <= eps2:
                    break
                if j > 0 and abs(delta[0] + prev_delta[0]) < 0.01 and abs(delta[1] + prev_delta[1]) < 0.01:
                    out[i] = (out[i] - delta * f32(0.5)).astype(f32)
                    break
                prev_delta = delta
            if status[i] and level == 0 and not (flags & 8):
                np_ = (out[i] - half).astype(f32)
                jx, jy = int(np.floor(np_[0])), int(np.floor(np_[1]))
                if jx < -ww or jx >= cols or jy < -wh or jy >= rows:
                    status[i] = 0
                    continue
                jw = _weights(np_[0] - f32(jx), np_[1] - f32(jy))
                diff = _patch(Jp, jx, jy, win, jw, W_BITS - 5, border) - Iw
                err[i] = f32(np.abs(diff).astype(f32).sum(dtype=f32)) * f32(1.0 / (32 * ww * wh))
    return out.reshape(-1, 1, 2), status.reshape(-1, 1), err.reshape(-1, 1)
