"""NumPy restatement of OpenCV's ``calcOpticalFlowFarneback`` (CPU, non-OpenCL path).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Stage-level oracle: cv2
exposes only the final flow; this restatement exposes every intermediate field
(level images I, polynomial coefficients R, matrix field M, per-level flow) so
each CUDA kernel can be checked on its own.

Follows upstream OpenCV 4.x ``modules/video/src/optflowgf.cpp`` (not on disk —
the reference declares OpenCV un-vendored: ``ros2_ws/src/nueflow/setup.py:29``;
the call sits behind ``ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:194``):

* ``level_schedule``      — ``FarnebackOpticalFlowImpl::calc`` level loop
* ``prepare_gaussian``    — ``FarnebackPrepareGaussian``
* ``pyramid_level``       — ``convertTo(CV_32F)`` + ``GaussianBlur`` + ``resize(INTER_LINEAR)``
                            (``imgproc/src/smooth.dispatch.cpp``, ``resize.cpp``)
* ``poly_exp``            — ``FarnebackPolyExp``
* ``update_matrices``     — ``FarnebackUpdateMatrices``
* ``blur_solve``          — ``FarnebackUpdateFlow_Blur`` / ``FarnebackUpdateFlow_GaussianBlur``
* ``upsample_flow``       — ``resize(prevFlow, INTER_LINEAR) * (1/pyr_scale)``
* ``farneback``           — the whole call

Pinned against the cv2 4.13.0 wheel in ``tests/test_oracle_farneback.py`` and
``tests/golden/farneback_*.npz``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_FARNEBACK_GAUSSIAN = 256

f32 = np.float32
f64 = np.float64


def cv_round(x: float) -> int:
    """cvRound: round half to even (SSE cvtsd2si)."""
    return int(np.rint(x))


# ----------------------------------------------------------------------------- schedule
@dataclass
class Level:
    k: int
    scale: float
    sigma: float
    ksize: int
    width: int
    height: int


def level_schedule(width: int, height: int, pyr_scale: float, levels: int) -> List[Level]:
    """Coarse→fine list of levels.  ``levels`` is clamped so the coarsest level is
    ≥ 32 px on both sides; note cv2 runs ``levels_eff + 1`` scales."""
    min_size = 32
    scale = 1.0
    k = 0
    while k < levels:
        scale *= pyr_scale
        if width * scale < min_size or height * scale < min_size:
            break
        k += 1
    levels_eff = k
    out = []
    for k in range(levels_eff, -1, -1):
        scale = 1.0
        for _ in range(k):
            scale *= pyr_scale
        sigma = (1.0 / scale - 1.0) * 0.5
        ksize = max(cv_round(sigma * 5) | 1, 3)
        out.append(Level(k, scale, sigma, ksize, cv_round(width * scale), cv_round(height * scale)))
    return out


# ----------------------------------------------------------------------------- gaussian kernels
def gaussian_kernel_f32(ksize: int, sigma: float) -> np.ndarray:
    """``cv::getGaussianKernel(ksize, sigma, CV_32F)``: fixed table for sigma<=0
    and small ksize, else exp(-x^2/2s^2) normalised in double, cast to f32."""
    if sigma <= 0:
        table = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
                 7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
        if ksize in table:
            return np.array(table[ksize], f32)
        sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8
    x = np.arange(ksize, dtype=f64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 * x * x / (sigma * sigma))
    k /= k.sum()
    return k.astype(f32)


def prepare_gaussian(n: int, sigma: float):
    """``FarnebackPrepareGaussian`` → (g, xg, xxg) f32 arrays indexed [0..n] for x>=0
    (g is even, xg odd, xxg even) and the four doubles ig11, ig03, ig33, ig55."""
    if sigma < np.finfo(f32).eps:
        sigma = n * 0.3
    xs = np.arange(-n, n + 1)
    g = np.exp(-(xs * xs) / (2.0 * sigma * sigma)).astype(f32)
    s = 1.0 / float(g.astype(f64).sum())
    g = (g.astype(f64) * s).astype(f32)
    xg = (xs * g.astype(f64)).astype(f32)
    xxg = (xs * xs * g.astype(f64)).astype(f32)
    G = np.zeros((6, 6), f64)
    gd = g.astype(f64)
    for iy, y in enumerate(xs):
        for ix, x in enumerate(xs):
            w = gd[iy] * gd[ix]
            G[0, 0] += w
            G[1, 1] += w * x * x
            G[3, 3] += w * x * x * x * x
            G[5, 5] += w * x * x * y * y
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    invG = np.linalg.inv(G)
    return (g[n:].copy(), xg[n:].copy(), xxg[n:].copy(),
            float(invG[1, 1]), float(invG[0, 3]), float(invG[3, 3]), float(invG[5, 5]))


# ----------------------------------------------------------------------------- pyramid stage
def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(idx)
    p = 2 * (n - 1)
    idx = np.mod(idx, p)
    return np.where(idx >= n, p - idx, idx)


def gaussian_blur_f32(img: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """Separable GaussianBlur on f32 with BORDER_REFLECT_101, rows then columns."""
    k = gaussian_kernel_f32(ksize, sigma)
    r = ksize // 2
    h, w = img.shape
    xi = _reflect101(np.arange(-r, w + r), w)
    tmp = np.zeros((h, w), f32)
    padded = img[:, xi]
    for i in range(ksize):
        tmp += k[i] * padded[:, i:i + w]
    yi = _reflect101(np.arange(-r, h + r), h)
    padded = tmp[yi, :]
    out = np.zeros((h, w), f32)
    for i in range(ksize):
        out += k[i] * padded[i:i + h, :]
    return out


def _linear_tab(dst: int, src: int):
    scale = 1.0 / (float(dst) / float(src))  # cv2: scale_x = 1./inv_scale_x
    d = np.arange(dst, dtype=f64)
    fx = ((d + 0.5) * scale - 0.5).astype(f32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(f32)).astype(f32)
    lo = sx < 0
    fx[lo] = 0
    sx[lo] = 0
    hi = sx >= src - 1
    fx[hi] = 0
    sx[hi] = src - 1
    sx1 = np.minimum(sx + 1, src - 1)
    return sx, sx1, (f32(1.0) - fx).astype(f32), fx


def resize_linear_f32(img: np.ndarray, width: int, height: int) -> np.ndarray:
    """``cv::resize(..., INTER_LINEAR)`` for f32 (any channel count in last axis)."""
    h, w = img.shape[:2]
    if (w, h) == (width, height):
        return img.copy()
    x0, x1, a0, a1 = _linear_tab(width, w)
    y0, y1, b0, b1 = _linear_tab(height, h)
    if img.ndim == 3:
        a0 = a0[None, :, None]; a1 = a1[None, :, None]
        b0 = b0[:, None, None]; b1 = b1[:, None, None]
    else:
        a0 = a0[None, :]; a1 = a1[None, :]
        b0 = b0[:, None]; b1 = b1[:, None]
    r0 = img[y0][:, x0] * a0 + img[y0][:, x1] * a1
    r1 = img[y1][:, x0] * a0 + img[y1][:, x1] * a1
    return (r0 * b0 + r1 * b1).astype(f32)


def _area_tab(dst: int, src: int) -> np.ndarray:
    """Dense [dst, src] weight matrix of ``computeResizeAreaTab``."""
    scale = float(src) / float(dst)
    tab = np.zeros((dst, src), f64)
    for d in range(dst):
        fsx1 = d * scale
        fsx2 = fsx1 + scale
        cell = min(scale, src - fsx1)
        sx1 = int(math.ceil(fsx1))
        sx2 = int(math.floor(fsx2))
        sx2 = min(sx2, src - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab[d, sx1 - 1] += (sx1 - fsx1) / cell
        for sx in range(sx1, sx2):
            tab[d, sx] += 1.0 / cell
        if fsx2 - sx2 > 1e-3:
            tab[d, sx2] += min(min(fsx2 - sx2, 1.0), cell) / cell
    return tab


def resize_area_f32(img: np.ndarray, width: int, height: int) -> np.ndarray:
    """``cv::resize(..., INTER_AREA)`` (down-scaling) for f32 [h,w,c]."""
    h, w = img.shape[:2]
    tx = _area_tab(width, w).astype(f32)
    ty = _area_tab(height, h).astype(f32)
    tmp = np.einsum("dw,hwc->hdc", tx, img.astype(f32)).astype(f32)
    return np.einsum("eh,hdc->edc", ty, tmp).astype(f32)


def pyramid_level(img_u8: np.ndarray, lv: Level) -> np.ndarray:
    f = img_u8.astype(f32)
    f = gaussian_blur_f32(f, lv.ksize, lv.sigma)
    return resize_linear_f32(f, lv.width, lv.height)


# ----------------------------------------------------------------------------- PolyExp
def poly_exp(I: np.ndarray, n: int, sigma: float) -> np.ndarray:
    """``FarnebackPolyExp``: f32 [h,w] → f32 [h,w,5] = (y, x, yy, xx, xy) coefficients."""
    g, xg, xxg, ig11, ig03, ig33, ig55 = prepare_gaussian(n, sigma)
    h, w = I.shape
    I = I.astype(f32)
    ys = np.arange(h)
    # vertical pass, f32 accumulation in cv2's order
    r0 = I * g[0]
    r1 = np.zeros_like(I)
    r2 = np.zeros_like(I)
    for k in range(1, n + 1):
        a = I[np.maximum(ys - k, 0)]
        b = I[np.minimum(ys + k, h - 1)]
        p = a + b
        r0 = r0 + g[k] * p
        r1 = r1 + xg[k] * (b - a)
        r2 = r2 + xxg[k] * p
    # horizontal pass, double accumulation
    xs = np.arange(w)
    r0d, r1d, r2d = r0.astype(f64), r1.astype(f64), r2.astype(f64)
    b1 = r0d * f64(g[0]); b2 = np.zeros_like(r0d); b3 = r1d * f64(g[0])
    b4 = np.zeros_like(r0d); b5 = r2d * f64(g[0]); b6 = np.zeros_like(r0d)
    for k in range(1, n + 1):
        xm = np.maximum(xs - k, 0)
        xp = np.minimum(xs + k, w - 1)
        # cv2 forms tg in float: row[] is float*, (float + float) then * float → promoted to double on +=
        tg = (r0[:, xp] + r0[:, xm]).astype(f32)
        b1 += (tg * g[k]).astype(f32)
        b4 += (tg * xxg[k]).astype(f32)
        b2 += ((r0[:, xp] - r0[:, xm]).astype(f32) * xg[k]).astype(f32)
        b3 += ((r1[:, xp] + r1[:, xm]).astype(f32) * g[k]).astype(f32)
        b6 += ((r1[:, xp] - r1[:, xm]).astype(f32) * xg[k]).astype(f32)
        b5 += ((r2[:, xp] + r2[:, xm]).astype(f32) * g[k]).astype(f32)
    R = np.empty((h, w, 5), f32)
    R[..., 1] = (b2 * ig11).astype(f32)
    R[..., 0] = (b3 * ig11).astype(f32)
    R[..., 3] = (b1 * ig03 + b4 * ig33).astype(f32)
    R[..., 2] = (b1 * ig03 + b5 * ig33).astype(f32)
    R[..., 4] = (b6 * ig55).astype(f32)
    return R


# ----------------------------------------------------------------------------- UpdateMatrices
_BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], f32)


def border_scale(n: int) -> np.ndarray:
    """Per-coordinate attenuation (product over the two sides)."""
    s = np.ones(n, f32)
    idx = np.arange(n)
    for i in range(n):
        v = f32(1.0)
        if i < 5:
            v = f32(v * _BORDER[i])
        if i >= n - 5:
            v = f32(v * _BORDER[n - i - 1])
        s[i] = v
    return s


def update_matrices(R0: np.ndarray, R1: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """``FarnebackUpdateMatrices`` over the whole frame → M f32 [h,w,5]
    = (g11, g12, g22, h1, h2)."""
    h, w = flow.shape[:2]
    x = np.arange(w, dtype=f32)[None, :]
    y = np.arange(h, dtype=f32)[:, None]
    dx = flow[..., 0].astype(f32)
    dy = flow[..., 1].astype(f32)
    fx = (x + dx).astype(f32)
    fy = (y + dy).astype(f32)
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(f32)).astype(f32)
    fy = (fy - y1.astype(f32)).astype(f32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    x1c = np.clip(x1, 0, max(w - 2, 0))
    y1c = np.clip(y1, 0, max(h - 2, 0))
    one = f32(1.0)
    a00 = ((one - fx) * (one - fy)).astype(f32)
    a01 = (fx * (one - fy)).astype(f32)
    a10 = ((one - fx) * fy).astype(f32)
    a11 = (fx * fy).astype(f32)
    x2 = np.minimum(x1c + 1, w - 1)
    y2 = np.minimum(y1c + 1, h - 1)
    r = []
    for c in range(5):
        ch = R1[..., c]
        v = (a00 * ch[y1c, x1c] + a01 * ch[y1c, x2] + a10 * ch[y2, x1c] + a11 * ch[y2, x2]).astype(f32)
        r.append(v)
    r2, r3, r4, r5, r6 = r
    r4 = np.where(inside, (R0[..., 2] + r4) * f32(0.5), R0[..., 2]).astype(f32)
    r5 = np.where(inside, (R0[..., 3] + r5) * f32(0.5), R0[..., 3]).astype(f32)
    r6 = np.where(inside, (R0[..., 4] + r6) * f32(0.25), R0[..., 4] * f32(0.5)).astype(f32)
    r2 = np.where(inside, r2, f32(0)).astype(f32)
    r3 = np.where(inside, r3, f32(0)).astype(f32)
    r2 = ((R0[..., 0] - r2) * f32(0.5)).astype(f32)
    r3 = ((R0[..., 1] - r3) * f32(0.5)).astype(f32)
    r2 = (r2 + r4 * dy + r6 * dx).astype(f32)
    r3 = (r3 + r6 * dy + r5 * dx).astype(f32)
    sc = (border_scale(h)[:, None] * border_scale(w)[None, :]).astype(f32) if (h >= 10 and w >= 10) else None
    if sc is None:
        sx = border_scale(w)[None, :]
        sy = border_scale(h)[:, None]
        sc = (sx * sy).astype(f32)
    r2 = r2 * sc; r3 = r3 * sc; r4 = r4 * sc; r5 = r5 * sc; r6 = r6 * sc
    M = np.empty((h, w, 5), f32)
    M[..., 0] = r4 * r4 + r6 * r6
    M[..., 1] = (r4 + r5) * r6
    M[..., 2] = r5 * r5 + r6 * r6
    M[..., 3] = r4 * r2 + r6 * r3
    M[..., 4] = r6 * r2 + r5 * r3
    return M


# ----------------------------------------------------------------------------- blur + solve
def box_blur(M: np.ndarray, winsize: int) -> np.ndarray:
    """Window [-m, m] (m = winsize//2), replicate border, double sums, × 1/winsize²."""
    m = winsize // 2
    h, w = M.shape[:2]
    yi = np.clip(np.arange(-m - 1, h + m), 0, h - 1)
    c = np.cumsum(M.astype(f64)[yi], axis=0)
    v = c[2 * m + 1:] - c[:h]
    xi = np.clip(np.arange(-m - 1, w + m), 0, w - 1)
    c = np.cumsum(v[:, xi], axis=1)
    s = c[:, 2 * m + 1:] - c[:, :w]
    return s * (1.0 / (winsize * winsize))


def gauss_blur(M: np.ndarray, winsize: int) -> np.ndarray:
    m = winsize // 2
    sigma = m * 0.3
    k = np.exp(-(np.arange(m + 1, dtype=f64) ** 2) / (2 * sigma * sigma)).astype(f32)
    s = float(k[0]) + 2.0 * float(k[1:].astype(f64).sum())
    k = (k.astype(f64) * (1.0 / s)).astype(f32)
    h, w = M.shape[:2]
    ys = np.arange(h)
    v = M * k[0]
    for i in range(1, m + 1):
        v = v + (M[np.minimum(ys + i, h - 1)] + M[np.maximum(ys - i, 0)]) * k[i]
    xs = np.arange(w)
    o = v * k[0]
    for i in range(1, m + 1):
        o = o + (v[:, np.minimum(xs + i, w - 1)] + v[:, np.maximum(xs - i, 0)]) * k[i]
    return o.astype(f32)


def solve(B: np.ndarray) -> np.ndarray:
    g11, g12, g22, h1, h2 = (B[..., i] for i in range(5))
    if B.dtype == f32:
        idet = f32(1.0) / (g11 * g22 - g12 * g12 + f32(1e-3))
    else:
        idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    out = np.empty(B.shape[:2] + (2,), f32)
    out[..., 0] = ((g11 * h2 - g12 * h1) * idet).astype(f32)
    out[..., 1] = ((g22 * h1 - g12 * h2) * idet).astype(f32)
    return out


def blur_solve(M: np.ndarray, winsize: int, gaussian: bool) -> np.ndarray:
    return solve(gauss_blur(M, winsize) if gaussian else box_blur(M, winsize))


def upsample_flow(prev_flow: np.ndarray, width: int, height: int, pyr_scale: float) -> np.ndarray:
    up = resize_linear_f32(prev_flow, width, height)
    return (up.astype(f64) * (1.0 / pyr_scale)).astype(f32)


# ----------------------------------------------------------------------------- whole call
def farneback(prev: np.ndarray, nxt: np.ndarray, flow0: Optional[np.ndarray] = None, pyr_scale=0.5,
              levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0,
              trace: Optional[dict] = None) -> np.ndarray:
    """Restatement of ``cv2.calcOpticalFlowFarneback``.  If ``trace`` is a dict it
    receives per-level intermediates keyed ``(k, name)``."""
    assert prev.shape == nxt.shape and prev.ndim == 2 and prev.dtype == np.uint8
    H, W = prev.shape
    sched = level_schedule(W, H, pyr_scale, levels)
    gaussian = bool(flags & OPTFLOW_FARNEBACK_GAUSSIAN)
    prev_flow = None
    for lv in sched:
        if prev_flow is None:
            if flags & OPTFLOW_USE_INITIAL_FLOW:
                assert flow0 is not None and flow0.shape == (H, W, 2)
                if (lv.width, lv.height) == (W, H):
                    flow = flow0.astype(f32).copy()
                else:
                    flow = resize_area_f32(flow0.astype(f32), lv.width, lv.height)
                flow = (flow.astype(f64) * lv.scale).astype(f32)
            else:
                flow = np.zeros((lv.height, lv.width, 2), f32)
        else:
            flow = upsample_flow(prev_flow, lv.width, lv.height, pyr_scale)
        I0 = pyramid_level(prev, lv)
        I1 = pyramid_level(nxt, lv)
        R0 = poly_exp(I0, poly_n, poly_sigma)
        R1 = poly_exp(I1, poly_n, poly_sigma)
        if trace is not None:
            trace[(lv.k, "I0")] = I0; trace[(lv.k, "I1")] = I1
            trace[(lv.k, "R0")] = R0; trace[(lv.k, "R1")] = R1
            trace[(lv.k, "flow_in")] = flow.copy()
        M = update_matrices(R0, R1, flow)
        if trace is not None:
            trace[(lv.k, "M0")] = M
        for i in range(iterations):
            flow = blur_solve(M, winsize, gaussian)
            if i < iterations - 1:
                M = update_matrices(R0, R1, flow)
        if trace is not None:
            trace[(lv.k, "flow_out")] = flow.copy()
        prev_flow = flow
    return prev_flow
