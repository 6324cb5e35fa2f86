"""NumPy restatement of ``cv2.cornerMinEigenVal`` / ``cv2.goodFeaturesToTrack`` (Shi-Tomasi).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows upstream OpenCV 4.x
``modules/imgproc/src/corner.cpp`` (``cornerEigenValsVecs``, ``calcMinEigenVal``),
``deriv.cpp`` (``Sobel`` with the scale folded into the smoothing kernel) and
``featureselect.cpp`` (``goodFeaturesToTrack``) — OpenCV is the reference's un-vendored dependency
(``ros2_ws/src/nueflow/setup.py:29``).  The float recipe is the one that reproduces the
*optimized* (SIMD/FMA) code path of the opencv-python-headless 4.13.0.92 wheel bit for bit
(SURVEY.md Appendix A.4); it is pinned against the wheel in ``tests/test_oracle_sparse.py``.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
f64 = np.float64
SIMD_BLOCK = 32   # column block of the wheel's vectorised Sobel row filter (AVX-512 host)


def _fma(a, b, c):
    """float32 fused multiply-add (product exact in float64; one rounding to float32)."""
    return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)


def _pad101(a, r):
    return np.pad(a, r, mode="reflect")


def _cov_box_sums(img_u8: np.ndarray, block_size: int = 3):
    """Unnormalised block_size^2 box sums (REFLECT_101, accumulated in double, cast to float32) of dx*dx, dx*dy, dy*dy,
    the Sobel derivatives scaled by 1 / (4 block_size 255) with the wheel's FMA placement and SIMD-tail rule."""
    h, w = img_u8.shape
    scale = 1.0 / (4.0 * block_size * 255.0)
    k = (np.array([1.0, 2.0, 1.0]) * scale).astype(f32)
    p = _pad101(img_u8.astype(f32), 1)                       # [h+2, w+2]
    # dx: row pass [-1,0,1] (exact), column pass [k0,k1,k0] as fma((r[y-1]+r[y+1]), k0, r[y]*k1)
    r = p[:, 2:] - p[:, :-2]                                 # [h+2, w]
    dx = _fma((r[:-2] + r[2:]).astype(f32), np.broadcast_to(k[0], (h, w)), (r[1:-1] * k[1]).astype(f32))
    # dy: row pass [k0,k1,k2] as fma(k2,p[x+1], fma(k1,p[x], k0*p[x-1])), column pass [-1,0,1]
    t = (p[:, :-2] * k[0]).astype(f32)
    t = _fma(np.broadcast_to(k[1], t.shape), p[:, 1:-1], t)
    rw = _fma(np.broadcast_to(k[2], t.shape), p[:, 2:], t)   # [h+2, w]
    # SIMD tail of the wheel's row filter (probe-verified): columns past the last full block of 32
    # are computed without FMA as (p[x-1]*k0 + p[x]*k1) + p[x+1]*k2
    wb = (w // SIMD_BLOCK) * SIMD_BLOCK
    if wb < w:
        tail = ((p[:, :-2] * k[0]).astype(f32) + (p[:, 1:-1] * k[1]).astype(f32)).astype(f32)
        tail = (tail + (p[:, 2:] * k[2]).astype(f32)).astype(f32)
        rw[:, wb:] = tail[:, wb:]
    dy = (rw[2:] - rw[:-2]).astype(f32)
    cxx = (dx * dx).astype(f32)
    cxy = (dx * dy).astype(f32)
    cyy = (dy * dy).astype(f32)
    rb = block_size // 2

    def box(c):
        cp = _pad101(c.astype(f64), rb) if block_size % 2 == 1 else None
        if cp is None:
            raise NotImplementedError("even block sizes")
        s = np.zeros((h, w), f64)
        for j in range(block_size):
            for i in range(block_size):
                s += cp[j:j + h, i:i + w]
        return s.astype(f32)

    return box(cxx), box(cxy), box(cyy)


def corner_min_eigenval(img_u8: np.ndarray, block_size: int = 3) -> np.ndarray:
    """== cv2.cornerMinEigenVal(img, block_size, ksize=3) for uint8 input (block_size 3 verified)."""
    sxx, sxy, syy = _cov_box_sums(img_u8, block_size)
    a = (sxx * f32(0.5)).astype(f32)
    b = sxy
    c = (syy * f32(0.5)).astype(f32)
    d = ((a - c).astype(f32) * (a - c).astype(f32)).astype(f32)
    d = (d + (b * b).astype(f32)).astype(f32)
    return ((a + c).astype(f32) - np.sqrt(d).astype(f32)).astype(f32)


def corner_harris(img_u8: np.ndarray, block_size: int = 3, k: float = 0.04) -> np.ndarray:
    """== cv2.cornerHarris(img, block_size, 3, k) for uint8 input.  Same derivative products and box sums as
    ``corner_min_eigenval`` (a, b, c NOT halved); the response as the wheel computes it over the image as ONE continuous
    row of n = w * h pixels (probe-verified): ``(a c - b b) - k ((a + c)(a + c))`` in float in the 8-wide body,
    ``(a c - b b) - (k (a + c)) (a + c)`` in the 4-wide step for pixels n - n % 8 .. n - n % 4, and the last n % 4 pixels
    in double with the double ``k``: ``float(double(float(a c - b b)) - (k (a + c)) (a + c))``."""
    a, b, c = _cov_box_sums(img_u8, block_size)
    h, w = a.shape
    a, b, c = a.ravel(), b.ravel(), c.ravel()
    kf = f32(k)
    t1 = ((a * c).astype(f32) - (b * b).astype(f32)).astype(f32)
    s = (a + c).astype(f32)
    out = (t1 - (kf * (s * s).astype(f32)).astype(f32)).astype(f32)
    n = h * w
    n8, n4 = n - n % 8, n - n % 4
    out[n8:n4] = (t1[n8:n4] - ((kf * s[n8:n4]).astype(f32) * s[n8:n4]).astype(f32)).astype(f32)
    sd = s[n4:].astype(f64)
    out[n4:] = (t1[n4:].astype(f64) - (float(k) * sd) * sd).astype(f32)
    return out.reshape(h, w)


def good_features(img_u8: np.ndarray, max_corners: int, quality_level: float, min_distance: float,
                  block_size: int = 3, eig: np.ndarray | None = None) -> np.ndarray:
    """== cv2.goodFeaturesToTrack(img, max_corners, quality_level, min_distance, blockSize=block_size)
    → float32 [N,1,2] (x, y)."""
    if eig is None:
        eig = corner_min_eigenval(img_u8, block_size)
    h, w = eig.shape
    max_val = float(eig.max())
    thr = f32(max_val * quality_level)
    e = np.where(eig > thr, eig, f32(0)).astype(f32)         # THRESH_TOZERO
    ep = np.pad(e, 1, mode="constant", constant_values=-np.inf)
    dil = np.full((h, w), -np.inf, f32)
    for j in range(3):
        for i in range(3):
            dil = np.maximum(dil, ep[j:j + h, i:i + w])
    cand = (e != 0) & (e == dil)
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    vals = e[ys, xs]
    # std::sort with greaterThanPtr: value descending, ties by DESCENDING address (y, then x)
    order = np.lexsort((-(ys * w + xs), -vals.astype(f64)))
    ys, xs = ys[order], xs[order]
    out = []
    if min_distance >= 1:
        cell = int(np.rint(min_distance))
        gw, gh = (w + cell - 1) // cell, (h + cell - 1) // cell
        grid = {}
        md2 = min_distance * min_distance
        for y, x in zip(ys.tolist(), xs.tolist()):
            xc, yc = x // cell, y // cell
            good = True
            for yy in range(max(0, yc - 1), min(gh - 1, yc + 1) + 1):
                for xx in range(max(0, xc - 1), min(gw - 1, xc + 1) + 1):
                    for (px, py) in grid.get((yy, xx), ()):
                        dx, dy = x - px, y - py
                        if dx * dx + dy * dy < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid.setdefault((yc, xc), []).append((x, y))
                out.append((x, y))
                if max_corners > 0 and len(out) == max_corners:
                    break
    else:
        n = len(xs) if max_corners <= 0 else min(max_corners, len(xs))
        out = list(zip(xs[:n].tolist(), ys[:n].tolist()))
    return np.array(out, f32).reshape(-1, 1, 2)
