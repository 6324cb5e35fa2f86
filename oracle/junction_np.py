"""ORACLE (test infrastructure only — never imported by the product path).

NumPy / SciPy restatement of the reference's junction detector
(``ros2_ws/src/junction_point_detector/src/junction_detector.cpp``):

  :3-28     ``dampenIntensity``          gain = clamp((R - B) * incline + intercept, 0, 1); channels * gain, truncated
  :46-53    ``cvtColor(BGR2GRAY)``, ``GaussianBlur(gray, (3, 3), 0)``
  :56       ``adaptiveThreshold(blur, 255, ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, 11, 2)``
  :72       ``findContours(thresh, RETR_TREE, CHAIN_APPROX_SIMPLE)``
  :76-127   per contour: ``contourArea`` window around grid_area, ``boundingRect`` fill ratio and aspect tests,
            the four box corners pushed out by one pixel become junction candidates
  :129-185  nanoflann KD-tree (leaf size 7), approximate radius search (eps 10), greedy clusters of >= 3 -> centres

The pixel stages follow the cv2 4.13.0 wheel (the reference links whatever OpenCV is installed) and are pinned against
it bit for bit in ``tests/test_oracle_junction.py``:
* 3x3 Gaussian on uint8: ``(sum of [1 2 1] x [1 2 1] + 8) >> 4``, BORDER_REFLECT_101;
* adaptive threshold: float 11x11 Gaussian (sigma 2.0, BORDER_REPLICATE), row pass then column pass in the wheel's
  exact operation order — rows: ``acc = k0 x0; acc = fma(x_i, k_i, acc)``, except the last ``width % 4`` columns where
  taps 1..8 are multiply-then-add and taps 9, 10 fused (the compiled scalar tail); columns: symmetric
  ``acc = k5 x5; acc = fma(x_{5+i} + x_{5-i}, k_{5+i}, acc)``, multiply-then-add in the last ``width % 8`` columns —
  rounded half-to-even to uint8; 255 where ``blur - mean > -2``;
* ``findContours(RETR_TREE)`` + ``contourArea`` + ``boundingRect`` restated WITHOUT border following, the way the
  device computes it: foreground components (8-connected) and background components (4-connected) of the zero-padded
  image; every border is the interface of one foreground and one background component — the OUTER border of F (its
  discovery pixel is F's first pixel in raster order) or the HOLE border around H (discovery pixel: left of H's first
  pixel); the polygon area is the shoelace sum over the cyclic order of the interface's pixel edges ("cracks", successor
  rule with 8-connectivity at saddle points); hole bounding boxes are H's box grown by one; parents follow from which
  component encloses which; cv2's output order is the depth-first pre-order of that tree with siblings in DESCENDING
  discovery order.
The clustering restates nanoflann 1.5 (``middleSplit_``, ``planeSplit``, ``searchLevel`` with ``epsError = 1 + eps``) and
is pinned against the reference's own vendored header compiled into ``oracle/_ref/junction_cluster`` (oracle/Makefile).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# ---- pixel stages ----
def dampen_intensity(bgr: np.ndarray, threshold_min: float, threshold_max: float) -> np.ndarray:
    incline = 1.0 / (threshold_max - threshold_min)
    intercept = -threshold_min * incline
    px = bgr.astype(np.float64)
    gain = np.clip((px[..., 2] - px[..., 0]) * incline + intercept, 0.0, 1.0)
    return np.trunc(px * gain[..., None]).astype(np.uint8)


def bgr2gray(bgr: np.ndarray) -> np.ndarray:
    p = bgr.astype(np.int64)
    return ((p[..., 0] * 3735 + p[..., 1] * 19235 + p[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


def blur3_u8(gray: np.ndarray) -> np.ndarray:
    p = np.pad(gray.astype(np.int64), 1, mode="reflect")
    hs = p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:]
    return ((hs[:-2] + 2 * hs[1:-1] + hs[2:] + 8) >> 4).astype(np.uint8)


def gaussian_kernel11() -> np.ndarray:
    """cv2.getGaussianKernel(11, 0, CV_32F): sigma = 0.3 * ((11 - 1) * 0.5 - 1) + 0.8 = 2.0, normalised in double."""
    x = np.arange(11, dtype=np.float64) - 5.0
    k = np.exp(-0.5 * x * x / 4.0)
    return (k / k.sum()).astype(f32)


def _fma(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(f32)


def _mul(a, b):
    return (a * f32(b)).astype(f32)


def gauss11_f32(f: np.ndarray) -> np.ndarray:
    K = gaussian_kernel11()
    h, w = f.shape
    pp = np.pad(f, ((0, 0), (5, 5)), mode="edge")
    X = [pp[:, i:i + w] for i in range(11)]
    a = _mul(X[0], K[0])
    b = a.copy()
    for i in range(1, 11):
        a = _fma(X[i], K[i], a)
        b = _fma(X[i], K[i], b) if i > 8 else (b + _mul(X[i], K[i])).astype(f32)
    t4 = w - (w % 4)
    a[:, t4:] = b[:, t4:]
    pp = np.pad(a, ((5, 5), (0, 0)), mode="edge")
    Y = [pp[i:i + h] for i in range(11)]
    a = _mul(Y[5], K[5])
    b = a.copy()
    for i in range(1, 6):
        s = (Y[5 + i] + Y[5 - i]).astype(f32)
        a = _fma(s, K[5 + i], a)
        b = (b + _mul(s, K[5 + i])).astype(f32)
    t8 = w - (w % 8)
    a[:, t8:] = b[:, t8:]
    return a


def adaptive_threshold(blur: np.ndarray) -> np.ndarray:
    mean = np.clip(np.rint(gauss11_f32(blur.astype(f32))), 0, 255).astype(np.int64)
    return np.where(blur.astype(np.int64) - mean > -2, 255, 0).astype(np.uint8)


# ---- contours as component interfaces ----
_DIRS = ((0, -1), (1, 0), (0, 1), (-1, 0))       # N, E, S, W; the walk leaves a crack towards the next direction


def contour_records(binary: np.ndarray):
    """binary: bool [h, w].  Returns the borders in cv2.findContours(RETR_TREE) order as dicts with ``key`` (discovery pixel
    y * w + x), ``hole``, ``area2`` (twice contourArea, an integer), ``bbox`` (x, y, w, h) and ``parent`` (key or -1)."""
    from scipy import ndimage as ndi
    h, w = binary.shape
    P = np.zeros((h + 2, w + 2), bool)
    P[1:-1, 1:-1] = binary
    H2, W2 = P.shape
    lf, nf = ndi.label(P, structure=np.ones((3, 3), int))
    lb, nb = ndi.label(~P, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    idx = np.arange(H2 * W2).reshape(H2, W2)
    first_f = np.atleast_1d(ndi.minimum(idx, lf, index=np.arange(1, nf + 1))).astype(np.int64) if nf else np.zeros(0, np.int64)
    first_b = np.atleast_1d(ndi.minimum(idx, lb, index=np.arange(1, nb + 1))).astype(np.int64)
    root = lb[0, 0]
    lbf, lff = lb.ravel(), lf.ravel()
    enc_bg = lbf[first_f - 1] if nf else np.zeros(0, np.int64)      # background component left of F's first pixel
    enc_fg = np.zeros(nb + 1, np.int64)
    for b in range(1, nb + 1):
        if b != root:
            enc_fg[b] = lff[first_b[b - 1] - 1]

    def key_of(bid):
        p = first_f[bid[1] - 1] if bid[0] == "o" else first_b[bid[1] - 1] - 1
        y, x = divmod(int(p), W2)
        return (y - 1) * w + (x - 1)

    a2, bb = {}, {}
    ys, xs = np.nonzero(P)
    for x, y in zip(xs.tolist(), ys.tolist()):
        F = lf[y, x]
        for d, (dx, dy) in enumerate(_DIRS):
            if P[y + dy, x + dx]:
                continue
            Hc = lb[y + dy, x + dx]
            bid = ("o", F) if Hc == enc_bg[F - 1] else ("h", Hc)
            tx, ty = _DIRS[(d + 1) % 4]
            if P[y + dy + ty, x + dx + tx]:
                nx, ny = x + dx + tx, y + dy + ty          # diagonal neighbour: 8-connectivity
            elif P[y + ty, x + tx]:
                nx, ny = x + tx, y + ty                    # straight on
            else:
                nx, ny = x, y                              # round the pixel's corner
            a2[bid] = a2.get(bid, 0) + (x * ny - nx * y)
            b = bb.get(bid)
            if b is None:
                bb[bid] = [x, y, x, y]
            else:
                b[0] = min(b[0], x); b[1] = min(b[1], y); b[2] = max(b[2], x); b[3] = max(b[3], y)
    rec = {}
    for bid in a2:
        if bid[0] == "o":
            Hc = enc_bg[bid[1] - 1]
            par = -1 if Hc == root else key_of(("h", Hc))
        else:
            par = key_of(("o", enc_fg[bid[1]]))
        b = bb[bid]
        k = key_of(bid)
        rec[k] = dict(key=k, hole=bid[0] == "h", area2=abs(int(a2[bid])), bbox=(b[0] - 1, b[1] - 1, b[2] - b[0] + 1, b[3] - b[1] + 1),
                      parent=par)
    children = {}
    for k, r in rec.items():
        children.setdefault(r["parent"], []).append(k)
    order = []
    stack = sorted(children.get(-1, []))                   # pop() takes the largest key first
    while stack:
        k = stack.pop()
        order.append(rec[k])
        stack.extend(sorted(children.get(k, [])))
    return order


def junction_candidates(records, grid_area: int, grid_area_threshold: float):
    """junction_detector.cpp:76-117 on (area, bounding box) pairs in contour order -> [n, 2] float32 box corners."""
    thr2 = f32(2) * f32(grid_area_threshold)
    lo = float(grid_area) * float(f32(1) / thr2)
    hi = float(grid_area) * float(thr2)
    out = []
    for r in records:
        area = r["area2"] * 0.5
        if not (lo < area < hi):
            continue
        x, y, bw, bh = r["bbox"]
        if area / float(bw * bh) >= 0.4 and 0.5 <= bw / bh <= 2.0:
            out += [(x - 1, y - 1), (x + bw + 1, y - 1), (x + bw + 1, y + bh + 1), (x - 1, y + bh + 1)]
    return np.asarray(out, f32).reshape(-1, 2)


# ---- nanoflann: KDTreeSingleIndexAdaptor<L2, float, 2>, leaf_max_size 7 ----
class _KDTree:
    def __init__(self, pts: np.ndarray, leaf_max: int = 7):
        self.p = pts.astype(f32)
        self.acc = list(range(len(pts)))
        self.leaf_max = leaf_max
        lo = [f32(self.p[:, d].min()) for d in range(2)]
        hi = [f32(self.p[:, d].max()) for d in range(2)]
        self.root_bbox = [[lo[0], hi[0]], [lo[1], hi[1]]]
        self.root = self._divide(0, len(pts), self.root_bbox)

    def _get(self, i, d):
        return self.p[self.acc[i], d]

    def _divide(self, left, right, bbox):
        if right - left <= self.leaf_max:
            for d in range(2):
                # (an empty leaf — possible with many duplicate points — takes the box of the point at `left`, as nanoflann's
                # loop does by initialising from vAcc_[left] before looking at the count)
                vals = [self._get(k, d) for k in range(left, max(right, left + 1))] if left < len(self.acc) else [self._get(left - 1, d)]
                bbox[d][0], bbox[d][1] = min(vals), max(vals)
            return ("leaf", left, right)
        idx, cutfeat, cutval = self._middle_split(left, right - left, bbox)
        lb = [list(b) for b in bbox]
        lb[cutfeat][1] = cutval
        c1 = self._divide(left, left + idx, lb)
        rb = [list(b) for b in bbox]
        rb[cutfeat][0] = cutval
        c2 = self._divide(left + idx, right, rb)
        for d in range(2):
            bbox[d][0] = min(lb[d][0], rb[d][0])
            bbox[d][1] = max(lb[d][1], rb[d][1])
        return ("node", cutfeat, lb[cutfeat][1], rb[cutfeat][0], c1, c2)

    def _middle_split(self, ind, count, bbox):
        eps = f32(0.00001)
        max_span = f32(bbox[0][1] - bbox[0][0])
        span1 = f32(bbox[1][1] - bbox[1][0])
        if span1 > max_span:
            max_span = span1
        max_spread, cutfeat, mn, mx = f32(-1), 0, f32(0), f32(0)
        for d in range(2):
            span = f32(bbox[d][1] - bbox[d][0])
            if span > f32(f32(1) - eps) * max_span:
                vals = [self._get(ind + k, d) for k in range(count)]
                lo, hi = min(vals), max(vals)
                spread = f32(hi - lo)
                if spread > max_spread:
                    cutfeat, max_spread, mn, mx = d, spread, lo, hi
        split = f32(f32(bbox[cutfeat][0] + bbox[cutfeat][1]) / f32(2))
        cutval = mn if split < mn else (mx if split > mx else split)
        lim1, lim2 = self._plane_split(ind, count, cutfeat, cutval)
        half = count // 2
        idx = lim1 if lim1 > half else (lim2 if lim2 < half else half)
        return idx, cutfeat, cutval

    def _plane_split(self, ind, count, cutfeat, cutval):
        a = self.acc
        left, right = 0, count - 1
        while True:
            while left <= right and self._get(ind + left, cutfeat) < cutval:
                left += 1
            while right and left <= right and self._get(ind + right, cutfeat) >= cutval:
                right -= 1
            if left > right or not right:
                break
            a[ind + left], a[ind + right] = a[ind + right], a[ind + left]
            left += 1
            right -= 1
        lim1 = left
        right = count - 1
        while True:
            while left <= right and self._get(ind + left, cutfeat) <= cutval:
                left += 1
            while right and left <= right and self._get(ind + right, cutfeat) > cutval:
                right -= 1
            if left > right or not right:
                break
            a[ind + left], a[ind + right] = a[ind + right], a[ind + left]
            left += 1
            right -= 1
        return lim1, left

    def radius_search(self, q, radius2, eps_error):
        found = []
        dists = [f32(0), f32(0)]
        dist = f32(0)
        for d in range(2):
            if q[d] < self.root_bbox[d][0]:
                dists[d] = f32((q[d] - self.root_bbox[d][0]) ** 2); dist = f32(dist + dists[d])
            if q[d] > self.root_bbox[d][1]:
                dists[d] = f32((q[d] - self.root_bbox[d][1]) ** 2); dist = f32(dist + dists[d])
        self._search(self.root, q, dist, dists, radius2, eps_error, found)
        return found

    def _search(self, node, q, mindist, dists, radius2, eps_error, found):
        if node[0] == "leaf":
            for i in range(node[1], node[2]):
                j = self.acc[i]
                dx, dy = f32(q[0] - self.p[j, 0]), f32(q[1] - self.p[j, 1])
                d = f32(f32(f32(0) + f32(dx * dx)) + f32(dy * dy))
                if d < radius2:
                    found.append(j)
            return
        _, feat, divlow, divhigh, c1, c2 = node
        val = q[feat]
        d1, d2 = f32(val - divlow), f32(val - divhigh)
        if f32(d1 + d2) < 0:
            best, other, cut = c1, c2, f32(f32(val - divhigh) * f32(val - divhigh))
        else:
            best, other, cut = c2, c1, f32(f32(val - divlow) * f32(val - divlow))
        self._search(best, q, mindist, dists, radius2, eps_error, found)
        dst = dists[feat]
        mindist = f32(f32(mindist + cut) - dst)
        dists[feat] = cut
        if f32(mindist * eps_error) <= radius2:
            self._search(other, q, mindist, dists, radius2, eps_error, found)
        dists[feat] = dst


def cluster_junctions(cand: np.ndarray, eps: int) -> np.ndarray:
    """junction_detector.cpp:123-185 -> [m, 2] float32 cluster centres."""
    cand = np.asarray(cand, f32).reshape(-1, 2)
    if len(cand) < 4:
        return np.zeros((0, 2), f32)
    tree = _KDTree(cand, 7)
    radius = f32(eps)
    visited = np.zeros(len(cand), bool)
    out = []
    for i in range(len(cand)):
        if visited[i]:
            continue
        nb = tree.radius_search(cand[i], f32(radius * radius), f32(1 + f32(10.0)))
        if len(nb) >= 3:
            sx, sy = f32(0), f32(0)
            for j in nb:
                sx = f32(sx + cand[j, 0]); sy = f32(sy + cand[j, 1])
            out.append((f32(sx / f32(len(nb))), f32(sy / f32(len(nb)))))
            visited[nb] = True
    return np.asarray(out, f32).reshape(-1, 2)


def threshold_image(img: np.ndarray) -> np.ndarray:
    gray = img if img.ndim == 2 else bgr2gray(img)
    return adaptive_threshold(blur3_u8(gray))


def find_junctions(img: np.ndarray, grid_area: int = 250, grid_area_threshold: float = 2.0, eps: int = 4) -> np.ndarray:
    """find_junctions_not_rotated(img, grid_area, grid_area_threshold, false, eps) -> [m, 2] float32."""
    rec = contour_records(threshold_image(img) > 0)
    return cluster_junctions(junction_candidates(rec, grid_area, grid_area_threshold), eps)
