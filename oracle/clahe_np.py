"""ORACLE (test infrastructure only — never imported by the product path).

NumPy restatement of ``cv2.createCLAHE(clipLimit, tileGridSize).apply(img)`` for uint8 images — the contrast
pre-filter of the reference's "adapt" node (``ros2_ws/src/liteflownet3/liteflownet3/lfn3_adapt_node.py:164-182``:
``self.clahe.setClipLimit(clip); v_enhanced = self.clahe.apply(v)``).  Follows OpenCV 4.x
``modules/imgproc/src/clahe.cpp`` (``CLAHE_CalcLut_Body``, ``CLAHE_Interpolation_Body``): per-tile histogram, integer
clip limit ``max((int)(clipLimit * tileArea / 256), 1)``, excess redistributed as ``clipped / 256`` to every bin plus one
to every ``max(256 / residual, 1)``-th bin, LUT = ``cvRound(cumsum * (255.f / tileArea))``, then a bilinear blend of the
four surrounding tiles' LUTs in float (separate roundings, no FMA) and ``cvRound``.  Images whose size is not a
multiple of the grid are extended to the right/bottom with BORDER_REFLECT_101 for the histograms.
Pinned bit for bit against the wheel in ``tests/test_oracle_clahe.py``.
"""
from __future__ import annotations

import numpy as np


def clahe_luts(img: np.ndarray, clip_limit: float, tiles_x: int, tiles_y: int):
    h, w = img.shape
    if w % tiles_x == 0 and h % tiles_y == 0:
        ext = img
    else:
        ext = np.pad(img, ((0, tiles_y - h % tiles_y), (0, tiles_x - w % tiles_x)), mode="reflect")   # REFLECT_101
    tw, th = ext.shape[1] // tiles_x, ext.shape[0] // tiles_y
    area = tw * th
    lut_scale = np.float32(255.0) / np.float32(area)
    clip = 0
    if clip_limit > 0.0:
        clip = max(int(clip_limit * area / 256), 1)
    luts = np.zeros((tiles_y, tiles_x, 256), np.uint8)
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            hist = np.bincount(ext[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(hist - clip, 0).sum())
                hist = np.minimum(hist, clip)
                batch = clipped // 256
                residual = clipped - batch * 256
                hist += batch
                if residual != 0:
                    step = max(256 // residual, 1)
                    i = 0
                    while i < 256 and residual > 0:
                        hist[i] += 1
                        i += step
                        residual -= 1
            cs = np.cumsum(hist).astype(np.float32) * lut_scale
            luts[ty, tx] = np.clip(np.rint(cs), 0, 255).astype(np.uint8)
    return luts, tw, th


def clahe_apply(img: np.ndarray, clip_limit: float = 40.0, tile_grid=(8, 8)) -> np.ndarray:
    img = np.asarray(img, np.uint8)
    tiles_x, tiles_y = int(tile_grid[0]), int(tile_grid[1])
    h, w = img.shape
    luts, tw, th = clahe_luts(img, clip_limit, tiles_x, tiles_y)
    f32 = np.float32
    inv_tw, inv_th = f32(1.0) / f32(tw), f32(1.0) / f32(th)
    txf = np.arange(w, dtype=np.float32) * inv_tw - f32(0.5)
    tx1 = np.floor(txf).astype(np.int64)
    xa = (txf - tx1.astype(np.float32)).astype(np.float32)
    xa1 = (f32(1.0) - xa).astype(np.float32)
    tx2 = np.minimum(tx1 + 1, tiles_x - 1)
    tx1 = np.maximum(tx1, 0)
    tyf = np.arange(h, dtype=np.float32) * inv_th - f32(0.5)
    ty1 = np.floor(tyf).astype(np.int64)
    ya = (tyf - ty1.astype(np.float32)).astype(np.float32)
    ya1 = (f32(1.0) - ya).astype(np.float32)
    ty2 = np.minimum(ty1 + 1, tiles_y - 1)
    ty1 = np.maximum(ty1, 0)
    v = img.astype(np.int64)
    Y1, Y2 = ty1[:, None], ty2[:, None]
    X1, X2 = tx1[None, :], tx2[None, :]
    l11 = luts[Y1, X1, v].astype(np.float32); l12 = luts[Y1, X2, v].astype(np.float32)
    l21 = luts[Y2, X1, v].astype(np.float32); l22 = luts[Y2, X2, v].astype(np.float32)
    XA, XA1 = xa[None, :], xa1[None, :]
    top = (l11 * XA1).astype(np.float32) + (l12 * XA).astype(np.float32)
    bot = (l21 * XA1).astype(np.float32) + (l22 * XA).astype(np.float32)
    res = (top.astype(np.float32) * ya1[:, None]).astype(np.float32) + (bot.astype(np.float32) * ya[:, None]).astype(np.float32)
    return np.clip(np.rint(res.astype(np.float32)), 0, 255).astype(np.uint8)
