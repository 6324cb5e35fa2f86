"""ORACLE (test infrastructure only — never imported by the product path).

NumPy restatement of the flow post-processing the reference's "adapt" node applies between the flow call and
the velocity scalar (``ros2_ws/src/liteflownet3/liteflownet3/lfn3_adapt_node.py:232-254``):

  :236-238  ``flow[0] = cv2.medianBlur(flow[0], k)``, same for ``flow[1]``  (k = 3 or 5 for float32 fields)
  :241-244  ``mag = sqrt(u**2 + v**2)``; ``u *= (mag >= threshold)``, ``v *= ...``           (float32 arithmetic)
  :247-251  ``gray = cvtColor(rgb, RGB2GRAY)``; ``u *= (gray < intensity_threshold)``, ``v *= ...``
  :254      ``u_avg = np.mean(u) / dt``   (mean over ALL pixels, masked ones contribute zeros)

``median_blur_np`` restates ``cv2.medianBlur`` for float32 (exact selection, BORDER_REPLICATE) and is pinned
bit-for-bit against the cv2 wheel in ``tests/test_oracle_postfilter.py``; the rest is the node's own NumPy.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def median_blur_np(a: np.ndarray, k: int) -> np.ndarray:
    """cv2.medianBlur(float32 [H,W], k) for k in (3, 5): median of the k*k replicate-padded neighbourhood."""
    if k not in (3, 5):
        raise ValueError("float32 medianBlur supports ksize 3 and 5")
    a = np.asarray(a, np.float32)
    r = k // 2
    h, w = a.shape
    p = np.pad(a, r, mode="edge")
    win = np.stack([p[i:i + h, j:j + w] for i in range(k) for j in range(k)], 0)
    return np.sort(win, axis=0)[k * k // 2]


def adapt_postfilter_np(flow_hw2: np.ndarray, median_ksize: int = 0, magnitude_threshold: Optional[float] = None,
                        gray: Optional[np.ndarray] = None, intensity_threshold: Optional[int] = None,
                        use_cv2_median: bool = False) -> np.ndarray:
    """lfn3_adapt_node.py:232-251 on a cv2-layout field [H,W,2] float32; returns the filtered field."""
    u = np.array(flow_hw2[..., 0], np.float32)
    v = np.array(flow_hw2[..., 1], np.float32)
    if median_ksize:
        if use_cv2_median:
            import cv2
            u, v = cv2.medianBlur(u, median_ksize), cv2.medianBlur(v, median_ksize)
        else:
            u, v = median_blur_np(u, median_ksize), median_blur_np(v, median_ksize)
    if magnitude_threshold is not None:
        mag = np.sqrt(u ** 2 + v ** 2)
        m = (mag >= np.float32(magnitude_threshold)).astype(np.float32)
        u = u * m
        v = v * m
    if intensity_threshold is not None:
        if gray is None:
            raise ValueError("intensity mask needs the gray frame")
        m = (np.asarray(gray) < intensity_threshold).astype(np.float32)
        u = u * m
        v = v * m
    return np.stack([u, v], -1)
