"""The reference implementation of the hot path: the installed OpenCV wheel.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see ``oracle/__init__.py``).

The reference repo declares OpenCV un-pinned (``ros2_ws/src/nueflow/setup.py:29``
``'opencv-python'``; ``ros2_ws/src/nueflow/package.xml:19`` ``python3-opencv``);
this image carries ``opencv-python-headless 4.13.0.92``.  The flow call a
Farneback node makes in place of ``self.net(t1, t2)``
(``ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:194``) is::

    cv2.calcOpticalFlowFarneback(prev_gray, gray, None, 0.5, 3, 15, 3, 5, 1.2, 0)
"""
from __future__ import annotations

import numpy as np

NODE_DEFAULTS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5,
                     poly_sigma=1.2, flags=0)

LK_DEFAULTS = dict(winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01), flags=0,
                   minEigThreshold=1e-4)

GFTT_DEFAULTS = dict(maxCorners=2000, qualityLevel=0.01, minDistance=7, blockSize=3)


def version() -> str:
    import cv2
    return cv2.__version__


def farneback(prev: np.ndarray, nxt: np.ndarray, flow=None, **kw) -> np.ndarray:
    import cv2
    p = dict(NODE_DEFAULTS)
    p.update(kw)
    return cv2.calcOpticalFlowFarneback(prev, nxt, flow, p["pyr_scale"], p["levels"], p["winsize"],
                                        p["iterations"], p["poly_n"], p["poly_sigma"], p["flags"])


def good_features(img: np.ndarray, **kw) -> np.ndarray:
    import cv2
    p = dict(GFTT_DEFAULTS)
    p.update(kw)
    r = cv2.goodFeaturesToTrack(img, p["maxCorners"], p["qualityLevel"], p["minDistance"],
                                mask=p.get("mask"), blockSize=p["blockSize"],
                                useHarrisDetector=p.get("useHarrisDetector", False), k=p.get("k", 0.04))
    if r is None:
        return np.zeros((0, 1, 2), np.float32)
    return r


def pyrlk(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray, **kw):
    import cv2
    p = dict(LK_DEFAULTS)
    p.update(kw)
    return cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=tuple(p["winSize"]),
                                    maxLevel=p["maxLevel"], criteria=tuple(p["criteria"]),
                                    flags=p["flags"], minEigThreshold=p["minEigThreshold"])


def epe(a: np.ndarray, b: np.ndarray):
    """(mean, max) endpoint error between two [H,W,2] fields."""
    d = np.sqrt(((a.astype(np.float64) - b.astype(np.float64)) ** 2).sum(-1))
    return float(d.mean()), float(d.max())
