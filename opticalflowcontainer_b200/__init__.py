"""opticalflowcontainer_b200 — B200-native drop-in for the per-frame-pair OpenCV flow call.

Same Python signatures as cv2 (``calcOpticalFlowFarneback``, ``calcOpticalFlowPyrLK``,
``goodFeaturesToTrack``) so a node shaped like
``ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py`` swaps ``import cv2`` for
``import opticalflowcontainer_b200 as ofb`` at the flow call (line 194) and nothing else.
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of ``libofb.so``
(``include/ofb.h``).  There is no CPU fallback: without the built library or a CUDA device
every call raises.
"""
from __future__ import annotations

import threading

import numpy as np

from ._lib import OfbError, LIB_PATH, EXPORTED_SYMBOLS  # noqa: F401
from .engine import (FlowEngine, OPTFLOW_FARNEBACK_GAUSSIAN, OPTFLOW_LK_GET_MIN_EIGENVALS,  # noqa: F401
                     OPTFLOW_USE_INITIAL_FLOW)

__all__ = ["FlowEngine", "OfbError", "calcOpticalFlowFarneback", "calcOpticalFlowPyrLK", "goodFeaturesToTrack",
           "buildOpticalFlowPyramid", "cornerMinEigenVal", "OPTFLOW_USE_INITIAL_FLOW", "OPTFLOW_FARNEBACK_GAUSSIAN",
           "OPTFLOW_LK_GET_MIN_EIGENVALS", "set_device", "imdecode", "find_junctions", "IMREAD_GRAYSCALE", "IMREAD_COLOR"]

IMREAD_GRAYSCALE = 0   # cv2.IMREAD_GRAYSCALE
IMREAD_COLOR = 1       # cv2.IMREAD_COLOR

_engines = {}
_engines_lock = threading.Lock()
_device = 0


def set_device(device: int):
    """Device used by the module-level cv2-style functions (one engine per (device, size))."""
    global _device
    _device = int(device)


def _engine_for(h: int, w: int) -> FlowEngine:
    key = (_device, h, w)
    with _engines_lock:
        e = _engines.get(key)
        if e is None:
            e = FlowEngine(w, h, max_batch=1, device=_device)
            _engines[key] = e
        return e


def calcOpticalFlowFarneback(prev, next, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags):
    """Drop-in for ``cv2.calcOpticalFlowFarneback`` → float32 [H,W,2]."""
    prev = np.asarray(prev)
    return _engine_for(prev.shape[0], prev.shape[1]).farneback(prev, next, flow, pyr_scale, levels, winsize,
                                                               iterations, poly_n, poly_sigma, flags)


def imdecode(buf, flags=IMREAD_COLOR):
    """Drop-in for ``cv2.imdecode(buf, flags)`` on baseline JPEG streams (the compressed-image node,
    opticalflow_comprerssed_node.py:43-46): ``IMREAD_COLOR`` -> uint8 [H,W,3] BGR, ``IMREAD_GRAYSCALE`` -> uint8 [H,W].
    Returns None — as cv2 does — for a stream the device path does not decode."""
    if flags not in (IMREAD_COLOR, IMREAD_GRAYSCALE):
        raise OfbError(1, "imdecode: flags must be IMREAD_COLOR or IMREAD_GRAYSCALE")
    eng = _engine_for(64, 64)
    try:
        return eng.imdecode(buf) if flags == IMREAD_COLOR else eng.imdecode_grayscale(buf)
    except OfbError as e:
        if e.status == 6:
            return None
        raise


def find_junctions(img, grid_area=250, grid_area_threshold=2.0, eps=4, dampen=None):
    """``find_junctions_not_rotated(img, grid_area, grid_area_threshold, false, eps)`` of the reference's junction detector
    (junction_detector.cpp:31-214; ``dampen=(min, max)``: ``dampenIntensity`` first) -> float32 [n, 2] junction centres."""
    return _engine_for(64, 64).find_junctions(img, grid_area, grid_area_threshold, eps, dampen=dampen)


def goodFeaturesToTrack(image, maxCorners, qualityLevel, minDistance, mask=None, blockSize=3,
                        useHarrisDetector=False, k=0.04):
    """Drop-in for ``cv2.goodFeaturesToTrack``: Shi-Tomasi or (``useHarrisDetector``) Harris response with parameter ``k``,
    optional ``mask``; the corner list is cv2's bit for bit (the response maps reproduce the wheel's arithmetic)."""
    image = np.asarray(image)
    r = _engine_for(image.shape[0], image.shape[1]).good_features(image, maxCorners, qualityLevel, minDistance,
                                                                  blockSize, mask, useHarrisDetector, k)
    return r if len(r) else None


def cornerMinEigenVal(src, blockSize, ksize=3):
    if ksize != 3:
        raise OfbError(1, "cornerMinEigenVal: only ksize=3 is supported")
    src = np.asarray(src)
    return _engine_for(src.shape[0], src.shape[1]).corner_min_eigenval(src, blockSize)


def buildOpticalFlowPyramid(img, winSize, maxLevel, withDerivatives=True):
    """Returns (maxLevel_built, [level0, deriv0, level1, deriv1, ...]) like cv2 (without the
    winSize border cv2 pads its pyramid Mats with)."""
    img = np.asarray(img)
    lv, dv = _engine_for(img.shape[0], img.shape[1]).lk_pyramid(img, winSize, maxLevel, withDerivatives)
    out = []
    for i, l in enumerate(lv):
        out.append(l)
        if withDerivatives:
            out.append(dv[i])
    return len(lv) - 1, out


def _pyramid_level0(img):
    """(image, depth): an image as it is (depth None), or level 0 of a buildOpticalFlowPyramid list and the index of its
    last level."""
    if isinstance(img, (list, tuple)):
        if not img:
            raise OfbError(1, "calcOpticalFlowPyrLK: empty pyramid")
        # with derivatives the list alternates [level, deriv, level, deriv, ...]; derivative images are int16 2-channel
        with_deriv = len(img) > 1 and np.asarray(img[1]).dtype == np.int16
        n_levels = len(img) // 2 if with_deriv else len(img)
        return np.asarray(img[0]), n_levels - 1
    return np.asarray(img), None


def calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts, status=None, err=None, winSize=(21, 21), maxLevel=3,
                         criteria=(3, 30, 0.01), flags=0, minEigThreshold=1e-4):
    """Drop-in for ``cv2.calcOpticalFlowPyrLK`` → (nextPts [N,1,2], status [N,1], err [N,1]).
    ``prevImg`` / ``nextImg`` may also be pyramids as ``buildOpticalFlowPyramid`` returns them (with or without the
    derivative images): level 0 is taken and ``maxLevel`` is clamped to the pyramid's depth, as cv2 does — the engine
    rebuilds the (bit-identical) uint8 pyramid on the GPU."""
    prevImg, d0 = _pyramid_level0(prevImg)
    nextImg, d1 = _pyramid_level0(nextImg)
    for d in (d0, d1):
        if d is not None:
            maxLevel = min(maxLevel, d)
    return _engine_for(prevImg.shape[0], prevImg.shape[1]).pyrlk(prevImg, nextImg, prevPts, nextPts, winSize, maxLevel,
                                                                criteria, flags, minEigThreshold)
