"""The junction detector on the device (ofb_find_junctions / ofb_junction_threshold) against the reference side: the cv2
wheel (adaptiveThreshold, findContours, contourArea, boundingRect) for everything up to the candidates, the reference's
nanoflann (golden fixture, live binary when present) for the clusters.  Anchor: junction_detector.cpp:3-214."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import junction_np as J  # noqa: E402
from oracle import synth  # noqa: E402
from test_oracle_junction import GOLD, REF_BIN, ref_cluster  # noqa: E402

pytestmark = pytest.mark.gpu


def cv2_candidates(th, grid_area, thr=2.0):
    cs, _ = cv2.findContours(th, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    thr2 = np.float32(2) * np.float32(thr)
    lo, hi = grid_area * float(np.float32(1) / thr2), grid_area * float(thr2)
    out = []
    for c in cs:
        area = cv2.contourArea(c)
        x, y, w, h = cv2.boundingRect(c)
        if lo < area < hi and area / float(w * h) >= 0.4 and 0.5 <= w / h <= 2.0:
            out += [(x - 1, y - 1), (x + w + 1, y - 1), (x + w + 1, y + h + 1), (x - 1, y + h + 1)]
    return np.asarray(out, np.float32).reshape(-1, 2)


def cv2_threshold(img):
    gray = img if img.ndim == 2 else cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    return cv2.adaptiveThreshold(cv2.GaussianBlur(gray, (3, 3), 0), 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_detector_equals_golden(engine_factory, name):
    eng = engine_factory(64, 64)
    img, (ga, eps) = GOLD[name + "_img"], GOLD[name + "_params"]
    assert np.array_equal(eng.junction_threshold(img), GOLD[name + "_thresh"])
    got, cand = eng.find_junctions(img, int(ga), 2.0, int(eps), return_candidates=True)
    assert np.array_equal(cand, GOLD[name + "_cand"])
    assert np.array_equal(got, GOLD[name + "_junctions"])


@pytest.mark.parametrize("size,gray", [((480, 640), False), ((481, 637), True), ((1080, 1920), False), ((1080, 1923), True)])
def test_threshold_and_candidates_equal_cv2(engine_factory, size, gray):
    """Full-size frames: binary image bit-exact, candidate list (contour order, areas, boxes) identical to cv2's; the
    clusters equal the restatement's (itself pinned against nanoflann) and the live nanoflann binary when it is here."""
    h, w = size
    img = synth.synth_net(h, w, h + w, bgr=not gray)
    eng = engine_factory(64, 64)
    th = cv2_threshold(img)
    assert np.array_equal(eng.junction_threshold(img), th)
    got, cand = eng.find_junctions(img, 200, 2.0, 6, return_candidates=True)
    want = cv2_candidates(th, 200)
    assert len(want) > 400
    assert np.array_equal(cand, want)
    assert np.array_equal(got, J.cluster_junctions(want, 6))
    if os.path.exists(REF_BIN):
        assert np.array_equal(got, ref_cluster(want, 6))


@pytest.mark.parametrize("seed", range(6))
def test_contours_on_noise_equal_cv2(engine_factory, seed):
    """Binary noise through the whole device path: the image is built so that the detector's own threshold reproduces a
    given random pattern is not possible, so the noise goes in as a gray frame and cv2 runs the same stages; a wide area
    window lets almost every contour through, which exercises the tree order (holes in holes, diagonal contacts)."""
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(40, 200)), int(rng.integers(40, 300))
    img = cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), float(rng.choice([0.7, 1.2, 2.0])))
    eng = engine_factory(64, 64)
    th = cv2_threshold(img)
    assert np.array_equal(eng.junction_threshold(img), th)
    got, cand = eng.find_junctions(img, 40, 8.0, 5, return_candidates=True)      # area window (2.5, 640)
    want = cv2_candidates(th, 40, 8.0)
    assert len(want) > 40
    assert np.array_equal(cand, want)
    assert np.array_equal(got, J.cluster_junctions(want, 5))


def test_dampen_and_node_parameters(engine_factory):
    """The ROS wrapper's call: dampenIntensity(img, -20, 15) then find_junctions_not_rotated(img, 200, 2.0, false, 6)
    (fishnet_detector_ros.cpp:49-58)."""
    img = synth.synth_net(240, 320, 5)
    img[..., 2] = np.clip(img[..., 2].astype(int) + 10, 0, 255)                 # some red-blue difference for the gain
    damp = J.dampen_intensity(img, -20, 15)
    eng = engine_factory(64, 64)
    th = cv2_threshold(damp)
    assert np.array_equal(eng.junction_threshold(img, dampen=(-20, 15)), th)
    got = eng.find_junctions(img, 200, 2.0, 6, dampen=(-20, 15))
    assert np.array_equal(got, J.cluster_junctions(cv2_candidates(th, 200), 6))


def test_junctions_feed_the_tracker(engine_factory):
    """Junctions of frame 1 as prevPts of calcOpticalFlowPyrLK into a shifted frame 2: same result as cv2 from the same points."""
    a = synth.synth_net(240, 320, 8, bgr=False)
    b = np.roll(a, (2, 3), axis=(0, 1))
    eng = engine_factory(320, 240)
    pts = eng.find_junctions(a, 200, 2.0, 6)
    assert len(pts) > 20
    p0 = pts.reshape(-1, 1, 2)
    nxt, st, err = eng.pyrlk(a, b, p0, None, (21, 21), 3, (3, 30, 0.01))
    rn, rs, re = cv2.calcOpticalFlowPyrLK(a, b, p0, None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
    assert np.array_equal(st.reshape(-1), rs.reshape(-1))
    ok = rs.reshape(-1) == 1
    assert np.abs(nxt.reshape(-1, 2)[ok] - rn.reshape(-1, 2)[ok]).max() <= 1e-2


def test_flat_frame_and_bad_arguments(engine_factory):
    from opticalflowcontainer_b200 import OfbError
    eng = engine_factory(64, 64)
    flat = np.full((60, 80), 128, np.uint8)
    assert eng.find_junctions(flat).shape == (0, 2)           # one contour (the frame), nothing passes
    with pytest.raises(OfbError):
        eng.find_junctions(flat, grid_area=0)
    with pytest.raises(OfbError):
        eng.find_junctions(np.zeros((10, 10, 4), np.uint8))
