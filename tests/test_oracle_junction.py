"""The junction-detector restatement (oracle/junction_np.py) against the reference side: the cv2 wheel for the pixel
stages and the contours (live and through tests/golden/junction.npz), the reference's own nanoflann header compiled into
oracle/_ref/junction_cluster for the clustering (live when the binary exists, and through the fixture).  The host half
of the product (``ofb_cluster_junctions``: KD-tree build + approximate radius search, no device) is checked the same way.
Anchor: ros2_ws/src/junction_point_detector/src/junction_detector.cpp:3-214."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import junction_np as J  # noqa: E402
from oracle import synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "junction.npz"))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "junction_cluster")


def ref_cluster(cand, eps):
    txt = "".join("%d %d\n" % (x, y) for x, y in cand)
    out = subprocess.run([REF_BIN, str(eps)], input=txt, capture_output=True, text=True, check=True).stdout
    return np.array([[float(a) for a in ln.split()] for ln in out.splitlines()], np.float32).reshape(-1, 2)


def product_cluster(cand, eps):
    from opticalflowcontainer_b200 import _lib
    lib = _lib.load()
    cand = np.ascontiguousarray(cand, np.float32)
    out = np.empty((max(len(cand), 1), 2), np.float32)
    n = C.c_int()
    assert lib.ofb_cluster_junctions(cand.ctypes.data, len(cand), eps, out.ctypes.data, len(out), C.byref(n)) == 0
    return out[:n.value].copy()


def test_gaussian_kernel_is_the_wheels():
    assert np.array_equal(J.gaussian_kernel11(), cv2.getGaussianKernel(11, 0, cv2.CV_32F).ravel())


@pytest.mark.parametrize("size", [(37, 53), (120, 163), (64, 1927), (100, 70), (480, 640)])
def test_pixel_stages_equal_cv2(size):
    """gray, 3x3 Gaussian, the float 11x11 Gaussian (bit for bit, including the wheel's scalar-tail operation order in the
    last width % 4 / width % 8 columns) and the adaptive threshold."""
    h, w = size
    img = synth.synth_net(h, w, w)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(J.bgr2gray(img), gray)
    blur = cv2.GaussianBlur(gray, (3, 3), 0)
    assert np.array_equal(J.blur3_u8(gray), blur)
    f = np.random.default_rng(h).integers(0, 256, (h, w)).astype(np.float32)
    ref = cv2.GaussianBlur(f, (11, 11), 0, 0, borderType=cv2.BORDER_REPLICATE | cv2.BORDER_ISOLATED)
    assert np.array_equal(J.gauss11_f32(f), ref)
    th = cv2.adaptiveThreshold(blur, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
    assert np.array_equal(J.adaptive_threshold(blur), th)


def _cv2_records(th):
    cs, hier = cv2.findContours(th, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    return [(int(round(cv2.contourArea(c) * 2)),) + tuple(cv2.boundingRect(c)) for c in cs], hier


@pytest.mark.parametrize("seed", range(12))
def test_contours_equal_findcontours_on_noise(seed):
    """Random binary images (the hardest case: holes in holes, one-pixel bridges, diagonal contacts): same number of
    contours, same areas and boxes, same order and same parents as cv2.findContours(RETR_TREE)."""
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(5, 48)), int(rng.integers(5, 60))
    if seed % 3 == 0:
        img = cv2.GaussianBlur((rng.random((h, w)) * 255).astype(np.uint8), (0, 0), 1.5) > 127
    else:
        img = rng.random((h, w)) < rng.choice([0.3, 0.5, 0.6, 0.8])
    th = img.astype(np.uint8) * 255
    ref, hier = _cv2_records(th)
    rec = J.contour_records(img)
    assert [(r["area2"],) + tuple(r["bbox"]) for r in rec] == ref
    keys = [r["key"] for r in rec]
    for i, r in enumerate(rec):
        par = hier[0, i, 3]
        assert r["parent"] == (-1 if par < 0 else keys[par])
        assert r["hole"] == (bool(par >= 0) and not rec[par]["hole"])


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_detector_restatement_equals_golden(name):
    img, (ga, eps) = GOLD[name + "_img"], GOLD[name + "_params"]
    th = J.threshold_image(img)
    assert np.array_equal(th, GOLD[name + "_thresh"])
    rec = J.contour_records(th > 0)
    assert np.array_equal(np.array([(r["area2"],) + tuple(r["bbox"]) for r in rec]), GOLD[name + "_contours"])
    cand = J.junction_candidates(rec, int(ga), 2.0)
    assert np.array_equal(cand, GOLD[name + "_cand"])
    assert np.array_equal(J.cluster_junctions(cand, int(eps)), GOLD[name + "_junctions"])
    assert np.array_equal(J.find_junctions(img, int(ga), 2.0, int(eps)), GOLD[name + "_junctions"])


@pytest.mark.parametrize("i", range(6))
def test_clustering_equals_nanoflann_golden(built_lib, i):
    cand, eps, want = GOLD["rand%d_cand" % i], int(GOLD["rand%d_eps" % i]), GOLD["rand%d_junctions" % i]
    assert np.array_equal(J.cluster_junctions(cand, eps), want)
    assert np.array_equal(product_cluster(cand, eps), want)


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/junction_cluster not built (needs /root/reference)")
def test_clustering_equals_nanoflann_live(built_lib):
    """Restatement and product host code against the reference's compiled nanoflann on fresh random point sets,
    including duplicates and collinear points (the KD-tree's degenerate splits)."""
    rng = np.random.default_rng(99)
    for t in range(60):
        n = int(rng.integers(4, 500))
        span = int(rng.integers(3, 300))
        pts = rng.integers(0, span, (n, 2)).astype(np.float32)
        if t % 5 == 0:
            pts[:, 1] = 7                              # all on one line
        eps = int(rng.integers(2, 9))
        want = ref_cluster(pts, eps)
        assert np.array_equal(J.cluster_junctions(pts, eps), want), t
        assert np.array_equal(product_cluster(pts, eps), want), t


def test_dampen_intensity_formula():
    """dampenIntensity (junction_detector.cpp:3-28): double gain, channels truncated."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)
    got = J.dampen_intensity(img, -20, 15)
    incline = 1.0 / 35.0
    for y, x in [(0, 0), (5, 7), (19, 29), (10, 3)]:
        b, g, r = (int(v) for v in img[y, x])
        gain = max(min((r - b) * incline + 20 * incline, 1.0), 0.0)
        assert tuple(got[y, x]) == (int(b * gain), int(g * gain), int(r * gain))
