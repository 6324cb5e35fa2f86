"""Pins the NumPy restatements of the sparse path (oracle/lk_np.py, oracle/features_np.py) against
the reference implementation — the cv2 4.13.0 wheel.  Integer stages and the corner list are
bit-exact; LK positions agree within 1e-3 px (different float accumulation order)."""
import numpy as np
import pytest

from oracle import features_np as FT
from oracle import lk_np as LK
from oracle import synth

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("shape", [(64, 96), (97, 131), (135, 240), (33, 35)])
def test_pyrdown_bit_exact(shape):
    a, _ = synth.synth_pair(shape[0], shape[1], 3)
    assert np.array_equal(LK.pyr_down(a), cv2.pyrDown(a))
    rng = np.random.default_rng(1)
    r = rng.integers(0, 256, size=shape, dtype=np.uint8)
    assert np.array_equal(LK.pyr_down(r), cv2.pyrDown(r))


@pytest.mark.parametrize("shape", [(64, 96), (97, 131)])
def test_scharr_bit_exact(shape):
    rng = np.random.default_rng(2)
    r = rng.integers(0, 256, size=shape, dtype=np.uint8)
    d = LK.scharr_deriv(r)
    assert np.array_equal(d[..., 0], cv2.Scharr(r, cv2.CV_16S, 1, 0))
    assert np.array_equal(d[..., 1], cv2.Scharr(r, cv2.CV_16S, 0, 1))


def test_pyramid_depth_clamp():
    a, _ = synth.synth_pair(60, 80, 3)
    n, pyr = cv2.buildOpticalFlowPyramid(a, (21, 21), 3, withDerivatives=False)
    mine = LK.build_pyramid(a, (21, 21), 3)
    assert n == len(mine) - 1 == 1
    for lv, m in enumerate(mine):
        ref = pyr[lv]
        assert np.array_equal(np.asarray(ref)[:m.shape[0], :m.shape[1]], m) or ref.shape == m.shape


@pytest.mark.parametrize("shape,seed", [((120, 160), 0), ((135, 241), 1)])
def test_min_eigenval_bit_exact(shape, seed):
    a, _ = synth.synth_pair(shape[0], shape[1], seed)
    ref = cv2.cornerMinEigenVal(a, 3, ksize=3)
    mine = FT.corner_min_eigenval(a, 3)
    assert np.array_equal(ref, mine), int((ref != mine).sum())


@pytest.mark.parametrize("shape,seed,maxc,q,md", [((240, 320), 0, 200, 0.01, 7), ((135, 241), 1, 50, 0.05, 3.5),
                                                    ((240, 320), 2, 0, 0.1, 10), ((120, 160), 3, 100, 0.01, 0)])
def test_good_features_exact(shape, seed, maxc, q, md):
    a, _ = synth.synth_pair(shape[0], shape[1], seed)
    ref = cv2.goodFeaturesToTrack(a, maxc, q, md, blockSize=3)
    mine = FT.good_features(a, maxc, q, md, 3)
    assert ref.shape == mine.shape and np.array_equal(ref, mine)


@pytest.mark.parametrize("shape,seed,k", [((120, 160), 0, 0.04), ((97, 131), 1, 0.04), ((77, 129), 2, 0.06), ((33, 36), 3, 0.04),
                                          ((31, 1921), 4, 0.06), ((90, 125), 5, 0.1)])
def test_corner_harris_bit_exact(shape, seed, k):
    """cv2.cornerHarris incl. the wheel's treatment of the image as one continuous row: the last (w*h) % 8 pixels of the
    image go through the 4-wide step ((k s) s instead of k (s s)) and the scalar double tail."""
    a, _ = synth.synth_pair(shape[0], shape[1], seed)
    ref = cv2.cornerHarris(a, 3, 3, k)
    mine = FT.corner_harris(a, 3, k)
    assert np.array_equal(ref, mine), int((ref != mine).sum())


@pytest.mark.parametrize("shape,seed,maxc,q,md,k", [((240, 320), 0, 200, 0.01, 7, 0.04), ((135, 241), 1, 50, 0.05, 3.5, 0.06)])
def test_good_features_harris_exact(shape, seed, maxc, q, md, k):
    a, _ = synth.synth_pair(shape[0], shape[1], seed)
    ref = cv2.goodFeaturesToTrack(a, maxc, q, md, blockSize=3, useHarrisDetector=True, k=k)
    mine = FT.good_features(a, maxc, q, md, 3, eig=FT.corner_harris(a, 3, k))
    assert ref.shape == mine.shape and np.array_equal(ref, mine)


def test_good_features_tie_order():
    t = np.zeros((64, 64), np.uint8)
    for y in range(8, 64, 16):
        for x in range(8, 64, 16):
            t[y:y + 4, x:x + 4] = 200
    ref = cv2.goodFeaturesToTrack(t, 50, 0.01, 0, blockSize=3)
    assert np.array_equal(ref, FT.good_features(t, 50, 0.01, 0, 3))
    ref = cv2.goodFeaturesToTrack(t, 50, 0.01, 5, blockSize=3)
    assert np.array_equal(ref, FT.good_features(t, 50, 0.01, 5, 3))


def test_pyrlk_matches_cv2():
    a, b = synth.synth_pair(240, 320, 4, (3.2, -1.7))
    pts = cv2.goodFeaturesToTrack(a, 60, 0.01, 7, blockSize=3)
    extra = np.array([[[1.5, 2.0]], [[318.2, 237.9]], [[0.0, 120.0]], [[160.0, 0.3]]], np.float32)
    pts = np.concatenate([pts, extra])
    rn, rs, re = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
    mn, ms, me = LK.calc_pyrlk(a, b, pts, None, (21, 21), 3, 30, 0.01)
    assert np.array_equal(rs, ms)
    ok = rs.ravel() == 1
    assert np.abs(rn - mn)[ok].max() < 1e-3
    assert np.abs(re - me)[ok].max() < 1e-2
    # other window / criteria / min-eig flag
    rn, rs, re = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(15, 11), maxLevel=2, criteria=(3, 10, 0.03),
                                          flags=8, minEigThreshold=1e-3)
    mn, ms, me = LK.calc_pyrlk(a, b, pts, None, (15, 11), 2, 10, 0.03, flags=8, min_eig_threshold=1e-3)
    assert np.array_equal(rs, ms)
    ok = rs.ravel() == 1
    assert np.abs(rn - mn)[ok].max() < 1e-3
    assert np.allclose(re, me, rtol=1e-3, atol=1e-6)


def test_sobel_simd_tail_rule_is_the_same_on_every_fma_dispatch_level():
    """The eigenvalue map depends on WHERE the wheel's Sobel row filter stops using FMA: columns past the last full block
    of 32 take the scalar tail.  That block size is a property of the wheel's dispatched code, so it is probed here on
    the wheel's AVX-512 path (if this host has it) and with AVX512-SKX disabled (its AVX2 path, what a host without
    AVX-512 runs): the block-of-32 restatement must match cv2 on both, at widths with every tail length class."""
    import os
    import subprocess
    import sys
    child = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import cv2\n"
        "from oracle import features_np as F, synth\n"
        "bad = 0\n"
        "for (h, w, seed) in [(60, 67, 1), (48, 100, 2), (40, 131, 3), (64, 250, 4), (33, 255, 5), (50, 96, 6)]:\n"
        "    a, _ = synth.synth_pair(h, w, seed)\n"
        "    bad += int((F.corner_min_eigenval(a, 3) != cv2.cornerMinEigenVal(a, 3, ksize=3)).sum())\n"
        "print('MISMATCH', bad, cv2.getCPUFeaturesLine())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for disable in ("", "AVX512-SKX"):
        env = dict(os.environ)
        if disable:
            env["OPENCV_CPU_DISABLE"] = disable
        out = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=300)
        line = [l for l in out.stdout.splitlines() if l.startswith("MISMATCH")]
        assert line, out.stderr[-2000:]
        assert line[0].split()[1] == "0", (disable, line[0])
        if disable and "AVX512-SKX" in line[0]:
            assert "AVX512-SKX?" in line[0]              # cv2 marks a disabled feature with '?': the level really was lowered
