"""GPU parity of the sparse path through the C ABI: uint8 pyrDown pyramid, Scharr derivatives and
the Shi-Tomasi corner list are BIT-EXACT against the cv2 wheel (BASELINE.json north_star);
pyramidal-LK status flags are equal and positions within the flow tolerance (mean EPE <= 0.01 px,
max <= 0.1 px) — in practice ~1e-4 px."""
import numpy as np
import pytest

from oracle import features_np as FT
from oracle import lk_np as LK
from oracle import synth

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("shape", [(480, 640), (135, 241), (97, 131), (1080, 1920)])
def test_pyramid_and_scharr_bit_exact(engine_factory, shape):
    a, _ = synth.synth_pair(shape[0], shape[1], 3)
    eng = engine_factory(shape[1], shape[0])
    levels, derivs = eng.lk_pyramid(a, (21, 21), 3, True)
    ref = a
    n_ref, _ = cv2.buildOpticalFlowPyramid(a, (21, 21), 3, withDerivatives=False)
    assert len(levels) == n_ref + 1
    for l, (lv, dv) in enumerate(zip(levels, derivs)):
        if l > 0:
            ref = cv2.pyrDown(ref)
        assert np.array_equal(lv, ref), "pyramid level %d differs" % l
        assert np.array_equal(dv[..., 0], cv2.Scharr(ref, cv2.CV_16S, 1, 0))
        assert np.array_equal(dv[..., 1], cv2.Scharr(ref, cv2.CV_16S, 0, 1))


def test_pyramid_random_bytes_and_depth_clamp(engine_factory):
    rng = np.random.default_rng(5)
    r = rng.integers(0, 256, size=(60, 80), dtype=np.uint8)
    eng = engine_factory(80, 60)
    levels, derivs = eng.lk_pyramid(r, (21, 21), 3, True)
    assert len(levels) == 2                      # 20x15 <= winSize is refused
    assert np.array_equal(levels[1], cv2.pyrDown(r))
    assert np.array_equal(levels[1], LK.pyr_down(r))
    assert np.array_equal(derivs[0], LK.scharr_deriv(r))


@pytest.mark.parametrize("shape,seed", [((480, 640), 0), ((270, 480), 1), ((135, 241), 2), ((100, 247), 3)])
def test_min_eigenval_bit_exact(engine_factory, shape, seed):
    a, _ = synth.synth_pair(shape[0], shape[1], seed)
    eng = engine_factory(shape[1], shape[0])
    got = eng.corner_min_eigenval(a, 3)
    assert np.array_equal(got, FT.corner_min_eigenval(a, 3))          # oracle restatement
    # live cv2 on this host, also where the width is not a multiple of 32 (the wheel's row filter takes its scalar,
    # FMA-free tail there — the same on its AVX2 and AVX-512 dispatch, tests/test_oracle_sparse.py)
    assert np.array_equal(got, cv2.cornerMinEigenVal(a, 3, ksize=3))


@pytest.mark.parametrize("shape,seed,maxc,q,md", [((480, 640), 0, 2000, 0.01, 7), ((480, 640), 1, 100, 0.05, 3.5),
                                                    ((270, 480), 2, 0, 0.1, 10), ((240, 320), 3, 100, 0.01, 0),
                                                    ((240, 320), 4, 500, 0.001, 1)])
def test_good_features_exact_list(engine_factory, shape, seed, maxc, q, md):
    a, _ = synth.synth_pair(shape[0], shape[1], seed)
    eng = engine_factory(shape[1], shape[0])
    got = eng.good_features(a, maxc, q, md, 3)
    ref = cv2.goodFeaturesToTrack(a, maxc, q, md, blockSize=3)
    assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.parametrize("shape,seed,maxc,q,md", [((481, 637), 7, 2000, 0.01, 7), ((270, 487), 8, 500, 0.01, 5),
                                                    ((1080, 1917), 9, 2000, 0.01, 7), ((135, 241), 2, 300, 0.02, 3)])
def test_good_features_exact_list_widths_not_multiple_of_32(engine_factory, shape, seed, maxc, q, md):
    """Corner list against LIVE cv2 on this host where the Sobel row filter has a SIMD tail (width % 32 != 0)."""
    a, _ = synth.synth_pair(shape[0], shape[1], seed)
    eng = engine_factory(shape[1], shape[0])
    got = eng.good_features(a, maxc, q, md, 3)
    ref = cv2.goodFeaturesToTrack(a, maxc, q, md, blockSize=3)
    assert got.shape == ref.shape and np.array_equal(got, ref)


@pytest.mark.parametrize("shape,seed,maxc,q,md,k", [((240, 320), 0, 200, 0.01, 7, 0.04), ((135, 241), 1, 50, 0.05, 3.5, 0.06),
                                                      ((481, 637), 7, 2000, 0.01, 7, 0.04), ((1080, 1917), 9, 2000, 0.01, 7, 0.04),
                                                      ((77, 129), 3, 0, 0.02, 2, 0.1)])
def test_good_features_harris_exact_list(engine_factory, shape, seed, maxc, q, md, k):
    """useHarrisDetector=True: cv2's corner list bit for bit (the Harris response reproduces the wheel's arithmetic,
    including its 4-wide step and double tail at the END of the image — sizes with w*h % 8 != 0 are in the list)."""
    import opticalflowcontainer_b200 as ofb
    a, _ = synth.synth_pair(shape[0], shape[1], seed)
    eng = engine_factory(shape[1], shape[0])
    ref = cv2.goodFeaturesToTrack(a, maxc, q, md, blockSize=3, useHarrisDetector=True, k=k)
    got = eng.good_features(a, maxc, q, md, 3, useHarrisDetector=True, k=k)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert np.array_equal(ofb.goodFeaturesToTrack(a, maxc, q, md, blockSize=3, useHarrisDetector=True, k=k), ref)


def test_good_features_mask(engine_factory):
    """cv2's mask argument: candidates outside the mask are dropped AND the quality threshold is relative to the
    strongest response inside the mask."""
    import opticalflowcontainer_b200 as ofb
    a, _ = synth.synth_pair(270, 480, 11)
    mask = np.zeros((270, 480), np.uint8)
    mask[40:200, 100:400] = 255
    mask[90:120, 150:300] = 0
    eng = engine_factory(480, 270)
    for (maxc, q, md) in [(300, 0.01, 7), (0, 0.05, 0), (50, 0.2, 12.5)]:
        ref = cv2.goodFeaturesToTrack(a, maxc, q, md, mask=mask, blockSize=3)
        got = eng.good_features(a, maxc, q, md, 3, mask=mask)
        assert ref is not None and np.array_equal(got, ref)
        assert np.array_equal(ofb.goodFeaturesToTrack(a, maxc, q, md, mask=mask, blockSize=3), ref)
    # a strided (non-contiguous) mask view
    big = np.zeros((270, 512), np.uint8); big[:, :480] = mask
    assert np.array_equal(eng.good_features(a, 300, 0.01, 7, 3, mask=big[:, :480]),
                          cv2.goodFeaturesToTrack(a, 300, 0.01, 7, mask=mask, blockSize=3))
    # mask and Harris response together
    assert np.array_equal(ofb.goodFeaturesToTrack(a, 100, 0.01, 7, mask=mask, useHarrisDetector=True),
                          cv2.goodFeaturesToTrack(a, 100, 0.01, 7, mask=mask, useHarrisDetector=True))


def test_good_features_many_equal_responses(engine_factory):
    """A checkerboard: thousands of candidates share one response value, so one value bucket of the candidate ordering
    is larger than a selection round (the CTA orders it in global memory) and cv2's tie rule — descending address —
    decides the whole list."""
    t = (((np.arange(512)[:, None] // 16) + (np.arange(512)[None, :] // 16)) % 2 * 200 + 20).astype(np.uint8)
    eng = engine_factory(512, 512)
    for (maxc, md) in [(0, 0), (700, 0), (0, 9), (300, 20)]:
        ref = cv2.goodFeaturesToTrack(t, maxc, 0.01, md, blockSize=3)
        got = eng.good_features(t, maxc, 0.01, md, 3)
        assert got.shape == ref.shape and np.array_equal(got, ref), (maxc, md)


def test_good_features_unbounded_and_tiny(engine_factory):
    """maxCorners <= 0 with minDistance < 1 keeps every candidate (corner buffers are sized like the candidate list);
    images without an interior pixel give an empty list, as cv2 does."""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, size=(240, 320), dtype=np.uint8)
    eng = engine_factory(320, 240)
    ref = cv2.goodFeaturesToTrack(a, 0, 0.001, 0, blockSize=3)
    got = eng.good_features(a, 0, 0.001, 0, 3)
    assert len(ref) > 240 * 320 // 16 and np.array_equal(got, ref)
    assert len(eng.good_features(np.zeros((2, 100), np.uint8), 10, 0.01, 3, 3)) == 0
    assert len(eng.good_features(np.zeros((100, 2), np.uint8), 10, 0.01, 3, 3)) == 0


@pytest.mark.parametrize("shape,win", [((270, 480), (21, 21)), ((251, 333), (15, 15)), ((251, 333), (11, 13))],
                         ids=["480x270_win21", "333x251_win15", "333x251_win11x13"])
def test_lk_stream_equals_pair_calls(built_lib, shape, win):
    """(The odd frame size takes the one-pixel-per-thread pyrDown / Scharr kernels, the 15x15 window the second
    instantiation of the CTA-per-point tracker, 11x13 the warp-per-point tracker with a run-time window; the tracking
    calls after the first two are CUDA-graph replays.)
    ofb_lk_stream keeps the previous frame's pyramid, derivatives and corner list on the GPU (one upload per frame):
    what it returns for (frame t-1 -> frame t) equals good_features(t-1) + pyrlk(t-1, t) bit for bit, and cv2 within
    the LK gate; a change of parameters or another sparse call on the handle re-primes it."""
    import opticalflowcontainer_b200 as ofb
    h, w = shape
    base = synth.synth_pair(h, w, 61, (0.0, 0.0))[0]
    fr = [synth.subpixel_shift(base, 1.7 * t, -0.8 * t) for t in range(5)]
    eng, ref = ofb.FlowEngine(w, h, 1, 0), ofb.FlowEngine(w, h, 1, 0)
    kw = dict(maxCorners=400, qualityLevel=0.01, minDistance=7, blockSize=3, winSize=win, maxLevel=3,
              criteria=(3, 30, 0.01))
    try:
        assert eng.lk_stream(fr[0], **kw) is None
        assert np.array_equal(eng.last_corners, ref.good_features(fr[0], 400, 0.01, 7, 3))
        for t in range(1, 5):
            prev, nxt, st, err = eng.lk_stream(fr[t], **kw)
            pts = ref.good_features(fr[t - 1], 400, 0.01, 7, 3)
            want = ref.pyrlk(fr[t - 1], fr[t], pts, None, win, 3, (3, 30, 0.01))
            assert np.array_equal(prev, pts)
            assert np.array_equal(nxt, want[0]) and np.array_equal(st, want[1]) and np.array_equal(err, want[2])
            cpts = cv2.goodFeaturesToTrack(fr[t - 1], 400, 0.01, 7, blockSize=3)
            assert np.array_equal(prev, cpts)
            _lk_check(cv2.calcOpticalFlowPyrLK(fr[t - 1], fr[t], cpts, None, winSize=win, maxLevel=3,
                                               criteria=(3, 30, 0.01)), (nxt, st, err))
        # another sparse call on the handle re-primes the stream; so does a parameter change
        eng.good_features(fr[0], 10, 0.01, 7, 3)
        assert eng.lk_stream(fr[1], **kw) is None
        assert eng.lk_stream(fr[2], **kw) is not None
        assert eng.lk_stream(fr[3], **dict(kw, maxCorners=200)) is None
        eng.lk_stream_reset()
        assert eng.lk_stream(fr[4], **dict(kw, maxCorners=200)) is None
        with pytest.raises(ofb.OfbError):
            eng.lk_stream(fr[0], **dict(kw, maxCorners=0))
    finally:
        eng.close()
        ref.close()


def test_pyrlk_accepts_pyramids(built_lib):
    """cv2.calcOpticalFlowPyrLK also takes pyramids from buildOpticalFlowPyramid (with or without derivative images)."""
    import opticalflowcontainer_b200 as ofb
    a, b = synth.synth_pair(270, 480, 14, (2.6, 1.4))
    pts = cv2.goodFeaturesToTrack(a, 200, 0.01, 7, blockSize=3)
    for with_d in (True, False):
        _, pa = cv2.buildOpticalFlowPyramid(a, (21, 21), 2, withDerivatives=with_d)
        _, pb = cv2.buildOpticalFlowPyramid(b, (21, 21), 2, withDerivatives=with_d)
        # (the Python binding of the wheel rejects pyramid tuples — "prevImg is not a numerical tuple" — so the reference is
        # the image call at the pyramid's depth, which is what the C++ function computes from such pyramids)
        ref = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(21, 21), maxLevel=2, criteria=(3, 30, 0.01))
        got = ofb.calcOpticalFlowPyrLK(pa, pb, pts, None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
        _lk_check(ref, got)
        # the pyramid's depth (2) limits maxLevel: the same as images with maxLevel=2
        same = ofb.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(21, 21), maxLevel=2, criteria=(3, 30, 0.01))
        assert np.array_equal(got[0], same[0]) and np.array_equal(got[1], same[1])


def test_good_features_1080p_2000_corners(engine_factory):
    """BASELINE config[3] front end: 2000 Shi-Tomasi corners at 1080p, exact list and order."""
    a, _ = synth.synth_pair(1080, 1920, 300)
    eng = engine_factory(1920, 1080)
    got = eng.good_features(a, 2000, 0.01, 7, 3)
    ref = cv2.goodFeaturesToTrack(a, 2000, 0.01, 7, blockSize=3)
    assert got.shape == (2000, 1, 2) and np.array_equal(got, ref)


def test_good_features_ties_and_flat_image(engine_factory):
    import opticalflowcontainer_b200 as ofb
    t = np.zeros((64, 64), np.uint8)
    for y in range(8, 64, 16):
        for x in range(8, 64, 16):
            t[y:y + 4, x:x + 4] = 200
    eng = engine_factory(64, 64)
    for md in (0, 5):
        assert np.array_equal(eng.good_features(t, 50, 0.01, md, 3), cv2.goodFeaturesToTrack(t, 50, 0.01, md, blockSize=3))
    flat = np.full((64, 64), 77, np.uint8)
    assert len(eng.good_features(flat, 10, 0.01, 3, 3)) == 0
    assert ofb.goodFeaturesToTrack(flat, 10, 0.01, 3) is None        # cv2 returns None


def _lk_check(ref, got, pos_tol=1e-2):
    rn, rs, re = ref
    gn, gs, ge = got
    assert np.array_equal(rs, gs)
    ok = rs.ravel() == 1
    d = np.sqrt(((rn - gn).reshape(-1, 2)[ok] ** 2).sum(-1))
    assert d.size == 0 or (d.mean() <= 1e-3 and d.max() <= pos_tol), (d.mean(), d.max())
    return ok


def test_pyrlk_matches_cv2(engine_factory):
    a, b = synth.synth_pair(480, 640, 4, (3.2, -1.7))
    eng = engine_factory(640, 480)
    pts = cv2.goodFeaturesToTrack(a, 400, 0.01, 7, blockSize=3)
    extra = np.array([[[1.5, 2.0]], [[638.2, 477.9]], [[0.0, 240.0]], [[320.0, 0.3]], [[-5.0, 10.0]], [[700.0, 100.0]]],
                     np.float32)
    pts = np.concatenate([pts, extra])
    ref = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
    got = eng.pyrlk(a, b, pts, None, (21, 21), 3, (3, 30, 0.01))
    ok = _lk_check(ref, got)
    assert np.abs(ref[2] - got[2])[ok].max() < 1e-2
    # true motion recovered
    mv = (got[0] - pts).reshape(-1, 2)[ok]
    assert abs(np.median(mv[:, 0]) - 3.2) < 0.1 and abs(np.median(mv[:, 1]) + 1.7) < 0.1


@pytest.mark.parametrize("win,lvl,crit,flags,thr", [((15, 11), 2, (3, 10, 0.03), 8, 1e-3), ((31, 31), 4, (1, 5, 0.0), 0, 1e-4),
                                                      ((9, 9), 0, (2, 0, 0.05), 0, 1e-4), ((21, 21), 3, (3, 30, 0.01), 4, 1e-4)])
def test_pyrlk_variants(engine_factory, win, lvl, crit, flags, thr):
    a, b = synth.synth_warp_pair(270, 480, 6, angle_deg=0.8, zoom=1.01)
    eng = engine_factory(480, 270)
    pts = cv2.goodFeaturesToTrack(a, 150, 0.01, 5, blockSize=3)
    init = (pts + np.float32([0.5, -0.5])).astype(np.float32) if flags & 4 else None
    ref = cv2.calcOpticalFlowPyrLK(a, b, pts, None if init is None else init.copy(), winSize=win, maxLevel=lvl,
                                   criteria=crit, flags=flags, minEigThreshold=thr)
    got = eng.pyrlk(a, b, pts, init, win, lvl, crit, flags, thr)
    ok = _lk_check(ref, got)
    if flags & 8:
        assert np.allclose(ref[2], got[2], rtol=1e-3, atol=1e-6)
    else:
        assert np.abs(ref[2] - got[2])[ok].max() < 2e-2


def test_pyrlk_1080p_2000_points_and_module_api(built_lib):
    """BASELINE config[3]: 2000 Shi-Tomasi corners tracked at 1080p through the cv2-signature functions."""
    import opticalflowcontainer_b200 as ofb
    a, b = synth.synth_pair(1080, 1920, 301, (5.5, 2.25))
    pts = ofb.goodFeaturesToTrack(a, 2000, 0.01, 7, blockSize=3)
    assert np.array_equal(pts, cv2.goodFeaturesToTrack(a, 2000, 0.01, 7, blockSize=3))
    got = ofb.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
    ref = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
    _lk_check(ref, got)
    n, pyr = ofb.buildOpticalFlowPyramid(a, (21, 21), 3, True)
    assert n == 3 and np.array_equal(pyr[2], cv2.pyrDown(a))


def test_handles_are_independent_across_threads(built_lib):
    """One handle and one host thread per camera stream (how bench.py --mode lk and a multi-camera node run): the
    concurrent results equal the sequential ones bit for bit."""
    from concurrent.futures import ThreadPoolExecutor
    import opticalflowcontainer_b200 as ofb
    h, w, n = 270, 480, 4
    pairs = [synth.synth_pair(h, w, 500 + s, (2.1 + 0.5 * s, -1.3 + 0.4 * s)) for s in range(n)]
    engines = [ofb.FlowEngine(w, h, 1, 0) for _ in range(n)]
    try:
        def work(i):
            a, b = pairs[i]
            pts = engines[i].good_features(a, 500, 0.01, 7, 3)
            nxt, st, err = engines[i].pyrlk(a, b, pts, None, (21, 21), 3, (3, 30, 0.01))
            flow = engines[i].farneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
            return pts, nxt, st, err, flow
        seq = [work(i) for i in range(n)]
        with ThreadPoolExecutor(max_workers=n) as pool:
            for _ in range(3):
                par = list(pool.map(work, range(n)))
                for s, p in zip(seq, par):
                    for x, y in zip(s, p):
                        assert np.array_equal(x, y)
    finally:
        for e in engines:
            e.close()
