"""Committed golden vectors of the cv2 wheel for the calls AROUND the flow call (tests/golden/around_the_path.npz,
made by tests/golden/make_golden.py): the NumPy restatements reproduce them without cv2 (CPU), and so does the CUDA
path through the C ABI (GPU).  Bit-exact everywhere except the LK positions (1e-3 px, as in test_sparse_gpu.py)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "around_the_path.npz"))


def test_oracle_restatements_match_golden():
    from oracle import clahe_np, features_np, lk_np, postfilter_np, prefilter_np, resize_np
    from opticalflowcontainer_b200.node import to_gray_u8
    bgr, gray = G["bgr"], G["gray"]
    assert np.array_equal(to_gray_u8(bgr, "bgr8"), gray)
    assert np.array_equal(to_gray_u8(bgr, "rgb8"), G["gray_rgb_order"])
    assert np.array_equal(resize_np.resize_linear_u8(bgr, 64, 48), G["resize_bgr_64x48"])
    assert np.array_equal(resize_np.resize_linear_u8(gray, 200, 133), G["resize_gray_200x133"])
    assert np.array_equal(to_gray_u8(resize_np.resize_linear_u8(bgr, 64, 48), "bgr8"), G["ingest_gray_64x48"])
    assert np.array_equal(clahe_np.clahe_apply(gray, 2.0, (8, 8)), G["clahe_2_8x8"])
    assert np.array_equal(clahe_np.clahe_apply(gray, 40.0, (4, 6)), G["clahe_40_4x6"])
    assert np.array_equal(prefilter_np.bgr2hsv_u8(bgr), G["hsv"])
    assert np.array_equal(prefilter_np.adapt_prefilter_np(bgr, True, None, (1.0, 4.0, 0.1, 0.8), (8, 8)), G["adapt_rgb"])
    assert abs(prefilter_np.adaptive_clip(G["hsv"][..., 2], 1.0, 4.0, 0.1, 0.8) - float(G["adapt_clip"])) < 1e-12
    for k in (3, 5):
        assert np.array_equal(postfilter_np.adapt_postfilter_np(G["flow"], k), G["median%d" % k])
    a = G["lk_prev"]
    assert np.array_equal(lk_np.pyr_down(a), G["pyrdown"])
    assert np.array_equal(features_np.corner_min_eigenval(a, 3), G["min_eig"])
    assert np.array_equal(features_np.good_features(a, 60, 0.01, 7.0, 3).reshape(-1, 1, 2), G["corners"])
    nxt, st, err = lk_np.calc_pyrlk(a, G["lk_next"], G["corners"])
    assert np.array_equal(np.asarray(st).reshape(-1), G["lk_status"].reshape(-1))
    ok = G["lk_status"].reshape(-1) == 1
    assert np.abs(np.asarray(nxt).reshape(-1, 2)[ok] - G["lk_next_pts"].reshape(-1, 2)[ok]).max() <= 1e-3


@pytest.mark.gpu
def test_cuda_path_matches_golden(engine_factory):
    eng = engine_factory(200, 160)
    bgr, gray = G["bgr"], G["gray"]
    assert np.array_equal(eng.cvt_gray(bgr), gray)
    assert np.array_equal(eng.cvt_gray(bgr, rgb=True), G["gray_rgb_order"])
    assert np.array_equal(eng.resize(bgr, (64, 48)), G["resize_bgr_64x48"])
    assert np.array_equal(eng.resize(gray, (200, 133)), G["resize_gray_200x133"])
    assert np.array_equal(eng.ingest_gray(bgr, (64, 48)), G["ingest_gray_64x48"])
    assert np.array_equal(eng.clahe(gray, 2.0, (8, 8)), G["clahe_2_8x8"])
    assert np.array_equal(eng.clahe(gray, 40.0, (4, 6)), G["clahe_40_4x6"])
    rgb, clip = eng.adapt_prefilter(bgr, None, (1.0, 4.0, 0.1, 0.8), (8, 8))
    assert np.array_equal(rgb, G["adapt_rgb"]) and abs(clip - float(G["adapt_clip"])) < 1e-9
    flow = G["flow"]
    h, w = flow.shape[:2]
    z = np.zeros((h, w), np.uint8)
    for k in (3, 5):
        eng.farneback(z, z, flow.copy(), 0.5, 0, 9, 0, 5, 1.1, 4)      # iterations = 0 + USE_INITIAL_FLOW: field = input
        eng.flow_postfilter(1, k)
        assert np.array_equal(eng.flow_download(1, h, w)[0], G["median%d" % k])
    a, b = G["lk_prev"], G["lk_next"]
    assert np.array_equal(eng.corner_min_eigenval(a, 3), G["min_eig"])
    pts = eng.good_features(a, 60, 0.01, 7, 3)
    assert np.array_equal(pts, G["corners"])
    nxt, st, err = eng.pyrlk(a, b, pts, None, (21, 21), 3, (3, 30, 0.01))
    assert np.array_equal(st.reshape(-1), G["lk_status"].reshape(-1))
    ok = G["lk_status"].reshape(-1) == 1
    assert np.abs(nxt.reshape(-1, 2)[ok] - G["lk_next_pts"].reshape(-1, 2)[ok]).max() <= 1e-3
