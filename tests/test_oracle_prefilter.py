"""Oracle pinning for the adapt node's image pre-filter chain: BGR2HSV, HSV2RGB and
bilateralFilter restated in NumPy equal the cv2 wheel bit for bit; so does the whole chain."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import prefilter_np as P


def _frame(h, w, seed):
    rng = np.random.default_rng(seed)
    a = cv2.GaussianBlur(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8), (0, 0), 1.5)
    a[: h // 4] = rng.integers(0, 256, size=(h // 4, w, 3), dtype=np.uint8)       # a noisy band: every hue and saturation
    return a


def test_bgr2hsv_equals_cv2_exhaustive_slices():
    g = np.arange(256, dtype=np.uint8)
    for fixed in (0, 17, 128, 255):
        img = np.stack(list(np.meshgrid(g, g, indexing="ij")) + [np.full((256, 256), fixed, np.uint8)], -1).astype(np.uint8)
        for perm in ([0, 1, 2], [2, 0, 1], [1, 2, 0]):
            x = np.ascontiguousarray(img[..., perm])
            assert np.array_equal(P.bgr2hsv_u8(x), cv2.cvtColor(x, cv2.COLOR_BGR2HSV))
            assert np.array_equal(P.bgr2hsv_u8(x, rgb_order=True), cv2.cvtColor(x, cv2.COLOR_RGB2HSV))


@pytest.mark.parametrize("width", [1000, 1024, 37, 31, 640])
def test_hsv2rgb_equals_cv2(width):
    """incl. the scalar tail of every row (width % 32 pixels), whose arithmetic differs from the vector body's"""
    rng = np.random.default_rng(width)
    hgt = 300
    n = hgt * width
    hsv = np.stack([rng.integers(0, 180, n), rng.integers(0, 256, n), rng.integers(0, 256, n)], -1).astype(np.uint8).reshape(hgt, width, 3)
    assert np.array_equal(P.hsv2rgb_u8(hsv), cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB))
    view = np.ascontiguousarray(hsv[:, : width - 3])                     # and a non-multiple width from the same data
    assert np.array_equal(P.hsv2rgb_u8(view), cv2.cvtColor(view, cv2.COLOR_HSV2RGB))


@pytest.mark.parametrize("params", [(9, 75.0, 75.0), (5, 50.0, 50.0), (9, 25.5, 10.0), (0, 30.0, 3.0), (7, 12.0, 2.5)])
def test_bilateral_restatement_is_not_pinned_yet(params):
    """The bilateral restatement agrees with cv2 except at rounding ties (values within one float ulp of x.5, where the
    wheel rounds up): at most a few values per 100 000, off by one.  NOT bit-exact (the wheel's IPP path; parity unpinned,
    DESIGN.md 7); this test records how close the restatement is — the device filter equals the restatement bit for bit."""
    img = _frame(150, 200, int(params[1]))
    d = np.abs(P.bilateral_u8c3(img, *params).astype(int) - cv2.bilateralFilter(img, *params).astype(int))
    assert d.max() <= 1 and (d > 0).sum() <= 1e-4 * d.size


@pytest.mark.parametrize("size", [(240, 320), (135, 243)])
def test_adapt_prefilter_chain_equals_cv2(size):
    bgr = _frame(size[0], size[1], 5)
    hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
    h, s, v = cv2.split(hsv)
    clip = P.adaptive_clip(v, 1.0, 4.0, 0.1, 0.8)
    clahe = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8))
    clahe.setClipLimit(clip)
    want = cv2.cvtColor(cv2.merge((h, s, clahe.apply(v))), cv2.COLOR_HSV2RGB)
    assert np.array_equal(P.adapt_prefilter_np(bgr, True, None, (1.0, 4.0, 0.1, 0.8), (8, 8)), want)
    assert np.array_equal(P.adapt_prefilter_np(bgr, False), cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
