"""Oracle pinning for the adapt node's flow post-filter: the NumPy restatement of
cv2.medianBlur(float32) equals the cv2 wheel bit for bit (incl. borders, ties, signed zeros)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import postfilter_np


@pytest.mark.parametrize("k", [3, 5])
@pytest.mark.parametrize("shape", [(37, 53), (5, 7), (64, 64), (3, 3)])
def test_median_blur_np_equals_cv2(k, shape):
    rng = np.random.default_rng(k * 100 + shape[0])
    a = (rng.normal(size=shape) * 4).astype(np.float32)
    a[rng.random(shape) < 0.2] = 0.0          # ties
    a[0, 0] = np.float32(-0.0)
    assert np.array_equal(postfilter_np.median_blur_np(a, k), cv2.medianBlur(a, k))


def test_adapt_postfilter_cv2_and_np_median_agree():
    rng = np.random.default_rng(3)
    f = (rng.normal(size=(48, 64, 2)) * 2).astype(np.float32)
    g = rng.integers(0, 256, size=(48, 64), dtype=np.uint8)
    a = postfilter_np.adapt_postfilter_np(f, 5, 1.5, g, 100)
    b = postfilter_np.adapt_postfilter_np(f, 5, 1.5, g, 100, use_cv2_median=True)
    assert np.array_equal(a, b)
    assert (a[g >= 100] == 0).all()
