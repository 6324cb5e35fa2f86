"""Oracle pinning for the adapt node's CLAHE pre-filter: the NumPy restatement equals
cv2.createCLAHE(...).apply bit for bit — divisible and non-divisible sizes, clip limits incl. 0, two grids."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import clahe_np


def _img(h, w, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == 0:
        return rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    a = cv2.GaussianBlur((rng.random((h, w)) * 255).astype(np.float32), (0, 0), 3)
    return ((a - a.min()) / (a.max() - a.min()) * 120 + 40).astype(np.uint8)      # low contrast: the clip limit bites


@pytest.mark.parametrize("size", [(240, 320), (241, 317), (96, 128), (33, 45)])
@pytest.mark.parametrize("clip", [2.0, 40.0, 0.0, 3.7])
@pytest.mark.parametrize("grid", [(8, 8), (4, 6)])
def test_clahe_np_equals_cv2(size, clip, grid):
    for kind in (0, 1):
        im = _img(size[0], size[1], kind, size[1] + int(clip * 10) + grid[0])
        assert np.array_equal(clahe_np.clahe_apply(im, clip, grid), cv2.createCLAHE(clipLimit=clip, tileGridSize=grid).apply(im))
