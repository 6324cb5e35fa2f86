"""Oracle pinning for the frame resize in front of the flow call: the NumPy restatement of
cv2.resize (INTER_LINEAR, uint8) equals the cv2 wheel bit for bit."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import resize_np

CASES = [((720, 1280), (480, 640)), ((1080, 1920), (480, 640)), ((480, 640), (1080, 1920)), ((481, 637), (240, 320)),
         ((100, 100), (50, 50)), ((48, 64), (108, 192)), ((48, 64), (96, 128)), ((37, 53), (37, 53)), ((64, 48), (5, 7)),
         ((5, 7), (64, 48)), ((300, 400), (301, 399))]


@pytest.mark.parametrize("src,dst", CASES)
@pytest.mark.parametrize("cn", [1, 3])
def test_resize_np_equals_cv2(src, dst, cn):
    rng = np.random.default_rng(src[0] * 7 + dst[1] + cn)
    shape = src if cn == 1 else src + (cn,)
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    assert np.array_equal(resize_np.resize_linear_u8(img, dst[1], dst[0]), cv2.resize(img, (dst[1], dst[0])))
