"""Pins the NumPy restatement (oracle/farneback_np.py) against the reference implementation of
the path — the cv2 4.13.0 wheel — live and through the committed golden fixtures."""
import os

import numpy as np
import pytest

from oracle import cv2_oracle as C
from oracle import farneback_np as F
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
cv2 = pytest.importorskip("cv2")


def test_level_schedule_matches_survey():
    s = F.level_schedule(640, 480, 0.5, 3)
    assert [(l.sigma, l.ksize, l.width, l.height) for l in s] == [(3.5, 19, 80, 60), (1.5, 9, 160, 120),
                                                                 (0.5, 3, 320, 240), (0.0, 3, 640, 480)]
    s = F.level_schedule(1920, 1080, 0.5, 3)
    assert [(l.width, l.height) for l in s] == [(240, 135), (480, 270), (960, 540), (1920, 1080)]
    assert len(F.level_schedule(640, 480, 0.5, 8)) == 4        # clamped by the 32-px rule
    assert len(F.level_schedule(640, 480, 0.5, 0)) == 1


def test_gaussian_kernel_matches_cv2():
    for ks, sg in [(3, 0.5), (9, 1.5), (19, 3.5), (3, 0.0)]:
        ref = cv2.getGaussianKernel(ks, sg, cv2.CV_32F).ravel()
        assert np.max(np.abs(ref - F.gaussian_kernel_f32(ks, sg))) <= 6e-8


def test_pyramid_level_matches_cv2():
    a, _ = synth.synth_pair(203, 317, 5)
    for lv in F.level_schedule(317, 203, 0.5, 2):
        f = cv2.GaussianBlur(a.astype(np.float32), (lv.ksize, lv.ksize), lv.sigma, lv.sigma)
        ref = cv2.resize(f, (lv.width, lv.height), interpolation=cv2.INTER_LINEAR)
        assert np.max(np.abs(ref - F.pyramid_level(a, lv))) < 1e-3  # values are 0..255: ~1e-6 relative


def test_resize_area_matches_cv2():
    rng = np.random.default_rng(0)
    f = rng.standard_normal((97, 131, 2)).astype(np.float32)
    for (w, h) in [(16, 12), (33, 25), (65, 49)]:
        ref = cv2.resize(f, (w, h), interpolation=cv2.INTER_AREA)
        assert np.max(np.abs(ref - F.resize_area_f32(f, w, h))) < 1e-5


CASES = [
    ("vga_small_shift", (480, 640), 0, (1.7, -0.9), {}),
    ("vga_large_shift", (480, 640), 1, (14.3, -9.6), {}),
    ("odd_size", (481, 637), 2, (3.1, 2.2), {}),
    ("gaussian", (240, 320), 3, (3, 2), dict(flags=256)),
    ("winsize16", (240, 320), 4, (3, 2), dict(winsize=16)),
    ("winsize5", (240, 320), 4, (2, 1), dict(winsize=5)),
    ("poly7", (240, 320), 5, (3, 2), dict(poly_n=7, poly_sigma=1.5)),
    ("pyr08", (240, 320), 6, (3, 2), dict(pyr_scale=0.8, levels=5)),
    ("levels0", (200, 300), 7, (2, 1), dict(levels=0)),
    ("iter1", (240, 320), 8, (2, 1), dict(iterations=1)),
]


@pytest.mark.parametrize("name,shape,seed,shift,kw", CASES, ids=[c[0] for c in CASES])
def test_restatement_matches_cv2(name, shape, seed, shift, kw):
    a, b = synth.synth_pair(shape[0], shape[1], seed, shift)
    ref = C.farneback(a, b, **kw)
    mine = F.farneback(a, b, **kw)
    mean, mx = C.epe(ref, mine)
    # gate written out: restatement vs cv2 wheel, two orders tighter than the product gate
    assert mean < 1e-4 and mx < 2e-2, (mean, mx)


def test_restatement_low_texture_and_initial_flow():
    a, b = synth.low_texture_pair(240, 320, 3)
    mean, mx = C.epe(C.farneback(a, b), F.farneback(a, b))
    assert mean < 1e-4 and mx < 2e-2
    a, b = synth.synth_pair(240, 320, 2, (5, 3))
    f0 = np.full((240, 320, 2), (4.5, 2.5), np.float32)
    mean, mx = C.epe(C.farneback(a, b, flow=f0.copy(), flags=4), F.farneback(a, b, flow0=f0, flags=4))
    assert mean < 1e-4 and mx < 2e-2


def test_golden_fixtures():
    """tests/golden/farneback_*.npz were produced by cv2 4.13.0 (make_golden.py); the restatement
    must reproduce them without cv2 in the loop."""
    import glob
    files = sorted(glob.glob(os.path.join(GOLDEN, "farneback_*.npz")))
    assert files, "golden fixtures missing"
    for f in files:
        z = np.load(f)
        kw = {k[3:]: z[k].item() for k in z.files if k.startswith("kw_")}
        mine = F.farneback(z["prev"], z["next"], **kw)
        mean, mx = C.epe(z["flow"], mine)
        assert mean < 1e-4 and mx < 2e-2, (f, mean, mx)
