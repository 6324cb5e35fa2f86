"""Generates the committed golden fixtures from the reference implementation of the path (the
cv2 wheel, opencv-python-headless 4.13.0.92).  Run from the repo root:
    python tests/golden/make_golden.py
Small frames only (fixtures stay a few hundred KB); flows stored as float16-safe float32 npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cv2_oracle as C  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    import cv2
    assert cv2.__version__ == "4.13.0", cv2.__version__
    cases = [
        ("default_96x128", synth.synth_pair(96, 128, 11, (1.7, -0.9)), {}),
        ("gauss_97x131", synth.synth_pair(97, 131, 12, (2.5, 1.25)), dict(flags=256, winsize=9)),
        ("poly7_lv1_80x112", synth.synth_pair(80, 112, 13, (-1.5, 2.0)), dict(poly_n=7, poly_sigma=1.5, levels=1)),
        ("lowtex_96x128", synth.low_texture_pair(96, 128, 14), {}),
    ]
    for name, (a, b), kw in cases:
        flow = C.farneback(a, b, **kw)
        d = {"prev": a, "next": b, "flow": flow}
        d.update({"kw_" + k: np.asarray(v) for k, v in kw.items()})
        np.savez_compressed(os.path.join(OUT, "farneback_%s.npz" % name), **d)
        print(name, flow.reshape(-1, 2).mean(0))
    around_the_path(cv2)


def around_the_path(cv2):
    """Golden vectors for the calls around the flow call (ingest, pre-filter, post-filter, sparse path): inputs and the
    wheel's outputs, small enough to commit.  tests/test_golden_around.py checks the NumPy restatements against them
    (CPU) and the CUDA path (GPU)."""
    rng = np.random.default_rng(2024)
    bgr = cv2.GaussianBlur(rng.integers(0, 256, size=(90, 136, 3), dtype=np.uint8), (0, 0), 1.2)
    bgr[:20] = rng.integers(0, 256, size=(20, 136, 3), dtype=np.uint8)
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    d = {"bgr": bgr, "gray": gray, "gray_rgb_order": cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY)}
    d["resize_bgr_64x48"] = cv2.resize(bgr, (64, 48))
    d["resize_gray_200x133"] = cv2.resize(gray, (200, 133))
    d["ingest_gray_64x48"] = cv2.cvtColor(cv2.resize(bgr, (64, 48)), cv2.COLOR_BGR2GRAY)
    d["clahe_2_8x8"] = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(gray)
    d["clahe_40_4x6"] = cv2.createCLAHE(clipLimit=40.0, tileGridSize=(4, 6)).apply(gray)
    hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
    d["hsv"] = hsv
    h, s, v = cv2.split(hsv)
    contrast = np.std(v) / (np.mean(v) + 1e-3)
    clip = float(np.clip(1.0 + (contrast - 0.1) / (0.8 - 0.1) * (4.0 - 1.0), 1.0, 4.0))
    cl = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8))
    cl.setClipLimit(clip)
    d["adapt_clip"] = np.float64(clip)
    d["adapt_rgb"] = cv2.cvtColor(cv2.merge((h, s, cl.apply(v))), cv2.COLOR_HSV2RGB)
    flow = (rng.normal(size=(60, 88, 2)) * 2).astype(np.float32)
    flow[rng.random((60, 88)) < 0.15] = 0.0
    d["flow"] = flow
    for k in (3, 5):
        d["median%d" % k] = np.stack([cv2.medianBlur(np.ascontiguousarray(flow[..., c]), k) for c in range(2)], -1)
    a, b = synth.synth_pair(120, 160, 21, (2.2, -1.4))
    pts = cv2.goodFeaturesToTrack(a, 60, 0.01, 7, blockSize=3)
    nxt, st, err = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(21, 21), maxLevel=3,
                                            criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01))
    d.update({"lk_prev": a, "lk_next": b, "corners": pts, "lk_next_pts": nxt, "lk_status": st, "lk_err": err,
              "min_eig": cv2.cornerMinEigenVal(a, 3), "pyrdown": cv2.pyrDown(a)})
    np.savez_compressed(os.path.join(OUT, "around_the_path.npz"), **d)
    print("around_the_path", {k: np.asarray(v).shape for k, v in d.items()})


if __name__ == "__main__":
    main()
