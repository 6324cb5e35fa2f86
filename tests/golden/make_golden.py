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


if __name__ == "__main__":
    main()
