"""Golden fixture of the junction detector (tests/golden/junction.npz), generated from the reference side:
the cv2 4.13.0 wheel for the pixel stages and contours, and oracle/_ref/junction_cluster — the reference's OWN vendored
nanoflann header compiled by oracle/Makefile — for the clustering.  Run from the repo root (needs /root/reference):
    make -C oracle && python tests/golden/make_golden_junction.py
Also stores JPEG fixtures (cv2.imencode streams and their cv2.imdecode output) for the compressed-image ingest."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "junction_cluster")


def ref_cluster(cand, eps):
    txt = "".join("%d %d\n" % (x, y) for x, y in cand)
    out = subprocess.run([REF_BIN, str(eps)], input=txt, capture_output=True, text=True, check=True).stdout
    return np.array([[float(a) for a in ln.split()] for ln in out.splitlines()], np.float32).reshape(-1, 2)


def cv2_candidates(cv2, thresh, grid_area, thr):
    """junction_detector.cpp:72-117 with the wheel's findContours / contourArea / boundingRect."""
    cs, _ = cv2.findContours(thresh, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    thr2 = np.float32(2) * np.float32(thr)
    lo, hi = grid_area * float(np.float32(1) / thr2), grid_area * float(thr2)
    out, recs = [], []
    for c in cs:
        area = cv2.contourArea(c)
        x, y, w, h = cv2.boundingRect(c)
        recs.append((int(round(area * 2)), x, y, w, h))
        if lo < area < hi and area / float(w * h) >= 0.4 and 0.5 <= w / h <= 2.0:
            out += [(x - 1, y - 1), (x + w + 1, y - 1), (x + w + 1, y + h + 1), (x - 1, y + h + 1)]
    return np.asarray(out, np.float32).reshape(-1, 2), np.asarray(recs, np.int64).reshape(-1, 5)


def main():
    import cv2
    assert cv2.__version__ == "4.13.0", cv2.__version__
    d = {}
    for name, (h, w, seed, ga, eps) in {"a": (120, 160, 0, 200, 6), "b": (97, 131, 1, 250, 4), "c": (150, 203, 2, 200, 6)}.items():
        img = synth.synth_net(h, w, seed)
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        blur = cv2.GaussianBlur(gray, (3, 3), 0)
        th = cv2.adaptiveThreshold(blur, 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
        cand, recs = cv2_candidates(cv2, th, ga, 2.0)
        d[name + "_img"], d[name + "_thresh"], d[name + "_contours"], d[name + "_cand"] = img, th, recs, cand
        d[name + "_junctions"] = ref_cluster(cand, eps)
        d[name + "_params"] = np.array([ga, eps])
        print(name, len(recs), "contours", len(cand), "candidates", len(d[name + "_junctions"]), "junctions")
    rng = np.random.default_rng(7)
    for i in range(6):                                          # clustering alone on random candidate sets
        pts = rng.integers(0, int(rng.integers(20, 200)), (int(rng.integers(4, 300)), 2)).astype(np.float32)
        eps = int(rng.integers(2, 9))
        d["rand%d_cand" % i], d["rand%d_eps" % i], d["rand%d_junctions" % i] = pts, np.array(eps), ref_cluster(pts, eps)
    np.savez_compressed(os.path.join(OUT, "junction.npz"), **d)

    j = {}
    for name, (h, w, q, samp) in {"s420": (37, 53, 85, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420),
                                  "s422": (40, 48, 60, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422),
                                  "s444": (24, 31, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)}.items():
        img = synth.synth_net(h, w, q)
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, samp])
        j[name + "_jpeg"], j[name + "_bgr"] = buf, cv2.imdecode(buf, cv2.IMREAD_COLOR)
    np.savez_compressed(os.path.join(OUT, "jpeg.npz"), **j)


if __name__ == "__main__":
    main()
