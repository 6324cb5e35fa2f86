"""Junction detector, host frame in -> junction list out, against the cv2 + nanoflann-restatement chain on one host core:
python tools/junction_bench.py   (1080p and 640x480 net frames)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cv2
import opticalflowcontainer_b200 as ofb
from oracle import synth

cv2.setNumThreads(1)
eng = ofb.FlowEngine(64, 64, 1, 0)


def t(fn, reps=10):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def cv_chain(img):
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    th = cv2.adaptiveThreshold(cv2.GaussianBlur(gray, (3, 3), 0), 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
    cs, _ = cv2.findContours(th, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    out = []
    for c in cs:
        a = cv2.contourArea(c)
        if 50 < a < 800:
            x, y, w, h = cv2.boundingRect(c)
            if a / (w * h) >= 0.4 and 0.5 <= w / h <= 2.0:
                out.append((x, y, w, h))
    return out


print("%-50s %10s %14s" % ("frame", "B200 ms", "cv2 ms (no clustering)"))
for h, w in ((480, 640), (1080, 1920)):
    img = synth.synth_net(h, w, 1)
    n = len(eng.find_junctions(img, 200, 2.0, 6))
    tl = eng.timing(True) if hasattr(eng, "timing") else None
    print("%-50s %10.3f %14.3f" % ("%dx%d net frame, %d junctions" % (w, h, n), t(lambda: eng.find_junctions(img, 200, 2.0, 6)), t(lambda: cv_chain(img))))
