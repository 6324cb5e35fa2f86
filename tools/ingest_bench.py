"""End-to-end time (host buffer in, host buffer out, pinned) of the calls around the flow call against cv2 on one host
core: python tools/ingest_bench.py  (1080p camera frame -> 640x480 node size, as the reference nodes are configured)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cv2
import torch
import opticalflowcontainer_b200 as ofb

cv2.setNumThreads(1)
rng = np.random.default_rng(0)
big = torch.from_numpy(cv2.GaussianBlur(rng.integers(0, 256, size=(1080, 1920, 3), dtype=np.uint8), (0, 0), 2.0)).pin_memory().numpy()
small = cv2.resize(big, (640, 480))
gray = cv2.cvtColor(big, cv2.COLOR_BGR2GRAY)
eng = ofb.FlowEngine(1920, 1080, 1, 0)
a, b = gray, np.roll(gray, 3, axis=1)
flow = eng.farneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
flow_t = np.transpose(flow, (2, 0, 1)).copy()


def t(fn, reps=20):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def cv_adapt(bgr):
    h, s, v = cv2.split(cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV))
    c = float(np.clip(1.0 + (np.std(v) / (np.mean(v) + 1e-3) - 0.1) / 0.7 * 3.0, 1.0, 4.0))
    cl = cv2.createCLAHE(clipLimit=c, tileGridSize=(8, 8))
    return cv2.cvtColor(cv2.merge((h, s, cl.apply(v))), cv2.COLOR_HSV2RGB)


def cv_post(f):
    u = cv2.medianBlur(f[0], 5); v = cv2.medianBlur(f[1], 5)
    m = (np.sqrt(u ** 2 + v ** 2) >= 0.5).astype(np.float32)
    return float(np.mean(u * m))


def gpu_post():
    eng.farneback(a, b, None, 0.5, 0, 15, 0, 5, 1.2, 0)      # (a cheap call to put a field on the device; not timed apart)
    eng.flow_postfilter(1, 5, 0.5)
    return eng.flow_u_stats(1, median=False)


jpg = cv2.imencode(".jpg", big, [cv2.IMWRITE_JPEG_QUALITY, 90])[1]
jpg_vga = cv2.imencode(".jpg", small, [cv2.IMWRITE_JPEG_QUALITY, 90])[1]
bil = lambda x: cv2.bilateralFilter(x, 9, 75.0, 75.0)

rows = [
    ("JPEG 1080p q90 (%d kB) -> bgr (imdecode)" % (jpg.size // 1000), lambda: eng.imdecode(jpg), lambda: cv2.imdecode(jpg, cv2.IMREAD_COLOR)),
    ("JPEG 1080p -> gray (imdecode + cvtColor)", lambda: eng.imdecode(jpg, gray=True), lambda: cv2.cvtColor(cv2.imdecode(jpg, cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY)),
    ("JPEG 1080p -> 640x480 gray (imdecode + resize + cvtColor)", lambda: eng.ingest_jpeg_gray(jpg, (640, 480)), lambda: cv2.cvtColor(cv2.resize(cv2.imdecode(jpg, cv2.IMREAD_COLOR), (640, 480)), cv2.COLOR_BGR2GRAY)),
    ("JPEG 640x480 q90 (%d kB) -> gray" % (jpg_vga.size // 1000), lambda: eng.ingest_jpeg_gray(jpg_vga), lambda: cv2.cvtColor(cv2.imdecode(jpg_vga, cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY)),
    ("bilateral 640x480 rgb (d 9, 75, 75)", lambda: eng.bilateral_filter(small, 9, 75.0, 75.0), lambda: bil(small)),
    ("bilateral 1080p rgb (d 9, 75, 75)", lambda: eng.bilateral_filter(big, 9, 75.0, 75.0), lambda: bil(big)),
    ("ingest 1080p bgr8 -> 640x480 gray (resize + cvtColor)", lambda: eng.ingest_gray(big, (640, 480)), lambda: cv2.cvtColor(cv2.resize(big, (640, 480)), cv2.COLOR_BGR2GRAY)),
    ("cvtColor BGR2GRAY 1080p", lambda: eng.ingest_gray(big), lambda: cv2.cvtColor(big, cv2.COLOR_BGR2GRAY)),
    ("resize 1080p bgr -> 640x480", lambda: eng.resize(big, (640, 480)), lambda: cv2.resize(big, (640, 480))),
    ("CLAHE 1080p gray", lambda: eng.clahe(gray, 2.0, (8, 8)), lambda: cv2.createCLAHE(2.0, (8, 8)).apply(gray)),
    ("adapt pre-filter 1080p (BGR2HSV, adaptive CLAHE, HSV2RGB)", lambda: eng.adapt_prefilter(big), lambda: cv_adapt(big)),
    ("adapt pre-filter 640x480", lambda: eng.adapt_prefilter(small), lambda: cv_adapt(small)),
]
print("%-62s %10s %10s" % ("call (host buffers in and out)", "B200 ms", "cv2 ms"))
for name, g, c in rows:
    print("%-62s %10.3f %10.3f" % (name, t(g), t(c)))
base = t(lambda: eng.farneback(a, b, None, 0.5, 0, 15, 0, 5, 1.2, 0))
print("%-62s %10.3f %10.3f" % ("flow post-filter 1080p (median 5x5 + magnitude mask + mean)", t(gpu_post) - base, t(lambda: cv_post(flow_t))))
