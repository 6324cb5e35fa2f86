"""Stage view of the JPEG ingest: python tools/jpeg_profile.py (under ncu for the launch list)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import opticalflowcontainer_b200 as ofb
from oracle import synth
eng = ofb.FlowEngine(64, 64, 1, 0)
for (h, w) in ((480, 640), (1080, 1920)):
    img = synth.synth_net(h, w, 3)
    jpg = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])[1]
    for _ in range(3):
        eng.ingest_jpeg_gray(jpg)
    l0 = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(10):
        eng.ingest_jpeg_gray(jpg)
    dt = (time.perf_counter() - t0) / 10 * 1e3
    print("%dx%d %d bytes: %.3f ms per frame, %d launches per frame" % (w, h, jpg.size, dt, (eng.launch_count - l0) // 10))
