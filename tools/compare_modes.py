"""Bitwise comparison of the flow fields two OFB_ITER_MODE settings produce (the schedules differ, the arithmetic must not).
Usage: python tools/compare_modes.py 2 3"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import opticalflowcontainer_b200 as ofb
from oracle import synth

ma, mb = sys.argv[1], sys.argv[2]
res = {}
for (h, w, shift) in [(1080, 1920, (6.2, 3.4)), (481, 637, (-14.3, 9.6)), (270, 480, (1.7, -0.9))]:
    a, b = synth.synth_pair(h, w, 1, shift)
    b = np.ascontiguousarray(np.roll(b, 3, axis=0))            # a discontinuity for the reuse test to fail on
    out = []
    for m in (ma, mb):
        os.environ["OFB_ITER_MODE"] = m
        eng = ofb.FlowEngine(w, h, 1, 0)
        out.append(eng.farneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0).copy())
        eng.close()
    d = np.abs(out[0] - out[1])
    print("%dx%d modes %s vs %s: identical=%s max|diff|=%.3g n_diff=%d max|flow|=%.2f" %
          (w, h, ma, mb, np.array_equal(out[0], out[1]), d.max(), int((d > 0).sum()), np.abs(out[0]).max()))
