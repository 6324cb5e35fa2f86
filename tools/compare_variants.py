"""Bitwise comparison of the flow fields two builds of libofb.so produce (experiment builds differ in schedule / memory
placement, never in arithmetic).  Usage: python tools/compare_variants.py <libA.so> <libB.so>
Each library runs in its own process (OFB_LIB is read at import)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import opticalflowcontainer_b200 as ofb
from oracle import synth
out = {}
for (h, w, shift) in [(1080, 1920, (6.2, 3.4)), (481, 637, (-14.3, 9.6)), (270, 480, (1.7, -0.9))]:
    a, b = synth.synth_pair(h, w, 1, shift)
    b = np.ascontiguousarray(np.roll(b, 3, axis=0))            # a discontinuity for the row-reuse test to fail on
    eng = ofb.FlowEngine(w, h, 1, 0)
    out["%%dx%%d" %% (w, h)] = eng.farneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0).copy()
    eng.close()
np.savez(sys.argv[1], **out)
''' % ROOT

def run(lib):
    f = tempfile.mktemp(suffix=".npz")
    env = dict(os.environ, OFB_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", CHILD, f], check=True, env=env)
    return f

if __name__ == "__main__":
    import numpy as np
    fa, fb = run(sys.argv[1]), run(sys.argv[2])
    za, zb = np.load(fa), np.load(fb)
    for k in za.files:
        d = np.abs(za[k] - zb[k])
        print("%s %s vs %s: identical=%s max|diff|=%.3g n_diff=%d max|flow|=%.2f" % (
            k, os.path.basename(sys.argv[1]), os.path.basename(sys.argv[2]), np.array_equal(za[k], zb[k]), d.max(),
            int((d > 0).sum()), np.abs(za[k]).max()))
