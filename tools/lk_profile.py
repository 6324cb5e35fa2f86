"""Single-camera LK stream call for profiling: python tools/lk_profile.py [frames]  (under ncu for the launch list /
a --set full capture of the sparse kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opticalflowcontainer_b200 as ofb
from oracle import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
base = synth.synth_pair(1080, 1920, 300, (0.0, 0.0))[0]
seq = [synth.subpixel_shift(base, 1.3 * t, -0.9 * t) for t in range(4)]
order = [0, 1, 2, 3, 2, 1]
eng = ofb.FlowEngine(1920, 1080, 1, 0)
for t in range(3):
    eng.lk_stream(seq[order[t % 6]], 2000, 0.01, 7, 3, (21, 21), 3, (3, 30, 0.01))
t0 = time.perf_counter()
for t in range(n):
    r = eng.lk_stream(seq[order[(3 + t) % 6]], 2000, 0.01, 7, 3, (21, 21), 3, (3, 30, 0.01))
dt = (time.perf_counter() - t0) / n * 1e3
print("lk_stream 1080p, 2000 corners: %.3f ms per frame host to host, %d tracked" % (dt, int(r[2].sum())))
eng.close()
