"""Small fixed workload for ncu: N calls of the device-resident Farneback path (1080p, batch B).
Usage: python tools/profile_run.py [calls] [batch] [w] [h]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import opticalflowcontainer_b200 as ofb
from oracle import synth

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
H = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
eng = ofb.FlowEngine(W, H, B, 0)
t = synth.cheap_texture(H, W, 1)
fr = np.empty((2 * B, H, W), np.uint8)
for i in range(B):
    fr[i] = np.roll(t, (3 * i, 5 * i), axis=(0, 1))
    if os.environ.get("OFB_SHIFT", "frac") == "int":
        fr[B + i] = np.roll(fr[i], (2 + i, -3 + i), axis=(0, 1))
    else:   # sub-pixel pan, like corpus C2
        fr[B + i] = synth.subpixel_shift(fr[i], -3.3 + 1.37 * i, 2.6 + 0.71 * i)
d = torch.from_numpy(fr).cuda()
flow = torch.empty((B, H, W, 2), dtype=torch.float32, device="cuda")
for _ in range(calls):
    eng.farneback_device(B, d.data_ptr(), d.data_ptr() + B * W * H, W, H, W, W * H, flow.data_ptr())
eng.synchronize()
print("ok", float(flow[0, H // 2, W // 2, 0]), eng.launch_count)
if os.environ.get("OFB_STAGES", "1") == "1":
    eng.timing_enable(True)
    K = 5
    for _ in range(K):
        eng.farneback_device(B, d.data_ptr(), d.data_ptr() + B * W * H, W, H, W, W * H, flow.data_ptr())
    st = eng.timing_read()
    eng.timing_enable(False)
    tot = sum(v[0] for v in st.values())
    print("per-call ms (batch %d): " % B + ", ".join("%s %.3f" % (k, v[0] / K) for k, v in st.items()) + " | total %.3f -> %.0f pairs/s" % (tot / K, B * K / tot * 1e3))
