"""Device-resident throughput for non-default parameter sets (1080p, 18 pairs): python tools/param_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import opticalflowcontainer_b200 as ofb
from oracle import synth

B, W, H = 18, 1920, 1080
eng = ofb.FlowEngine(W, H, B, 0)
t = synth.cheap_texture(H, W, 1)
fr = np.empty((2 * B, H, W), np.uint8)
for i in range(B):
    fr[i] = np.roll(t, (3 * i, 5 * i), axis=(0, 1))
    fr[B + i] = synth.subpixel_shift(fr[i], -3.3 + 1.37 * i, 2.6 + 0.71 * i)
d = torch.from_numpy(fr).cuda()
flow = torch.empty((B, H, W, 2), dtype=torch.float32, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream)
for kw in [dict(), dict(winsize=5), dict(winsize=9), dict(winsize=13), dict(winsize=17), dict(winsize=19), dict(winsize=21), dict(winsize=25), dict(winsize=31), dict(winsize=23), dict(poly_n=7, poly_sigma=1.5), dict(flags=256),
           dict(levels=5), dict(iterations=5), dict(pyr_scale=0.8, levels=3)]:
    p = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
    p.update(kw)
    for _ in range(2):
        eng.farneback_device(B, d.data_ptr(), d.data_ptr() + B * W * H, W, H, W, W * H, flow.data_ptr(), **p)
    eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    K = 5
    for _ in range(K):
        eng.farneback_device(B, d.data_ptr(), d.data_ptr() + B * W * H, W, H, W, W * H, flow.data_ptr(), **p)
    e1.record(stream)
    eng.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("%-40s %.3f ms / %d pairs -> %.0f pairs/s" % (kw or "defaults", ms, B, B / ms * 1e3))
