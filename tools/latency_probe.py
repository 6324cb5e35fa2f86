"""Single-pair latency of the host-buffer call (what a 30 Hz node sees) and of the device call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import opticalflowcontainer_b200 as ofb
from oracle import synth
for (W, H) in [(640, 480), (1920, 1080)]:
    eng = ofb.FlowEngine(W, H, 1, 0)
    t = synth.cheap_texture(H, W, 1)
    a = torch.from_numpy(t).pin_memory(); b = torch.from_numpy(synth.subpixel_shift(t, 2.3, -1.2)).pin_memory()
    out = torch.empty((H, W, 2), dtype=torch.float32).pin_memory()
    an, bn, on = a.numpy(), b.numpy(), out.numpy()
    for _ in range(5): eng.farneback(an, bn, on)
    K = 50
    t0 = time.perf_counter()
    for _ in range(K): eng.farneback(an, bn, on)
    host_ms = (time.perf_counter() - t0) / K * 1e3
    da, db = a.cuda(), b.cuda(); fl = torch.empty((1, H, W, 2), dtype=torch.float32, device="cuda")
    for _ in range(5): eng.farneback_device(1, da.data_ptr(), db.data_ptr(), W, H, W, W * H, fl.data_ptr())
    eng.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        eng.farneback_device(1, da.data_ptr(), db.data_ptr(), W, H, W, W * H, fl.data_ptr()); eng.synchronize()
    dev_ms = (time.perf_counter() - t0) / K * 1e3
    t0 = time.perf_counter()
    for _ in range(K):
        eng.farneback_device(1, da.data_ptr(), db.data_ptr(), W, H, W, W * H, fl.data_ptr())
    eng.synchronize()
    dev_async_ms = (time.perf_counter() - t0) / K * 1e3
    eng.timing_enable(True)
    eng.farneback_device(1, da.data_ptr(), db.data_ptr(), W, H, W, W * H, fl.data_ptr())
    st = eng.timing_read(); eng.timing_enable(False)
    print("%dx%d  host call %.3f ms | device call + sync %.3f ms | back-to-back %.3f ms | kernel time %.3f ms (%d launches)" % (
        W, H, host_ms, dev_ms, dev_async_ms, sum(v[0] for v in st.values()), sum(v[1] for v in st.values())))
    eng.close()
