"""Single-pair latency (what one camera node sees per frame): python tools/latency_bench.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import opticalflowcontainer_b200 as ofb
from oracle import synth

for (W, H) in [(640, 480), (1920, 1080)]:
    eng = ofb.FlowEngine(W, H, 1, 0)
    t = synth.cheap_texture(H, W, 1)
    frames = [torch.from_numpy(synth.subpixel_shift(t, 1.3 * i, -0.7 * i)).pin_memory() for i in range(6)]
    out = torch.empty((1, H, W, 2), dtype=torch.float32).pin_memory()
    f = [x.numpy() for x in frames]

    def timeit(fn, reps=30):
        for i in range(3):
            fn(i)
        t0 = time.perf_counter()
        for i in range(reps):
            fn(i)
        return (time.perf_counter() - t0) / reps * 1e3

    pair_full = timeit(lambda i: eng.farneback_batch_into(f[i % 5][None], f[i % 5 + 1][None], out.numpy()))
    pair_stat = timeit(lambda i: eng.farneback_batch_stats(f[i % 5][None], f[i % 5 + 1][None]))
    eng.stream_reset(); eng.farneback_stream(f[0][None])
    order = [1, 2, 3, 4, 5, 4, 3, 2]
    strm_full = timeit(lambda i: eng.farneback_stream(f[order[i % 8]][None], out=out.numpy()))
    def s2(i):
        eng.farneback_stream(f[order[i % 8]][None], download=False)
        return eng.flow_u_stats(1)
    strm_stat = timeit(s2)
    d = torch.from_numpy(np.stack([f[0], f[1]])).cuda()
    fl = torch.empty((1, H, W, 2), dtype=torch.float32, device="cuda")
    def dev(i):
        eng.farneback_device(1, d.data_ptr(), d.data_ptr() + W * H, W, H, W, W * H, fl.data_ptr())
        eng.synchronize()
    dev_ms = timeit(dev)
    print("%dx%d one pair, ms per call:  device-resident %.3f | host pair call, field back %.3f | host pair call, scalars back %.3f | "
          "stream call, field back %.3f | stream call, scalars back %.3f" % (W, H, dev_ms, pair_full, pair_stat, strm_full, strm_stat))
    eng.close()
