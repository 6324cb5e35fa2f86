"""Does running two half-batches on two handles (two compute streams) beat one launch sequence over the whole batch?
python tools/probes/split_batch_probe.py [pairs]   (device-resident 1080p pairs, CUDA events around K steps)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import opticalflowcontainer_b200 as ofb
from oracle import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 18
W, H = 1920, 1080
P = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
base = [synth.cheap_texture(H, W, i) for i in range(4)]
n_sets = 5
sets = []
for s in range(n_sets):
    fr = np.stack([np.roll(base[(s + i) % 4], (3 * i + s, 5 * i), (0, 1)) for i in range(B)] +
                  [synth.subpixel_shift(np.roll(base[(s + i) % 4], (3 * i + s, 5 * i), (0, 1)), 2.5, -1.5) for i in range(B)])
    sets.append(torch.from_numpy(fr).cuda())
istride = W * H


def run(splits, steps=20, warm=4):
    engs = [ofb.FlowEngine(W, H, B // splits, 0) for _ in range(splits)]
    flows = [torch.zeros((B // splits, H, W, 2), dtype=torch.float32, device="cuda") for _ in range(splits)]
    n = B // splits

    def step(i):
        d = sets[i % n_sets]
        for k, e in enumerate(engs):
            e.farneback_device(n, d.data_ptr() + k * n * istride, d.data_ptr() + (B + k * n) * istride, W, H, W, istride,
                               flows[k].data_ptr(), **P)
    for i in range(warm):
        step(i)
    for e in engs:
        e.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        step(warm + i)
    for e in engs:
        e.synchronize()
    dt = time.perf_counter() - t0
    for e in engs:
        e.close()
    return dt / steps * 1e3


for splits in (1, 2, 3, 1, 2):
    if B % splits:
        continue
    ms = run(splits)
    print("%d x %d pairs: %.3f ms per %d pairs = %.0f pairs/s" % (splits, B // splits, ms, B, B / ms * 1e3))
