"""Camera streams on ONE GPU: frames/s of n = 1, 2, 4, 8 streams, (a) in lock step (a barrier per frame, as bench.py's
lk record runs them) and (b) free-running threads.  python tools/lk_scaling.py"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
import opticalflowcontainer_b200 as ofb
from oracle import synth

W, H = 1920, 1080
order = [0, 1, 2, 3, 2, 1]
seqs = []
for s in range(8):
    base = synth.synth_pair(H, W, 300 + s, (0.0, 0.0))[0]
    seqs.append([synth.subpixel_shift(base, (1.3 + 0.2 * s) * t, (-0.9 + 0.1 * s) * t) for t in range(4)])
engines = [ofb.FlowEngine(W, H, 1, 0) for _ in range(8)]

def one(i, t):
    engines[i].lk_stream(seqs[i][order[t % 6]], 2000, 0.01, 7, 3, (21, 21), 3, (3, 30, 0.01))

for n in (1, 2, 4, 8):
    pool = ThreadPoolExecutor(max_workers=n)
    for t in range(4):
        list(pool.map(lambda i: one(i, t), range(n)))
    K = 40
    t0 = time.perf_counter()
    for t in range(K):
        list(pool.map(lambda i: one(i, 4 + t), range(n)))
    lock = n * K / (time.perf_counter() - t0)
    def run(i):
        for t in range(K):
            one(i, 4 + t)
    th = [threading.Thread(target=run, args=(i,)) for i in range(n)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    free = n * K / (time.perf_counter() - t0)
    print("%d streams: lock step %.0f frames/s, free-running %.0f frames/s" % (n, lock, free))
    pool.shutdown()
