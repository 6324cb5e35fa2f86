"""CPU baseline: the reference implementation of the path (cv2 wheel) timed on host cores.

TEST INFRASTRUCTURE / BASELINE ONLY — used by ``bench.py`` (``cpu_baseline`` block and
``--impl reference``).  Farneback in this wheel is single-threaded, so the fair
multi-core figure is process-parallel: one frame pair per worker, ``workers`` processes.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def _texture_pair(h, w, seed):
    from oracle import synth
    a = synth.cheap_texture(h, w, seed)
    b = np.roll(a, (2, 3), axis=(0, 1))
    return a, b


def _worker(args):
    h, w, seed, n_pairs, params = args
    import cv2
    cv2.setNumThreads(1)
    a, b = _texture_pair(h, w, seed)
    cv2.calcOpticalFlowFarneback(a, b, None, *params)  # warm-up (page-in, allocator)
    t0 = time.perf_counter()
    for _ in range(n_pairs):
        cv2.calcOpticalFlowFarneback(a, b, None, *params)
    return time.perf_counter() - t0


PARAMS = (0.5, 3, 15, 3, 5, 1.2, 0)


def farneback_cpu_throughput(h, w, target_seconds=12.0, workers=None, params=PARAMS):
    """Returns dict(value=pairs/s aggregate, cores, pairs_per_worker, seconds, single_pair_ms,
    cv2_version, cv2_threads)."""
    import cv2
    workers = workers or os.cpu_count() or 1
    a, b = _texture_pair(h, w, 0)
    cv2.calcOpticalFlowFarneback(a, b, None, *params)
    t0 = time.perf_counter()
    cv2.calcOpticalFlowFarneback(a, b, None, *params)
    t1 = time.perf_counter() - t0
    # one process, cv2.setNumThreads(1) and the default thread count, 5 timed repetitions after the warm-up,
    # best and median (Farneback is single-threaded in this wheel: the two agree)
    single = {}
    cv2.setNumThreads(-1)                       # back to the build's default, whatever the caller had set
    default_threads = cv2.getNumThreads()
    for label, nt in (("threads_1", 1), ("threads_default", default_threads)):
        cv2.setNumThreads(nt)
        cv2.calcOpticalFlowFarneback(a, b, None, *params)
        reps = []
        for _ in range(5):
            t0 = time.perf_counter()
            cv2.calcOpticalFlowFarneback(a, b, None, *params)
            reps.append((time.perf_counter() - t0) * 1e3)
        single[label] = dict(threads=nt, best_ms=min(reps), median_ms=float(np.median(reps)), reps=5)
    cv2.setNumThreads(default_threads)
    n_pairs = int(max(1, min(32, round(target_seconds / max(t1, 1e-3)))))
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        t0 = time.perf_counter()
        pool.map(_worker, [(h, w, 1000 + i, n_pairs, params) for i in range(workers)])
        wall = time.perf_counter() - t0
    # wall includes interpreter start + warm-up of each worker; use the slowest worker's own timer
    with ctx.Pool(workers) as pool:
        times = pool.map(_worker, [(h, w, 2000 + i, n_pairs, params) for i in range(workers)])
    busy = max(times)
    return dict(value=workers * n_pairs / busy, cores=workers, pairs_per_worker=n_pairs, seconds=busy,
                single_pair_ms=t1 * 1e3, cv2_version=cv2.__version__, cv2_threads=1, wall_first_pool=wall,
                single_process=single, per_core_pairs_per_s=n_pairs / busy)


def farneback_cpu_step(h, w, pairs_per_worker, workers, pool, params=PARAMS, seed=0):
    """One bounded 'step' for --impl reference: every worker runs `pairs_per_worker` pairs."""
    times = pool.map(_worker, [(h, w, seed + i, pairs_per_worker, params) for i in range(workers)])
    return max(times)
