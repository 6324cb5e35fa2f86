#!/usr/bin/env python
"""bench.py — Farneback 1080p frame-pairs/s on N B200s (BASELINE.json metric), one JSON line.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the cv2 CPU path on the box's host cores

A step = one pass of the hot path (all pyramid levels, all iterations) over one batch of
`--batch` synthetic 1920x1080 frame pairs per GPU.  `value` = pairs/s with inputs resident in
HBM; `e2e` = the same through the host-buffer C-ABI call (H2D of both frames from pinned memory
and D2H of the full flow field inside the timed region).  Frame pairs are independent, so ranks
shard them with no data-path collective (weak scaling); torch.distributed is used only for the
barrier and the max-over-ranks of the device time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "farneback_1080p_frame_pairs_per_s"
UNIT = "frame-pairs/s"
W_, H_ = 1920, 1080
PARAMS = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
L2_BYTES = 126 * 1024 * 1024


def level_pixels(w, h, pyr_scale=0.5, levels=3):
    out, scale = [], 1.0
    k = 0
    while k < levels:
        scale *= pyr_scale
        if w * scale < 32 or h * scale < 32:
            break
        k += 1
    for kk in range(k, -1, -1):
        s = pyr_scale ** kk
        out.append(int(np.rint(w * s)) * int(np.rint(h * s)))
    return out


def algorithmic_bytes_per_pair(w, h, iters=3):
    """SURVEY.md §8d fused-stage model: B = L*2N + (40 + 56*iters) * sum(n_l)."""
    nl = level_pixels(w, h)
    return len(nl) * 2 * w * h + (40 + 56 * iters) * sum(nl)


def bench_config():
    """The workload both arms run (same dict in the `ours` and the `reference` line)."""
    return {"workload": "farneback_%dx%d_independent_frame_pairs" % (W_, H_), "params": PARAMS}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "note": "no NVML samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """cv2.calcOpticalFlowFarneback (the reference implementation of the path) on the host cores,
    process-parallel (the wheel's Farneback is single-threaded).  Rank 0 only."""
    if rank != 0:
        return
    from oracle import cpu_bench
    import multiprocessing as mp
    workers = os.cpu_count() or 1
    pairs_per_worker = max(1, args.ref_pairs)
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        for _ in range(max(args.warmup, 1) if args.warmup else 0):
            cpu_bench.farneback_cpu_step(H_, W_, 1, workers, pool)
        t_total = 0.0
        for s in range(args.steps):
            t_total += cpu_bench.farneback_cpu_step(H_, W_, pairs_per_worker, workers, pool, seed=100 * s)
    pairs = workers * pairs_per_worker * args.steps
    value = pairs / t_total
    import cv2
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "run": {"pairs_per_step": workers * pairs_per_worker},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "reference",
                         "sample": "%d worker processes x %d pairs x %d steps of cv2 %s calcOpticalFlowFarneback "
                                   "(cv2.setNumThreads(1) per worker; the algorithm is single-threaded)"
                                   % (workers, pairs_per_worker, args.steps, cv2.__version__)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ sparse path (config 4)
def _lk_cpu_worker(args):
    """One camera stream on one host core: goodFeaturesToTrack + calcOpticalFlowPyrLK per frame."""
    seed, reps = args
    import cv2
    from oracle import synth
    cv2.setNumThreads(1)
    a, b = synth.synth_pair(H_LK, W_LK, seed, (3.3, -2.1))
    crit = (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 30, 0.01)

    def frame():
        p = cv2.goodFeaturesToTrack(a, 2000, 0.01, 7, blockSize=3)
        cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=(21, 21), maxLevel=3, criteria=crit)

    frame()                                   # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        frame()
        ts.append(time.perf_counter() - t0)
    return ts


W_LK, H_LK = 1920, 1080


def lk_cpu_baseline(reps=5):
    """cv2 on the host cores, 8 independent streams process-parallel (one stream per worker), >= 5 timed repetitions
    after a warm-up; aggregate frames/s from the slowest worker's median / best frame time."""
    import multiprocessing as mp
    import cv2
    workers = min(8, os.cpu_count() or 1)
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        per = pool.map(_lk_cpu_worker, [(300 + i, reps) for i in range(workers)])
    med = max(float(np.median(t)) for t in per)
    best = max(float(np.min(t)) for t in per)
    return {"value": workers / med, "unit": "frames/s", "cores": workers, "kind": "reference",
            "best_frames_per_s": workers / best, "ms_per_frame_one_core_median": med * 1e3,
            "sample": "cv2 %s goodFeaturesToTrack(2000, 0.01, 7, blockSize 3) + calcOpticalFlowPyrLK(21x21, maxLevel 3, "
                      "(30, 0.01)), %d streams process-parallel (cv2.setNumThreads(1) each), %d timed frames per stream "
                      "after 1 warm-up; value from the median frame time of the slowest worker" % (cv2.__version__, workers, reps)}


def around_record(local_rank, cpu=True):
    """The calls in front of the flow call that round 2 added, host buffer in -> host result out,
    with cv2 on one host core beside them and a bit-exactness check of each: JPEG CompressedImage ingest
    (opticalflow_comprerssed_node.py:43-46) and the junction detector (junction_detector.cpp:31-214)."""
    import cv2
    import opticalflowcontainer_b200 as ofb
    from oracle import synth
    cv2.setNumThreads(1)
    eng = ofb.FlowEngine(64, 64, 1, local_rank)

    def ms(fn, reps=8):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    rec = {}
    frame = synth.synth_net(1080, 1920, 3)
    jpg = cv2.imencode(".jpg", frame, [cv2.IMWRITE_JPEG_QUALITY, 90])[1]
    ref = cv2.cvtColor(cv2.imdecode(jpg, cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY)
    l0 = eng.launch_count
    got = eng.ingest_jpeg_gray(jpg)
    n_launch = int(eng.launch_count - l0)
    rec["jpeg_ingest_1080p"] = {
        "ms_per_frame": ms(lambda: eng.ingest_jpeg_gray(jpg)), "jpeg_bytes": int(jpg.size), "bit_exact_vs_cv2": bool(np.array_equal(got, ref)),
        "gpu_launches_per_frame": n_launch,
        "api": "ofb_ingest_jpeg_gray (markers + byte unstuffing on the host; Huffman decoding, IDCT, up-sampling, YCbCr->BGR->gray on the device)",
        "cv2_ms_per_frame": ms(lambda: cv2.cvtColor(cv2.imdecode(jpg, cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY)) if cpu else None}
    net = synth.synth_net(480, 640, 1)
    th = cv2.adaptiveThreshold(cv2.GaussianBlur(cv2.cvtColor(net, cv2.COLOR_BGR2GRAY), (3, 3), 0), 255,
                               cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)

    def cv_chain():
        g = cv2.cvtColor(net, cv2.COLOR_BGR2GRAY)
        t = cv2.adaptiveThreshold(cv2.GaussianBlur(g, (3, 3), 0), 255, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, 11, 2)
        cs, _ = cv2.findContours(t, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        return [(cv2.contourArea(c), cv2.boundingRect(c)) for c in cs]

    l0 = eng.launch_count
    pts = eng.find_junctions(net, 200, 2.0, 6)
    n_launch = int(eng.launch_count - l0)
    rec["junction_detector_640x480"] = {
        "ms_per_frame": ms(lambda: eng.find_junctions(net, 200, 2.0, 6)), "junctions": int(len(pts)),
        "threshold_bit_exact_vs_cv2": bool(np.array_equal(eng.junction_threshold(net), th)),
        "gpu_launches_per_frame": n_launch,
        "api": "ofb_find_junctions (pixel stages + contours by component labelling on the device; ordering + nanoflann-exact clustering on the host)",
        "cv2_ms_per_frame_without_clustering": ms(cv_chain) if cpu else None}
    eng.close()
    return rec


def lk_record(args, rank, local_rank, world, steps, warmup, cpu=True):
    """BASELINE.json config 4: Shi-Tomasi (2000 corners) + pyramidal LK on 1080p camera streams, 8 streams sharded over
    the ranks.  A step = one new frame of every stream of this rank through the camera-stream call of the sparse path
    (ofb_lk_stream: one upload per frame, the previous frame's pyramid and the corner list stay on the GPU).
    Returns the record on rank 0 (None elsewhere).  The process group, if any, is already up."""
    import torch
    import torch.distributed as dist
    import opticalflowcontainer_b200 as ofb
    from opticalflowcontainer_b200 import sharding
    from oracle import synth
    from concurrent.futures import ThreadPoolExecutor

    streams = sharding.shard_indices(8, rank, world)
    base = [synth.synth_pair(H_LK, W_LK, 300 + s, (0.0, 0.0))[0] for s in streams]
    seqs = [[synth.subpixel_shift(base[i], (1.3 + 0.2 * s) * t, (-0.9 + 0.1 * s) * t) for t in range(4)]
            for i, s in enumerate(streams)]
    # one handle (= one CUDA stream) and one host thread per camera stream, as a multi-camera node would run them:
    # handles are independent and ctypes releases the GIL during a call, so the latency-bound kernels of different
    # cameras overlap on the GPU
    engines = [ofb.FlowEngine(W_LK, H_LK, 1, local_rank) for _ in streams]
    pool = ThreadPoolExecutor(max_workers=max(1, len(streams)))
    order = [0, 1, 2, 3, 2, 1]

    def one(arg):
        i, t = arg
        r = engines[i].lk_stream(seqs[i][order[t % len(order)]], 2000, 0.01, 7, 3, (21, 21), 3, (3, 30, 0.01))
        return 0 if r is None else int(r[2].sum())

    def step(t):
        return sum(pool.map(one, [(i, t) for i in range(len(streams))]))

    tracked = 0
    for t in range(max(warmup, 2)):
        tracked = step(t)
    if world > 1:
        dist.barrier()
    # timed region: every camera thread runs its own `steps` frames at its own pace — cameras are not in lock step, and
    # the cv2 arm's 8 worker processes are not either (a barrier per frame makes every frame wait for the slowest of
    # the 8 calls: tools/lk_scaling.py measured 5 540 frames/s in lock step against 9 140 free-running)
    def camera(i):
        n = 0
        for t in range(steps):
            n = one((i, max(warmup, 2) + t))
        return n
    t0 = time.perf_counter()
    tracked = sum(pool.map(camera, range(len(streams))))
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    frames = torch.tensor([len(streams) * steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(frames, op=dist.ReduceOp.SUM)
    launches = int(sum(e.launch_count for e in engines))
    # single camera, one frame per call, host to host
    lat_ms = None
    if streams:
        e = engines[0]
        t0 = time.perf_counter()
        for t in range(20):
            e.lk_stream(seqs[0][order[t % len(order)]], 2000, 0.01, 7, 3, (21, 21), 3, (3, 30, 0.01))
        lat_ms = (time.perf_counter() - t0) / 20 * 1e3
    pool.shutdown()
    for e in engines:
        e.close()
    if rank != 0:
        return None
    value = float(frames.item()) / float(tt.item())
    rec = {"metric": "shi_tomasi_pyrlk_1080p_2000pt_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world,
           "steps": steps, "ms_per_step": float(tt.item()) / steps * 1e3, "scaling": "strong (8 streams over the ranks)",
           "workload": "goodFeaturesToTrack(2000, 0.01, 7) + calcOpticalFlowPyrLK(21x21, maxLevel 3, (30, 0.01)) on 1920x1080, "
                       "8 camera streams sharded over %d GPU(s), one host thread per camera, each running its frames back to back" % world,
           "api": "ofb_lk_stream: one frame up per call (pinned staging), corners of the previous frame tracked into the new "
                  "one, new corners detected for the next call; points, status and error back; one handle and host thread "
                  "per stream",
           "tracked_points_last_step_rank0": tracked, "single_camera_ms_per_frame_host_to_host": lat_ms,
           "gpu_launches": launches}
    if cpu:
        rec["cpu_baseline"] = lk_cpu_baseline()
    return rec


def run_lk(args, rank, local_rank, world):
    """--mode lk: the sparse-path record alone, as one JSON line."""
    import torch
    import torch.distributed as dist
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from opticalflowcontainer_b200 import build as ofb_build
    if rank == 0:
        ofb_build.build()
    if world > 1:
        dist.barrier()
    rec = lk_record(args, rank, local_rank, world, args.steps, args.warmup, cpu=not args.no_cpu_baseline)
    if rank == 0:
        rec.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "s32/f32", "data": "synthetic"})
        real_stdout.write(json.dumps(rec) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ tiled mode (config 5)
def tiled_record(args, rank, local_rank, world, size, steps, warmup, parity=True, whole=True):
    """One frame pair spatially tiled over the N ranks (BASELINE.json config 5: 7680x4320 over 8 B200).
    A step = one tiled pair.  Parity: EVERY rank's rows against cv2 on the same pair (rank 0 runs cv2, the field is
    broadcast, each rank compares the rows it owns).  Returns the record on rank 0."""
    import torch
    import torch.distributed as dist
    import opticalflowcontainer_b200 as ofb
    from opticalflowcontainer_b200 import tiled
    from oracle import synth

    W, H = {"8k": (7680, 4320), "4k": (3840, 2160), "1080p": (1920, 1080)}[size]
    eng = ofb.FlowEngine(W, H, 1, local_rank)
    tiled.setup_distributed(eng, rank, world)
    t = synth.cheap_texture(H, W, 400)
    n_sets = 3
    host_frames, frames = [], []
    for s in range(n_sets):
        a = np.roll(t, (17 * s, 29 * s), axis=(0, 1))
        b = synth.subpixel_shift(a, 9.5 - s, -4.25 + s)
        host_frames.append((a, b))
        frames.append((torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()))
    d_flow = torch.zeros((H, W, 2), dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(eng.stream)

    def step(i):
        a, b = frames[i % n_sets]
        return tiled.farneback_tiled_device(eng, a.data_ptr(), b.data_ptr(), W, H, W, d_flow.data_ptr(), **PARAMS)

    def barrier():
        torch.cuda.synchronize()
        eng.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rows = (0, 0)
    for i in range(warmup):
        rows = step(i)
    timed_out = tiled.tiled_status(eng)
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count
    ev0.record(stream)
    for i in range(steps):
        step(warmup + i)
    ev1.record(stream)
    timed_out = tiled.tiled_status(eng) or timed_out
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count - l0
    tms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    # per-stage event times of one more (untimed) step on this rank; "other" = the cross-GPU barriers + halo pulls
    eng.timing_enable(True)
    last = warmup + steps
    rows = step(last)
    stage = eng.timing_read()
    barrier()
    all_stage = [None] * world
    mine = {k: round(v[0], 3) for k, v in stage.items()}
    if world > 1:
        dist.all_gather_object(all_stage, mine)
    else:
        all_stage = [mine]
    eng.timing_enable(False)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    # ---- parity of the tiled field against cv2, every rank's rows
    par = None
    if parity:
        ref = torch.empty((H, W, 2), dtype=torch.float32, device="cuda")
        cv2_s = 0.0
        if rank == 0:
            import cv2
            a, b = host_frames[last % n_sets]
            t0 = time.perf_counter()
            f = cv2.calcOpticalFlowFarneback(a, b, None, PARAMS["pyr_scale"], PARAMS["levels"], PARAMS["winsize"],
                                             PARAMS["iterations"], PARAMS["poly_n"], PARAMS["poly_sigma"], PARAMS["flags"])
            cv2_s = time.perf_counter() - t0
            ref.copy_(torch.from_numpy(f))
        if world > 1:
            dist.broadcast(ref, 0)
        yb, ye = rows
        d = (ref[yb:ye] - d_flow[yb:ye]).double()
        epe = torch.sqrt((d * d).sum(-1))
        acc = torch.tensor([float(epe.sum().item()), float(epe.numel())], dtype=torch.float64, device="cuda")
        mx = torch.tensor([float(epe.max().item()) if epe.numel() else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        import cv2
        par = {"mean_epe": float(acc[0].item() / max(acc[1].item(), 1.0)), "max_epe": float(mx.item()),
               "pixels_compared": int(acc[1].item()), "vs": "cv2 %s calcOpticalFlowFarneback, same pair" % cv2.__version__,
               "rows": "every rank's own rows (all %d ranks)" % world, "cv2_seconds": cv2_s}
        del ref
    # single-GPU whole-frame time on rank 0 (same engine, untiled) for the speed-up
    whole_ms = None
    if whole and rank == 0:
        tmp = torch.zeros((H, W, 2), dtype=torch.float32, device="cuda")
        a, b = frames[last % n_sets]
        for _ in range(2):
            eng.farneback_device(1, a.data_ptr(), b.data_ptr(), W, H, W, W * H, tmp.data_ptr(), **PARAMS)
        eng.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            eng.farneback_device(1, a.data_ptr(), b.data_ptr(), W, H, W, W * H, tmp.data_ptr(), **PARAMS)
        e1.record(stream)
        eng.synchronize()
        whole_ms = e0.elapsed_time(e1) / 3
    if world > 1:
        dist.barrier()
    eng.close()
    if rank != 0:
        return None
    peaks, peak_src = measured_peaks()
    value = steps / (ms * 1e-3)
    bytes_pair = algorithmic_bytes_per_pair(W, H)
    return {"metric": "farneback_%s_tiled_frame_pairs_per_s" % size, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": steps, "ms_per_pair": ms / steps, "scaling": "strong",
            "workload": "farneback_%dx%d_spatially_tiled_x%d" % (W, H, world),
            "parallelism": "row strips x%d, halo rows pulled over NVLink peer pointers, neighbour flag barriers in peer "
                           "memory, no collective" % world,
            "roofline": {"bound": "hbm", "achieved": bytes_pair * value / 1e9, "peak": peaks["hbm_gbs"] * world,
                         "unit": "GB/s", "frac": bytes_pair * value / 1e9 / (peaks["hbm_gbs"] * world),
                         "peak_source": peak_src + " x n_gpus", "algorithmic_bytes_per_pair": bytes_pair},
            "stage_ms_per_rank": all_stage, "whole_frame_single_gpu_ms": whole_ms,
            "speedup_vs_whole_frame_single_gpu": (whole_ms / (ms / steps)) if whole_ms else None,
            "parity": par, "barrier_timed_out": bool(timed_out), "gpu_launches": int(launches), "clocks": clocks}


def run_tiled(args, rank, local_rank, world):
    """--mode tiled: the tiled record alone, as one JSON line."""
    import torch
    import torch.distributed as dist
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from opticalflowcontainer_b200 import build as ofb_build
    if rank == 0:
        ofb_build.build()
    if world > 1:
        dist.barrier()
    rec = tiled_record(args, rank, local_rank, world, args.tile_size, args.steps, args.warmup,
                       parity=not args.no_tiled_check, whole=not args.no_tiled_check)
    if rank == 0:
        rec.update({"warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic"})
        real_stdout.write(json.dumps(rec) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ extras of the main line
def batched_4k_record(rank, local_rank, world, B=18, steps=6, warmup=3):
    """BASELINE.json config 3: Farneback on 3840x2160 frame pairs, B independent pairs per step per GPU, device-resident,
    sharded over the ranks without a collective (weak scaling, like the headline).  One pair of the last step is checked
    against cv2 on rank 0 (about 1.6 s of cv2).  Returns the record on rank 0."""
    import torch
    import torch.distributed as dist
    import opticalflowcontainer_b200 as ofb
    from oracle import synth
    W, H = 3840, 2160
    eng = ofb.FlowEngine(W, H, B, local_rank)
    t = synth.cheap_texture(H, W, 700 + rank)
    n_sets = 2                                   # 2 x 199 MB of frames > L2; the step's working set is several GB anyway
    host = []
    for s in range(n_sets):
        fr = np.empty((2 * B, H, W), np.uint8)
        for i in range(B):
            fr[i] = np.roll(t, (5 * i + 3 * s, 9 * i + 7 * s), axis=(0, 1))
        for i in range(B):
            fr[B + i] = np.roll(fr[i], (1 + (i + s) % 5, -2 - (i % 3)), axis=(0, 1)) if i else synth.subpixel_shift(fr[0], 6.3 - s, -3.6)
        host.append(fr)
    dev = [torch.from_numpy(f).cuda() for f in host]
    d_flow = torch.empty((B, H, W, 2), dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(eng.stream)

    def step(i):
        p0 = dev[i % n_sets].data_ptr()
        eng.farneback_device(B, p0, p0 + B * W * H, W, H, W, W * H, d_flow.data_ptr(), **PARAMS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(steps):
        step(warmup + i)
    ev1.record(stream)
    barrier()
    tms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    rec = None
    if rank == 0:
        ms = float(tms.item()) / steps
        value = B * world / ms * 1e3
        peaks, _ = measured_peaks()
        frac = None
        if peaks.get("hbm_gbs"):
            frac = algorithmic_bytes_per_pair(W, H) * value / world / 1e9 / peaks["hbm_gbs"]
        rec = {"metric": "farneback_4k_frame_pairs_per_s", "value": value, "unit": "frame-pairs/s", "n_gpus": world,
               "steps": steps, "warmup": warmup, "ms_per_step": ms, "pairs_per_step_per_gpu": B, "scaling": "weak",
               "workload": "farneback_3840x2160_independent_frame_pairs, device-resident, default node params",
               "pipeline_roofline_frac": frac}
        try:
            import cv2
            last = (warmup + steps - 1) % n_sets
            a, b = host[last][0], host[last][B]
            t0 = time.perf_counter()
            ref = cv2.calcOpticalFlowFarneback(a, b, None, PARAMS["pyr_scale"], PARAMS["levels"], PARAMS["winsize"],
                                               PARAMS["iterations"], PARAMS["poly_n"], PARAMS["poly_sigma"], PARAMS["flags"])
            cv_s = time.perf_counter() - t0
            got = d_flow[0].cpu().numpy()
            epe = np.sqrt(((got.astype(np.float64) - ref) ** 2).sum(-1))
            rec["parity"] = {"mean_epe": float(epe.mean()), "max_epe": float(epe.max()), "vs": "cv2 %s, pair 0 of the last timed step" % cv2.__version__,
                             "cv2_seconds": cv_s, "ok": bool(epe.mean() <= 0.01 and epe.max() <= 0.1)}
        except ImportError:
            pass
    eng.close()
    del dev, d_flow
    torch.cuda.empty_cache()
    return rec


def e2e_copy_ceiling(torch, dist, world, h2d_bytes, d2h_bytes, steps=10):
    """Bare pinned-memory copies of one step's traffic (H2D of the frames, D2H of the fields, on two streams, all ranks
    at once): the pairs/s the host link allows with NO kernel at all — the ceiling of the full-field e2e number."""
    up = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    down = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    d_down = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def one():
        with torch.cuda.stream(s1):
            d_up.copy_(up, non_blocking=True)
        with torch.cuda.stream(s2):
            down.copy_(d_down, non_blocking=True)

    one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item()) / steps


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    # keep stdout clean for the ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from opticalflowcontainer_b200 import build as ofb_build
    if rank == 0:
        ofb_build.build()
    if world > 1:
        dist.barrier()
    import opticalflowcontainer_b200 as ofb
    from oracle import synth  # input generator only (test infrastructure), never on the timed path

    B = args.batch
    seq = args.mode == "sequence"
    strm = args.mode == "stream"
    eng = ofb.FlowEngine(W_, H_, B, local_rank)
    pitch = W_
    istride = pitch * H_
    frames_per_set = (B + 1) if seq else (B if strm else 2 * B)
    n_sets = max(2, int(np.ceil(2.0 * L2_BYTES / (frames_per_set * istride))) + 1)
    # synthetic frames: a few distinct textures, shifted copies as "next"
    # (corpus C2: a panning texture, per-frame shift ~U(-8, 8) px — real-valued, so the
    # flow is sub-pixel as it is for a real camera)
    base = [synth.cheap_texture(H_, W_, 1000 * rank + i) for i in range(4)]
    rng = np.random.default_rng(rank)
    host_sets = []
    for s in range(n_sets):
        fr = np.empty((frames_per_set, H_, W_), np.uint8)
        if strm:
            # set s = frame s of each of the B streams (stream i pans its own texture by a fixed sub-pixel velocity)
            for i in range(B):
                vx, vy = 0.9 + 0.37 * (i % 7), -0.6 + 0.29 * (i % 5)
                fr[i] = synth.subpixel_shift(base[i % 4], vx * s, vy * s)
        elif seq:
            t = base[s % 4]
            ox = oy = 0.0
            for i in range(frames_per_set):
                fr[i] = synth.subpixel_shift(t, ox, oy)
                ox += float(rng.uniform(-8, 8)); oy += float(rng.uniform(-8, 8))
        else:
            for i in range(B):
                t = base[(s + i) % 4]
                fr[i] = np.roll(t, (i * 7 % 13, i * 5 % 11), axis=(0, 1))
                fr[B + i] = synth.subpixel_shift(fr[i], float(rng.uniform(-8, 8)), float(rng.uniform(-8, 8)))
        host_sets.append(fr)
    dev_sets = [torch.from_numpy(fr).cuda() for fr in host_sets]
    d_flow = torch.empty((B, H_, W_, 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    # stream mode walks the time axis back and forth so that consecutive calls always see consecutive frames
    order = list(range(n_sets)) + list(range(n_sets - 2, 0, -1))

    def step(i):
        if strm:
            d = dev_sets[order[i % len(order)]]
            eng.farneback_stream_device(B, d.data_ptr(), W_, H_, pitch, istride, d_flow.data_ptr(), **PARAMS)
            return
        d = dev_sets[i % n_sets]
        p0 = d.data_ptr()
        eng.farneback_device(B, p0, p0 + B * istride, W_, H_, pitch, istride, d_flow.data_ptr(), sequence=seq, **PARAMS)

    if strm:
        eng.farneback_stream_device(B, dev_sets[1].data_ptr(), W_, H_, pitch, istride, d_flow.data_ptr(), **PARAMS)  # prime

    stream = torch.cuda.ExternalStream(eng.stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    # ---- timed region: exactly K steps, CUDA events on the engine's stream
    sampler = ClockSampler(local_rank)
    l0 = eng.launch_count
    eng.timing_enable(True)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    sampler.start()
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record(stream)
    eng.synchronize()
    barrier()
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    stage = eng.timing_read()
    iter_samples = eng.timing_samples("iteration")
    eng.timing_enable(False)
    launches = eng.launch_count - l0
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())

    # the stage timers add event records between kernels; re-time the K steps without them for `value`
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + args.steps + i)
    ev1.record(stream)
    eng.synchronize()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    plain_ms_max = float(t.item())
    pairs = B * args.steps * world
    value = pairs / (plain_ms_max * 1e-3)

    # ---- e2e: host-buffer C-ABI call, pinned host memory, H2D + D2H inside the timed region
    pin_prev = [torch.from_numpy(host_sets[s][:B].copy()).pin_memory() for s in range(0 if strm else min(n_sets, 3))]
    pin_next = [torch.from_numpy((host_sets[s][1:B + 1] if seq else host_sets[s][B:2 * B]).copy()).pin_memory()
                for s in range(0 if strm else min(n_sets, 3))]
    # two result buffers: the asynchronous batch call pipelines across calls (upload + kernels of step
    # i+1 overlap the download of step i); a step's result is complete before its buffer is reused
    # (the library orders that) and everything is complete at eng.wait() inside the timed region
    pin_flow = [torch.empty((B, H_, W_, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_steps = args.steps if args.e2e_steps <= 0 else args.e2e_steps

    pin_frames = [torch.from_numpy(fr).pin_memory() for fr in host_sets] if strm else []

    def e2e_step(i):
        if strm:   # one new frame per stream up, the field of (previous, new) down; synchronous per call
            eng.farneback_stream(pin_frames[order[i % len(order)]].numpy(), out=pin_flow[i % 2].numpy(), **PARAMS)
            return
        s = i % len(pin_prev)
        eng.farneback_batch_into(pin_prev[s].numpy(), pin_next[s].numpy(), pin_flow[i % 2].numpy(),
                                 wait=args.e2e_sync, **PARAMS)

    for i in range(min(2, args.warmup)):
        e2e_step(i)
    eng.wait()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    eng.wait()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * e2e_steps * world / float(t.item())

    # ---- e2e through the node contract: frames up, flow on the device, median/mean of u back (what every
    # node of the reference does right after the flow call) — the field never crosses PCIe
    def node_step(i):
        if strm:
            eng.farneback_stream(pin_frames[order[i % len(order)]].numpy(), download=False, **PARAMS)
            return eng.flow_u_stats(B, wait=args.e2e_sync or i < 0)
        s_ = i % len(pin_prev)
        return eng.farneback_batch_stats(pin_prev[s_].numpy(), pin_next[s_].numpy(), wait=args.e2e_sync or i < 0, **PARAMS)

    for i in range(2):
        node_step(i - 2)
    barrier()
    t0 = time.perf_counter()
    node_out = []
    for i in range(e2e_steps):
        node_out.append(node_step(i))
    eng.wait()                                  # every step's mean/median of u is on the host here
    node_s = time.perf_counter() - t0
    t = torch.tensor([node_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    node_value = B * e2e_steps * world / float(t.item())

    # ---- extras (all ranks take part; rank 0 reports) ---------------------------------------------------------
    extras = {}
    if not args.no_extras and args.mode == "pairs":
        # (1) parity of the timed workload: pair 0 of the last timed step's frame set against cv2 on this host
        last = (args.warmup + 2 * args.steps - 1) % n_sets
        if rank == 0:
            import cv2
            eng.farneback_device(B, dev_sets[last].data_ptr(), dev_sets[last].data_ptr() + B * istride, W_, H_, pitch, istride,
                                 d_flow.data_ptr(), **PARAMS)
            eng.synchronize()
            got = d_flow[0].cpu().numpy().astype(np.float64)
            t0 = time.perf_counter()
            ref = cv2.calcOpticalFlowFarneback(host_sets[last][0], host_sets[last][B], None, PARAMS["pyr_scale"],
                                               PARAMS["levels"], PARAMS["winsize"], PARAMS["iterations"], PARAMS["poly_n"],
                                               PARAMS["poly_sigma"], PARAMS["flags"])
            cv2_s = time.perf_counter() - t0
            epe = np.sqrt(((got - ref) ** 2).sum(-1))
            extras["parity"] = {"mean_epe": float(epe.mean()), "max_epe": float(epe.max()),
                                "vs": "cv2 %s calcOpticalFlowFarneback on the same pair (pair 0 of the last timed step, "
                                      "computed by the same %d-pair launch sequence)" % (cv2.__version__, B),
                                "gate": "mean <= 0.01 px, max <= 0.1 px", "cv2_seconds": cv2_s,
                                "ok": bool(epe.mean() <= 0.01 and epe.max() <= 0.1)}
        # (2) one pair per call: what a single camera node sees (BASELINE.json configs[1] read literally)
        if rank == 0:
            one = dev_sets[0]
            for _ in range(5):
                eng.farneback_device(1, one.data_ptr(), one.data_ptr() + B * istride, W_, H_, pitch, istride, d_flow.data_ptr(), **PARAMS)
            eng.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 50
            e0.record(stream)
            for i in range(reps):
                d = dev_sets[i % n_sets]
                eng.farneback_device(1, d.data_ptr(), d.data_ptr() + B * istride, W_, H_, pitch, istride, d_flow.data_ptr(), **PARAMS)
            e1.record(stream)
            eng.synchronize()
            dev1 = e0.elapsed_time(e1) / reps
            p1, n1, f1 = pin_prev[0][0].numpy(), pin_next[0][0].numpy(), pin_flow[0][0].numpy()
            for _ in range(3):
                eng.farneback(p1, n1, f1, **PARAMS)
            t0 = time.perf_counter()
            for _ in range(reps):
                eng.farneback(p1, n1, f1, **PARAMS)
            host1 = (time.perf_counter() - t0) / reps * 1e3
            extras["batch1"] = {"device_resident_ms_per_pair": dev1, "device_resident_pairs_per_s": 1e3 / dev1,
                                "host_to_host_full_field_ms_per_pair": host1, "host_to_host_pairs_per_s": 1e3 / host1,
                                "calls": reps, "api": "ofb_farneback_device / ofb_farneback (one pair per call, CUDA-graph replay)"}
        # (3) bare-copy ceiling of the full-field e2e number, all ranks at once
        ceil_s = e2e_copy_ceiling(torch, dist, world, 2 * B * W_ * H_, 8 * B * W_ * H_)
        extras["e2e_ceiling"] = {"value": B * world / ceil_s, "unit": UNIT, "seconds_per_step": ceil_s,
                                 "what": "pinned cudaMemcpyAsync of one step's frames up (%d B) and fields down (%d B) on two "
                                         "streams, all %d rank(s) concurrently, no kernels: the pairs/s the host link allows "
                                         "for the full-field path" % (2 * B * W_ * H_, 8 * B * W_ * H_, world),
                                 "aggregate_d2h_gbs": 8 * B * W_ * H_ * world / ceil_s / 1e9,
                                 "aggregate_h2d_gbs": 2 * B * W_ * H_ * world / ceil_s / 1e9}
    eng.close()
    del dev_sets, d_flow
    torch.cuda.empty_cache()
    if not args.no_extras and args.mode == "pairs" and args.frame == "1080p":
        # (4) BASELINE.json config 4: sparse path, 8 camera streams sharded over the ranks
        rec = lk_record(args, rank, local_rank, world, steps=30, warmup=3, cpu=(world == 1 and not args.no_cpu_baseline))
        if rank == 0:
            extras["lk_8_streams"] = rec
        if rank == 0:
            try:
                extras["around_the_path"] = around_record(local_rank, cpu=not args.no_cpu_baseline)
            except Exception as e:     # (never lose the headline line to a side record)
                extras["around_the_path"] = {"error": repr(e)}
        # (3) BASELINE.json config 3: 3840x2160 frame pairs, batched, sharded over the ranks like the 1080p pairs
        try:
            rec = batched_4k_record(rank, local_rank, world)
        except Exception as e:         # (never lose the headline line to a side record)
            rec = {"error": repr(e)}
        if rank == 0:
            extras["batched_4k"] = rec
        # (5) BASELINE.json config 5: one 8K pair tiled over the ranks (needs >= 2 GPUs)
        if world >= 2:
            rec = tiled_record(args, rank, local_rank, world, "8k", steps=20, warmup=3, parity=True, whole=True)
            if rank == 0:
                extras["tiled_8k"] = rec

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    nl = level_pixels(W_, H_)
    it_ms, it_cnt = stage["iteration"]
    iters = PARAMS["iterations"]
    # dominant kernel = the level-0 (1920x1080) launch of the fused iteration kernel k_iter_v: the last
    # `iters` iteration launches of every call.  achieved = algorithmic bytes of one such launch
    # (56 B x N x B pairs) / its average CUDA-event duration in the timed region.
    per_call = len(nl) * iters
    l0 = [t for i, t in enumerate(iter_samples) if i % per_call >= per_call - iters]
    l0_ms = float(np.mean(l0)) if l0 else 0.0
    alg_launch_bytes = 56.0 * nl[-1] * B
    it_gbs = alg_launch_bytes / (l0_ms * 1e-3) / 1e9 if l0_ms > 0 else 0.0
    stage_gbs = 56.0 * sum(nl) * iters * B * args.steps / (it_ms * 1e-3) / 1e9 if it_ms > 0 else 0.0
    pipe_gbs = algorithmic_bytes_per_pair(W_, H_) * (value / world) / 1e9
    # DRAM bytes per launch from the committed ncu capture of the same kernel (a 1080p level-0 launch), scaled by the
    # pixels of this run's launch; None for other frame sizes (the halo / L2 behaviour is not the same there)
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r2_iter_v_ncu.json")))
        if prof.get("pairs_per_launch") and (W_, H_) == (prof.get("width", 1920), prof.get("height", 1080)):
            traffic = prof["dram_bytes_per_launch"] * B / prof["pairs_per_launch"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": plain_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "run": {"pairs_per_step_per_gpu": B, "mode": args.mode,
                "parallelism": "frame-pair sharding x%d, no collective" % world,
                "l2": "inputs rotate over %d frame sets (%.0f MB > 2x L2); per-step working set %.0f MB >> L2"
                      % (n_sets, n_sets * frames_per_set * istride / 1e6, B * 232.0 * W_ * H_ / 2073600.0),
                "note": "a step = ONE launch sequence over %d independent pairs (36 pairs x 8 strips = 288 of the 296 "
                        "resident CTA slots of the iteration kernel at level 0, one row segment each; 18 pairs per step "
                        "run 2-3 %% slower, 54 no faster); one pair per call is reported under batch1" % B},
        "roofline": {"bound": "hbm", "kernel": "k_iter_v, level-0 launch (UpdateMatrices + box blur + 2x2 solve fused)",
                     "achieved": it_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": it_gbs / peaks["hbm_gbs"],
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_launch_bytes,
                     "algorithmic_bytes_model": "56 B x %d px x %d pairs" % (nl[-1], B),
                     "launch_ms": l0_ms, "launches_timed": len(l0),
                     "level0_launch_share_of_step": (sum(l0) / dev_ms) if dev_ms > 0 else None,
                     "iteration_stage_all_levels_gbs": stage_gbs,
                     "share_of_step": it_ms / dev_ms if dev_ms > 0 else None,
                     "pipeline_achieved": pipe_gbs, "pipeline_frac": pipe_gbs / peaks["hbm_gbs"],
                     "pipeline_bytes_per_pair": algorithmic_bytes_per_pair(W_, H_),
                     "stage_ms": {k: v[0] for k, v in stage.items()}, "stage_launch_groups": {k: v[1] for k, v in stage.items()},
                     "timed_region_ms_with_stage_events": dev_ms_max},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (1 if strm else 2) * B * W_ * H_,
                "d2h_bytes_per_step": 8 * B * W_ * H_, "steps": e2e_steps,
                "api": ("ofb_farneback_stream (one new frame per stream per call, synchronous)" if strm else
                        "ofb_farneback_batch" if args.e2e_sync else "ofb_farneback_batch_async + ofb_wait") +
                       " (host buffers, pinned; full float32 [H,W,2] flow of every pair returned to the host)"},
        "e2e_node_contract": {"value": node_value, "unit": UNIT, "h2d_bytes_per_step": (1 if strm else 2) * B * W_ * H_,
                              "d2h_bytes_per_step": 12 * B, "steps": e2e_steps,
                              "api": ("ofb_farneback_stream(flow=NULL) + ofb_flow_u_stats_async + ofb_wait" if strm else
                                      "ofb_farneback_batch_stats" if args.e2e_sync else "ofb_farneback_batch_stats_async + ofb_wait") +
                                     " (host frames in, on-device mean + exact median of u out: "
                                     "the reduction every node applies, lfn3_sub_node.py:205-212)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_bench
        cb = cpu_bench.farneback_cpu_throughput(H_, W_, target_seconds=args.cpu_seconds)
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": "reference",
                                "sample": "cv2 %s calcOpticalFlowFarneback, %d processes x %d pairs of %dx%d "
                                          "(single-threaded algorithm; 1 pair = %.0f ms on one core)"
                                          % (cb["cv2_version"], cb["cores"], cb["pairs_per_worker"], W_, H_,
                                             cb["single_pair_ms"]),
                                "per_core_pairs_per_s": cb["per_core_pairs_per_s"],
                                "single_process": cb["single_process"]}
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # 18 pairs: the level-0 launch of the iteration kernel is 8 strips x 2 row segments x 18 pairs = 288 CTAs on
    # the 296 resident CTA slots of a B200 (16 pairs leave 40 slots, i.e. 40 SMs half empty)
    ap.add_argument("--batch", type=int, default=36, help="frame pairs per step per GPU")
    ap.add_argument("--frame", default="1080p", choices=["vga", "1080p", "4k"],
                    help="frame size of the pairs/sequence modes (1080p = the BASELINE.json metric; vga = config 0, 4k = config 3)")
    ap.add_argument("--mode", default="pairs", choices=["pairs", "sequence", "stream", "tiled", "lk"],
                    help="pairs: independent pairs (the headline); sequence: B+1 consecutive frames of one stream per call; "
                         "stream: one new frame of each of B camera streams per call, temporal state kept on the GPU")
    ap.add_argument("--tile-size", default="8k", choices=["8k", "4k", "1080p"], help="--mode tiled: frame size")
    ap.add_argument("--no-tiled-check", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--e2e-sync", action="store_true", help="e2e through the synchronous batch call (no cross-call overlap)")
    ap.add_argument("--ref-pairs", type=int, default=2, help="--impl reference: pairs per worker per step")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records of the main line (parity vs cv2, batch-1 latency, copy ceiling, LK streams, tiled 8K)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    global W_, H_, METRIC
    W_, H_ = {"vga": (640, 480), "1080p": (1920, 1080), "4k": (3840, 2160)}[args.frame]
    if args.frame != "1080p":
        METRIC = "farneback_%s_frame_pairs_per_s" % args.frame
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.mode == "tiled":
        run_tiled(args, rank, local_rank, world)
    elif args.mode == "lk":
        run_lk(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
