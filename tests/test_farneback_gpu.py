"""GPU parity of the dense path, called through the C ABI (ctypes → libofb.so), against
(i) the reference implementation of the path (cv2 wheel), (ii) the NumPy restatement and
(iii) the committed golden fixtures.

Tolerance (BASELINE.json north_star): mean endpoint error <= 0.01 px and max <= 0.1 px vs cv2 on
identical inputs.  The tests use a 10x tighter gate (1e-3 / 1e-2) on well-conditioned inputs.
"""
import glob
import os

import numpy as np
import pytest

from oracle import cv2_oracle as C
from oracle import farneback_np as F
from oracle import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

MEAN_GATE, MAX_GATE = 0.01, 0.1          # the contract
TIGHT_MEAN, TIGHT_MAX = 1e-3, 1e-2       # what we hold ourselves to


def _check(ref, got, mean_gate=TIGHT_MEAN, max_gate=TIGHT_MAX):
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.isfinite(got).all()
    mean, mx = C.epe(ref, got)
    assert mean <= mean_gate and mx <= max_gate, (mean, mx)
    return mean, mx


CASES = [
    ("vga_small_shift", (480, 640), 0, (1.7, -0.9), {}),
    ("vga_large_shift", (480, 640), 1, (14.3, -9.6), {}),
    ("odd_size", (481, 637), 2, (3.1, 2.2), {}),
    ("odd_size2", (539, 959), 2, (-4.1, 1.2), {}),
    ("gaussian", (240, 320), 3, (3, 2), dict(flags=256)),
    ("gaussian31", (240, 320), 3, (3, 2), dict(flags=256, winsize=31)),
    ("winsize16", (240, 320), 4, (3, 2), dict(winsize=16)),
    ("winsize5", (240, 320), 4, (2, 1), dict(winsize=5)),
    ("winsize31", (240, 320), 4, (2, 1), dict(winsize=31)),
    # radii 8..12: the vertical ring packed into tensor memory (fb_iter_v.cuh, tmem_ring_packed); 1080p rows: two segments
    ("winsize17", (240, 320), 4, (2, 1), dict(winsize=17)),
    ("winsize19", (240, 320), 4, (2, 1), dict(winsize=19)),
    ("winsize21_odd_size", (203, 331), 5, (-3, 2), dict(winsize=21)),
    ("winsize23", (240, 320), 6, (2, -1), dict(winsize=23)),
    ("winsize25_vga", (480, 640), 7, (4, 3), dict(winsize=25)),
    ("winsize27", (240, 320), 4, (2, 1), dict(winsize=27)),
    ("poly7", (240, 320), 5, (3, 2), dict(poly_n=7, poly_sigma=1.5)),
    ("pyr08", (240, 320), 6, (3, 2), dict(pyr_scale=0.8, levels=5)),
    ("levels0", (200, 300), 7, (2, 1), dict(levels=0)),
    ("levels1", (200, 300), 7, (2, 1), dict(levels=1)),
    ("levels8_clamped", (200, 300), 7, (2, 1), dict(levels=8)),
    ("iter1", (240, 320), 8, (2, 1), dict(iterations=1)),
    ("tiny_33x35", (33, 35), 9, (0.5, 0.25), {}),
]


@pytest.mark.parametrize("name,shape,seed,shift,kw", CASES, ids=[c[0] for c in CASES])
def test_farneback_matches_cv2(engine_factory, name, shape, seed, shift, kw):
    a, b = synth.synth_pair(shape[0], shape[1], seed, shift)
    eng = engine_factory(shape[1], shape[0])
    got = eng.farneback(a, b, None, **kw)
    _check(C.farneback(a, b, **kw), got)


def test_module_level_cv2_signature(built_lib):
    import opticalflowcontainer_b200 as ofb
    a, b = synth.synth_pair(120, 160, 21, (1.0, 0.5))
    got = ofb.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    _check(C.farneback(a, b), got)


def test_rotation_zoom_and_low_texture(engine_factory):
    eng = engine_factory(640, 480)
    a, b = synth.synth_warp_pair(480, 640, 31)
    _check(C.farneback(a, b), eng.farneback(a, b))
    a, b = synth.low_texture_pair(480, 640, 3)
    # low-texture frames are ill-conditioned (the 1e-3 regulariser dominates): contract gate
    _check(C.farneback(a, b), eng.farneback(a, b), MEAN_GATE, MAX_GATE)


def test_high_contrast_edges(engine_factory):
    """Strong edges beside flat areas: the float (van Herk / Gil-Werman) vertical sums of the fused
    iteration kernel replace cv2's double running sums here — contract gate."""
    eng = engine_factory(640, 480)
    for seed in (5, 6):
        a, b = synth.high_contrast_pair(480, 640, seed)
        _check(C.farneback(a, b), eng.farneback(a, b), MEAN_GATE, MAX_GATE)


def test_initial_flow(engine_factory):
    a, b = synth.synth_pair(240, 320, 2, (5, 3))
    f0 = np.full((240, 320, 2), (4.5, 2.5), np.float32)
    eng = engine_factory(320, 240)
    got = eng.farneback(a, b, f0.copy(), flags=4)
    _check(C.farneback(a, b, flow=f0.copy(), flags=4), got)
    # odd size: INTER_AREA with fractional cells
    a, b = synth.synth_pair(203, 317, 2, (5, 3))
    rng = np.random.default_rng(0)
    f0 = (np.array([4.5, 2.5], np.float32) + rng.standard_normal((203, 317, 2)).astype(np.float32) * 0.2)
    eng = engine_factory(317, 203)
    _check(C.farneback(a, b, flow=f0.copy(), flags=4), eng.farneback(a, b, f0.copy(), flags=4))


def test_against_numpy_restatement(engine_factory):
    a, b = synth.synth_pair(200, 264, 41, (2.2, -1.1))
    eng = engine_factory(264, 200)
    _check(F.farneback(a, b), eng.farneback(a, b))


def test_golden_fixtures(engine_factory):
    files = sorted(glob.glob(os.path.join(GOLDEN, "farneback_*.npz")))
    assert files
    for f in files:
        z = np.load(f)
        kw = {k[3:]: z[k].item() for k in z.files if k.startswith("kw_")}
        hgt, wid = z["prev"].shape
        eng = engine_factory(wid, hgt)
        gate = (MEAN_GATE, MAX_GATE) if "lowtex" in f else (TIGHT_MEAN, TIGHT_MAX)
        _check(z["flow"], eng.farneback(z["prev"], z["next"], None, **kw), *gate)


def test_batch_equals_single(engine_factory):
    eng = engine_factory(320, 240, 4)
    pairs = [synth.synth_pair(240, 320, 50 + i, (1.0 + i, -0.25 - 0.5 * i)) for i in range(4)]
    out = eng.farneback_batch([p[0] for p in pairs], [p[1] for p in pairs])
    for i, (a, b) in enumerate(pairs):
        single = eng.farneback(a, b)
        assert np.array_equal(single, out[i])          # batching must not change a bit
        _check(C.farneback(a, b), out[i])


def test_integer_shift_knife_edge(engine_factory):
    """A whole-pixel vertical shift of zero puts the last row's displaced position exactly ON the inside test of
    UpdateMatrices (floor(y + dy) == h - 1 is outside, h - 2 inside): a 1e-7 difference in dy flips one matrix element,
    which the 15x15 window spreads over its footprint (measured: one element, 0.011 px over 8 x 15 pixels, depending on
    the summation order of the window sums).  cv2 itself is discontinuous there, so this case is held to the contract
    gate, not to the 10x tighter one."""
    eng = engine_factory(320, 240, 1)
    a, b = synth.synth_pair(240, 320, 50, (1.0, -0.0))
    got = eng.farneback(a, b)
    mean, mx = _check(C.farneback(a, b), got, MEAN_GATE, MAX_GATE)
    assert mean <= 1e-4        # everything but the knife-edge footprint agrees to float rounding


def test_strided_input_and_capacity(engine_factory):
    import opticalflowcontainer_b200 as ofb
    a, b = synth.synth_pair(120, 200, 60, (1.5, 0.5))
    big_a = np.zeros((120, 256), np.uint8); big_a[:, :200] = a
    big_b = np.zeros((120, 256), np.uint8); big_b[:, :200] = b
    eng = engine_factory(200, 120)
    got = eng.farneback(big_a[:, :200], big_b[:, :200])       # row stride 256
    assert np.array_equal(got, eng.farneback(a, b))
    with pytest.raises(ofb.OfbError) as ei:
        eng.farneback(np.zeros((300, 300), np.uint8), np.zeros((300, 300), np.uint8))
    assert ei.value.status == 4
    with pytest.raises(ofb.OfbError):
        eng.farneback(a, b, pyr_scale=1.5)
    with pytest.raises(ofb.OfbError):
        eng.farneback(a, b[:100])
    with pytest.raises(ofb.OfbError):
        eng.farneback(a.astype(np.float32), b)


def test_full_size_1080p_properties(engine_factory):
    """BASELINE config[1] size.  cv2 at 1080p takes ~0.8 s — affordable once; plus size-independent
    properties: a pure translation is recovered in the interior, and swapping the frames negates it."""
    a, b = synth.synth_pair(1080, 1920, 1, (6.2, 3.4))
    eng = engine_factory(1920, 1080)
    got = eng.farneback(a, b)
    _check(C.farneback(a, b), got)
    inner = got[100:-100, 100:-100]
    assert abs(float(np.median(inner[..., 0])) - 6.2) < 0.1 and abs(float(np.median(inner[..., 1])) - 3.4) < 0.1
    back = eng.farneback(b, a)[100:-100, 100:-100]
    assert abs(float(np.median(back[..., 0])) + 6.2) < 0.1 and abs(float(np.median(back[..., 1])) + 3.4) < 0.1
    # determinism: same input twice -> identical bits
    assert np.array_equal(got, eng.farneback(a, b))


@pytest.mark.parametrize("n", [3, 8])
def test_pinned_pipelined_batch_equals_single(engine_factory, n):
    """Pinned host buffers take the chunked copy/compute/copy pipeline; results must be bit-identical
    to one-pair calls."""
    import torch
    eng = engine_factory(320, 240, 8)
    pairs = [synth.synth_pair(240, 320, 80 + i, (1.0 + 0.5 * i, -0.25 * i)) for i in range(n)]
    prevs = torch.from_numpy(np.stack([p[0] for p in pairs])).pin_memory()
    nexts = torch.from_numpy(np.stack([p[1] for p in pairs])).pin_memory()
    out = torch.empty((n, 240, 320, 2), dtype=torch.float32).pin_memory()
    eng.farneback_batch_into(prevs.numpy(), nexts.numpy(), out.numpy())
    for i, (a, b) in enumerate(pairs):
        assert np.array_equal(eng.farneback(a, b), out[i].numpy())
    # the on-device reduction sees all n fields of the pipelined call
    eng.farneback_batch_into(prevs.numpy(), nexts.numpy(), out.numpy())
    _, med = eng.flow_u_stats(n)
    for i in range(n):
        assert med[i] == np.float32(np.median(out[i].numpy()[..., 0]))


def test_async_batch_pipelines_across_calls(built_lib):
    """ofb_farneback_batch_async + ofb_wait: three back-to-back calls on page-locked buffers, two result
    buffers in flight, results equal to the synchronous call's.  The two paths split the batch into different chunks, and
    the iteration kernel cuts a level into row segments by batch size (fb_iter_launch.cuh): its float window sums restart
    with a segment, so the last bits of the flow follow the chunking — equal to 1e-4 px (measured 2e-6), not bit for bit."""
    import torch
    import opticalflowcontainer_b200 as ofb
    n, h, w = 8, 240, 320
    eng = ofb.FlowEngine(w, h, n, 0)
    try:
        sets = []
        for s in range(3):
            prs = [synth.synth_pair(h, w, 40 + 8 * s + i, (2.5 + 0.3 * i, -1.5 + 0.2 * s)) for i in range(n)]
            a = torch.from_numpy(np.stack([p[0] for p in prs])).pin_memory()
            b = torch.from_numpy(np.stack([p[1] for p in prs])).pin_memory()
            sets.append((a, b))
        ref = [eng.farneback_batch_into(a.numpy(), b.numpy(), np.empty((n, h, w, 2), np.float32)).copy() for a, b in sets]
        outs = [torch.empty((n, h, w, 2), dtype=torch.float32).pin_memory() for _ in range(3)]
        for s, (a, b) in enumerate(sets):
            eng.farneback_batch_into(a.numpy(), b.numpy(), outs[s].numpy(), wait=False)
        eng.wait()
        for s in range(3):
            assert np.abs(outs[s].numpy() - ref[s]).max() <= 1e-4
        # reuse of a result buffer by a later call is ordered by the library
        eng.farneback_batch_into(sets[0][0].numpy(), sets[0][1].numpy(), outs[0].numpy(), wait=False)
        eng.farneback_batch_into(sets[1][0].numpy(), sets[1][1].numpy(), outs[0].numpy(), wait=False)
        eng.wait()
        assert np.abs(outs[0].numpy() - ref[1]).max() <= 1e-4
    finally:
        eng.close()


def test_flow_discontinuity_and_odd_size_match_cv2(engine_factory):
    """A torn frame (large |flow| differences between neighbouring rows: the row-reuse gather of k_iter_v has to fall
    back to the full 2x2 gather, and the fused inter-level upsample crosses the tear) on an even and an odd size (general
    resize tables in the fused upsample): contract gate against cv2 (the tear itself is ill-conditioned)."""
    for (h, w, shift) in [(270, 480, (1.7, -0.9)), (213, 317, (-6.3, 4.2))]:
        a, b = synth.synth_pair(h, w, 5, shift)
        b = np.ascontiguousarray(np.roll(b, 5, axis=0))
        eng = engine_factory(w, h)
        _check(C.farneback(a, b), eng.farneback(a, b), MEAN_GATE, MAX_GATE)


@pytest.mark.parametrize("seed", range(8))
def test_corpus_c1_seeds_0_to_7(engine_factory, seed):
    """SURVEY 8d corpus C1: seeds 0..7 of the VGA translation corpus, shifts spread over +-12 px."""
    shift = (((seed * 37) % 25) - 12 + 0.3, ((seed * 53) % 17) - 8 - 0.4)
    a, b = synth.synth_pair(480, 640, seed, shift)
    eng = engine_factory(640, 480)
    _check(C.farneback(a, b), eng.farneback(a, b))


def test_full_size_4k_matches_cv2(engine_factory):
    """BASELINE config[2] size (3840x2160): one pair against cv2 (about 3 s of cv2)."""
    a, b = synth.synth_pair(2160, 3840, 3, (7.3, -4.6))
    eng = engine_factory(3840, 2160)
    _check(C.farneback(a, b), eng.farneback(a, b))


def test_full_size_8k_matches_cv2(engine_factory):
    """BASELINE config[4] size (7680x4320), whole frame on one GPU: one pair against cv2 (about 13 s of cv2).  The
    multi-GPU tiled run of the same size is checked against cv2 by tests/test_tiled_gpu.py and by bench.py."""
    a, b = synth.synth_pair(4320, 7680, 4, (-9.4, 5.2))
    eng = engine_factory(7680, 4320)
    _check(C.farneback(a, b), eng.farneback(a, b))


def test_stream_cache_grow_then_shrink(built_lib):
    """n_streams 1 -> 2 -> 1 with identical frames, sizes and staging pointers: the stream cache is reallocated when it
    grows, so the CUDA graphs captured for the first configuration must not be replayed on the freed pool (the pool
    base is part of the graph key)."""
    import opticalflowcontainer_b200 as ofb
    h, w = 135, 240
    eng = ofb.FlowEngine(w, h, 2, 0)
    ref = ofb.FlowEngine(w, h, 2, 0)
    try:
        base = synth.synth_pair(h, w, 77, (0.0, 0.0))[0]
        fr = [synth.subpixel_shift(base, 1.1 * t, -0.6 * t) for t in range(8)]
        kw = dict(pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2, flags=0)
        t = 0
        for n in (1, 2, 1, 2, 1):
            assert eng.farneback_stream(np.stack([fr[t % 8]] * n), **kw) is None      # n changed: priming call
            for _ in range(4):                                                        # every graph key comes round twice
                out = eng.farneback_stream(np.stack([fr[(t + 1) % 8]] * n), **kw)
                want = ref.farneback_batch([fr[t % 8]] * n, [fr[(t + 1) % 8]] * n, **kw)
                assert out is not None and np.array_equal(out, want), (n, t)
                t += 1
    finally:
        eng.close()
        ref.close()


@pytest.mark.parametrize("flags,winsize", [(0, 15), (0, 9), (256, 15)])
def test_stream_call_equals_pair_calls(built_lib, flags, winsize):
    """ofb_farneback_stream: one new frame per stream per call, the previous frame's state kept on the device.  Every
    field equals ofb_farneback(previous, frame) bit for bit — for the cached-expansion path (box window) and for the
    kept-frame path (Gaussian window) — across priming, a parameter change and a reset."""
    import opticalflowcontainer_b200 as ofb
    n, h, w = 3, 135, 240
    eng = ofb.FlowEngine(w, h, n, 0)
    ref = ofb.FlowEngine(w, h, n, 0)       # same batch size: same launch geometry, hence the same float summation order
    try:
        seqs = []
        for s in range(n):
            base = synth.synth_pair(h, w, 40 + s, (0.0, 0.0))[0]
            seqs.append([synth.subpixel_shift(base, 1.3 * t * (s + 1), -0.7 * t) for t in range(4)])
        kw = dict(pyr_scale=0.5, levels=3, winsize=winsize, iterations=3, poly_n=5, poly_sigma=1.2, flags=flags)
        out = eng.farneback_stream(np.stack([seqs[s][0] for s in range(n)]), **kw)
        assert out is None                                   # priming
        for t in range(1, 4):
            out = eng.farneback_stream(np.stack([seqs[s][t] for s in range(n)]), **kw)
            assert out is not None and out.shape == (n, h, w, 2)
            want = ref.farneback_batch([seqs[s][t - 1] for s in range(n)], [seqs[s][t] for s in range(n)], **kw)
            assert np.array_equal(out, want)
        # fields can stay on the device for the node reduction
        assert eng.farneback_stream(np.stack([seqs[s][2] for s in range(n)]), download=False, **kw) is True
        mean, med = eng.flow_u_stats(n)
        want = ref.farneback_batch([seqs[s][3] for s in range(n)], [seqs[s][2] for s in range(n)], **kw)
        assert med[0] == np.float32(np.median(want[0][..., 0]))
        # a parameter change primes again; so does a reset
        kw2 = dict(kw, levels=2)
        assert eng.farneback_stream(np.stack([seqs[s][0] for s in range(n)]), **kw2) is None
        assert eng.farneback_stream(np.stack([seqs[s][1] for s in range(n)]), **kw2) is not None
        eng.stream_reset()
        assert eng.farneback_stream(np.stack([seqs[s][1] for s in range(n)]), **kw2) is None
    finally:
        eng.close(); ref.close()


def test_repeated_calls_are_deterministic(built_lib):
    """compute-sanitizer's racecheck is not available on the GPU pool: a race in the producer/consumer hand-over of the
    fused kernels would show up as run-to-run differences.  Same batch, ten calls, three window sizes: identical bits."""
    import opticalflowcontainer_b200 as ofb
    n, h, w = 6, 270, 480
    eng = ofb.FlowEngine(w, h, n, 0)
    try:
        prs = [synth.synth_pair(h, w, 70 + i, (2.0 + i, -1.0 - 0.5 * i)) for i in range(n)]
        a = [p[0] for p in prs]; b = [p[1] for p in prs]
        for winsize in (15, 9, 21):
            first = eng.farneback_batch(a, b, winsize=winsize)
            for _ in range(9):
                assert np.array_equal(eng.farneback_batch(a, b, winsize=winsize), first)
    finally:
        eng.close()


def test_stream_call_asynchronous_reduction(built_lib):
    """ofb_farneback_stream(flow=NULL) + ofb_flow_u_stats_async: a camera loop that never waits — page-locked frames,
    several calls in flight, results handed over at ofb_wait — gives the scalars of the synchronous loop."""
    import torch
    import opticalflowcontainer_b200 as ofb
    n, h, w = 2, 135, 240
    eng = ofb.FlowEngine(w, h, n, 0)
    ref = ofb.FlowEngine(w, h, n, 0)
    try:
        base = [synth.synth_pair(h, w, 90 + s, (0.0, 0.0))[0] for s in range(n)]
        T = 7
        frames = [torch.from_numpy(np.stack([synth.subpixel_shift(base[s], 0.8 * t * (s + 1), 0.3 * t) for s in range(n)])).pin_memory()
                  for t in range(T)]
        want = []
        for t in range(T):
            r = ref.farneback_stream(frames[t].numpy(), download=False)
            if r is not None:
                want.append(ref.flow_u_stats(n))
        got = []
        for t in range(T):
            r = eng.farneback_stream(frames[t].numpy(), download=False)
            if r is not None:
                got.append(eng.flow_u_stats(n, wait=False))
        eng.wait()
        assert len(got) == len(want) == T - 1
        for (m0, d0), (m1, d1) in zip(want, got):
            assert np.array_equal(np.asarray(m0), m1) and np.array_equal(np.asarray(d0, np.float32), d1)
    finally:
        eng.close(); ref.close()
