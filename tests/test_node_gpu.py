"""Node-level contract (frame in -> Vector3Stamped out) and the on-device flow reduction,
against the same arithmetic done with NumPy on the cv2 flow (what the reference nodes do:
lfn3_sub_node.py:205-222, opticalflow_node.py:97-109, sub_n_pub_lfn3_node.py:195-210)."""
from collections import deque

import numpy as np
import pytest

from oracle import cv2_oracle as C
from oracle import synth

pytestmark = pytest.mark.gpu


def test_flow_u_stats_matches_numpy(engine_factory):
    a, b = synth.synth_warp_pair(240, 320, 5)
    eng = engine_factory(320, 240, 2)
    flow = eng.farneback(a, b)
    mean, med = eng.flow_u_stats(1)
    assert abs(mean[0] - float(np.mean(flow[..., 0].astype(np.float64)))) < 1e-6
    assert med[0] == np.float32(np.median(flow[..., 0]))          # exact selection
    # odd count + mask
    mask = np.zeros((240, 320), np.uint8)
    mask[10:51, 20:61] = 1                                          # 41*41 = 1681 (odd)
    mean, med = eng.flow_u_stats(1, mask=mask)
    sel = flow[..., 0][mask.astype(bool)]
    assert abs(mean[0] - float(sel.astype(np.float64).mean())) < 1e-6
    assert med[0] == np.float32(np.median(sel))
    # batch of 2
    pairs = [synth.synth_pair(240, 320, 70 + i, (2.0 + i, -1.0)) for i in range(2)]
    out = eng.farneback_batch([p[0] for p in pairs], [p[1] for p in pairs])
    mean, med = eng.flow_u_stats(2)
    for i in range(2):
        assert abs(mean[i] - float(out[i, ..., 0].astype(np.float64).mean())) < 1e-6
        assert med[i] == np.float32(np.median(out[i, ..., 0]))


def test_flow_u_stats_negative_and_empty_mask(engine_factory):
    a, b = synth.synth_pair(120, 160, 9, (-3.5, 1.0))
    eng = engine_factory(160, 120)
    flow = eng.farneback(a, b)
    _, med = eng.flow_u_stats(1)
    assert med[0] == np.float32(np.median(flow[..., 0])) and med[0] < 0
    mean, med = eng.flow_u_stats(1, mask=np.zeros((120, 160), np.uint8))
    assert np.isnan(mean[0]) and np.isnan(med[0])


@pytest.mark.parametrize("reduce,on_device", [("median", False), ("mean", False), ("median", True), ("mean", True)])
def test_node_velocity_contract(built_lib, reduce, on_device):
    from opticalflowcontainer_b200.node import FarnebackVelocityNode
    frames = synth.panning_sequence(120, 160, 5, seed=3, max_shift=3.0)
    node = FarnebackVelocityNode(width=160, height=120, pixel_to_meter=0.0011, reduce=reduce,
                                 on_device_reduce=on_device)
    stamps = [0.0, 0.033, 0.066, 0.066, 0.1]            # a repeated stamp -> dt <= 0 -> 1e-3
    buf = deque(maxlen=5)
    prev, prev_t = None, None
    for f, t in zip(frames, stamps):
        bgr = np.repeat(f[..., None], 3, axis=2)
        out = node.image_callback(bgr, t, "bgr8")
        if prev is None:
            assert out is None                           # first frame only primes the state
            prev, prev_t = f, t
            continue
        dt = t - prev_t
        if dt <= 0:
            dt = 1e-3
        flow = C.farneback(prev, f)
        u = np.mean(flow[..., 0]) if reduce == "mean" else np.median(flow[..., 0])
        vx = float(u / dt * 0.0011)
        buf.append(vx)
        raw, smooth = out
        assert raw.frame_id == "camera_link" and raw.stamp == t and raw.vector[1:] == (0.0, 0.0)
        assert abs(raw.vector[0] - vx) <= 1e-3 * max(1.0, abs(vx)) + 1e-6, (raw.vector[0], vx)
        assert abs(smooth.vector[0] - float(np.mean(buf))) <= 1e-3 * max(1.0, abs(vx)) + 1e-6
        prev, prev_t = f, t


def test_gray_conversion_is_cv2_exact():
    cv2 = pytest.importorskip("cv2")
    from opticalflowcontainer_b200.node import to_gray_u8
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(97, 131, 3), dtype=np.uint8)
    assert np.array_equal(to_gray_u8(img, "bgr8"), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(to_gray_u8(img, "rgb8"), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))


def test_junction_mask_matches_reference_logic():
    from opticalflowcontainer_b200.node import junction_mask
    pts = [(5.2, 7.9), (159, 119), (-3, 4), (80, 60)]
    m = junction_mask(pts, 120, 160)
    ref = np.zeros((120, 160), bool)
    for p in pts:
        x, y = int(p[0]), int(p[1])
        if 0 <= x < 160 and 0 <= y < 120:
            ref[max(0, y - 5):min(120, y + 6), max(0, x - 5):min(160, x + 6)] = True
    assert np.array_equal(m, ref)


def test_cvt_gray_bit_exact(engine_factory):
    """Frame ingest: BGR/RGB -> gray on the device equals cv2.cvtColor bit for bit,
    including odd widths and padded rows (sensor_msgs/Image `step`)."""
    import cv2
    rng = np.random.default_rng(5)
    eng = engine_factory(641, 481)
    for (h, w) in [(480, 640), (481, 641), (7, 5), (64, 3)]:
        bgr = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        assert np.array_equal(eng.cvt_gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
        assert np.array_equal(eng.cvt_gray(bgr, rgb=True), cv2.cvtColor(bgr, cv2.COLOR_RGB2GRAY))
    padded = rng.integers(0, 256, size=(100, 200, 3), dtype=np.uint8)
    view = padded[:, :150]                      # row step 600 bytes, 450 used
    assert np.array_equal(eng.cvt_gray(view), cv2.cvtColor(np.ascontiguousarray(view), cv2.COLOR_BGR2GRAY))


def test_batch_stats_matches_numpy(built_lib):
    """ofb_farneback_batch_stats = flow + on-device mean/median of u without downloading the field."""
    import opticalflowcontainer_b200 as ofb
    n, h, w = 4, 240, 320
    eng = ofb.FlowEngine(w, h, n, 0)
    try:
        prs = [synth.synth_pair(h, w, 60 + i, (1.5 + 0.7 * i, -0.5 * i)) for i in range(n)]
        a = np.stack([p[0] for p in prs]); b = np.stack([p[1] for p in prs])
        flows = eng.farneback_batch_into(a, b, np.empty((n, h, w, 2), np.float32))
        mean, med = eng.farneback_batch_stats(a, b)
        for i in range(n):
            u = flows[i, :, :, 0]
            assert abs(mean[i] - float(u.astype(np.float64).mean())) <= 1e-6
            assert med[i] == np.float32(np.median(u))
    finally:
        eng.close()


def test_batch_stats_async_pipelines_across_calls(built_lib):
    """ofb_farneback_batch_stats_async: several reductions in flight on page-locked frames, results handed over
    by ofb_wait, equal to the synchronous call's (and to numpy on the downloaded field)."""
    import torch
    import opticalflowcontainer_b200 as ofb
    n, h, w = 3, 200, 264
    eng = ofb.FlowEngine(w, h, n, 0)
    try:
        sets = []
        for s in range(6):       # more calls than the library keeps pending slots for
            prs = [synth.synth_pair(h, w, 80 + 10 * s + i, (1.1 + 0.6 * i + 0.2 * s, 0.4 * i - 0.3 * s)) for i in range(n)]
            a = torch.from_numpy(np.stack([p[0] for p in prs])).pin_memory()
            b = torch.from_numpy(np.stack([p[1] for p in prs])).pin_memory()
            sets.append((a, b))
        sync = [eng.farneback_batch_stats(a.numpy(), b.numpy()) for a, b in sets]
        pend = [eng.farneback_batch_stats(a.numpy(), b.numpy(), wait=False) for a, b in sets]
        eng.wait()
        for (m0, d0), (m1, d1) in zip(sync, pend):
            assert np.array_equal(m0, m1) and np.array_equal(d0, d1)
        # pageable frames through the async entry point are served synchronously
        a, b = sets[0]
        m2, d2 = eng.farneback_batch_stats(a.numpy().copy(), b.numpy().copy(), wait=False)
        assert np.array_equal(m2, sync[0][0]) and np.array_equal(d2, sync[0][1])
        eng.wait()
    finally:
        eng.close()


@pytest.mark.parametrize("k", [0, 3, 5])
def test_flow_postfilter_bit_exact(engine_factory, k):
    """ofb_flow_postfilter == the adapt node's medianBlur + magnitude mask + intensity mask
    (lfn3_adapt_node.py:236-251) computed by cv2 / NumPy on the same field, bit for bit; odd sizes included."""
    from oracle import postfilter_np
    for (h, w) in [(120, 160), (67, 93)]:
        eng = engine_factory(w, h)
        a, b = synth.synth_pair(h, w, 31 + k, (2.3, -1.2))
        flow = eng.farneback(a, b, None, 0.5, 2, 9, 2, 5, 1.1, 0)
        flow[::7, ::5] *= -3.0                      # outliers for the median to remove (re-upload below)
        # put the modified field on the device: run the filter on a field we control exactly
        eng.farneback(a, b, flow.copy(), 0.5, 0, 9, 0, 5, 1.1, 4)   # iterations=0 + USE_INITIAL_FLOW: field = input
        base = eng.flow_download(1, h, w)[0]
        assert np.array_equal(base, flow)
        rng = np.random.default_rng(9)
        gray = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        thr = float(np.median(np.hypot(flow[..., 0], flow[..., 1])))
        eng.flow_postfilter(1, k, thr, gray, 128)
        got = eng.flow_download(1, h, w)[0]
        ref = postfilter_np.adapt_postfilter_np(flow, k, thr, gray, 128, use_cv2_median=True)
        assert np.array_equal(got, ref)
        mean, _ = eng.flow_u_stats(1, median=False)
        assert abs(mean[0] - float(ref[..., 0].astype(np.float64).mean())) <= 1e-6
        # median only / masks only
        eng.farneback(a, b, flow.copy(), 0.5, 0, 9, 0, 5, 1.1, 4)
        eng.flow_postfilter(1, k)
        got = eng.flow_download(1, h, w)[0]
        assert np.array_equal(got, postfilter_np.adapt_postfilter_np(flow, k, use_cv2_median=True))


def test_adapt_node_contract(engine_factory):
    """FarnebackVelocityNode with the adapt node's parameters: same velocity whether the post-filter + mean run
    on the device or in NumPy/cv2 on the downloaded field."""
    import cv2
    from opticalflowcontainer_b200.node import FarnebackVelocityNode
    h, w = 120, 160
    frames = [synth.synth_pair(h, w, 77, (1.0 + 0.5 * i, 0.2))[1] for i in range(3)]
    kw = dict(width=w, height=h, reduce="mean", median_kernel_size=5, flow_magnitude_threshold=0.3, intensity_threshold=140)
    n_dev = FarnebackVelocityNode(engine=engine_factory(w, h), on_device_reduce=True, **kw)
    n_ref = FarnebackVelocityNode(engine=engine_factory(w, h), width=w, height=h, reduce="mean")
    for i, f in enumerate(frames):
        out = n_dev.image_callback(f, 0.1 * i)
        ref = n_ref.image_callback(f, 0.1 * i)
        if i == 0:
            assert out is None and ref is None
            continue
        fl = np.transpose(n_ref.last_flow, (2, 0, 1)).copy()
        fl[0] = cv2.medianBlur(fl[0], 5); fl[1] = cv2.medianBlur(fl[1], 5)
        m = (np.sqrt(fl[0] ** 2 + fl[1] ** 2) >= 0.3).astype(np.float32)
        fl[0] *= m; fl[1] *= m
        mi = (f < 140).astype(np.float32)
        fl[0] *= mi
        vx = float(np.mean(fl[0]) / 0.1 * 0.0011)
        assert abs(out[0].vector[0] - vx) <= 1e-6 * max(1.0, abs(vx))


def test_junction_velocity_matches_reference_logic(engine_factory):
    """flow_sample + junction_velocity == the junction node's Python (lfn3_junction_node.py:203-231) run on the
    downloaded field."""
    from opticalflowcontainer_b200.node import junction_velocity
    h, w = 120, 160
    eng = engine_factory(w, h)
    a, b = synth.synth_pair(h, w, 12, (3.2, -1.4))
    flow = eng.farneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    arr = np.transpose(flow, (2, 0, 1))
    rng = np.random.default_rng(4)
    prev = np.concatenate([rng.uniform([8, 8], [w - 8, h - 8], size=(12, 2)), [[-3.0, 5.0], [w + 2.0, 10.0]]])
    curr = prev[:12] + np.array([3.2, -1.4]) + rng.normal(scale=0.4, size=(12, 2))
    s = eng.flow_sample(prev.astype(np.int64))
    assert np.isnan(s[12:]).all()
    for i in range(12):
        x, y = int(prev[i, 0]), int(prev[i, 1])
        assert s[i, 0] == arr[0, y, x] and s[i, 1] == arr[1, y, x]
    # the node's own loop
    matches = []
    for p in prev:
        x, y = int(p[0]), int(p[1])
        if 0 <= x < w and 0 <= y < h:
            pp = np.array([p[0] + arr[0, y, x], p[1] + arr[1, y, x]])
            dist = np.sqrt(((curr - pp) ** 2).sum(1))
            j = int(dist.argmin())
            if dist[j] < 5.0:
                matches.append((p, curr[j]))
    assert len(matches) >= 4
    vx_ref = float(np.mean([c - p for p, c in matches], axis=0)[0] / 0.05 * 0.0011)
    vx = junction_velocity(eng, prev, curr, 0.05, 0.0011)
    assert vx is not None and abs(vx - vx_ref) <= 1e-12 + 1e-9 * abs(vx_ref)
    assert junction_velocity(eng, prev[:2], curr, 0.05, 0.0011) is None      # fewer than 4 matches


def test_node_stream_mode_equals_pair_mode(engine_factory):
    """FarnebackVelocityNode(use_stream=True): previous frame's state on the GPU, same messages as the pair calls."""
    from opticalflowcontainer_b200.node import FarnebackVelocityNode
    h, w = 120, 160
    base = synth.synth_pair(h, w, 21, (0.0, 0.0))[0]
    frames = [synth.subpixel_shift(base, 1.1 * i, 0.4 * i) for i in range(4)]
    a = FarnebackVelocityNode(engine=engine_factory(w, h), width=w, height=h, use_stream=True)
    b = FarnebackVelocityNode(engine=engine_factory(w, h), width=w, height=h)
    for i, f in enumerate(frames):
        ra, rb = a.image_callback(f, 0.05 * i), b.image_callback(f, 0.05 * i)
        assert (ra is None) == (rb is None)
        if ra is not None:
            assert ra[0].vector == rb[0].vector and ra[1].vector == rb[1].vector
            assert np.array_equal(a.last_flow, b.last_flow)


@pytest.mark.parametrize("src,dst", [((720, 1280), (480, 640)), ((480, 640), (1080, 1920)), ((481, 637), (240, 320)),
                                     ((100, 100), (50, 50)), ((48, 64), (108, 192)), ((37, 53), (37, 53)), ((64, 48), (5, 7))])
def test_resize_bit_exact(engine_factory, src, dst):
    """ofb_resize_u8 == cv2.resize (INTER_LINEAR, uint8), gray and 3-channel, down- and up-scaling, odd sizes."""
    import cv2
    eng = engine_factory(64, 64)           # frames of any size: not limited by the handle's capacity
    rng = np.random.default_rng(src[0] + dst[1])
    for cn in (1, 3):
        img = rng.integers(0, 256, size=src if cn == 1 else src + (3,), dtype=np.uint8)
        assert np.array_equal(eng.resize(img, (dst[1], dst[0])), cv2.resize(img, (dst[1], dst[0])))


def test_ingest_gray_matches_node_ingest(engine_factory):
    """ofb_ingest_gray == cv2.resize + cv2.cvtColor as the nodes run them (lfn3_sub_node.py:148-159); the node mirror
    accepts camera frames larger than its configured size."""
    import cv2
    from opticalflowcontainer_b200.node import FarnebackVelocityNode
    eng = engine_factory(160, 120)
    rng = np.random.default_rng(2)
    frame = rng.integers(0, 256, size=(360, 640, 3), dtype=np.uint8)
    want = cv2.cvtColor(cv2.resize(frame, (160, 120)), cv2.COLOR_BGR2GRAY)
    assert np.array_equal(eng.ingest_gray(frame, (160, 120)), want)
    assert np.array_equal(eng.ingest_gray(frame, (160, 120), rgb=True), cv2.cvtColor(cv2.resize(frame, (160, 120)), cv2.COLOR_RGB2GRAY))
    assert np.array_equal(eng.ingest_gray(frame), cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY))
    view = frame[:, :500]                                   # row step larger than the row
    assert np.array_equal(eng.ingest_gray(view, (160, 120)), cv2.cvtColor(cv2.resize(np.ascontiguousarray(view), (160, 120)), cv2.COLOR_BGR2GRAY))
    node = FarnebackVelocityNode(engine=eng, width=160, height=120)
    big = [cv2.resize(np.dstack([synth.synth_pair(120, 160, 3, (1.5 * i, 0.5))[1]] * 3), (640, 360)) for i in range(2)]
    assert node.image_callback(big[0], 0.0) is None
    out = node.image_callback(big[1], 0.1)
    assert out is not None and np.isfinite(out[0].vector[0])


@pytest.mark.parametrize("size", [(480, 640), (241, 317), (1080, 1920)])
def test_clahe_bit_exact(engine_factory, size):
    """ofb_clahe == cv2.createCLAHE(clip, grid).apply (lfn3_adapt_node.py:164-182), incl. the adaptive clip limit."""
    import cv2
    from opticalflowcontainer_b200.node import adaptive_clip_limit
    eng = engine_factory(64, 64)
    rng = np.random.default_rng(size[0])
    a = cv2.GaussianBlur((rng.random(size) * 255).astype(np.float32), (0, 0), 3)
    low = ((a - a.min()) / (a.max() - a.min()) * 120 + 40).astype(np.uint8)
    noise = rng.integers(0, 256, size=size, dtype=np.uint8)
    for im in (low, noise):
        for clip, grid in [(2.0, (8, 8)), (40.0, (8, 8)), (0.0, (4, 6)), (adaptive_clip_limit(im, 1.0, 4.0, 0.1, 0.8), (8, 8))]:
            assert np.array_equal(eng.clahe(im, clip, grid), cv2.createCLAHE(clipLimit=clip, tileGridSize=grid).apply(im))


@pytest.mark.parametrize("size", [(480, 640), (243, 317), (96, 100)])
def test_adapt_prefilter_bit_exact(engine_factory, size):
    """ofb_adapt_prefilter == the adapt node's BGR2HSV -> adaptive CLAHE on V -> HSV2RGB in cv2 (lfn3_adapt_node.py:164-184),
    bit for bit, for widths with and without a scalar tail in cv2's HSV2RGB."""
    import cv2
    eng = engine_factory(64, 64)
    rng = np.random.default_rng(size[1])
    bgr = cv2.GaussianBlur(rng.integers(0, 256, size=size + (3,), dtype=np.uint8), (0, 0), 1.5)
    bgr[: size[0] // 4] = rng.integers(0, 256, size=(size[0] // 4, size[1], 3), dtype=np.uint8)
    hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
    h, s, v = cv2.split(hsv)
    for clip in (None, 2.0, 40.0):
        c = clip
        if clip is None:       # the node's adaptive clip limit
            contrast = np.std(v) / (np.mean(v) + 1e-3)
            c = float(np.clip(1.0 + (contrast - 0.1) / (0.8 - 0.1) * (4.0 - 1.0), 1.0, 4.0))
        clahe = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8))
        clahe.setClipLimit(c)
        want = cv2.cvtColor(cv2.merge((h, s, clahe.apply(v))), cv2.COLOR_HSV2RGB)
        got, used = eng.adapt_prefilter(bgr, clip, (1.0, 4.0, 0.1, 0.8), (8, 8))
        assert abs(used - c) <= 1e-9 * max(1.0, abs(c))
        assert np.array_equal(got, want)


def test_flow_to_color_matches_cv2_recipe(built_lib):
    """ofb_flow_to_bgr: the nodes' flow_to_color (sub_n_pub_lfn3_node.py:132-140) on the device, bit for bit against the
    cv2 recipe applied to the downloaded field — at a width with a 32-pixel vector body and a scalar tail."""
    import cv2
    import opticalflowcontainer_b200 as ofb
    from oracle import visual_np as V
    for (h, w) in [(135, 240), (97, 131)]:
        a, b = synth.synth_warp_pair(h, w, 23, angle_deg=2.0, zoom=1.03)
        eng = ofb.FlowEngine(w, h, 1, 0)
        try:
            flow = eng.farneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
            got = eng.flow_to_color(h, w)
            hsv = np.zeros((h, w, 3), np.uint8)
            hsv[..., 1] = 255
            mag, ang = cv2.cartToPolar(flow[..., 0], flow[..., 1])
            hsv[..., 0] = (ang * 180 / np.pi / 2).astype(np.uint8)
            hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
            assert np.array_equal(got, cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR))
            assert np.array_equal(got, V.flow_to_color(flow))
        finally:
            eng.close()


@pytest.mark.parametrize("size,d,sc,ss", [((120, 160), 9, 75.0, 75.0), ((97, 131), 5, 30.0, 3.0), ((64, 80), 0, 40.0, 2.0)])
def test_bilateral_filter(engine_factory, size, d, sc, ss):
    """ofb_bilateral_u8c3 (adapt node, lfn3_adapt_node.py:186-190): bit-exact with the restatement of OpenCV's own 8UC3
    algorithm; against the installed wheel — which sends 8-bit images through Intel IPP — equal up to one grey level at
    rounding ties in at most 1 value per 10 000 (measured: 13 of 1.9 M on a 1080p frame)."""
    import cv2
    from oracle import prefilter_np as P
    h, w = size
    rng = np.random.default_rng(5)
    base = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 2.0)
    img = np.clip(base.astype(np.int32) + rng.integers(-12, 13, (h, w, 3)), 0, 255).astype(np.uint8)
    eng = engine_factory(w, h)
    got = eng.bilateral_filter(img, d, sc, ss)
    assert np.array_equal(got, P.bilateral_u8c3(img, d, sc, ss))
    ref = cv2.bilateralFilter(img, d, sc, ss)
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff > 0).mean() <= 1e-4, (diff.max(), (diff > 0).mean())


def test_compressed_callback_equals_image_callback_on_decoded_frames(built_lib):
    """CompressedImage (JPEG) frames through the node contract: same velocities as feeding cv2.imdecode's frames to
    image_callback (opticalflow_comprerssed_node.py:41-62)."""
    import cv2
    from opticalflowcontainer_b200.node import FarnebackVelocityNode
    from oracle import synth
    frames = [synth.synth_pair(240, 320, 3 + i, (1.5 + i, -0.5))[0] for i in range(3)]
    jpgs = [cv2.imencode(".jpg", cv2.cvtColor(f, cv2.COLOR_GRAY2BGR), [cv2.IMWRITE_JPEG_QUALITY, 90])[1] for f in frames]
    a = FarnebackVelocityNode(width=320, height=240)
    b = FarnebackVelocityNode(width=320, height=240, engine=a.engine)
    try:
        for i, j in enumerate(jpgs):
            ra = a.compressed_callback(j.tobytes(), 0.1 * i)
            rb = b.image_callback(cv2.imdecode(j, cv2.IMREAD_COLOR), 0.1 * i)
            assert (ra is None) == (rb is None)
            if ra is not None:
                assert ra[0].vector == rb[0].vector and ra[1].vector == rb[1].vector
        assert a.compressed_callback(b"garbage", 1.0) is None
    finally:
        a.engine.close()


def test_junction_detector_node_contract(built_lib):
    """fishnet_detector_ros.cpp:30-80: dampenIntensity(-20, 15) + find_junctions_not_rotated(200, 2.0, false, 6) -> a
    PointCloud with z = 0; nothing for fewer than 4 junctions."""
    import cv2
    from opticalflowcontainer_b200.node import JunctionDetectorNode
    from oracle import junction_np as J
    from oracle import synth
    node = JunctionDetectorNode()
    try:
        img = synth.synth_net(240, 320, 6)
        msg = node.image_callback(img, 12.5, "cam")
        want = J.find_junctions(J.dampen_intensity(img, -20, 15), 200, 2.0, 6)
        assert msg is not None and msg.stamp == 12.5 and msg.frame_id == "cam"
        assert np.array_equal(msg.points[:, :2], want) and not msg.points[:, 2].any()
        assert node.image_callback(np.full((120, 160, 3), 90, np.uint8), 13.0) is None
    finally:
        node.engine.close()


def test_dense_view_and_arrow_grid(built_lib):
    """The sub node's two debug views of the field (lfn3_sub_node.py:225-262) from the device: the speed-scaled HSV image
    bit-exact against the recipe's restatement (pinned against cv2 + NumPy in tests/test_oracle_visual.py) and the arrow
    segments against the node's loop on the downloaded field."""
    import opticalflowcontainer_b200 as ofb
    from opticalflowcontainer_b200 import node
    from oracle import synth
    from oracle import visual_np as V
    h, w = 123, 208
    a, b = synth.synth_pair(h, w, 21, (5.5, -3.25))
    eng = ofb.FlowEngine(w, h, 1, 0)
    try:
        flow = eng.farneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
        for dt, p2m, vmax in ((0.033, 0.0011, 0.25), (0.1, 0.000566, 0.01)):
            assert np.array_equal(eng.flow_to_color_speed(h, w, dt, p2m, vmax), V.flow_to_color_speed(flow, dt, p2m, vmax))
        assert np.array_equal(eng.flow_to_color(h, w), V.flow_to_color(flow))          # (the other mode still holds)
        assert np.array_equal(node.flow_arrows(eng, h, w, 20), V.flow_arrows(flow, 20))
    finally:
        eng.close()
