"""Host-side node contract (no GPU): FarnebackVelocityNode around a stub engine that answers with cv2's flow — the
logic every reference node wraps around its flow call (lfn3_sub_node.py:141-222, opticalflow_node.py:41-128,
sub_n_pub_lfn3_node.py:195-210, lfn3_junction_node.py:203-231)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from opticalflowcontainer_b200 import node as N
from oracle import synth


class StubEngine:
    """Duck-typed FlowEngine: the flow call is cv2's, everything else NumPy."""

    def __init__(self):
        self.flow = None

    def farneback(self, prev, nxt, flow, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags):
        self.flow = cv2.calcOpticalFlowFarneback(prev, nxt, None, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)
        return self.flow

    def ingest_gray(self, frame, size, rgb=False):
        if (frame.shape[1], frame.shape[0]) != tuple(size):
            frame = cv2.resize(frame, tuple(size))
        return cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY if rgb else cv2.COLOR_BGR2GRAY)

    def resize(self, img, size):
        return cv2.resize(img, tuple(size))

    def flow_sample(self, pts, pair=0):
        pts = np.asarray(pts).reshape(-1, 2)
        out = np.full((len(pts), 2), np.nan, np.float32)
        h, w = self.flow.shape[:2]
        for i, (x, y) in enumerate(pts):
            if 0 <= x < w and 0 <= y < h:
                out[i] = self.flow[int(y), int(x)]
        return out


def _frames(n, h=96, w=128):
    base = synth.synth_pair(h, w, 3, (0.0, 0.0))[0]
    return [synth.subpixel_shift(base, 1.5 * i, -0.5 * i) for i in range(n)]


@pytest.mark.parametrize("reduce", ["median", "mean"])
def test_node_contract(reduce):
    fr = _frames(8)
    node = N.FarnebackVelocityNode(engine=StubEngine(), width=128, height=96, reduce=reduce, pixel_to_meter=0.002)
    stamps = [0.0, 0.1, 0.1, 0.05, 0.3, 0.4, 0.5, 0.6]          # a repeated and a backwards stamp: dt <= 0 -> 1e-3
    vx_all = []
    prev = None
    for f, t in zip(fr, stamps):
        out = node.image_callback(f, t, encoding="mono8")
        if prev is None:
            assert out is None                                   # first frame only primes (lfn3_sub_node.py:164-167)
            prev = (f, t)
            continue
        dt = t - prev[1]
        if dt <= 0:
            dt = 1e-3
        flow = cv2.calcOpticalFlowFarneback(prev[0], f, None, 0.5, 3, 15, 3, 5, 1.2, 0)
        u = np.transpose(flow, (2, 0, 1))[0]
        vx = float(np.median(u) if reduce == "median" else np.mean(u)) / dt * 0.002   # (float64 arithmetic, as NumPy 1.x nodes)
        vx_all.append(vx)
        raw, smooth = out
        assert raw.frame_id == "camera_link" and raw.stamp == t and raw.vector[1:] == (0.0, 0.0)
        assert raw.vector[0] == pytest.approx(vx, rel=1e-12, abs=1e-15)
        assert smooth.vector[0] == pytest.approx(float(np.mean(vx_all[-5:])), rel=1e-12, abs=1e-15)   # deque(maxlen=5)
        prev = (f, t)


def test_node_masked_median_and_colour_ingest():
    fr = _frames(2)
    bgr = [np.dstack([f, f, f]) for f in fr]
    big = [cv2.resize(b, (256, 192)) for b in bgr]
    node = N.FarnebackVelocityNode(engine=StubEngine(), width=128, height=96)
    mask = N.junction_mask([(20.7, 30.2), (100, 50), (500, 500)], 96, 128, radius=5)
    assert mask.sum() == 2 * 11 * 11 and mask[30, 20] and not mask[0, 0]
    assert node.image_callback(big[0], 0.0) is None
    raw, _ = node.image_callback(big[1], 0.1, mask=mask)
    g0 = cv2.cvtColor(cv2.resize(big[0], (128, 96)), cv2.COLOR_BGR2GRAY)
    g1 = cv2.cvtColor(cv2.resize(big[1], (128, 96)), cv2.COLOR_BGR2GRAY)
    u = cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)[..., 0]
    assert raw.vector[0] == pytest.approx(float(np.median(u[mask])) / 0.1 * node.pixel_to_meter, rel=1e-12)
    with pytest.raises(ValueError):
        node.image_callback(big[1], 0.2, encoding="yuv422")


def test_to_gray_matches_cv2():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    assert np.array_equal(N.to_gray_u8(img, "bgr8"), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    assert np.array_equal(N.to_gray_u8(img, "rgb8"), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY))


def test_junction_velocity_logic():
    eng = StubEngine()
    h, w = 96, 128
    eng.flow = np.zeros((h, w, 2), np.float32)
    eng.flow[..., 0] = 2.5
    eng.flow[..., 1] = -1.0
    prev = np.array([[10.2, 10.9], [50.0, 40.0], [100.5, 80.5], [64.0, 20.0], [-5.0, 3.0]])
    curr = prev[:4] + [2.4, -1.1]
    vx = N.junction_velocity(eng, prev, curr, 0.05, 0.001)
    assert vx == pytest.approx(2.4 / 0.05 * 0.001, rel=1e-9)
    assert N.junction_velocity(eng, prev, curr + 100.0, 0.05, 0.001) is None       # nothing within 5 px
    assert N.junction_velocity(eng, prev[:3], curr, 0.05, 0.001) is None           # fewer than 4 matches
    assert N.junction_velocity(eng, [], curr, 0.05, 0.001) is None
