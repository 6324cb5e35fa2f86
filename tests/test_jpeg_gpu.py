"""JPEG frame ingest on the device (ofb_jpeg_decode / ofb_ingest_jpeg_gray) against ``cv2.imdecode`` of the wheel:
bit-exact BGR, gray and resized-gray frames.  Anchor: opticalflow_comprerssed_node.py:43-46."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from test_oracle_jpeg import encode, jpeg_frame  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,sampling,quality,rst", [
    ((48, 64), "420", 95, 0), ((37, 53), "420", 50, 3), ((37, 53), "422", 90, 0), ((16, 16), "444", 10, 0),
    ((9, 21), "440", 75, 2), ((481, 637), "420", 85, 0), ((1080, 1920), "420", 90, 0), ((1080, 1920), "422", 60, 16),
    ((539, 959), "444", 97, 0)])
def test_imdecode_equals_cv2(engine_factory, size, sampling, quality, rst):
    h, w = size
    buf = encode(jpeg_frame(h, w, quality), sampling, quality, rst)
    ref = cv2.imdecode(buf, cv2.IMREAD_COLOR)
    eng = engine_factory(64, 64)
    assert eng.jpeg_info(buf) == (w, h, 3)
    got = eng.imdecode(buf)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert np.array_equal(eng.imdecode(buf.tobytes(), gray=True), cv2.cvtColor(ref, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("size,sampling,quality", [((1080, 1920), "420", 90), ((481, 637), "422", 35), ((539, 959), "444", 98),
                                                   ((64, 4000), "420", 75), ((2160, 3840), "420", 80)])
def test_device_and_host_entropy_decoders_agree(engine_factory, size, sampling, quality):
    """Scans without restart intervals are Huffman-decoded on the device (self-synchronising subsequences); forcing the
    host walk must give the same frame, and both equal cv2."""
    h, w = size
    buf = encode(jpeg_frame(h, w, quality + 1), sampling, quality, 0)
    ref = cv2.imdecode(buf, cv2.IMREAD_COLOR)
    eng = engine_factory(64, 64)
    dev = eng.imdecode(buf)
    eng.jpeg_host_entropy(True)
    try:
        host = eng.imdecode(buf)
    finally:
        eng.jpeg_host_entropy(False)
    assert np.array_equal(dev, ref) and np.array_equal(host, ref)


def test_device_entropy_on_smooth_and_flat_frames(engine_factory):
    """Few bits per block (flat frames: an MCU is a handful of bits, a subsequence holds dozens of blocks) and a gray JPEG."""
    eng = engine_factory(64, 64)
    flat = np.full((240, 320, 3), 77, np.uint8)
    ramp = np.dstack([np.tile(np.linspace(0, 255, 640).astype(np.uint8), (480, 1))] * 3)
    for img in (flat, ramp, ramp[..., 0]):
        for q in (95, 20):
            buf = encode(img, "420", q, 0)
            assert np.array_equal(eng.imdecode(buf), cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_module_level_imdecode_color_and_grayscale(built_lib):
    """cv2.imdecode's two flags the nodes could pass; None for an undecodable stream, like cv2."""
    import opticalflowcontainer_b200 as ofb
    for (h, w), samp in (((97, 131), "420"), ((64, 80), "444"), ((50, 66), "422")):
        buf = encode(jpeg_frame(h, w, h), samp, 88, 0)
        assert np.array_equal(ofb.imdecode(buf, ofb.IMREAD_COLOR), cv2.imdecode(buf, cv2.IMREAD_COLOR))
        assert np.array_equal(ofb.imdecode(buf, ofb.IMREAD_GRAYSCALE), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE))
    g = encode(jpeg_frame(40, 56, 3, channels=1), quality=80)
    assert np.array_equal(ofb.imdecode(g, ofb.IMREAD_GRAYSCALE), cv2.imdecode(g, cv2.IMREAD_GRAYSCALE))
    ok, prog = cv2.imencode(".jpg", jpeg_frame(32, 32, 1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert ofb.imdecode(prog) is None


def test_gray_jpeg_and_repeated_calls(engine_factory):
    eng = engine_factory(64, 64)
    for seed, (h, w) in enumerate([(40, 56), (200, 312), (33, 35)]):      # staging grows and is reused
        buf = encode(jpeg_frame(h, w, seed, channels=1), quality=80)
        assert np.array_equal(eng.imdecode(buf), cv2.imdecode(buf, cv2.IMREAD_COLOR))
    buf = encode(jpeg_frame(64, 64, 9), "420", 70)
    assert np.array_equal(eng.imdecode(buf), cv2.imdecode(buf, cv2.IMREAD_COLOR))


@pytest.mark.parametrize("size,dst", [((480, 640), None), ((1080, 1920), (640, 480)), ((300, 400), (512, 384))])
def test_ingest_jpeg_gray_equals_node_chain(engine_factory, size, dst):
    """decode -> resize (colour frame, if the size differs) -> gray, as lfn3_sub_node.py:148-159 does after imdecode."""
    buf = encode(jpeg_frame(size[0], size[1], 4), "420", 88)
    ref = cv2.imdecode(buf, cv2.IMREAD_COLOR)
    if dst is not None:
        ref = cv2.resize(ref, dst)
    ref = cv2.cvtColor(ref, cv2.COLOR_BGR2GRAY)
    eng = engine_factory(64, 64)
    assert np.array_equal(eng.ingest_jpeg_gray(buf, dst), ref)


def test_decoded_frames_feed_the_flow_call(engine_factory):
    """Two JPEG frames -> gray on the device -> Farneback: same field as cv2 on cv2-decoded frames."""
    from oracle import synth
    a, b = synth.synth_pair(240, 320, 2, (2.5, -1.25))
    ja, jb = (cv2.imencode(".jpg", cv2.cvtColor(x, cv2.COLOR_GRAY2BGR), [cv2.IMWRITE_JPEG_QUALITY, 92])[1] for x in (a, b))
    eng = engine_factory(320, 240)
    ga, gb = eng.ingest_jpeg_gray(ja), eng.ingest_jpeg_gray(jb)
    ra, rb = (cv2.cvtColor(cv2.imdecode(j, cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY) for j in (ja, jb))
    assert np.array_equal(ga, ra) and np.array_equal(gb, rb)
    got = eng.farneback(ga, gb, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    ref = cv2.calcOpticalFlowFarneback(ra, rb, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    epe = np.sqrt(((got - ref) ** 2).sum(-1))
    assert epe.mean() <= 0.01 and epe.max() <= 0.1


def test_unsupported_streams_fail_loudly(engine_factory):
    from opticalflowcontainer_b200 import OfbError
    eng = engine_factory(64, 64)
    ok, prog = cv2.imencode(".jpg", jpeg_frame(32, 32, 1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(OfbError) as e:
        eng.imdecode(prog)
    assert e.value.status == 6
    with pytest.raises(OfbError):
        eng.imdecode(b"\xff\xd8\xff\xd9")


def test_corrupt_streams_do_not_take_the_device_down(engine_factory):
    """Bit flips, truncation and garbage in the entropy-coded segment: the call returns a frame of the right size or an
    OfbError — never a CUDA fault — and the next good frame still decodes bit-exactly (both entropy decoders)."""
    from opticalflowcontainer_b200 import OfbError
    eng = engine_factory(64, 64)
    good = encode(jpeg_frame(120, 176, 2), "420", 85, 0)
    ref = cv2.imdecode(good, cv2.IMREAD_COLOR)
    rng = np.random.default_rng(11)
    sos = good.tobytes().rfind(b"\xff\xda")
    for host in (False, True):
        eng.jpeg_host_entropy(host)
        try:
            for trial in range(24):
                bad = good.copy()
                if trial % 3 == 0:
                    bad = bad[: int(rng.integers(sos + 20, bad.size - 2))]                      # truncated inside the scan
                elif trial % 3 == 1:
                    for k in rng.integers(sos + 14, bad.size - 2, 12):
                        bad[k] ^= 1 << int(rng.integers(0, 8))                                   # bit flips
                else:
                    a = int(rng.integers(sos + 14, bad.size - 40))
                    bad[a:a + 32] = rng.integers(0, 255, 32, dtype=np.uint8)                     # a burst of garbage
                try:
                    out = eng.imdecode(np.ascontiguousarray(bad))
                    assert out.shape == ref.shape
                except OfbError:
                    pass
                assert np.array_equal(eng.imdecode(good), ref)
        finally:
            eng.jpeg_host_entropy(False)
