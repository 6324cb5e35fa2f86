"""The JPEG restatement (oracle/jpeg_np.py: entropy decoding, islow IDCT, fancy up-sampling, YCbCr -> BGR) equals
``cv2.imdecode(buf, cv2.IMREAD_COLOR)`` of the wheel (libjpeg-turbo 3.1.2) bit for bit, and the host half of the product
decoder (``ofb_jpeg_entropy_decode``: marker parsing + Huffman decoding, no device) equals the restatement's
coefficients.  Anchor: ros2_ws/src/optical_flow/optical_flow/opticalflow_comprerssed_node.py:43-46."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import jpeg_np as J  # noqa: E402

SAMPLING = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
            "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440}


def jpeg_frame(h, w, seed, channels=3):
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.integers(0, 256, (h, w, channels), dtype=np.uint8), (0, 0), 1.5)
    img = np.clip(base.astype(int).reshape(h, w, channels) * 2 - 128 + rng.integers(-20, 21, (h, w, channels)), 0, 255)
    return img.astype(np.uint8) if channels == 3 else img.astype(np.uint8)[..., 0]


def encode(img, sampling="420", quality=90, rst=0):
    ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                         SAMPLING[sampling], cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
    assert ok
    return buf


CASES = [((48, 64), "420", 95, 0), ((37, 53), "420", 50, 3), ((37, 53), "422", 90, 0), ((16, 16), "444", 10, 0),
         ((9, 21), "440", 75, 2), ((33, 47), "420", 10, 0), ((40, 40), "444", 100, 5), ((25, 70), "422", 30, 1)]


@pytest.mark.parametrize("size,sampling,quality,rst", CASES)
def test_restatement_equals_imdecode(size, sampling, quality, rst):
    buf = encode(jpeg_frame(size[0], size[1], quality + rst), sampling, quality, rst)
    assert np.array_equal(J.imdecode_color(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_restatement_gray_jpeg():
    buf = encode(jpeg_frame(40, 56, 3, channels=1), quality=80)
    assert np.array_equal(J.imdecode_color(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_progressive_is_refused_by_the_restatement():
    ok, buf = cv2.imencode(".jpg", jpeg_frame(32, 32, 1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(J.JpegError):
        J.parse(buf.tobytes())


def _host_coefficients(lib, buf):
    n = C.c_size_t()
    assert lib.ofb_jpeg_entropy_decode(buf.ctypes.data, buf.size, None, 0, C.byref(n)) == 0
    out = np.empty(n.value, np.int16)
    assert lib.ofb_jpeg_entropy_decode(buf.ctypes.data, buf.size, out.ctypes.data, out.size, None) == 0
    return out


@pytest.mark.parametrize("size,sampling,quality,rst", CASES + [((120, 200), "420", 85, 0), ((64, 64), "420", 100, 7)])
def test_host_entropy_decoder_equals_restatement(built_lib, size, sampling, quality, rst):
    """The product's host half (no GPU needed): same coefficient blocks as the restatement's Huffman walk."""
    from opticalflowcontainer_b200 import _lib
    lib = _lib.load()
    buf = encode(jpeg_frame(size[0], size[1], quality), sampling, quality, rst)
    want, _ = J.decode_coefficients(J.parse(buf.tobytes()))
    got = _host_coefficients(lib, buf)
    assert np.array_equal(got, np.concatenate([w.reshape(-1) for w in want]).astype(np.int16))


def test_host_decoder_header_and_refusals(built_lib):
    from opticalflowcontainer_b200 import _lib
    lib = _lib.load()
    buf = encode(jpeg_frame(37, 53, 0), "420", 90)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    assert lib.ofb_jpeg_info(buf.ctypes.data, buf.size, C.byref(w), C.byref(h), C.byref(c)) == 0
    assert (w.value, h.value, c.value) == (53, 37, 3)
    ok, prog = cv2.imencode(".jpg", jpeg_frame(32, 32, 1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    n = C.c_size_t()
    assert lib.ofb_jpeg_entropy_decode(prog.ctypes.data, prog.size, None, 0, C.byref(n)) == 6      # OFB_ERR_UNSUPPORTED
    junk = np.frombuffer(b"not a jpeg at all", np.uint8)
    assert lib.ofb_jpeg_info(junk.ctypes.data, junk.size, C.byref(w), C.byref(h), C.byref(c)) == 6
    trunc = np.ascontiguousarray(buf[:200])                                                       # cut inside the tables
    assert lib.ofb_jpeg_entropy_decode(trunc.ctypes.data, trunc.size, None, 0, C.byref(n)) == 6
    # a stream cut inside the scan still decodes (zeros are fed, as libjpeg does): no crash, same size
    cut = np.ascontiguousarray(buf[: buf.size - 40])
    out = np.empty(_host_coefficients(lib, buf).size, np.int16)
    assert lib.ofb_jpeg_entropy_decode(cut.ctypes.data, cut.size, out.ctypes.data, out.size, None) == 0
