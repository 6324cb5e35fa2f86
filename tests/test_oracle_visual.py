"""The restatement of the nodes' flow_to_color (oracle/visual_np.py) against the cv2 wheel, bit for bit."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import visual_np as V


def _node_flow_to_color(flow_chw):
    """sub_n_pub_lfn3_node.py:132-140, verbatim arithmetic."""
    h, w = flow_chw.shape[1:]
    hsv = np.zeros((h, w, 3), dtype=np.uint8)
    hsv[..., 1] = 255
    mag, ang = cv2.cartToPolar(flow_chw[0], flow_chw[1])
    hsv[..., 0] = (ang * 180 / np.pi / 2).astype(np.uint8)
    hsv[..., 2] = cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    return cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)


@pytest.mark.parametrize("shape,seed", [((97, 131), 0), ((64, 96), 1), ((33, 250), 2)])
def test_flow_to_color_bit_exact(shape, seed):
    rng = np.random.default_rng(seed)
    f = (rng.standard_normal(shape + (2,)) * 4).astype(np.float32)
    f[0, :5, 0] = 0; f[1, :7, 1] = 0; f[2, :4] = (-1.0, 0.0); f[3, :4] = 0
    mag, ang = cv2.cartToPolar(f[..., 0], f[..., 1])
    m2, a2 = V.cart_to_polar(f[..., 0], f[..., 1])
    assert np.array_equal(mag, m2) and np.array_equal(ang, a2)
    assert np.array_equal(cv2.normalize(mag, None, 0, 255, cv2.NORM_MINMAX), V.normalize_minmax_0_255(mag))
    assert np.array_equal(_node_flow_to_color(np.ascontiguousarray(f.transpose(2, 0, 1))), V.flow_to_color(f))


def test_flow_to_color_constant_field():
    f = np.full((40, 64, 2), 1.5, np.float32)
    assert np.array_equal(_node_flow_to_color(np.ascontiguousarray(f.transpose(2, 0, 1))), V.flow_to_color(f))


def _node_dense_view(flow_chw, dt, pixel_to_meter, max_speed):
    """lfn3_sub_node.py:244-262, verbatim arithmetic."""
    flow_u, flow_v = flow_chw[0], flow_chw[1]
    h, w = flow_u.shape
    mag, ang = cv2.cartToPolar(flow_u, flow_v, angleInDegrees=False)
    mag_mps = (mag / dt) * pixel_to_meter
    mag_norm = np.clip(mag_mps / max_speed, 0.0, 1.0)
    hsv = np.zeros((h, w, 3), dtype=np.uint8)
    hsv[..., 0] = np.uint8((ang * 90.0 / np.pi))
    hsv[..., 1] = 255
    hsv[..., 2] = np.uint8(mag_norm * 255)
    return cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)


@pytest.mark.parametrize("shape,seed,dt,p2m,vmax", [((97, 131), 0, 0.033, 0.0011, 0.25), ((64, 96), 1, 0.1, 0.000566, 0.05),
                                                    ((33, 250), 2, 1e-3, 0.0011, 2.0)])
def test_dense_view_bit_exact(shape, seed, dt, p2m, vmax):
    rng = np.random.default_rng(seed)
    f = (rng.standard_normal(shape + (2,)) * 6).astype(np.float32)
    f[0, :5] = 0
    want = _node_dense_view(np.ascontiguousarray(f.transpose(2, 0, 1)), dt, p2m, vmax)
    assert np.array_equal(V.flow_to_color_speed(f, dt, p2m, vmax), want)
