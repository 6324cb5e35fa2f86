"""Experiment (CPU): how far from cv2 does the flow move if the vertical box sums of
FarnebackUpdateFlow_Blur are done in float32 instead of cv2's double running sums?
Variants: 'double' (the oracle), 'vhgw' (block prefix sums, v = (B_prev - P_prev[k]) + P_cur[k]),
'running' (naive float add/subtract running sum)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
from oracle import farneback_np as F
from oracle import synth

f32 = np.float32


def hsum_f32(M, m):
    h, w = M.shape[:2]
    xi = np.clip(np.arange(-m, w + m), 0, w - 1)
    E = M.astype(f32)[:, xi]
    out = np.zeros_like(M, dtype=f32)
    for d in range(2 * m + 1):
        out = (out + E[:, d:d + w]).astype(f32)
    return out


def v_vhgw(H, m):
    R = 2 * m + 1
    h = H.shape[0]
    yi = np.clip(np.arange(-m, h + m), 0, h - 1)
    E = H[yi]
    n = E.shape[0]
    nb = (n + R - 1) // R
    pad = nb * R - n
    if pad:
        E = np.concatenate([E, np.zeros((pad,) + E.shape[1:], f32)], 0)
    Eb = E.reshape((nb, R) + E.shape[1:])
    P = np.cumsum(Eb, axis=1, dtype=f32)
    Bs = P[:, R - 1]
    out = np.empty_like(H)
    for y in range(h):
        i = y + 2 * m
        b, k = divmod(i, R)
        if k == R - 1:
            out[y] = P[b, k]
        else:
            out[y] = ((Bs[b - 1] - P[b - 1, k]).astype(f32) + P[b, k]).astype(f32)
    return out


def v_running(H, m):
    R = 2 * m + 1
    h = H.shape[0]
    yi = np.clip(np.arange(-m, h + m), 0, h - 1)
    E = H[yi]
    out = np.empty_like(H)
    v = np.zeros(H.shape[1:], f32)
    for i in range(E.shape[0]):
        v = (v + E[i]).astype(f32)
        if i >= R:
            v = (v - E[i - R]).astype(f32)
        if i >= R - 1:
            out[i - (R - 1)] = v
    return out


def make_box(variant):
    def box(M, winsize):
        m = winsize // 2
        H = hsum_f32(M, m)
        V = v_vhgw(H, m) if variant == 'vhgw' else v_running(H, m)
        return (V * f32(1.0 / (winsize * winsize))).astype(f32)
    return box


def epe(a, b):
    d = np.sqrt(((a.astype(np.float64) - b) ** 2).sum(-1))
    return d.mean(), d.max()


def checker(h, w, seed):
    rng = np.random.default_rng(seed)
    a = np.full((h, w), 20, np.uint8)
    for _ in range(40):
        y, x = rng.integers(0, h - 40), rng.integers(0, w - 40)
        a[y:y + rng.integers(8, 40), x:x + rng.integers(8, 40)] = rng.integers(0, 2) * 235 + 10
    a = (a + rng.integers(0, 3, size=a.shape)).astype(np.uint8)
    b = np.roll(a, (2, -3), axis=(0, 1))
    return a, b


cases = {
    'vga_shift': synth.synth_pair(240, 320, 0, (1.7, -0.9)),
    'low_texture': synth.low_texture_pair(240, 320, 3),
    'checker': checker(240, 320, 5),
    'warp': synth.synth_warp_pair(240, 320, 31),
}
orig = F.box_blur
for name, (a, b) in cases.items():
    ref = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    F.box_blur = orig
    print(name, 'double  ', '%.2e %.2e' % epe(F.farneback(a, b), ref))
    for v in ('vhgw', 'running'):
        F.box_blur = make_box(v)
        print(name, v.ljust(8), '%.2e %.2e' % epe(F.farneback(a, b), ref))
F.box_blur = orig
