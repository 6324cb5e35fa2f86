"""ORACLE (test infrastructure only — never imported by the product path).

NumPy restatement of ``cv2.imdecode(buf, cv2.IMREAD_COLOR)`` for baseline JPEG, the first step of the reference's
compressed-image node (``ros2_ws/src/optical_flow/optical_flow/opticalflow_comprerssed_node.py:43-46``).  The
arithmetic lives in the wheel's bundled **libjpeg-turbo 3.1.2** (``cv2.getBuildInformation()``), driven by OpenCV's
``grfmt_jpeg.cpp`` with the library defaults: ``dct_method = JDCT_ISLOW``, ``do_fancy_upsampling = TRUE``,
``out_color_space = JCS_EXT_BGR``.  Restated from the published algorithm:

* entropy decoding (ITU-T T.81 F.2.2): Huffman tables from DHT, DC prediction per component, restart intervals;
* ``jidctint.c::jpeg_idct_islow`` — the 13-bit fixed-point LL&M inverse DCT, columns then rows, ``PASS1_BITS = 2``,
  dequantisation folded into the column pass, result ``clamp(descale(x, 18) + 128)``;
* ``jdsample.c`` — ``h2v1_fancy_upsample`` (3/4, 1/4 with biases 1, 2), ``h2v2_fancy_upsample`` (9/16, 3/16, 3/16, 1/16
  with biases 8, 7) and ``h1v2_fancy_upsample``; chroma planes replicate their last real row/column
  (``jdmainct.c::set_bottom_pointers``), the column ends use the undivided neighbour;
* ``jdcolor.c::ycc_rgb_convert`` — 16-bit fixed-point tables (FIX(1.40200) …), ``G = Y + ((Cb_g[cb] + Cr_g[cr]) >> 16)``.

Pinned bit for bit against the wheel in ``tests/test_oracle_jpeg.py`` on frames encoded by ``cv2.imencode`` with 4:2:0,
4:2:2, 4:4:4 and 4:4:0 sampling, grey-scale, odd sizes, several qualities and restart intervals.
The Huffman loop is pure Python: small frames only.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63], np.int64)


class JpegError(ValueError):
    pass


def parse(buf: bytes):
    """Markers of a baseline (SOF0 / SOF1, 8-bit, Huffman, one interleaved scan) JPEG -> dict."""
    b = bytes(buf)
    if len(b) < 4 or b[0] != 0xFF or b[1] != 0xD8:
        raise JpegError("not a JPEG")
    pos = 2
    qt = {}
    huff = {}
    frame = None
    dri = 0
    while True:
        while pos < len(b) and b[pos] != 0xFF:
            pos += 1
        while pos < len(b) and b[pos] == 0xFF:
            pos += 1
        if pos >= len(b):
            raise JpegError("no scan")
        m = b[pos]
        pos += 1
        if m == 0xD8 or (0xD0 <= m <= 0xD7) or m == 0x01:
            continue
        ln = (b[pos] << 8) | b[pos + 1]
        seg = b[pos + 2: pos + ln]
        if m == 0xDB:
            i = 0
            while i < len(seg):
                pq, tq = seg[i] >> 4, seg[i] & 15
                i += 1
                if pq:
                    t = np.frombuffer(seg[i:i + 128], ">u2").astype(np.int64)
                    i += 128
                else:
                    t = np.frombuffer(seg[i:i + 64], np.uint8).astype(np.int64)
                    i += 64
                nat = np.zeros(64, np.int64)
                nat[ZIGZAG] = t
                qt[tq] = nat
        elif m == 0xC4:
            i = 0
            while i < len(seg):
                tc, th = seg[i] >> 4, seg[i] & 15
                counts = list(seg[i + 1:i + 17])
                n = sum(counts)
                vals = list(seg[i + 17:i + 17 + n])
                i += 17 + n
                huff[(tc, th)] = (counts, vals)
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise JpegError("precision")
            h, w, nc = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4], seg[5]
            comps = []
            for c in range(nc):
                cid, hv, tq = seg[6 + 3 * c], seg[7 + 3 * c], seg[8 + 3 * c]
                comps.append(dict(id=cid, h=hv >> 4, v=hv & 15, tq=tq))
            frame = dict(width=w, height=h, comps=comps)
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise JpegError("not baseline (SOF%d)" % (m - 0xC0))
        elif m == 0xDD:
            dri = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            if frame is None:
                raise JpegError("SOS before SOF")
            ns = seg[0]
            if ns != len(frame["comps"]):
                raise JpegError("non-interleaved scans")
            for c in range(ns):
                cid, tt = seg[1 + 2 * c], seg[2 + 2 * c]
                comp = [k for k in frame["comps"] if k["id"] == cid][0]
                comp["td"], comp["ta"] = tt >> 4, tt & 15
            frame.update(qt=qt, huff=huff, dri=dri, data=b[pos + ln:])
            return frame
        pos += ln


def _build_lookup(counts, vals):
    """code length / value lookup keyed by (length, code)."""
    table = {}
    code = 0
    k = 0
    for ln in range(1, 17):
        for _ in range(counts[ln - 1]):
            table[(ln, code)] = vals[k]
            code += 1
            k += 1
        code <<= 1
    return table


class _Bits:
    def __init__(self, data: bytes):
        self.d = data
        self.p = 0
        self.acc = 0
        self.n = 0

    def _fill(self):
        if self.p < len(self.d):
            v = self.d[self.p]
            if v == 0xFF:
                nxt = self.d[self.p + 1] if self.p + 1 < len(self.d) else 0xD9
                if nxt == 0:
                    self.p += 2
                else:
                    v = 0                     # marker: feed zeros (as libjpeg does)
            else:
                self.p += 1
        else:
            v = 0
        self.acc = (self.acc << 8) | v
        self.n += 8

    def bit(self):
        if self.n == 0:
            self._fill()
        self.n -= 1
        return (self.acc >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def restart(self):
        self.n = 0
        self.acc = 0
        while self.p + 1 < len(self.d) and not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _decode_symbol(bits, table):
    code = 0
    for ln in range(1, 17):
        code = (code << 1) | bits.bit()
        v = table.get((ln, code))
        if v is not None:
            return v
    raise JpegError("bad Huffman code")


def _extend(v, t):
    return v - ((1 << t) - 1) if t and v < (1 << (t - 1)) else v


def decode_coefficients(fr):
    """-> per component: int array [blocks_y, blocks_x, 64] of quantised coefficients in natural order."""
    comps = fr["comps"]
    hmax = max(c["h"] for c in comps)
    vmax = max(c["v"] for c in comps)
    if len(comps) == 1:
        comps[0]["h"] = comps[0]["v"] = 1      # a single-component scan is never interleaved: MCU = one block
        hmax = vmax = 1
    mcux = -(-fr["width"] // (8 * hmax))
    mcuy = -(-fr["height"] // (8 * vmax))
    out = [np.zeros((mcuy * c["v"], mcux * c["h"], 64), np.int64) for c in comps]
    tabs = {k: _build_lookup(*v) for k, v in fr["huff"].items()}
    bits = _Bits(fr["data"])
    pred = [0] * len(comps)
    n = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if fr["dri"] and n and n % fr["dri"] == 0:
                bits.restart()
                pred = [0] * len(comps)
            n += 1
            for ci, c in enumerate(comps):
                dct, act = tabs[(0, c["td"])], tabs[(1, c["ta"])]
                for by in range(c["v"]):
                    for bx in range(c["h"]):
                        blk = out[ci][my * c["v"] + by, mx * c["h"] + bx]
                        t = _decode_symbol(bits, dct)
                        pred[ci] += _extend(bits.bits(t), t)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _decode_symbol(bits, act)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            if k > 63:
                                raise JpegError("coefficient index out of range")
                            blk[ZIGZAG[k]] = _extend(bits.bits(s), s)
                            k += 1
    return out, (hmax, vmax)


# ---- jidctint.c: jpeg_idct_islow ----
_F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
          f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _idct_1d(x0, x1, x2, x3, x4, x5, x6, x7, shift, first):
    """one pass of the LL&M network over eight int64 arrays; `first`: even part scaled by 2**13, descale by `shift`."""
    F = _F
    z2, z3 = x2, x6
    z1 = (z2 + z3) * F["f0_541"]
    tmp2 = z1 + z3 * (-F["f1_847"])
    tmp3 = z1 + z2 * F["f0_765"]
    tmp0 = (x0 + x4) << 13
    tmp1 = (x0 - x4) << 13
    tmp10, tmp13 = tmp0 + tmp3, tmp0 - tmp3
    tmp11, tmp12 = tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = x7, x5, x3, x1
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F["f1_175"]
    t0 = t0 * F["f0_298"]
    t1 = t1 * F["f2_053"]
    t2 = t2 * F["f3_072"]
    t3 = t3 * F["f1_501"]
    z1 = z1 * (-F["f0_899"])
    z2 = z2 * (-F["f2_562"])
    z3 = z3 * (-F["f1_961"]) + z5
    z4 = z4 * (-F["f0_390"]) + z5
    t0 = t0 + z1 + z3
    t1 = t1 + z2 + z4
    t2 = t2 + z2 + z3
    t3 = t3 + z1 + z4
    r = 1 << (shift - 1)
    return [(tmp10 + t3 + r) >> shift, (tmp11 + t2 + r) >> shift, (tmp12 + t1 + r) >> shift, (tmp13 + t0 + r) >> shift,
            (tmp13 - t0 + r) >> shift, (tmp12 - t1 + r) >> shift, (tmp11 - t2 + r) >> shift, (tmp10 - t3 + r) >> shift]


def idct_islow(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """coef [..., 64] quantised (natural order), q [64] -> uint8 samples [..., 8, 8]."""
    d = (coef * q).reshape(coef.shape[:-1] + (8, 8)).astype(np.int64)
    cols = _idct_1d(*[d[..., r, :] for r in range(8)], shift=11, first=True)          # column pass: rows of the block
    ws = np.stack(cols, -2)                                                           # [..., 8 (row), 8 (col)]
    rows = _idct_1d(*[ws[..., :, c] for c in range(8)], shift=18, first=False)
    out = np.stack(rows, -1) + 128
    return np.clip(out, 0, 255).astype(np.uint8)


def component_planes(fr):
    coefs, (hmax, vmax) = decode_coefficients(fr)
    planes = []
    for c, cf in zip(fr["comps"], coefs):
        s = idct_islow(cf, fr["qt"][c["tq"]])                                         # [by, bx, 8, 8]
        by, bx = s.shape[:2]
        full = s.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)
        dw = -(-fr["width"] * c["h"] // hmax)
        dh = -(-fr["height"] * c["v"] // vmax)
        planes.append(full[:dh, :dw])
    return planes, (hmax, vmax)


# ---- jdsample.c ----
def _h2v1_fancy(p: np.ndarray) -> np.ndarray:
    p = p.astype(np.int64)
    h, w = p.shape
    out = np.zeros((h, 2 * w), np.int64)
    prev = np.concatenate([p[:, :1], p[:, :-1]], 1)
    nxt = np.concatenate([p[:, 1:], p[:, -1:]], 1)
    out[:, 0::2] = (3 * p + prev + 1) >> 2
    out[:, 1::2] = (3 * p + nxt + 2) >> 2
    out[:, 0] = p[:, 0]
    out[:, -1] = p[:, -1]
    return out


def _h2v2_fancy(p: np.ndarray) -> np.ndarray:
    p = p.astype(np.int64)
    h, w = p.shape
    up = np.concatenate([p[:1], p[:-1]], 0)
    dn = np.concatenate([p[1:], p[-1:]], 0)
    out = np.zeros((2 * h, 2 * w), np.int64)
    for v, other in ((0, up), (1, dn)):
        cs = 3 * p + other                                                            # column sums
        last = np.concatenate([cs[:, :1], cs[:, :-1]], 1)
        nxt = np.concatenate([cs[:, 1:], cs[:, -1:]], 1)
        ev = (3 * cs + last + 8) >> 4
        od = (3 * cs + nxt + 7) >> 4
        ev[:, 0] = (4 * cs[:, 0] + 8) >> 4
        od[:, -1] = (4 * cs[:, -1] + 7) >> 4
        out[v::2, 0::2] = ev
        out[v::2, 1::2] = od
    return out


def _h1v2_fancy(p: np.ndarray) -> np.ndarray:
    p = p.astype(np.int64)
    up = np.concatenate([p[:1], p[:-1]], 0)
    dn = np.concatenate([p[1:], p[-1:]], 0)
    out = np.zeros((2 * p.shape[0], p.shape[1]), np.int64)
    out[0::2] = (3 * p + up + 1) >> 2
    out[1::2] = (3 * p + dn + 2) >> 2
    return out


def upsample(p: np.ndarray, fh: int, fv: int, width: int, height: int) -> np.ndarray:
    if (fh, fv) == (1, 1):
        o = p.astype(np.int64)
    elif (fh, fv) == (2, 1):
        o = _h2v1_fancy(p) if p.shape[1] > 2 else np.repeat(p.astype(np.int64), 2, 1)
    elif (fh, fv) == (2, 2):
        o = _h2v2_fancy(p) if p.shape[1] > 2 else np.repeat(np.repeat(p.astype(np.int64), 2, 0), 2, 1)
    elif (fh, fv) == (1, 2):
        o = _h1v2_fancy(p)
    else:
        raise JpegError("sampling factors %dx%d" % (fh, fv))
    return o[:height, :width]


# ---- jdcolor.c ----
def ycc_to_bgr(y, cb, cr) -> np.ndarray:
    x_cb, x_cr = cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    r = y + ((91881 * x_cr + 32768) >> 16)
    b = y + ((116130 * x_cb + 32768) >> 16)
    g = y + ((-22554 * x_cb + 32768 - 46802 * x_cr) >> 16)
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def imdecode_color(buf) -> np.ndarray:
    """cv2.imdecode(buf, cv2.IMREAD_COLOR) for a baseline JPEG."""
    fr = parse(bytes(buf))
    planes, (hmax, vmax) = component_planes(fr)
    w, h = fr["width"], fr["height"]
    if len(planes) == 1:
        y = planes[0][:h, :w]
        return np.stack([y, y, y], -1)
    if len(planes) != 3:
        raise JpegError("component count")
    full = [upsample(p, hmax // c["h"], vmax // c["v"], w, h) for p, c in zip(planes, fr["comps"])]
    return ycc_to_bgr(*full)
