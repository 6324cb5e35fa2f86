"""Host-side engine object over the C ABI: one handle = one GPU + one CUDA stream.

``FlowEngine`` owns an ``ofb_handle``; its methods take/return NumPy arrays (host path, what a
ROS node calls) or raw device pointers / torch CUDA tensors (device path, used by ``bench.py``).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import FarnebackParams, GfttParams, JunctionParams, LKParams, OfbError

OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_LK_GET_MIN_EIGENVALS = 8
OPTFLOW_FARNEBACK_GAUSSIAN = 256


def _u8_image(a, name):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise OfbError(1, "%s must be a single-channel uint8 image [H,W] (got %s %s)" % (name, a.dtype, a.shape))
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


class FlowEngine:
    """A libofb handle.  Not re-entrant: calls are serialised with a lock (nodes call from one
    thread at a time, possibly a non-main thread — ``lfn3_node.py:84-89``)."""

    def __init__(self, max_width: int, max_height: int, max_batch: int = 1, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        st = self._lib.ofb_create(int(device), int(max_width), int(max_height), int(max_batch), C.byref(h))
        _lib.check(st, None)
        self._h = h
        self.max_width, self.max_height, self.max_batch, self.device = max_width, max_height, max_batch, device
        self._lock = threading.Lock()
        self._pending = []   # arrays the library still writes into (asynchronous reductions), released by wait()

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.ofb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self) -> int:
        return int(self._lib.ofb_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.ofb_launch_count(self._h))

    STAGES = ("pyramid", "polyexp", "iteration", "flow_init", "other")

    def timing_enable(self, on: bool = True):
        _lib.check(self._lib.ofb_timing_enable(self._h, 1 if on else 0), self._h)

    def timing_read(self):
        """{stage: (ms_total, launches)} of the CUDA-event stage timers since timing_enable()."""
        ms = (C.c_double * len(self.STAGES))()
        cnt = (C.c_uint64 * len(self.STAGES))()
        _lib.check(self._lib.ofb_timing_read(self._h, ms, cnt), self._h)
        return {s: (ms[i], int(cnt[i])) for i, s in enumerate(self.STAGES)}

    def timing_samples(self, stage: str):
        """Per-launch event times (ms, launch order) of one stage since timing_enable()."""
        idx = self.STAGES.index(stage)
        n = C.c_int(0)
        cap = 1 << 16
        buf = (C.c_double * cap)()
        _lib.check(self._lib.ofb_timing_read_samples(self._h, idx, buf, cap, C.byref(n)), self._h)
        return [buf[i] for i in range(min(n.value, cap))]

    def synchronize(self):
        _lib.check(self._lib.ofb_synchronize(self._h), self._h)

    # -- dense
    @staticmethod
    def _fb_params(pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags) -> FarnebackParams:
        return FarnebackParams(float(pyr_scale), int(levels), int(winsize), int(iterations), int(poly_n),
                               float(poly_sigma), int(flags))

    def farneback(self, prev, nxt, flow=None, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5,
                  poly_sigma=1.2, flags=0) -> np.ndarray:
        """cv2.calcOpticalFlowFarneback with host arrays → float32 [H,W,2]."""
        prev = _u8_image(prev, "prev")
        nxt = _u8_image(nxt, "next")
        if prev.shape != nxt.shape:
            raise OfbError(1, "prev and next must have the same size")
        hgt, wid = prev.shape
        if flags & OPTFLOW_USE_INITIAL_FLOW:
            if flow is None or np.asarray(flow).shape != (hgt, wid, 2) or np.asarray(flow).dtype != np.float32:
                raise OfbError(1, "OPTFLOW_USE_INITIAL_FLOW needs a float32 [H,W,2] flow")
            out = np.ascontiguousarray(flow)
        else:
            if (isinstance(flow, np.ndarray) and flow.shape == (hgt, wid, 2) and flow.dtype == np.float32
                    and flow.flags.c_contiguous):
                out = flow
            else:
                out = np.empty((hgt, wid, 2), np.float32)
        p = self._fb_params(pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags)
        if prev.strides[0] != nxt.strides[0]:
            prev, nxt = np.ascontiguousarray(prev), np.ascontiguousarray(nxt)
        with self._lock:
            st = self._lib.ofb_farneback(self._h, prev.ctypes.data, nxt.ctypes.data, wid, hgt, prev.strides[0],
                                         out.ctypes.data, 0, C.byref(p))
            _lib.check(st, self._h)
        return out

    def farneback_batch(self, prevs: Sequence[np.ndarray], nexts: Sequence[np.ndarray], **kw) -> np.ndarray:
        """n independent pairs in one batched launch sequence → float32 [n,H,W,2]."""
        n = len(prevs)
        if n != len(nexts) or n == 0:
            raise OfbError(1, "prevs and nexts must be equally long and non-empty")
        prevs = [np.ascontiguousarray(_u8_image(a, "prev")) for a in prevs]
        nexts = [np.ascontiguousarray(_u8_image(a, "next")) for a in nexts]
        hgt, wid = prevs[0].shape
        for a in prevs + nexts:
            if a.shape != (hgt, wid):
                raise OfbError(1, "all images of a batch must have the same size")
        flags = int(kw.get("flags", 0))
        out = np.empty((n, hgt, wid, 2), np.float32)
        if flags & OPTFLOW_USE_INITIAL_FLOW:
            out[...] = np.asarray(kw["flow"], np.float32)
        p = self._fb_params(kw.get("pyr_scale", 0.5), kw.get("levels", 3), kw.get("winsize", 15),
                            kw.get("iterations", 3), kw.get("poly_n", 5), kw.get("poly_sigma", 1.2), flags)
        pp = (C.c_void_p * n)(*[a.ctypes.data for a in prevs])
        nn = (C.c_void_p * n)(*[a.ctypes.data for a in nexts])
        ff = (C.c_void_p * n)(*[out[i].ctypes.data for i in range(n)])
        with self._lock:
            st = self._lib.ofb_farneback_batch(self._h, n, pp, nn, wid, hgt, wid, ff, 0, C.byref(p))
            _lib.check(st, self._h)
        return out

    def wait(self):
        """Block until everything enqueued by the asynchronous calls has finished (ofb_wait)."""
        with self._lock:
            _lib.check(self._lib.ofb_wait(self._h), self._h)
            self._pending.clear()

    def farneback_batch_into(self, prevs: np.ndarray, nexts: np.ndarray, out: np.ndarray, wait: bool = True,
                             **kw) -> np.ndarray:
        """Zero-copy variant of farneback_batch: prevs/nexts uint8 [n,H,W], out float32 [n,H,W,2], all
        C-contiguous (pinned memory is DMA'd directly, pageable memory is staged by the library).
        ``wait=False`` (page-locked buffers only) returns as soon as the work is enqueued: successive
        calls then pipeline across calls — keep the buffers untouched until :meth:`wait`."""
        n, hgt, wid = prevs.shape
        if (nexts.shape != prevs.shape or out.shape != (n, hgt, wid, 2) or prevs.dtype != np.uint8
                or nexts.dtype != np.uint8 or out.dtype != np.float32
                or not (prevs.flags.c_contiguous and nexts.flags.c_contiguous and out.flags.c_contiguous)):
            raise OfbError(1, "farneback_batch_into: need contiguous uint8 [n,H,W] x2 and float32 [n,H,W,2]")
        p = self._fb_params(kw.get("pyr_scale", 0.5), kw.get("levels", 3), kw.get("winsize", 15),
                            kw.get("iterations", 3), kw.get("poly_n", 5), kw.get("poly_sigma", 1.2),
                            kw.get("flags", 0))
        ist, fst = hgt * wid, hgt * wid * 8
        pp = (C.c_void_p * n)(*[prevs.ctypes.data + i * ist for i in range(n)])
        nn = (C.c_void_p * n)(*[nexts.ctypes.data + i * ist for i in range(n)])
        ff = (C.c_void_p * n)(*[out.ctypes.data + i * fst for i in range(n)])
        fn = self._lib.ofb_farneback_batch if wait else self._lib.ofb_farneback_batch_async
        with self._lock:
            st = fn(self._h, n, pp, nn, wid, hgt, wid, ff, 0, C.byref(p))
            _lib.check(st, self._h)
        return out

    def farneback_batch_stats(self, prevs: np.ndarray, nexts: np.ndarray, mask: Optional[np.ndarray] = None,
                              wait: bool = True, **kw):
        """The node contract in one call: flow of n pairs (uint8 [n,H,W] each) reduced on the device to
        (mean_u [n] float64, median_u [n] float32); the field itself is not downloaded.
        ``wait=False`` (page-locked frames) returns as soon as the work is enqueued
        (ofb_farneback_batch_stats_async): the two returned arrays are filled by :meth:`wait`, and successive
        calls pipeline across calls — keep the frames untouched until then."""
        n, hgt, wid = prevs.shape
        if (nexts.shape != prevs.shape or prevs.dtype != np.uint8 or nexts.dtype != np.uint8
                or not (prevs.flags.c_contiguous and nexts.flags.c_contiguous)):
            raise OfbError(1, "farneback_batch_stats: need contiguous uint8 [n,H,W] x2")
        p = self._fb_params(kw.get("pyr_scale", 0.5), kw.get("levels", 3), kw.get("winsize", 15),
                            kw.get("iterations", 3), kw.get("poly_n", 5), kw.get("poly_sigma", 1.2),
                            kw.get("flags", 0))
        ist = hgt * wid
        pp = (C.c_void_p * n)(*[prevs.ctypes.data + i * ist for i in range(n)])
        nn = (C.c_void_p * n)(*[nexts.ctypes.data + i * ist for i in range(n)])
        mptr = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            if mask.shape != (hgt, wid):
                raise OfbError(1, "mask must be uint8 [H,W]")
            mptr = mask.ctypes.data
        mean = np.zeros(n, np.float64)
        med = np.zeros(n, np.float32)
        with self._lock:
            if wait:
                st = self._lib.ofb_farneback_batch_stats(self._h, n, pp, nn, wid, hgt, wid, C.byref(p), mptr,
                                                         mean.ctypes.data_as(C.POINTER(C.c_double)),
                                                         med.ctypes.data_as(C.POINTER(C.c_float)))
            else:
                # the library writes into these arrays at ofb_wait: keep them (and the frames) alive until then
                self._pending.append((mean, med, prevs, nexts))
                st = self._lib.ofb_farneback_batch_stats_async(self._h, n, pp, nn, wid, hgt, wid, C.byref(p), mptr,
                                                               mean.ctypes.data, med.ctypes.data)
            _lib.check(st, self._h)
        return mean, med

    def farneback_stream(self, frames: np.ndarray, download: bool = True, out: Optional[np.ndarray] = None,
                         **kw) -> Optional[np.ndarray]:
        """Camera-stream call (ofb_farneback_stream): ``frames`` uint8 [n,H,W] = the new frame of each of n streams.
        Returns None while priming (first call, or after a change of n / size / parameters / :meth:`stream_reset`),
        then the flow of (previous frame, frame) per stream as float32 [n,H,W,2] — or True with ``download=False``
        (fields stay on the device for :meth:`flow_u_stats` / :meth:`flow_postfilter`)."""
        frames = np.asarray(frames)
        if frames.ndim == 2:
            frames = frames[None]
        if frames.dtype != np.uint8 or frames.ndim != 3 or not frames.flags.c_contiguous:
            raise OfbError(1, "farneback_stream: need contiguous uint8 [n,H,W]")
        n, hgt, wid = frames.shape
        p = self._fb_params(kw.get("pyr_scale", 0.5), kw.get("levels", 3), kw.get("winsize", 15),
                            kw.get("iterations", 3), kw.get("poly_n", 5), kw.get("poly_sigma", 1.2),
                            kw.get("flags", 0))
        fp = (C.c_void_p * n)(*[frames.ctypes.data + i * hgt * wid for i in range(n)])
        if download and out is None:
            out = np.empty((n, hgt, wid, 2), np.float32)
        if download and (out.shape != (n, hgt, wid, 2) or out.dtype != np.float32 or not out.flags.c_contiguous):
            raise OfbError(1, "farneback_stream: out must be contiguous float32 [n,H,W,2]")
        op = (C.c_void_p * n)(*[out.ctypes.data + i * hgt * wid * 8 for i in range(n)]) if download else None
        got = C.c_int(0)
        with self._lock:
            if not download:
                self._pending.append((frames,))      # asynchronous: the upload may still be running when the call returns
            st = self._lib.ofb_farneback_stream(self._h, n, fp, wid, hgt, wid, op, 0, C.byref(p), C.byref(got))
            _lib.check(st, self._h)
        if got.value == 0:
            return None
        return out if download else True

    def farneback_stream_device(self, n: int, d_frames: int, width: int, height: int, pitch: int, image_stride: int,
                                d_flow: int, **kw) -> int:
        """Asynchronous device-pointer form (ofb_farneback_stream_device); returns the number of fields produced."""
        p = self._fb_params(kw.get("pyr_scale", 0.5), kw.get("levels", 3), kw.get("winsize", 15),
                            kw.get("iterations", 3), kw.get("poly_n", 5), kw.get("poly_sigma", 1.2),
                            kw.get("flags", 0))
        got = C.c_int(0)
        with self._lock:
            st = self._lib.ofb_farneback_stream_device(self._h, n, d_frames, width, height, pitch, image_stride, d_flow,
                                                       C.byref(p), C.byref(got))
            _lib.check(st, self._h)
        return got.value

    def stream_reset(self):
        with self._lock:
            _lib.check(self._lib.ofb_stream_reset(self._h), self._h)

    def farneback_device(self, n: int, d_prev: int, d_next: int, width: int, height: int, pitch: int,
                         image_stride: int, d_flow: int, sequence: bool = False, **kw):
        """Asynchronous device-pointer call on the handle's stream (see ofb_farneback_device /
        ofb_farneback_sequence_device)."""
        p = self._fb_params(kw.get("pyr_scale", 0.5), kw.get("levels", 3), kw.get("winsize", 15),
                            kw.get("iterations", 3), kw.get("poly_n", 5), kw.get("poly_sigma", 1.2),
                            kw.get("flags", 0))
        with self._lock:
            if sequence:
                st = self._lib.ofb_farneback_sequence_device(self._h, n, d_prev, width, height, pitch, image_stride,
                                                             d_flow, C.byref(p))
            else:
                st = self._lib.ofb_farneback_device(self._h, n, d_prev, d_next, width, height, pitch, image_stride,
                                                    d_flow, C.byref(p))
            _lib.check(st, self._h)

    def flow_u_stats(self, n: int = 1, mask: Optional[np.ndarray] = None, mean=True, median=True, wait: bool = True):
        """Mean / median of the u component of the most recent flow field(s), reduced on the
        device (the node contract, lfn3_sub_node.py:205-212).  ``wait=False`` (ofb_flow_u_stats_async) returns two
        NumPy arrays that :meth:`wait` fills."""
        mp = None
        if mask is not None:
            mask = np.ascontiguousarray(np.asarray(mask).astype(np.uint8))
            mp = mask.ctypes.data
        if not wait:
            om = np.zeros(n, np.float64)
            od = np.zeros(n, np.float32)
            with self._lock:
                self._pending.append((om, od, mask))
                st = self._lib.ofb_flow_u_stats_async(self._h, n, mp, om.ctypes.data if mean else None,
                                                      od.ctypes.data if median else None)
                _lib.check(st, self._h)
            return (om if mean else None), (od if median else None)
        om = (C.c_double * n)() if mean else None
        od = (C.c_float * n)() if median else None
        with self._lock:
            st = self._lib.ofb_flow_u_stats(self._h, n, mp, om, od)
            _lib.check(st, self._h)
        return (list(om) if mean else None), (list(od) if median else None)

    # -- sparse
    def flow_postfilter(self, n: int = 1, median_ksize: int = 0, magnitude_threshold: Optional[float] = None,
                        gray: Optional[np.ndarray] = None, intensity_threshold: Optional[int] = None):
        """The adapt node's post-processing (lfn3_adapt_node.py:236-251) of the last flow result, on the device:
        ``cv2.medianBlur`` of u and v (ksize 3 | 5), magnitude threshold mask, intensity mask from ``gray``
        (uint8 [H,W] or [n,H,W]).  The filtered field replaces the handle's current one (see
        :meth:`flow_u_stats`, :meth:`flow_download`)."""
        gp = None
        stride = 0
        if intensity_threshold is not None:
            if gray is None:
                raise OfbError(1, "intensity mask needs the gray frame(s)")
            g = np.ascontiguousarray(gray, dtype=np.uint8)
            if g.ndim == 2:
                g = g[None]
            if g.ndim != 3 or g.shape[0] < n:
                raise OfbError(1, "gray must be uint8 [H,W] or [n,H,W]")
            gp = (C.c_void_p * n)(*[g.ctypes.data + i * g.shape[1] * g.shape[2] for i in range(n)])
            stride = g.shape[2]
        with self._lock:
            st = self._lib.ofb_flow_postfilter(self._h, n, int(median_ksize),
                                               -1.0 if magnitude_threshold is None else float(magnitude_threshold),
                                               gp, stride, 0 if intensity_threshold is None else int(intensity_threshold))
            _lib.check(st, self._h)

    def flow_sample(self, points, pair: int = 0) -> np.ndarray:
        """(dx, dy) of the handle's current field at integer pixel positions ``points`` [n,2] (x, y); NaN rows for
        positions outside the frame (ofb_flow_sample)."""
        pts = np.ascontiguousarray(np.asarray(points).reshape(-1, 2), dtype=np.int32)
        out = np.empty((pts.shape[0], 2), np.float32)
        with self._lock:
            _lib.check(self._lib.ofb_flow_sample(self._h, int(pair), pts.shape[0], pts.ctypes.data, out.ctypes.data), self._h)
        return out

    def flow_download(self, n: int, height: int, width: int) -> np.ndarray:
        """The handle's current field(s) as float32 [n,H,W,2] (ofb_flow_download)."""
        out = np.empty((n, height, width, 2), np.float32)
        pp = (C.c_void_p * n)(*[out.ctypes.data + i * height * width * 8 for i in range(n)])
        with self._lock:
            _lib.check(self._lib.ofb_flow_download(self._h, n, pp, 0), self._h)
        return out

    def cvt_gray(self, frame, rgb: bool = False) -> np.ndarray:
        """cv2.cvtColor(frame, COLOR_BGR2GRAY) (or RGB2GRAY with ``rgb=True``) of a uint8 [H,W,3] frame on
        the device, bit-exact with cv2 (the ingest step of the nodes: lfn3_sub_node.py:148-159)."""
        frame = np.asarray(frame)
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise OfbError(1, "cvt_gray needs a uint8 [H,W,3] frame")
        if frame.strides[2] != 1 or frame.strides[1] != 3:
            frame = np.ascontiguousarray(frame)
        hgt, wid = frame.shape[:2]
        out = np.empty((hgt, wid), np.uint8)
        with self._lock:
            st = self._lib.ofb_cvt_gray(self._h, frame.ctypes.data, wid, hgt, frame.strides[0], 1 if rgb else 0,
                                        out.ctypes.data, 0)
            _lib.check(st, self._h)
        return out

    def resize(self, image, size) -> np.ndarray:
        """``cv2.resize(image, size)`` (INTER_LINEAR) of a uint8 [H,W] or [H,W,3] image on the device, bit-exact with
        cv2; ``size`` = (width, height) as in cv2."""
        image = np.ascontiguousarray(image)
        if image.dtype != np.uint8 or image.ndim not in (2, 3) or (image.ndim == 3 and image.shape[2] != 3):
            raise OfbError(1, "resize needs a uint8 [H,W] or [H,W,3] image")
        cn = 1 if image.ndim == 2 else 3
        dw, dh = int(size[0]), int(size[1])
        out = np.empty((dh, dw) if cn == 1 else (dh, dw, 3), np.uint8)
        with self._lock:
            st = self._lib.ofb_resize_u8(self._h, image.ctypes.data, image.shape[1], image.shape[0], 0, cn,
                                         out.ctypes.data, dw, dh, 0)
            _lib.check(st, self._h)
        return out

    def ingest_gray(self, frame, size=None, rgb: bool = False) -> np.ndarray:
        """The nodes' ingest in one call: a uint8 [H,W,3] bgr8 (``rgb=True``: rgb8) frame → ``cv2.resize`` to ``size`` =
        (width, height) if it has another size → ``cv2.cvtColor(..., COLOR_BGR2GRAY)``; bit-exact with cv2."""
        frame = np.asarray(frame)
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise OfbError(1, "ingest_gray needs a uint8 [H,W,3] frame")
        if frame.strides[2] != 1 or frame.strides[1] != 3:
            frame = np.ascontiguousarray(frame)
        hgt, wid = frame.shape[:2]
        dw, dh = (wid, hgt) if size is None else (int(size[0]), int(size[1]))
        out = np.empty((dh, dw), np.uint8)
        with self._lock:
            st = self._lib.ofb_ingest_gray(self._h, frame.ctypes.data, wid, hgt, frame.strides[0], 1 if rgb else 0,
                                           out.ctypes.data, dw, dh, 0)
            _lib.check(st, self._h)
        return out

    @staticmethod
    def _jpeg_bytes(buf):
        buf = np.ascontiguousarray(np.frombuffer(buf, np.uint8) if isinstance(buf, (bytes, bytearray, memoryview)) else
                                   np.asarray(buf, np.uint8).reshape(-1))
        return buf

    def jpeg_info(self, buf):
        """(width, height, components) of a baseline JPEG stream (header parse on the host, no device work)."""
        buf = self._jpeg_bytes(buf)
        w, h, c = C.c_int(), C.c_int(), C.c_int()
        st = self._lib.ofb_jpeg_info(buf.ctypes.data, buf.size, C.byref(w), C.byref(h), C.byref(c))
        if st:
            raise OfbError(st, "not a baseline JPEG stream the device path decodes")
        return w.value, h.value, c.value

    def jpeg_host_entropy(self, on: bool):
        """Walk the Huffman stream on the host (True) instead of on the device (default for scans without restart
        intervals); same result."""
        _lib.check(self._lib.ofb_jpeg_set_host_entropy(self._h, 1 if on else 0), self._h)

    def imdecode_grayscale(self, buf) -> np.ndarray:
        """``cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE)`` of a baseline JPEG: the luma plane, uint8 [H,W]."""
        buf = self._jpeg_bytes(buf)
        w, h, _ = self.jpeg_info(buf)
        out = np.empty((h, w), np.uint8)
        with self._lock:
            _lib.check(self._lib.ofb_jpeg_decode_luma(self._h, buf.ctypes.data, buf.size, out.ctypes.data, 0), self._h)
        return out

    def imdecode(self, buf, gray: bool = False) -> np.ndarray:
        """``cv2.imdecode(buf, cv2.IMREAD_COLOR)`` of a baseline JPEG (``sensor_msgs/CompressedImage.data``; the
        compressed-image node, opticalflow_comprerssed_node.py:43-46): uint8 [H,W,3] BGR, bit-exact with the wheel's
        libjpeg-turbo.  ``gray=True``: ``cv2.cvtColor(that, COLOR_BGR2GRAY)`` instead, uint8 [H,W] (a third of the bytes
        back over PCIe).  Raises ``OfbError`` (status 6) for streams the device path does not decode."""
        buf = self._jpeg_bytes(buf)
        w, h, _ = self.jpeg_info(buf)
        out = np.empty((h, w) if gray else (h, w, 3), np.uint8)
        with self._lock:
            st = self._lib.ofb_jpeg_decode(self._h, buf.ctypes.data, buf.size, None if gray else out.ctypes.data, 0,
                                           out.ctypes.data if gray else None, 0)
            _lib.check(st, self._h)
        return out

    def ingest_jpeg_gray(self, buf, size=None) -> np.ndarray:
        """``ingest_gray`` for a compressed frame: decode → ``cv2.resize`` of the colour frame to ``size`` = (width, height)
        if it has another size → ``cv2.cvtColor(..., COLOR_BGR2GRAY)``; the uint8 [H,W] frame the flow calls take."""
        buf = self._jpeg_bytes(buf)
        w, h, _ = self.jpeg_info(buf)
        dw, dh = (w, h) if size is None else (int(size[0]), int(size[1]))
        out = np.empty((dh, dw), np.uint8)
        with self._lock:
            st = self._lib.ofb_ingest_jpeg_gray(self._h, buf.ctypes.data, buf.size, out.ctypes.data, dw, dh, 0)
            _lib.check(st, self._h)
        return out

    @staticmethod
    def _junction_args(img, grid_area, grid_area_threshold, eps, dampen):
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and img.shape[2] != 3):
            raise OfbError(1, "the junction detector needs a uint8 [H,W] or [H,W,3] frame")
        cn = 1 if img.ndim == 2 else 3
        if img.strides[-1] != 1 or (cn == 3 and img.strides[1] != 3):
            img = np.ascontiguousarray(img)
        p = JunctionParams(int(grid_area), float(grid_area_threshold), int(eps), 0, 0.0, 0.0)
        if dampen is not None:
            p.dampen, p.dampen_min, p.dampen_max = 1, float(dampen[0]), float(dampen[1])
        return img, cn, p

    def find_junctions(self, img, grid_area: int = 250, grid_area_threshold: float = 2.0, eps: int = 4, dampen=None,
                       return_candidates: bool = False):
        """``find_junctions_not_rotated(img, grid_area, grid_area_threshold, false, eps)`` of the reference's junction
        detector (junction_detector.cpp:31-214) on a uint8 gray or bgr8 frame; ``dampen=(min, max)`` applies
        ``dampenIntensity`` first (:3-28; the ROS node uses (-20, 15), grid_area 200, eps 6).  Returns the junction
        centres, float32 [n, 2] (x, y) in the reference's order (and the box corners before clustering if asked)."""
        img, cn, p = self._junction_args(img, grid_area, grid_area_threshold, eps, dampen)
        hgt, wid = img.shape[:2]
        cap = max(4096, (wid * hgt) // 128)         # a junction per ~128 px is far beyond any net; grown on demand
        while True:
            out = np.empty((cap, 2), np.float32)
            cand = np.empty((4 * cap, 2), np.float32)
            n, nc = C.c_int(), C.c_int()
            with self._lock:
                st = self._lib.ofb_find_junctions(self._h, img.ctypes.data, wid, hgt, img.strides[0], cn, C.byref(p),
                                                  out.ctypes.data, cap, C.byref(n), cand.ctypes.data, 4 * cap, C.byref(nc))
                if st == 4 and cap < (1 << 20) and max(n.value, nc.value // 4) > cap:
                    cap *= 8
                    continue
                _lib.check(st, self._h)
            break
        res = out[:n.value].copy()
        return (res, cand[:nc.value].copy()) if return_candidates else res

    def junction_threshold(self, img, dampen=None) -> np.ndarray:
        """The detector's binary image: gray -> ``GaussianBlur(3x3)`` -> ``adaptiveThreshold(GAUSSIAN_C, BINARY, 11, 2)``."""
        img, cn, p = self._junction_args(img, 250, 2.0, 4, dampen)
        hgt, wid = img.shape[:2]
        out = np.empty((hgt, wid), np.uint8)
        with self._lock:
            st = self._lib.ofb_junction_threshold(self._h, img.ctypes.data, wid, hgt, img.strides[0], cn, C.byref(p),
                                                  out.ctypes.data, 0)
            _lib.check(st, self._h)
        return out

    def clahe(self, image, clipLimit: float = 40.0, tileGridSize=(8, 8)) -> np.ndarray:
        """``cv2.createCLAHE(clipLimit, tileGridSize).apply(image)`` of a uint8 [H,W] image on the device, bit-exact with
        cv2 (the adapt node's pre-filter, lfn3_adapt_node.py:164-182)."""
        image = np.asarray(image)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise OfbError(1, "clahe needs a uint8 [H,W] image")
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        out = np.empty(image.shape, np.uint8)
        with self._lock:
            st = self._lib.ofb_clahe(self._h, image.ctypes.data, image.shape[1], image.shape[0], image.strides[0],
                                     float(clipLimit), int(tileGridSize[0]), int(tileGridSize[1]), out.ctypes.data, 0)
            _lib.check(st, self._h)
        return out

    def adapt_prefilter(self, bgr, clipLimit: Optional[float] = None, clip_range=(1.0, 4.0, 0.1, 0.8), tileGridSize=(8, 8)):
        """The adapt node's colour pre-filter (lfn3_adapt_node.py:164-184) on the device: BGR2HSV, CLAHE on V with the
        adaptive clip limit (``clipLimit=None``; ``clip_range`` = (clip_min, clip_max, C_min, C_max)) or a fixed one,
        HSV2RGB.  Returns (rgb uint8 [H,W,3], clip limit used); bit-exact with cv2."""
        bgr = np.asarray(bgr)
        if bgr.dtype != np.uint8 or bgr.ndim != 3 or bgr.shape[2] != 3:
            raise OfbError(1, "adapt_prefilter needs a uint8 [H,W,3] frame")
        if bgr.strides[2] != 1 or bgr.strides[1] != 3:
            bgr = np.ascontiguousarray(bgr)
        hgt, wid = bgr.shape[:2]
        out = np.empty((hgt, wid, 3), np.uint8)
        cp = _lib.ClaheParams(1 if clipLimit is None else 0, 0.0 if clipLimit is None else float(clipLimit),
                              float(clip_range[0]), float(clip_range[1]), float(clip_range[2]), float(clip_range[3]),
                              int(tileGridSize[0]), int(tileGridSize[1]))
        used = C.c_double(0.0)
        with self._lock:
            st = self._lib.ofb_adapt_prefilter(self._h, bgr.ctypes.data, wid, hgt, bgr.strides[0], C.byref(cp),
                                               out.ctypes.data, 0, C.byref(used))
            _lib.check(st, self._h)
        return out, used.value

    def bilateral_filter(self, src, d: int, sigmaColor: float, sigmaSpace: float) -> np.ndarray:
        """``cv2.bilateralFilter(src, d, sigmaColor, sigmaSpace)`` on a uint8 [H,W,3] frame (the last, optional step of the
        adapt node's pre-filter, lfn3_adapt_node.py:186-190).  OpenCV's own algorithm, bit-exact with
        ``oracle/prefilter_np.py::bilateral_u8c3``; the installed wheel (Intel IPP for 8-bit images) differs by one at
        rounding ties in a few values per 100 000."""
        src = np.asarray(src)
        if src.dtype != np.uint8 or src.ndim != 3 or src.shape[2] != 3:
            raise OfbError(1, "bilateral_filter needs a uint8 [H,W,3] frame")
        if src.strides[2] != 1 or src.strides[1] != 3:
            src = np.ascontiguousarray(src)
        hgt, wid = src.shape[:2]
        out = np.empty((hgt, wid, 3), np.uint8)
        with self._lock:
            st = self._lib.ofb_bilateral_u8c3(self._h, src.ctypes.data, wid, hgt, src.strides[0], int(d), float(sigmaColor),
                                              float(sigmaSpace), out.ctypes.data, 0)
            _lib.check(st, self._h)
        return out

    def good_features(self, image, maxCorners, qualityLevel, minDistance, blockSize=3, mask=None, useHarrisDetector=False,
                      k=0.04) -> np.ndarray:
        image = _u8_image(image, "image")
        hgt, wid = image.shape
        cap = int(maxCorners) if maxCorners > 0 else hgt * wid
        out = np.empty((max(cap, 1), 2), np.float32)
        n = C.c_int(0)
        p = GfttParams(int(maxCorners), float(qualityLevel), float(minDistance), int(blockSize), 1 if useHarrisDetector else 0,
                       float(k))
        mptr, mstride = None, 0
        if mask is not None:
            mask = _u8_image(mask, "mask")
            if mask.shape != image.shape:
                raise OfbError(1, "goodFeaturesToTrack: mask must have the image's size")
            mptr, mstride = mask.ctypes.data, mask.strides[0]
        with self._lock:
            st = self._lib.ofb_good_features_masked(self._h, image.ctypes.data, wid, hgt, image.strides[0], mptr, mstride,
                                                    C.byref(p), out.ctypes.data, C.byref(n))
            _lib.check(st, self._h)
        return out[:n.value].reshape(-1, 1, 2).copy()

    def flow_to_color(self, height: int, width: int, pair: int = 0) -> np.ndarray:
        """The nodes' ``flow_to_color`` (sub_n_pub_lfn3_node.py:132-140) of the handle's current field ``pair`` (of
        ``height`` x ``width``, as for :meth:`flow_download`), computed on the device → uint8 [H,W,3] BGR."""
        out = np.empty((height, width, 3), np.uint8)
        with self._lock:
            st = self._lib.ofb_flow_to_bgr(self._h, int(pair), out.ctypes.data, 0)
            _lib.check(st, self._h)
        return out

    def flow_to_color_speed(self, height: int, width: int, dt: float, pixel_to_meter: float, max_speed: float,
                            pair: int = 0) -> np.ndarray:
        """The sub node's dense view (lfn3_sub_node.py:244-262) of the current field: hue from the angle, value =
        ``clip(|flow| / dt * pixel_to_meter / max_speed, 0, 1) * 255`` → uint8 [H,W,3] BGR, on the device."""
        out = np.empty((height, width, 3), np.uint8)
        with self._lock:
            st = self._lib.ofb_flow_to_bgr_speed(self._h, int(pair), float(dt), float(pixel_to_meter), float(max_speed),
                                                 out.ctypes.data, 0)
            _lib.check(st, self._h)
        return out

    def lk_stream(self, frame, maxCorners=2000, qualityLevel=0.01, minDistance=7, blockSize=3, winSize=(21, 21),
                  maxLevel=3, criteria=(3, 30, 0.01), minEigThreshold=1e-4, flags=0):
        """Camera-stream form of goodFeaturesToTrack + calcOpticalFlowPyrLK (ofb_lk_stream): one new frame per call, the
        previous frame's pyramid, derivatives and corners stay on the GPU.  Returns None on the priming call, then
        ``(prevPts [N,1,2], nextPts [N,1,2], status [N,1], err [N,1])`` for (previous frame -> frame); the corners of
        ``frame`` (tracked by the next call) are in :attr:`last_corners`."""
        frame = _u8_image(frame, "frame")
        hgt, wid = frame.shape
        cap = int(maxCorners)
        if cap <= 0:
            raise OfbError(1, "lk_stream: maxCorners must be > 0")
        gp = GfttParams(cap, float(qualityLevel), float(minDistance), int(blockSize), 0, 0.04)
        ctype, max_count, eps = criteria
        if not (ctype & 1):
            max_count = 30
        if not (ctype & 2):
            eps = 0.01
        lp = LKParams(int(winSize[0]), int(winSize[1]), int(maxLevel), int(max_count), float(eps), int(flags),
                      float(minEigThreshold))
        prev = np.empty((cap, 2), np.float32)
        nxt = np.empty((cap, 2), np.float32)
        status = np.zeros((cap,), np.uint8)
        err = np.zeros((cap,), np.float32)
        new = np.empty((cap, 2), np.float32)
        n, n_new = C.c_int(0), C.c_int(0)
        with self._lock:
            st = self._lib.ofb_lk_stream(self._h, frame.ctypes.data, wid, hgt, frame.strides[0], C.byref(gp), C.byref(lp),
                                         prev.ctypes.data, nxt.ctypes.data, status.ctypes.data, err.ctypes.data,
                                         C.byref(n), new.ctypes.data, C.byref(n_new))
            _lib.check(st, self._h)
        self.last_corners = new[:n_new.value].reshape(-1, 1, 2).copy()
        if n.value < 0:
            return None
        k = n.value
        return (prev[:k].reshape(-1, 1, 2).copy(), nxt[:k].reshape(-1, 1, 2).copy(), status[:k].reshape(-1, 1).copy(),
                err[:k].reshape(-1, 1).copy())

    def lk_stream_reset(self):
        with self._lock:
            _lib.check(self._lib.ofb_lk_stream_reset(self._h), self._h)

    def corner_min_eigenval(self, image, blockSize=3) -> np.ndarray:
        image = _u8_image(image, "image")
        hgt, wid = image.shape
        out = np.empty((hgt, wid), np.float32)
        with self._lock:
            st = self._lib.ofb_corner_min_eigenval(self._h, image.ctypes.data, wid, hgt, image.strides[0],
                                                   int(blockSize), out.ctypes.data)
            _lib.check(st, self._h)
        return out

    def lk_pyramid(self, image, winSize=(21, 21), maxLevel=3, with_derivatives=True):
        """(levels, derivs): uint8 pyrDown chain and int16 Scharr (dx,dy) per level."""
        image = _u8_image(image, "image")
        hgt, wid = image.shape
        sizes = [(hgt, wid)]
        for _ in range(maxLevel):
            sizes.append(((sizes[-1][0] + 1) // 2, (sizes[-1][1] + 1) // 2))
        lv = [np.empty(s, np.uint8) for s in sizes]
        dv = [np.empty(s + (2,), np.int16) for s in sizes] if with_derivatives else None
        lp = (C.c_void_p * len(lv))(*[a.ctypes.data for a in lv])
        dp = (C.c_void_p * len(lv))(*[a.ctypes.data for a in dv]) if dv else None
        n = C.c_int(0)
        with self._lock:
            st = self._lib.ofb_lk_pyramid(self._h, image.ctypes.data, wid, hgt, image.strides[0], int(winSize[0]),
                                          int(winSize[1]), int(maxLevel), lp, dp, C.byref(n))
            _lib.check(st, self._h)
        return lv[:n.value], (dv[:n.value] if dv else None)

    def pyrlk(self, prev, nxt, prevPts, nextPts=None, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01), flags=0,
              minEigThreshold=1e-4):
        prev = _u8_image(prev, "prevImg")
        nxt = _u8_image(nxt, "nextImg")
        if prev.shape != nxt.shape:
            raise OfbError(1, "prevImg and nextImg must have the same size")
        if prev.strides[0] != nxt.strides[0]:
            prev, nxt = np.ascontiguousarray(prev), np.ascontiguousarray(nxt)
        pts = np.ascontiguousarray(np.asarray(prevPts, np.float32).reshape(-1, 2))
        n = pts.shape[0]
        if flags & OPTFLOW_USE_INITIAL_FLOW:
            nxp = np.ascontiguousarray(np.asarray(nextPts, np.float32).reshape(-1, 2)).copy()
            if nxp.shape != pts.shape:
                raise OfbError(1, "nextPts must match prevPts with OPTFLOW_USE_INITIAL_FLOW")
        else:
            nxp = np.empty_like(pts)
        status = np.zeros((n,), np.uint8)
        err = np.zeros((n,), np.float32)
        ctype, max_count, eps = criteria
        if not (ctype & 1):
            max_count = 30
        if not (ctype & 2):
            eps = 0.01
        p = LKParams(int(winSize[0]), int(winSize[1]), int(maxLevel), int(max_count), float(eps), int(flags),
                     float(minEigThreshold))
        hgt, wid = prev.shape
        with self._lock:
            st = self._lib.ofb_pyrlk(self._h, prev.ctypes.data, nxt.ctypes.data, wid, hgt, prev.strides[0],
                                     pts.ctypes.data, n, nxp.ctypes.data, status.ctypes.data, err.ctypes.data,
                                     C.byref(p))
            _lib.check(st, self._h)
        return nxp.reshape(-1, 1, 2), status.reshape(-1, 1), err.reshape(-1, 1)
