"""Host-side mirror of the reference nodes' frame → velocity contract, without rclpy.

Every reference node does the same thing around its flow call:
``ros2_ws/src/liteflownet3/liteflownet3/lfn3_sub_node.py:141-222`` (median),
``ros2_ws/src/optical_flow/optical_flow/opticalflow_node.py:41-128`` (mean),
``ros2_ws/src/liteflownet3/liteflownet3/sub_n_pub_lfn3_node.py:195-210`` (masked median):
first frame primes the state; then ``dt = stamp - prev_stamp`` (``<= 0 → 1e-3``), flow,
``u_avg = median|mean(flow[0]) / dt``, ``vx = u_avg * pixel_to_meter``, a ``deque(maxlen=5)``
mean for the smooth topic, and a ``Vector3Stamped{stamp = image stamp, frame_id 'camera_link',
vector = (vx, 0, 0)}``.  ``FarnebackVelocityNode`` is that logic with the flow call swapped for
the B200 engine; messages are plain dataclasses so the class runs (and is tested) without ROS.
``examples/farneback_sub_node.py`` shows the same class wired to rclpy.

``FarnebackVelocityNode.compressed_callback`` is the same for ``sensor_msgs/CompressedImage`` (JPEG) frames
(``ros2_ws/src/optical_flow/optical_flow/opticalflow_comprerssed_node.py:41-62``: ``cv2.imdecode`` first), and
``JunctionDetectorNode`` mirrors the C++ detector node (``ros2_ws/src/junction_point_detector/src/fishnet_detector_ros.cpp:30-80``:
``dampenIntensity(img, -20, 15)``, ``find_junctions_not_rotated(img, 200, 2.0, false, 6)``, a ``PointCloud`` with z = 0,
nothing published for fewer than 4 junctions).
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

from ._lib import OfbError
from .engine import FlowEngine


@dataclass
class Vector3Stamped:
    """geometry_msgs/Vector3Stamped look-alike."""
    stamp: float
    frame_id: str = "camera_link"
    vector: Tuple[float, float, float] = (0.0, 0.0, 0.0)


def to_gray_u8(image: np.ndarray, encoding: str = "bgr8") -> np.ndarray:
    """sensor_msgs/Image payload → gray uint8 with cv2's fixed-point BGR2GRAY
    ``(B*3735 + G*19235 + R*9798 + 16384) >> 15`` (probe-verified)."""
    image = np.asarray(image)
    if image.ndim == 2:
        return np.ascontiguousarray(image.astype(np.uint8, copy=False))
    if encoding == "bgr8":
        b, g, r = image[..., 0], image[..., 1], image[..., 2]
    elif encoding == "rgb8":
        r, g, b = image[..., 0], image[..., 1], image[..., 2]
    else:
        raise ValueError("Unsupported image encoding: %s" % encoding)
    y = (b.astype(np.uint32) * 3735 + g.astype(np.uint32) * 19235 + r.astype(np.uint32) * 9798 + 16384) >> 15
    return np.ascontiguousarray(y.astype(np.uint8))


@dataclass
class FarnebackVelocityNode:
    width: int = 640
    height: int = 480
    pixel_to_meter: float = 0.0011
    reduce: str = "median"        # 'median' (lfn3_sub_node.py:207) or 'mean' (opticalflow_node.py:98)
    device: int = 0
    pyr_scale: float = 0.5
    levels: int = 3
    winsize: int = 15
    iterations: int = 3
    poly_n: int = 5
    poly_sigma: float = 1.2
    flags: int = 0
    on_device_reduce: bool = False
    use_stream: bool = False      # keep the previous frame's state on the GPU (ofb_farneback_stream): one upload per frame
    # the "adapt" node's post-processing of the field (lfn3_adapt_node.py:236-251), applied on the device when set
    median_kernel_size: int = 0                       # 3 | 5: cv2.medianBlur of u and v
    flow_magnitude_threshold: Optional[float] = None  # u, v *= (|flow| >= threshold)
    intensity_threshold: Optional[int] = None         # u, v *= (gray < threshold)
    engine: Optional[FlowEngine] = None
    prev_gray: Optional[np.ndarray] = field(default=None, repr=False)
    prev_time: Optional[float] = None
    velocity_buffer: deque = field(default_factory=lambda: deque(maxlen=5), repr=False)
    last_flow: Optional[np.ndarray] = field(default=None, repr=False)
    _stream_primed: bool = field(default=False, repr=False)

    def __post_init__(self):
        if self.engine is None:
            self.engine = FlowEngine(self.width, self.height, 1, self.device)

    def image_callback(self, image: np.ndarray, stamp: float, encoding: str = "bgr8",
                       mask: Optional[np.ndarray] = None) -> Optional[Tuple[Vector3Stamped, Vector3Stamped]]:
        """One camera frame in → (raw, smooth) velocity messages out (None on the priming frame)."""
        image = np.asarray(image)
        if encoding not in ("bgr8", "rgb8", "mono8"):
            raise ValueError("Unsupported image encoding: %s" % encoding)
        if image.ndim == 3 and image.dtype == np.uint8 and image.shape[2] == 3:
            # the nodes' ingest, on the device: cv2.resize to the configured size if needed (lfn3_sub_node.py:152-153),
            # then cv2.cvtColor(..., BGR2GRAY) — one upload, bit-exact with cv2
            gray = self.engine.ingest_gray(image, (self.width, self.height), rgb=encoding == "rgb8")
        else:
            gray = to_gray_u8(image, encoding)
            if gray.shape != (self.height, self.width):
                gray = self.engine.resize(gray, (self.width, self.height))
        return self._on_gray(gray, stamp, mask)

    def compressed_callback(self, data, stamp: float, mask: Optional[np.ndarray] = None):
        """One ``sensor_msgs/CompressedImage`` (JPEG bytes) in → the same messages: ``cv2.imdecode(..., IMREAD_COLOR)``
        (opticalflow_comprerssed_node.py:43-46), resize to the configured size and gray conversion on the device, bit-exact
        with cv2; only the compressed bytes go up and the gray frame comes back.  A stream that cannot be decoded returns
        None, as the node logs and returns when ``cv2.imdecode`` gives None."""
        try:
            gray = self.engine.ingest_jpeg_gray(data, (self.width, self.height))
        except OfbError:
            return None
        return self._on_gray(gray, stamp, mask)

    def _on_gray(self, gray: np.ndarray, stamp: float, mask: Optional[np.ndarray]):
        if self.prev_gray is None:
            self.prev_gray, self.prev_time = gray, stamp
            return None
        dt = stamp - self.prev_time
        if dt <= 0:
            dt = 1e-3
        self.prev_time = stamp
        if self.use_stream:
            kw = dict(pyr_scale=self.pyr_scale, levels=self.levels, winsize=self.winsize, iterations=self.iterations,
                      poly_n=self.poly_n, poly_sigma=self.poly_sigma, flags=self.flags)
            if not self._stream_primed:                   # the engine primes on the frame the node primed on
                self.engine.stream_reset()
                self.engine.farneback_stream(self.prev_gray, **kw)
                self._stream_primed = True
            flow = self.engine.farneback_stream(gray, **kw)[0]
        else:
            flow = self.engine.farneback(self.prev_gray, gray, None, self.pyr_scale, self.levels, self.winsize,
                                         self.iterations, self.poly_n, self.poly_sigma, self.flags)
        if self.median_kernel_size or self.flow_magnitude_threshold is not None or self.intensity_threshold is not None:
            self.engine.flow_postfilter(1, self.median_kernel_size, self.flow_magnitude_threshold,
                                        gray if self.intensity_threshold is not None else None, self.intensity_threshold)
            if not self.on_device_reduce:
                flow = self.engine.flow_download(1, self.height, self.width)[0]
        self.last_flow = flow
        flow_np = np.transpose(flow, (2, 0, 1))      # [2,H,W] as the reference nodes consume it
        if self.on_device_reduce:
            mean, med = self.engine.flow_u_stats(1, mask, mean=self.reduce == "mean", median=self.reduce == "median")
            u = (mean if self.reduce == "mean" else med)[0]
        else:
            u_field = flow_np[0] if mask is None else flow_np[0][np.asarray(mask, bool)]
            if u_field.size == 0:
                self.prev_gray = gray
                return None
            u = float(np.mean(u_field)) if self.reduce == "mean" else float(np.median(u_field))
        vx = float(u / dt * self.pixel_to_meter)
        self.velocity_buffer.append(vx)
        vx_smooth = float(np.mean(self.velocity_buffer))
        self.prev_gray = gray
        return (Vector3Stamped(stamp, "camera_link", (vx, 0.0, 0.0)),
                Vector3Stamped(stamp, "camera_link", (vx_smooth, 0.0, 0.0)))


@dataclass
class PointCloud:
    """sensor_msgs/PointCloud look-alike: points [n, 3] float32 (x, y, 0)."""
    stamp: float
    frame_id: str
    points: np.ndarray


@dataclass
class JunctionDetectorNode:
    """The junction detector node (fishnet_detector_ros.cpp:30-80) on the B200 engine.  The C++ node converts the message
    to rgb8 and hands that buffer to functions written for BGR; the bytes are taken here exactly as the node passes them."""
    grid_area: int = 200
    grid_area_threshold: float = 2.0
    eps: int = 6
    dampen: Optional[Tuple[float, float]] = (-20.0, 15.0)
    device: int = 0
    engine: Optional[FlowEngine] = None

    def __post_init__(self):
        if self.engine is None:
            self.engine = FlowEngine(64, 64, 1, self.device)

    def image_callback(self, image: np.ndarray, stamp: float, frame_id: str = "camera_link") -> Optional[PointCloud]:
        image = np.asarray(image)
        pts = self.engine.find_junctions(image, self.grid_area, self.grid_area_threshold, self.eps,
                                         dampen=self.dampen if image.ndim == 3 else None)
        if len(pts) < 4:
            return None                                   # "No junctions found": nothing is published
        cloud = np.zeros((len(pts), 3), np.float32)
        cloud[:, :2] = pts
        return PointCloud(stamp, frame_id, cloud)


def junction_mask(points: Sequence[Sequence[float]], height: int, width: int, radius: int = 5) -> np.ndarray:
    """±radius squares around junction points (sub_n_pub_lfn3_node.py:195-204)."""
    mask = np.zeros((height, width), dtype=bool)
    for p in points:
        x, y = int(p[0]), int(p[1])
        if 0 <= x < width and 0 <= y < height:
            mask[max(0, y - radius):min(height, y + radius + 1), max(0, x - radius):min(width, x + radius + 1)] = True
    return mask


def junction_velocity(engine: FlowEngine, prev_junctions: Sequence[Sequence[float]],
                      curr_junctions: Sequence[Sequence[float]], dt: float, pixel_to_meter: float,
                      max_dist: float = 5.0, min_matches: int = 4, pair: int = 0) -> Optional[float]:
    """The junction node's velocity (``lfn3_junction_node.py:203-231``) from the engine's current field: every previous
    junction inside the frame is moved by the flow at its integer position (looked up on the device,
    :meth:`FlowEngine.flow_sample`), matched to the nearest current junction closer than ``max_dist`` px, and the mean
    x displacement of at least ``min_matches`` matches gives ``vx = mean_dx / dt * pixel_to_meter`` (None otherwise)."""
    prev = np.asarray(prev_junctions, np.float64).reshape(-1, 2)
    curr = np.asarray(curr_junctions, np.float64).reshape(-1, 2)
    if prev.shape[0] == 0 or curr.shape[0] == 0:
        return None
    d = engine.flow_sample(prev.astype(np.int64), pair).astype(np.float64)   # int(p) truncates, as the node does
    ok = ~np.isnan(d[:, 0])
    if not ok.any():
        return None
    pred = prev[ok] + d[ok]
    dist = np.sqrt(((pred[:, None, :] - curr[None, :, :]) ** 2).sum(-1))      # exact nearest neighbour (KDTree.query k=1)
    idx = dist.argmin(1)
    hit = dist[np.arange(pred.shape[0]), idx] < max_dist
    if int(hit.sum()) < min_matches:
        return None
    disp = curr[idx[hit]] - prev[ok][hit]
    return float(disp[:, 0].mean() / dt * pixel_to_meter)


def flow_arrows(engine: FlowEngine, height: int, width: int, step: int = 20, pair: int = 0) -> np.ndarray:
    """End points of the sub node's arrow overlay (``lfn3_sub_node.py:225-238``: a grid of ``cv2.arrowedLine`` segments
    (x, y) -> (x + int(u), y + int(v)) every ``step`` pixels) from the engine's current field, sampled on the device
    (:meth:`FlowEngine.flow_sample`) instead of downloading the field: int32 [n, 4] (x1, y1, x2, y2) in the node's loop
    order.  Drawing stays with the caller (``cv2.arrowedLine(img, (x1, y1), (x2, y2), (0, 255, 0), 1, tipLength=0.4)``)."""
    ys, xs = np.meshgrid(np.arange(0, height, step), np.arange(0, width, step), indexing="ij")
    pts = np.stack([xs.ravel(), ys.ravel()], -1).astype(np.int64)
    d = engine.flow_sample(pts, pair)
    end = pts + np.trunc(d).astype(np.int64)             # int() truncates towards zero
    return np.concatenate([pts, end], 1).astype(np.int32)


def adaptive_clip_limit(v: np.ndarray, clip_min: float, clip_max: float, c_min: float, c_max: float) -> float:
    """The adapt node's clip limit from the contrast of the V channel (``lfn3_adapt_node.py:170-175``):
    ``contrast = std(v) / (mean(v) + 1e-3)``, mapped linearly from [c_min, c_max] to [clip_min, clip_max] and clipped."""
    contrast = np.std(v) / (np.mean(v) + 1e-3)
    return float(np.clip(clip_min + (contrast - c_min) / (c_max - c_min) * (clip_max - clip_min), clip_min, clip_max))
